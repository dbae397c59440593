"""The fused kernel decodes integers without the IEEE divide sequence (qd_fast.cu div_exact).  The
three-operation form is only valid because it is exact on every value the formats can hold: check all
of them on the CPU (software fma, so the result does not depend on the host having FMA hardware)."""
import ctypes as C

import oracle_lib as O


def test_divide_free_decode_is_exact_for_every_input():
    L = O.lib()
    L.qo_check_div_trick.restype = C.c_int
    assert L.qo_check_div_trick() == 0


def test_offset_formats_decode_with_one_fused_operation():
    """cu8 / cs16 (lib.rs:252-253): fl(fl(x / den) - off) == fl(x * fl(1/den) - off) rounded ONCE, for every u8 / 255
    and every i16 / 65535 -- what dec_fused2 (qd_fir_kernel.cuh) and decode_sample_packed (qd_device_math.cuh)
    compute with a single packed FMA.  The candidate is evaluated in exact rational arithmetic and rounded to f32
    (nearest, ties to even) by comparing the neighbouring floats exactly; the reference side is numpy's f32 divide
    and subtract, i.e. the two IEEE operations of the reference."""
    from fractions import Fraction

    import numpy as np

    def rn32(fr):
        f = np.float32(float(fr))  # within one ulp; pick the exact nearest among the neighbours
        cands = [np.nextafter(f, np.float32(-np.inf)), f, np.nextafter(f, np.float32(np.inf))]
        return min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(np.float32(c).view(np.uint32)) & 1))

    for values, den, off in ((range(256), 255.0, 127.5), (range(-32768, 32768), 65535.0, 32767.5)):
        c = Fraction(float(np.float32(1) / np.float32(den)))
        x = np.array(list(values), dtype=np.float32)
        ref = (x / np.float32(den) - np.float32(off)).astype(np.float32)
        for v, r in zip(values, ref):
            got = np.float32(rn32(Fraction(v) * c - Fraction(off)))
            assert got.view(np.uint32) == r.view(np.uint32), (v, den, float(got), float(r))
