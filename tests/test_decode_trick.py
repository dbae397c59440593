"""The fused kernel decodes integers without the IEEE divide sequence (qd_fast.cu div_exact).  The
three-operation form is only valid because it is exact on every value the formats can hold: check all
of them on the CPU (software fma, so the result does not depend on the host having FMA hardware)."""
import ctypes as C

import oracle_lib as O


def test_divide_free_decode_is_exact_for_every_input():
    L = O.lib()
    L.qo_check_div_trick.restype = C.c_int
    assert L.qo_check_div_trick() == 0
