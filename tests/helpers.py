"""Shared helpers: build the same chain in the oracle (checker) and in the product (checked)."""
from __future__ import annotations

import numpy as np

import oracle_lib as O


def synth_raw(fmt, n, seed=0x5EED0000, first=0, tones=None, noise=6, rate=20e6):
    if tones is None:
        tones = [(1.6e6, 45, 0), (-4.1e6, 30, 0), (0.3e6, 20, 3000)]
    scale = {O.CS8: 1, O.CU8: 1, O.CS16: 200, O.CF32: 300}[fmt]
    p = O.make_synth(seed, [(O.tone_step(f, rate), a * scale, k) for f, a, k in tones], noise * scale)
    return O.synth_fill(p, fmt, first, n), p


def oracle_chain(raw, fmt, rate, stages, base=0, total=0):
    s = O.Samples.from_window(raw, fmt, rate, base, total) if total else O.Samples.from_bytes(raw, fmt, rate)
    for st in stages:
        s = s.shift(st[1]) if st[0] == "shift" else s.lowpass(st[1], st[2], st[3])
    return s


def gpu_chain(raw, fmt, rate, stages, base=0, total=0, precision=None):
    import quadrs_b200 as Q

    s = Q.Samples.from_bytes(raw, fmt, rate, base_sample=base, total_samples=total)
    for st in stages:
        s = s.shift(st[1]) if st[0] == "shift" else s.lowpass(st[1], st[2], st[3])
    if precision is not None:
        s = s.with_precision(precision)
    return s


def bits(a: np.ndarray) -> np.ndarray:
    """f32 words with every NaN canonicalised: x86 and sm_100 generate different default NaN patterns
    (0xFFC00000 vs 0x7FFFFFFF) for the same invalid operation (e.g. the 0/0 centre tap of an odd-length
    filter, filter.rs:87-89); NaN payloads copied from the input are compared by test_decode_bit_exact."""
    w = np.ascontiguousarray(a).view(np.uint32).copy()
    f = w.view(np.float32)
    w[np.isnan(f)] = 0x7FC00000
    return w


def assert_bit_equal(got, want, what=""):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if not np.array_equal(bits(got), bits(want)):
        bad = np.nonzero(bits(got).reshape(-1) != bits(want).reshape(-1))[0]
        raise AssertionError(f"{what}: {len(bad)} of {bits(want).size} words differ, first at {bad[:8]}: "
                             f"got {got.reshape(-1).view(np.float32)[bad[:4]]} want {want.reshape(-1).view(np.float32)[bad[:4]]}")


def rel_err(got, want):
    """max-norm relative error (per comparison block), the metric SURVEY 8c/BASELINE.md use."""
    d = np.abs(got.astype(np.complex128) - want.astype(np.complex128)).max() if got.size else 0.0
    return d / max(float(np.abs(want).max()) if want.size else 0.0, 1e-30)


class kept_only:
    """Oracle evaluates only the kept FIR outputs (bit-identical to the literal loop, tested)."""

    def __enter__(self):
        O.set_kept_only(True)

    def __exit__(self, *a):
        O.set_kept_only(False)
