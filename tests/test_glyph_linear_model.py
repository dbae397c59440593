"""CPU model of the linear glyph form (quadrs_b200/csrc/qd_stft_epilogue.cuh glyph_lin, qd_stft.cu stft_linear_glyphs).

The kernel decides a bin's glyph from g = a*sqrt(re^2 + im^2) + b in f32 and hands every bin whose g lies within eps of
an integer to the exact thresholds.  This test restates the host-side acceptance and the device arithmetic in numpy
f32 -- with the square root pushed one ulp either way, the worst the approximate instruction may do -- and checks, on
magnitudes packed around every glyph boundary, that a DECIDED bin always equals the oracle's glyph (fft.rs:53-60)."""
import numpy as np
import pytest

import oracle_lib as O

RANGES = [(0.05, 2.0), (0.001, 0.01), (0.5, 50.0), (100.0, 100.5), (0.0, 1.0), (3.0, 4.0e6), (0.08, 1.0), (7.0, 7.5e3)]
MAGIC = np.float32(12582912.0)


def linear_form(lo, hi):
    """stft_linear_glyphs: (a, bh, bl, lo, hi, eps) in f32, or None when the form is refused."""
    dist = float(np.float32((np.float32(hi) - np.float32(lo)) / np.float32(7)))
    if not (dist > 0 and np.isfinite(dist)):
        return None
    a = np.float32(1.0 / dist)
    b = np.float32(1.0 - float(np.float32(lo)) / dist)
    if not (np.isfinite(a) and np.isfinite(b) and a > 0 and a <= 1e9 and abs(b) <= 8192):
        return None
    A, B = float(a), float(b)
    eps = (9.0 + abs(B)) * 2.0 ** -20
    lo_c = np.float32((0.5 - B) / A) if B < 0.5 else np.float32(0)
    hi_c = np.float32((8.5 - B) / A)
    return a, np.float32(B - 0.5 + eps), np.float32(B - 0.5 - eps), lo_c, hi_c, eps


def device_glyph(form, re, im, ulp_push):
    """glyph_lin in numpy f32; ulp_push in {-1, 0, +1} moves the square root by that many ulps."""
    a, bh, bl, lo_c, hi_c, _ = form
    re, im = np.float32(re), np.float32(im)
    s = np.float32(np.float64(im) * np.float64(im) + np.float64(np.float32(re * re)))  # fmaf(im, im, re*re)
    r = np.sqrt(s, dtype=np.float32)
    for _ in range(abs(ulp_push)):
        r = np.nextafter(r, np.float32(np.inf if ulp_push > 0 else 0))
    r = np.minimum(np.maximum(r, lo_c), hi_c)
    h = np.float32(np.float64(r) * np.float64(a) + np.float64(bh))  # fused
    l = np.float32(np.float64(r) * np.float64(a) + np.float64(bl))
    mh, ml = np.float32(h + MAGIC), np.float32(l + MAGIC)
    return (mh.view(np.uint32) & 0xFF), mh != ml


@pytest.mark.parametrize("rng_", RANGES)
def test_decided_bins_equal_the_oracle(rng_):
    lo, hi = rng_
    form = linear_form(lo, hi)
    assert form is not None, "these ranges are all within the form's domain"
    d = np.float32((np.float32(hi) - np.float32(lo)) / np.float32(7))
    mags = []
    for k in range(9):
        e = np.float32(np.float32(lo) + np.float32(k) * d) if k < 8 else np.float32(hi)
        for near in (1.0, 1 - 2e-6, 1 + 2e-6, 1 - 3e-5, 1 + 3e-5, 1 - 1e-3, 1 + 1e-3):
            v = np.float32(float(e) * near)
            up = dn = v
            mags.append(v)
            for _ in range(12):
                up, dn = np.nextafter(up, np.float32(np.inf)), np.nextafter(dn, np.float32(0))
                mags += [up, dn]
    r = np.random.default_rng(5)
    mags += list(r.uniform(lo, hi, 300).astype(np.float32)) + [np.float32(0), np.float32(hi * 100), np.float32(lo / 3)]
    decided = undecided = 0
    for v in mags:
        for ang in (0.0, 0.3, 1.2):
            re, im = np.float32(v * np.cos(ang)), np.float32(v * np.sin(ang))
            norm = np.float32(np.sqrt(np.float64(re) ** 2 + np.float64(im) ** 2))  # hypotf of finite inputs
            want = O.glyph_index(float(norm), lo, hi)
            for push in (-1, 0, 1):
                g, und = device_glyph(form, re, im, push)
                if und:
                    undecided += 1
                    continue
                decided += 1
                assert want >= 0 and g == want, (rng_, float(v), ang, push, int(g), want)
    assert decided > undecided  # the band is narrow: most bins are decided without the thresholds
