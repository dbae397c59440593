"""CPU model of fk_tcfir's formulation (rows of 64 samples, frequency-translated complex taps split hi + lo in f16,
row phasors, sum over the rows that meet in an output) checked against the oracle.  Mirrors tcfir_geometry /
tcfir_b_image of quadrs_b200/csrc/qd_tcfir.cu; run by tests/test_tcfir_model.py."""
import math
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import oracle_lib as O  # noqa: E402


def geometry(L, D):
    i0 = L - L // 2
    opr = 64 // D
    gmax = (63 - i0) // D  # python floor division
    gmin = -((i0 + L - 1) // D)
    nout = gmax - gmin + 1
    nh = (2 * nout + 7) // 8 * 8
    if (2 * nh) % 16:
        nh += 8
    return dict(i0=i0, OPR=opr, c0=gmin, NOUT=nout, NH=nh, N=2 * nh, DMAX=-(-nout // opr), cown=gmin + nout - opr)


def b_matrix(g, taps, L, D, rsum):
    """[128][N] float64 pair (hi, lo) as f16-representable values, plus scales"""
    gm = float(np.abs(taps.astype(np.float64) / 127.0).max())
    _, e = math.frexp(gm)
    sigma = math.ldexp(1.0, 13 - e)
    B = np.zeros((128, g["N"]), np.float64)
    for n in range(g["N"]):
        half, cn = divmod(n, g["NH"])
        if cn >= 2 * g["NOUT"]:
            continue
        ip, ri = divmod(cn, 2)
        for k in range(128):
            kk, c = divmod(k, 2)
            j = kk - (g["c0"] + ip) * D - g["i0"]
            if j < 0 or j >= L:
                continue
            f = float(taps[j]) / 127.0 * sigma
            gr, gi = f * math.cos(rsum * kk), f * math.sin(rsum * kk)
            v = (gr if c == 0 else -gi) if ri == 0 else (gi if c == 0 else gr)
            hi = float(np.float16(np.float32(v)))
            lo = float(np.float16(np.float32((v - hi) * 2048.0)))
            B[k, n] = hi if half == 0 else lo
    return B, 1.0 / sigma, 1.0 / sigma / 2048.0


def model(raw_i8, rate, freq, cutoff, D, L, g0, g1):
    """outputs [g0, g1) of shift(freq) | lowpass(cutoff, D, L), untruncated, from int8 I/Q bytes (sample 0 first)"""
    taps = O.taps(cutoff, rate, L)
    rsum = O.shift_ratio(freq, rate) if freq is not None else 0.0
    g = geometry(L, D)
    B, s_hi, s_lo = b_matrix(g, taps, L, D, rsum)
    x = raw_i8.astype(np.float64).reshape(-1)  # re, im interleaved
    nrows = len(x) // 128
    A = x[: nrows * 128].reshape(nrows, 128)
    Dm = (A @ B).astype(np.float32)  # f32 accumulators
    P = Dm[:, : g["NH"]] * np.float32(s_hi) + Dm[:, g["NH"] :] * np.float32(s_lo)
    P = (P[:, 0 : 2 * g["NOUT"] : 2] + 1j * P[:, 1 : 2 * g["NOUT"] : 2]).astype(np.complex64)
    rot = np.exp(1j * (rsum * 64.0 * np.arange(nrows))).astype(np.complex64)
    P = P * rot[:, None]
    out = np.zeros(g1 - g0, np.complex64)
    for gg in range(g0, g1):
        b = (gg * D + g["i0"]) // 64
        t = gg - (g["OPR"] * b + g["cown"])
        assert 0 <= t < g["OPR"], (gg, b, t)
        acc = np.complex64(0)
        for d in range(g["DMAX"]):
            ip = g["NOUT"] - g["OPR"] * (d + 1) + t
            if ip < 0:
                break
            if b + d < nrows:
                acc += P[b + d, ip]
        out[gg - g0] = acc
    return out


if __name__ == "__main__":
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
    from helpers import kept_only, oracle_chain, rel_err, synth_raw

    for D, L, f in [(8, 40, 1_500_000), (16, 100, 3_000_000), (4, 24, -2_000_000), (32, 40, 700_000), (2, 18, 1_000_000), (8, 64, None)]:
        n = 40_000
        raw, _ = synth_raw(O.CS8, n, rate=20e6)
        st = ([("shift", f)] if f is not None else []) + [("lowpass", 1_000_000, D, L)]
        with kept_only():
            want = oracle_chain(raw, O.CS8, 20_000_000, st).read_at(0, 1000)
        T = (L - L // 2 + D - 1) // D - 1
        got = model(np.frombuffer(raw, np.int8), 20_000_000, f, 1_000_000, D, L, 0, 1000 - T)
        print(D, L, f, geometry(L, D), "rel err", rel_err(got, want[: 1000 - T]))
