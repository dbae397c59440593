"""The fused kernel (qd_fast.cu) against the oracle and against the general unit-local executor, the
pipelined host path, sharding, and size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import assert_bit_equal, gpu_chain, kept_only, oracle_chain, rel_err, synth_raw

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Q():
    import quadrs_b200

    return quadrs_b200


CASES = [
    # fmt, stages, sink
    (O.CS8, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], ("write", 0x1000)),          # config 2
    (O.CS8, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], ("spark", 64, 16, (0.05, 2.0))),
    (O.CU8, [("lowpass", 1_000_000, 8, 40)], ("write", 0x1000)),
    (O.CS16, [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)], ("spark", 128, 128, (0.5, 50.0))),  # config 4
    (O.CF32, [("shift", 280_000), ("lowpass", 200_000, 32, 400)], ("spark", 64, 16, None)),       # config 1
    (O.CF32, [("lowpass", 3_000_000, 4, 22)], ("write", 0x1000)),
    (O.CS16, [("shift", -1_000_000), ("shift", 2_500_000), ("lowpass", 2_000_000, 2, 9 * 2)], ("write", 512)),
    (O.CS8, [("lowpass", 1_000_000, 8, 37)], ("write", 64)),  # odd filter length: NaN taps, still the same order
    (O.CS8, [("shift", 3_000_000), ("lowpass", 500_000, 16, 100)], ("spark", 32, 7, (0.01, 1.0))),
    (O.CU8, [("shift", 3_000_000), ("lowpass", 500_000, 32, 40)], ("spark", 4, 2, (0.01, 1.0))),
    # two-stage lowpass shared as streams (config 5 shape): the outer stage never sees the inner truncated tail
    (O.CF32, [("lowpass", 5_000_000, 8, 40), ("lowpass", 125_000, 32, 40)], ("spark", 4, 2, (0.001, 0.01))),
    (O.CS8, [("shift", 1_000_000), ("lowpass", 5_000_000, 8, 40), ("lowpass", 125_000, 32, 40)], ("write", 0x1000)),
    (O.CS16, [("lowpass", 5_000_000, 4, 24), ("lowpass", 1_000_000, 16, 30)], ("spark", 32, 8, (0.1, 10.0))),
    # single lowpass without truncated positions (D >= L - L/2): overlapping windows cut from one stream
    (O.CS16, [("lowpass", 1_000_000, 8, 8)], ("spark", 16, 3, (0.1, 10.0))),
    (O.CS8, [("shift", -2_000_000), ("lowpass", 3_000_000, 32, 40)], ("spark", 8, 1, (0.01, 1.0))),
    # no stage at all: windows decoded straight from the capture bytes
    (O.CU8, [], ("spark", 4096, 1024, (2.0, 500.0))),  # config 3 shape
    (O.CF32, [], ("spark", 64, 16, (0.01, 3.0))),
    (O.CS8, [], ("spark", 4, 1, (0.01, 1.0))),
    (O.CS16, [], ("spark", 2, 2, None)),
]


def _run(chain, sink):
    if sink[0] == "write":
        data, rc = chain.write_mem(chunk=sink[1])
        return data, rc
    idx, mag = chain.spark_fft(sink[1], sink[2], sink[3], want_mag=True)
    return (idx, mag), 0


@pytest.mark.parametrize("fmt,stages,sink", CASES)
def test_fused_matches_oracle_and_general_executor(Q, fmt, stages, sink):
    n = 150_000
    raw, _ = synth_raw(fmt, n, rate=100e6)
    fused = gpu_chain(raw, fmt, 100_000_000, stages)
    general = gpu_chain(raw, fmt, 100_000_000, stages).set_option("use_fast", 0)
    launches0 = Q._lib.lib().qd_kernel_launches()
    got, rc = _run(fused, sink)
    assert Q._lib.lib().qd_kernel_launches() > launches0
    ref, rc2 = _run(general, sink)
    with kept_only():
        want, rc3 = _run_oracle(oracle_chain(raw, fmt, 100_000_000, stages), sink)
    assert rc == rc2 == rc3
    if sink[0] == "write":
        assert_bit_equal(got, ref, "fused vs general executor")
        assert_bit_equal(got, want, "fused vs oracle")
    else:
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[0], want[0])
        assert_bit_equal(got[1], ref[1], "magnitudes fused vs general")
        assert_bit_equal(got[1], want[1], "magnitudes fused vs oracle")


def _run_oracle(chain, sink):
    if sink[0] == "write":
        return chain.write_mem(chunk=sink[1])
    idx, mag = chain.spark_fft(sink[1], sink[2], sink[3])
    return (idx, mag), 0


@pytest.mark.parametrize("seg", [40_000, 1_000_000])
def test_pipelined_host_and_file_sources(Q, tmp_path, seg):
    n = 700_000
    raw, _ = synth_raw(O.CS8, n)
    st = [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]
    with kept_only():
        want, want_rc = oracle_chain(raw, O.CS8, 20_000_000, st).write_mem()
        widx, _ = oracle_chain(raw, O.CS8, 20_000_000, st).spark_fft(64, 64, (0.05, 2.0))
    host = gpu_chain(raw, O.CS8, 20_000_000, st).set_option("segment_bytes", seg)
    got, rc = host.write_mem()
    assert rc == want_rc
    assert_bit_equal(got, want, "pipelined host source")
    idx, _ = host.spark_fft(64, 64, (0.05, 2.0))
    assert np.array_equal(idx, widx)
    path = tmp_path / "cap.sr20M.cs8"
    raw.tofile(path)
    f = Q.Samples.from_file(path, Q.CS8, 20_000_000).shift(1_500_000).lowpass(1_000_000, 8, 40)
    f.set_option("segment_bytes", seg)
    got, rc = f.write_mem()
    assert rc == want_rc
    assert_bit_equal(got, want, "pipelined file source")


def test_sharded_equals_unsharded(Q):
    # contiguous unit ranges, halo of (taps, window) samples, absolute indices: no hand-off
    n = 1_000_000
    raw, _ = synth_raw(O.CS16, n, rate=100e6)
    st = [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)]
    whole = gpu_chain(raw, O.CS16, 100_000_000, st)
    idx, _ = whole.spark_fft(128, 128, (0.5, 50.0))
    data, _ = whole.write_mem()
    for n_shards in (2, 3):
        rows, outs = [], []
        for r in range(n_shards):
            p = Q.shard_plan(Q.CS16, 100_000_000, n, st, Q.shard.SINK_SPARKFFT, 128, 128, n_shards, r)
            part = raw[p.first_sample * 4 : (p.first_sample + p.n_samples) * 4]
            g = gpu_chain(part, O.CS16, 100_000_000, st, base=p.first_sample, total=n)
            got, _ = g.spark_fft(128, 128, (0.5, 50.0), first_row=p.first_unit, max_rows=p.n_units)
            rows.append(got)
            w = Q.shard_plan(Q.CS16, 100_000_000, n, st, Q.shard.SINK_WRITE, 0x1000, 0x1000, n_shards, r)
            part = raw[w.first_sample * 4 : (w.first_sample + w.n_samples) * 4]
            g = gpu_chain(part, O.CS16, 100_000_000, st, base=w.first_sample, total=n)
            o, _ = g.write_mem(first_chunk=w.first_unit, max_chunks=w.n_units)
            outs.append(o)
        assert np.array_equal(np.concatenate(rows), idx)
        assert_bit_equal(np.concatenate(outs), data, f"{n_shards} shards")


def _oracle_rows_from_device(Q, torch, d_in, fmt, rate, total, base, stages, W, S, rng, rows, need):
    """Oracle rows computed from the device input's own bytes around each sampled row."""
    out = []
    pb = O.FORMAT_BYTES[fmt]
    mult = 1
    for st in stages:
        if st[0] == "lowpass":
            mult *= st[2]
    for r in rows:
        lo = r * S * mult
        hi = min(lo + need, total)
        raw = d_in[(lo - base) * pb : (hi - base) * pb].cpu().numpy()
        with kept_only():
            idx, mag = oracle_chain(raw, fmt, rate, stages, lo, total).spark_fft(W, S, rng, first_row=r, max_rows=1)
        out.append((idx[0], mag[0]))
    return out


def test_full_size_config2_checks(Q):
    """BASELINE.json configs[1] at full size (2^30 cs8 samples): sampled chunks against the oracle, and
    the whole 1 GiB output against a two-shard run (equality of every byte = a checksum of checksums)."""
    import torch

    n = 2**30
    fmt, rate = Q.CS8, 20_000_000
    st = [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]
    synth = Q.make_synth(0x5EED0002, [(Q.tone_step(1.6e6, rate), 45, 0), (Q.tone_step(-4.1e6, rate), 30, 0),
                                      (Q.tone_step(0.3e6, rate), 20, 3000)], 6)
    d_in = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    Q.synth_fill_device(synth, fmt, 0, n, d_in.data_ptr())
    torch.cuda.synchronize()
    chain = Q.Samples.from_device(d_in.data_ptr(), 2 * n, fmt, rate, keep=(d_in,)).shift(1_500_000).lowpass(1_000_000, 8, 40)
    ln = chain.len()
    assert ln == 1 + (n - 40) // 8
    chunks = -(-ln // 0x1000)
    d_out = torch.zeros(2 * chunks * 0x1000, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own (non-blocking) stream
    # one read more than there are chunks: the reference's loop makes it too, gets 0 samples and panics
    got_n, rc = chain.write_into(0x1000, 0, chunks + 1, d_out.data_ptr(), chunks * 0x1000, Q._lib.SPACE_DEVICE)
    chain.synchronize()
    assert rc == Q._lib.E_WRITE_SHORT and got_n == ln - 1  # lib.rs:203 fires after the last readable sample
    # sampled chunks vs the oracle (first, interior at 2^k boundaries, last full, ragged tail)
    for c in (0, 1, 12_345, chunks // 2, chunks - 3, chunks - 2, chunks - 1):
        lo = c * 0x1000 * 8
        hi = min(lo + 0x1000 * 8 + 40, n)
        raw = d_in[2 * lo : 2 * hi].cpu().numpy()
        with kept_only():
            want, _ = oracle_chain(raw, O.CS8, rate, st, lo, n).write_mem(first_chunk=c, max_chunks=1)
        got = d_out[2 * c * 0x1000 : 2 * (c * 0x1000 + len(want))].cpu().numpy().view(np.complex64)
        assert_bit_equal(got, want, f"chunk {c}")
    # two shards reproduce every output sample of the unsharded run
    d_out2 = torch.zeros_like(d_out)
    torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own (non-blocking) stream
    for r in range(2):
        p = Q.shard_plan(fmt, rate, n, st, Q.shard.SINK_WRITE, 0x1000, 0x1000, 2, r)
        view = d_in[2 * p.first_sample : 2 * (p.first_sample + p.n_samples)]
        g = Q.Samples.from_device(view.data_ptr(), view.numel(), fmt, rate, base_sample=p.first_sample, total_samples=n,
                                  keep=(d_in,)).shift(1_500_000).lowpass(1_000_000, 8, 40)
        off = 2 * p.first_unit * 0x1000
        g.write_into(0x1000, p.first_unit, p.n_units, d_out2.data_ptr() + 4 * off, p.n_units * 0x1000, Q._lib.SPACE_DEVICE)
        g.synchronize()
    assert torch.equal(d_out, d_out2)
    out_head = d_out[: 2 * 65536].cpu().numpy()
    assert np.isfinite(out_head).all() and np.abs(out_head).max() > 0


def test_full_size_config4_shape_sampled_rows(Q):
    """BASELINE.json configs[3] shape on a 2^28-sample shard of the 2^33 capture, at its absolute place."""
    import torch

    total, n = 2**33, 2**28
    base = total - n
    fmt, rate = Q.CS16, 100_000_000
    st = [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)]
    synth = Q.make_synth(0x5EED0004, [(Q.tone_step(7.3e6, rate), 9000, 0), (Q.tone_step(6.2e6, rate), 6000, 50_000),
                                      (Q.tone_step(-20e6, rate), 4000, 0)], 1200)
    d_in = torch.empty(4 * n, dtype=torch.uint8, device="cuda")
    Q.synth_fill_device(synth, fmt, base, n, d_in.data_ptr())
    torch.cuda.synchronize()
    chain = Q.Samples.from_device(d_in.data_ptr(), 4 * n, fmt, rate, base_sample=base, total_samples=total, keep=(d_in,))
    chain = chain.shift(7_000_000).lowpass(2_000_000, 16, 800)
    rows_total = chain.spark_rows(128, 128)
    assert rows_total == 4_194_303
    first = -(-base // (128 * 16))
    n_rows = rows_total - first
    d_idx = torch.zeros(n_rows * 128, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own (non-blocking) stream
    assert chain.spark_fft_device(128, 128, (0.5, 50.0), first, n_rows, d_idx.data_ptr()) == n_rows
    chain.synchronize()
    idx = d_idx.cpu().numpy().reshape(n_rows, 128)
    assert idx.max() <= 8
    rows = [first, first + 1, first + 77_777, rows_total - 2, rows_total - 1]
    want = _oracle_rows_from_device(Q, torch, d_in, O.CS16, rate, total, base, st, 128, 128, (0.5, 50.0), rows, 128 * 16 + 800)
    for r, (widx, _) in zip(rows, want):
        assert np.array_equal(idx[r - first], widx), f"row {r}"


TAIL_CASES = [
    # fmt, stages, W, S: overlapping windows behind one filter with truncated positions (stream + tail patch)
    (O.CS8, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], 64, 16),      # T = 2
    (O.CF32, [("shift", 280_000), ("lowpass", 200_000, 32, 400)], 64, 16),       # config 1: T = 6
    (O.CU8, [("lowpass", 3_000_000, 2, 132)], 128, 32),                           # T = 32, the most the patch takes
    (O.CS16, [("lowpass", 3_000_000, 2, 140)], 128, 32),                          # T = 34: every window on its own
    (O.CS8, [("shift", -2_000_000), ("shift", 700_000), ("lowpass", 500_000, 4, 64)], 16, 1),   # T = 7, stride 1
    (O.CS16, [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)], 256, 255),  # T = 24, one sample of overlap
    (O.CF32, [("lowpass", 1_000_000, 8, 10)], 8, 3),                              # T = 0: plain stream
    (O.CF32, [("lowpass", 20_000_000, 8, 40), ("lowpass", 500_000, 32, 40)], 4, 2),  # config 5: two stages, 4-point windows
    (O.CS8, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], 256, 64),      # windows half a tile wide
]


@pytest.mark.parametrize("fmt,stages,W,S", TAIL_CASES)
def test_overlapping_windows_stream_plus_tail(Q, fmt, stages, W, S):
    """Bucket indices and magnitudes of overlapping windows are bit-identical to the oracle's per-window reads,
    from the first window to the ragged ones at the end of the capture."""
    n = _mult(stages) * (W + 300 * S) + 5000
    raw, _ = synth_raw(fmt, n, rate=100e6)
    scale = 3e4 if fmt == O.CS16 else (100.0 if fmt == O.CU8 else 1.0)
    rng = (0.02 * scale * np.sqrt(W), 3.0 * scale * np.sqrt(W))
    o = oracle_chain(raw, fmt, 100_000_000, stages)
    g = gpu_chain(raw, fmt, 100_000_000, stages)
    with kept_only():
        try:
            widx, wmag = o.spark_fft(W, S, rng)
        except O.OracleError as e:
            with pytest.raises(Q.QdError) as ge:
                g.spark_fft(W, S, rng)
            assert ge.value.code == e.code
            return
    idx, mag = g.spark_fft(W, S, rng, want_mag=True)
    idx2, _ = g.spark_fft(W, S, rng)
    assert idx.shape == widx.shape and idx.shape[0] > 250
    assert np.array_equal(idx, widx) and np.array_equal(idx2, widx)
    assert_bit_equal(mag, wmag, "magnitudes")
    # the same with the STFT never / always inside the filter kernel (fk_fir FUSE = 2: windows carried from tile to
    # tile, snapshots for the truncated tails), in several host-path segments
    for fuse, seg in ((0, 0), (2, 0), (2, 9_000 * O.FORMAT_BYTES[fmt])):
        h = gpu_chain(raw, fmt, 100_000_000, stages)
        h.set_option("fuse_stft", fuse)
        if seg:
            h.set_option("segment_bytes", seg)
        fidx, fmag = h.spark_fft(W, S, rng, want_mag=True)
        fidx2, _ = h.spark_fft(W, S, rng)
        assert np.array_equal(fidx, widx) and np.array_equal(fidx2, widx), (fuse, seg)
        assert_bit_equal(fmag, wmag, f"magnitudes, fuse_stft={fuse}, segment_bytes={seg}")


# ---------------------------------------------------------------- FAST arithmetic mode
FAST_CASES = [
    (O.CS8, 20_000_000, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], 0),                      # config 2
    (O.CS8, 20_000_000, [("shift", 9_999_999), ("lowpass", 1_000_000, 8, 40)], 2**30 - 600_000),        # near Nyquist, n ~ 2^30
    (O.CF32, 21_000_000, [("shift", 280_000), ("lowpass", 200_000, 32, 400)], 0),                       # config 1
    (O.CF32, 400_000_000, [("shift", -150_000_000), ("shift", 199_999_999), ("lowpass", 20_000_000, 8, 40)], 2**34 - 600_000 - 2**34 % 0x8000),
    (O.CF32, 400_000_000, [("lowpass", 20_000_000, 8, 40), ("lowpass", 500_000, 32, 40)], 0),           # config 5
]


@pytest.mark.parametrize("fmt,rate,stages,base", FAST_CASES)
def test_fast_mode_within_1e5_of_the_oracle(Q, fmt, rate, stages, base):
    """north_star tolerance: cf32 samples within 1e-5 relative (max-norm per 0x1000-sample chunk).  FAST keeps
    integer decode exact, evaluates the f64 phase once per thread and tile, and contracts the FIR to FMA."""
    n = 500_000 if _mult(stages) <= 8 else 3_000_000
    total = base + n
    raw, _ = synth_raw(fmt, n, first=base, rate=rate)
    fast = gpu_chain(raw, fmt, rate, stages, base, total if base else 0, precision=Q.FAST)
    first = -(-base // (0x1000 * _mult(stages)))
    with kept_only():
        want, _ = oracle_chain(raw, fmt, rate, stages, base, total if base else 0).write_mem(first_chunk=first, max_chunks=10)
    got, _ = fast.write_mem(first_chunk=first, max_chunks=10)
    assert len(got) == len(want) and len(want) >= 8_000
    worst = 0.0
    for c in range(0, len(want), 0x1000):
        worst = max(worst, rel_err(got[c : c + 0x1000], want[c : c + 0x1000]))
    print(f"FAST worst chunk rel err {worst:.3e}")
    assert worst <= 1e-5, worst


def _lean_case(seed):
    """cs8 FAST through the lean decode loop: random shift (sign, tiny, zero, near Nyquist), decimation, filter
    length and capture offset, including offsets above 2^32 (more than 32 product bits rounded away) and offsets
    placed so that a tile straddles a binade of n*ratio (where the number of rounded bits changes)."""
    import math

    rng = np.random.default_rng(0xFA57 + seed)
    rate = int(rng.choice([2_400_000, 20_000_000, 100_000_000]))
    kind = seed % 6
    if kind == 0:
        f = int(rng.integers(-rate // 2 + 1, rate // 2 - 1))
    elif kind == 1:
        f = int(rng.choice([-1, 1])) * (rate // 2 - 1)
    elif kind == 2:
        f = int(rng.choice([-3, 1, 7]))
    elif kind == 3:
        f = 0
    else:
        f = int(rng.integers(-rate // 2 + 1, rate // 2 - 1))
    D = int(rng.choice([2, 4, 8, 8, 16, 32]))
    L = int(rng.choice([40, 40, 40, 24, 64]))
    stages = [("shift", f), ("lowpass", int(rng.integers(rate // 64, rate // 8)), D, L)]
    if seed % 7 == 3:
        stages = stages[1:]  # no shift at all: decode only
    unit = 0x1000 * D
    if kind == 4 and f != 0:  # straddle: bitlen(n * M) changes at n = 2^j / m, |ratio| = m * 2^ex
        m, _ = math.frexp(abs(2.0 * math.pi * f / rate))
        j = int(rng.integers(24, 36))
        edge = int((1 << j) / m)
        base = max(0, (edge - 3 * unit) // unit * unit)
    elif kind == 5:
        base = int(rng.integers(2**32, 2**36)) // unit * unit
    else:
        base = int(rng.integers(0, 2**31)) // unit * unit
    return rate, stages, base


@pytest.mark.parametrize("seed", range(int(__import__("os").environ.get("QD_LEAN_SEEDS", "48"))))
def test_fast_cs8_lean_loop_random(Q, seed):
    rate, stages, base = _lean_case(seed)
    n = 0x1000 * _mult(stages) * 9 + 4096
    total = base + n
    raw, _ = synth_raw(O.CS8, n, first=base, rate=rate)
    fast = gpu_chain(raw, O.CS8, rate, stages, base, total if base else 0, precision=Q.FAST)
    first = base // (0x1000 * _mult(stages))
    with kept_only():
        want, _ = oracle_chain(raw, O.CS8, rate, stages, base, total if base else 0).write_mem(first_chunk=first, max_chunks=8)
    got, _ = fast.write_mem(first_chunk=first, max_chunks=8)
    assert len(got) == len(want) == 8 * 0x1000
    worst = max(rel_err(got[c : c + 0x1000], want[c : c + 0x1000]) for c in range(0, len(want), 0x1000))
    assert worst <= 1e-5, (worst, rate, stages, base)


@pytest.mark.parametrize("seed", range(8))
def test_fast_cs8_lean_loop_under_overlapping_windows(Q, seed):
    """FAST through sparkfft: overlapping windows are separate reads (one unit each, their own truncated tails),
    so the fused kernel runs in its per-unit tiling; magnitudes stay within 1e-5 of the oracle's (max-norm per row)."""
    rate, stages, base = _lean_case(seed * 5 + 1)
    W, S = [(64, 16), (128, 128), (32, 7), (256, 64)][seed % 4]
    n = _mult(stages) * (W + 40 * S) + 4096
    total = base + n
    raw, _ = synth_raw(O.CS8, n, first=base, rate=rate)
    fast = gpu_chain(raw, O.CS8, rate, stages, base, total if base else 0, precision=Q.FAST)
    first = -(-base // (_mult(stages) * S))
    rng = (1e-4, 1e4)
    with kept_only():
        _, want = oracle_chain(raw, O.CS8, rate, stages, base, total if base else 0).spark_fft(W, S, rng, first_row=first, max_rows=24)
    _, got = fast.spark_fft(W, S, rng, first_row=first, max_rows=24, want_mag=True)
    assert got.shape == want.shape and want.shape[0] == 24
    worst = max(float(np.abs(got[r] - want[r]).max() / max(np.abs(want[r]).max(), 1e-30)) for r in range(24))
    assert worst <= 1e-5, (worst, rate, stages, base, W, S)


def _mult(stages):
    m = 1
    for st in stages:
        if st[0] == "lowpass":
            m *= st[2]
    return m


def test_fast_mode_is_refused_for_offset_formats(Q):
    raw, _ = synth_raw(O.CS16, 10_000)
    for fmt in (O.CS16, O.CU8):
        with pytest.raises(Q.QdError) as e:
            gpu_chain(raw, fmt, 20_000_000, [("shift", 1_000_000), ("lowpass", 1_000_000, 8, 40)], precision=Q.FAST)
        assert e.value.code == Q._lib.E_INVALID_ARG and "1e-5" in e.value.msg


def test_fast_mode_full_size_config2_against_exact(Q):
    """The bench workload itself (2^30 cs8 samples, config 2): every 0x1000-sample chunk of the FAST output is
    within 1e-5 (max-norm relative) of the bit-exact EXACT output, which the tests above tie to the oracle."""
    import torch

    n = 2**30
    fmt, rate = Q.CS8, 20_000_000
    synth = Q.make_synth(0x5EED0002, [(Q.tone_step(1.6e6, rate), 45, 0), (Q.tone_step(-4.1e6, rate), 30, 0),
                                      (Q.tone_step(0.3e6, rate), 20, 3000)], 6)
    d_in = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    Q.synth_fill_device(synth, fmt, 0, n, d_in.data_ptr())
    torch.cuda.synchronize()
    chunks = 32767
    outs = []
    for prec in (Q.EXACT, Q.FAST):
        chain = Q.Samples.from_device(d_in.data_ptr(), 2 * n, fmt, rate, keep=(d_in,)).shift(1_500_000)
        chain = chain.lowpass(1_000_000, 8, 40).with_precision(prec)
        d_out = torch.zeros(2 * chunks * 0x1000, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own (non-blocking) stream
        got_n, rc = chain.write_into(0x1000, 0, chunks, d_out.data_ptr(), chunks * 0x1000, Q._lib.SPACE_DEVICE)
        chain.synchronize()
        assert got_n == chunks * 0x1000 and rc == 0
        outs.append(torch.view_as_complex(d_out.view(-1, 2)).view(chunks, 0x1000))
    exact, fast = outs
    err = (fast - exact).abs().amax(dim=1) / exact.abs().amax(dim=1)
    worst = float(err.max())
    print(f"FAST vs EXACT over 2^30 samples: worst chunk rel err {worst:.3e}, mean {float(err.mean()):.3e}")
    assert worst <= 1e-5, worst
    assert not torch.equal(exact, fast)


def test_shards_and_pointers_at_awkward_alignments(Q):
    """Shard bases that are not multiples of 8 samples, host pointers at odd byte offsets, and a device
    pointer that breaks the 16-byte rule (falls back to the general executor): same bits every time."""
    import torch

    n = 300_000
    raw, _ = synth_raw(O.CS8, n)
    st = [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]
    with kept_only():
        want, _ = oracle_chain(raw, O.CS8, 20_000_000, st).write_mem()
    for base in (3, 8 * 1237 + 5, 0x1000 * 8 * 3 + 1):
        # the shard starts at raw sample `base`; chunks whose span lies inside it reproduce the unsharded bits
        part = np.ascontiguousarray(raw[2 * base :])
        first_chunk = -(-base // (0x1000 * 8))
        g = gpu_chain(part, O.CS8, 20_000_000, st, base=base, total=n)
        got, _ = g.write_mem(first_chunk=first_chunk, max_chunks=4)
        assert_bit_equal(got, want[first_chunk * 0x1000 : first_chunk * 0x1000 + len(got)], f"host shard base {base}")
        assert len(got) == 4 * 0x1000
        # the same bytes at an odd device address: the fused kernel's bulk copies need absolute sample 0 on a
        # 16-byte boundary, so this goes through the general executor
        buf = torch.zeros(part.size + 64, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own (non-blocking) stream
        off = 2 + 2 * (base % 7)  # sample-aligned, not 16-byte aligned
        buf[off : off + part.size] = torch.from_numpy(part).cuda()
        d = Q.Samples.from_device(buf.data_ptr() + off, part.size, Q.CS8, 20_000_000, base_sample=base, total_samples=n,
                                  keep=(buf,)).shift(1_500_000).lowpass(1_000_000, 8, 40)
        got, _ = d.write_mem(first_chunk=first_chunk, max_chunks=4)
        assert_bit_equal(got, want[first_chunk * 0x1000 : first_chunk * 0x1000 + len(got)], f"device shard base {base}")
        with pytest.raises(Q.QdError) as e:  # a device pointer in the middle of a sample is refused
            Q.Samples.from_device(buf.data_ptr() + 1, part.size, Q.CS8, 20_000_000)
        assert e.value.code == Q._lib.E_INVALID_ARG


# ---------------------------------------------------------------- FAST results never depend on how a run is cut
FAST_CUT_CASES = [
    (O.CS8, 20_000_000, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]),    # lean loop, tiles of 508
    (O.CS8, 20_000_000, [("shift", -9_999_999), ("lowpass", 1_000_000, 8, 40)]),   # near Nyquist
    (O.CS8, 20_000_000, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 56)]),    # run-time filter length
    (O.CF32, 21_000_000, [("shift", 280_000), ("lowpass", 200_000, 32, 400)]),     # general FAST loop
    (O.CS8, 20_000_000, [("shift", 700_000), ("shift", 800_000), ("lowpass", 1_000_000, 4, 40)]),
]


@pytest.mark.parametrize("fmt,rate,stages", FAST_CUT_CASES)
def test_fast_output_is_independent_of_segments_and_shards(Q, fmt, rate, stages):
    """shift.rs:49 makes the phase a function of the absolute index alone; the FAST kernel's tiles, anchors and
    recurrences are numbered from absolute output 0, so host-path segments of any size, shards and the
    device-resident path all produce the same bits."""
    import torch

    n = 1_500_000
    pb = O.FORMAT_BYTES[fmt]
    raw, _ = synth_raw(fmt, n, rate=rate)
    d_raw = torch.from_numpy(raw).cuda()
    dev = Q.Samples.from_device(d_raw.data_ptr(), raw.size, fmt, rate, keep=(d_raw,))
    for st in stages:
        dev = dev.shift(st[1]) if st[0] == "shift" else dev.lowpass(st[1], st[2], st[3])
    dev = dev.with_precision(Q.FAST)
    want, want_rc = dev.write_mem()
    assert len(want) > 10 * 0x1000 or stages[-1][2] >= 32
    # host path, two segment sizes (neither a multiple of the tile size)
    for seg in (100_000, 1 << 20):
        h = gpu_chain(raw, fmt, rate, stages, precision=Q.FAST)
        h.set_option("segment_bytes", seg)
        got, rc = h.write_mem()
        assert rc == want_rc
        assert_bit_equal(got, want, f"host path, segment_bytes {seg}")
    # shards
    for n_shards in (2, 3):
        outs = []
        for r in range(n_shards):
            w = Q.shard_plan(fmt, rate, n, stages, Q.shard.SINK_WRITE, 0x1000, 0x1000, n_shards, r)
            part = raw[w.first_sample * pb : (w.first_sample + w.n_samples) * pb]
            g = gpu_chain(part, fmt, rate, stages, base=w.first_sample, total=n, precision=Q.FAST)
            o, _ = g.write_mem(first_chunk=w.first_unit, max_chunks=w.n_units)
            outs.append(o)
        assert_bit_equal(np.concatenate(outs), want, f"{n_shards} shards")
    # a read that starts in the middle of a tile and of a chunk: its untruncated part is the stream's
    D, L = stages[-1][2], stages[-1][3]
    T = -(-(L - L // 2) // D) - 1
    for off, cnt in ((4096 + 130, 1000), (509, 2 * 508 + 4), (12, 3000)):
        if off + cnt >= len(want):
            continue
        if cnt % 8:
            cnt -= cnt % 8
        got = dev.read_at(off, cnt)
        chunk_end = (off // 0x1000 + 1) * 0x1000
        keep = min(cnt - T, chunk_end - T - off)  # positions that are untruncated both in the read and in its write chunk
        assert_bit_equal(got[:keep], want[off : off + keep], f"read_at({off}, {cnt})")


def test_fast_sparkfft_is_independent_of_segments(Q):
    n = 1_200_000
    raw, _ = synth_raw(O.CS8, n)
    st = [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]
    for W, S in ((64, 16), (64, 64), (16, 6)):
        ref = None
        for seg in (0, 70_000, 1 << 19):
            g = gpu_chain(raw, O.CS8, 20_000_000, st, precision=Q.FAST)
            if seg:
                g.set_option("segment_bytes", seg)
            idx, mag = g.spark_fft(W, S, (0.01, 3.0), want_mag=True)
            if ref is None:
                ref = (idx, mag)
            else:
                assert np.array_equal(idx, ref[0]), (W, S, seg)
                assert_bit_equal(mag, ref[1], f"sparkfft {W}/{S} segment_bytes {seg}")


def test_fast_full_size_config2_shards_and_host_path_bitwise(Q):
    """2^30 cs8 samples in FAST arithmetic: 2- and 3-shard concatenations and the pipelined host path (two segment
    sizes) equal the unsharded device-resident run in every bit."""
    import torch

    n = 2**30
    fmt, rate = Q.CS8, 20_000_000
    st = [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]
    synth = Q.make_synth(0x5EED0002, [(Q.tone_step(1.6e6, rate), 45, 0), (Q.tone_step(-4.1e6, rate), 30, 0),
                                      (Q.tone_step(0.3e6, rate), 20, 3000)], 6)
    d_in = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
    Q.synth_fill_device(synth, fmt, 0, n, d_in.data_ptr())
    torch.cuda.synchronize()

    def build(s):
        return s.shift(1_500_000).lowpass(1_000_000, 8, 40).with_precision(Q.FAST)

    chunks = 32767
    d_out = torch.zeros(2 * chunks * 0x1000, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own (non-blocking) stream
    whole = build(Q.Samples.from_device(d_in.data_ptr(), 2 * n, fmt, rate, keep=(d_in,)))
    got_n, _ = whole.write_into(0x1000, 0, chunks, d_out.data_ptr(), chunks * 0x1000, Q._lib.SPACE_DEVICE)
    whole.synchronize()
    assert got_n == chunks * 0x1000
    for n_shards in (2, 3):
        d_out2 = torch.zeros_like(d_out)
        torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own (non-blocking) stream
        for r in range(n_shards):
            p = Q.shard_plan(fmt, rate, n, st, Q.shard.SINK_WRITE, 0x1000, 0x1000, n_shards, r)
            view = d_in[2 * p.first_sample : 2 * (p.first_sample + p.n_samples)]
            g = build(Q.Samples.from_device(view.data_ptr(), view.numel(), fmt, rate, base_sample=p.first_sample,
                                            total_samples=n, keep=(d_in,)))
            nu = min(p.n_units, chunks - p.first_unit)
            g.write_into(0x1000, p.first_unit, nu, d_out2.data_ptr() + 8 * p.first_unit * 0x1000, nu * 0x1000,
                         Q._lib.SPACE_DEVICE)
            g.synchronize()
        assert torch.equal(d_out, d_out2), f"{n_shards} shards"
    h_in = torch.empty(2 * n, dtype=torch.uint8, pin_memory=True)
    h_in.copy_(d_in)
    h_out = torch.empty(2 * chunks * 0x1000, dtype=torch.float32, pin_memory=True)
    for seg in (32 << 20, 24_000_000):
        h = build(Q.Samples.from_host_ptr(h_in.data_ptr(), 2 * n, fmt, rate, keep=(h_in,)))
        h.set_option("segment_bytes", seg)
        h_out.zero_()
        h.write_into(0x1000, 0, chunks, h_out.data_ptr(), chunks * 0x1000, Q._lib.SPACE_HOST)
        h.synchronize()
        assert torch.equal(h_out, d_out.cpu()), f"host path, segment_bytes {seg}"


# ---------------------------------------------------------------- BASELINE sizes, whole captures
def _full_size_sparkfft(Q, fmt, rate, total, stages, W, S, rng, synth_args, need, sample_rows):
    """The whole capture resident on the device: every row computed, sampled rows against the oracle (from the
    device's own bytes), and a two-shard run equal to the unsharded one in every byte of the output."""
    import torch

    pb = O.FORMAT_BYTES[fmt]
    synth = Q.make_synth(*synth_args)
    d_in = torch.empty(pb * total + 64, dtype=torch.uint8, device="cuda")
    Q.synth_fill_device(synth, fmt, 0, total, d_in.data_ptr())
    torch.cuda.synchronize()

    def build(src):
        for st in stages:
            src = src.shift(st[1]) if st[0] == "shift" else src.lowpass(st[1], st[2], st[3])
        return src

    whole = build(Q.Samples.from_device(d_in.data_ptr(), pb * total, fmt, rate, keep=(d_in,)))
    rows_total = whole.spark_rows(W, S)
    d_idx = torch.zeros(rows_total * W, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own
    assert whole.spark_fft_device(W, S, rng, 0, rows_total, d_idx.data_ptr()) == rows_total
    whole.synchronize()
    rows = [r if r >= 0 else rows_total + r for r in sample_rows]
    want = _oracle_rows_from_device(Q, torch, d_in, fmt, rate, total, 0, stages, W, S, rng, rows, need)
    for r, (widx, _) in zip(rows, want):
        got = d_idx[r * W : (r + 1) * W].cpu().numpy()
        assert np.array_equal(got, widx), f"row {r}"
    assert int(d_idx.max()) <= 8 and int((d_idx > 0).sum()) > 0
    d_idx2 = torch.zeros_like(d_idx)
    torch.cuda.synchronize()
    for r in range(2):
        p = Q.shard_plan(fmt, rate, total, stages, Q.shard.SINK_SPARKFFT, W, S, 2, r)
        view = d_in[pb * p.first_sample : pb * (p.first_sample + p.n_samples)]
        g = build(Q.Samples.from_device(view.data_ptr(), view.numel(), fmt, rate, base_sample=p.first_sample,
                                        total_samples=total, keep=(d_in,)))
        assert g.spark_fft_device(W, S, rng, p.first_unit, p.n_units, d_idx2.data_ptr() + p.first_unit * W) == p.n_units
        g.synchronize()
    assert torch.equal(d_idx, d_idx2)
    return rows_total


def test_full_size_config3_whole_capture(Q):
    """BASELINE.json configs[2]: cu8, 2^30 samples, sparkfft -width 4096 -stride 1024 (4 GiB of glyph rows)."""
    rate = 2_400_000
    synth = (0x5EED0003, [(Q.tone_step(-800e3, rate), 40, 0), (Q.tone_step(-123_456, rate), 30, 0),
                          (Q.tone_step(300e3, rate), 25, 0), (Q.tone_step(1_000_001, rate), 20, 0)], 4)
    rows = _full_size_sparkfft(Q, Q.CU8, rate, 2**30, [], 4096, 1024, (2.0, 500.0), synth, 4096,
                               [0, 1, 524_287, 524_288, -2, -1])
    assert rows == (2**30 - 4096 + 1023) // 1024


def test_full_size_config4_whole_capture(Q):
    """BASELINE.json configs[3] at its full 2^33 samples (32 GiB resident): filter and STFT in one kernel."""
    rate = 100_000_000
    st = [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)]
    synth = (0x5EED0004, [(Q.tone_step(7.3e6, rate), 9000, 0), (Q.tone_step(6.2e6, rate), 6000, 50_000),
                          (Q.tone_step(-20e6, rate), 4000, 0)], 1200)
    rows = _full_size_sparkfft(Q, Q.CS16, rate, 2**33, st, 128, 128, (0.5, 50.0), synth, 128 * 16 + 800,
                               [0, 1, 2_097_151, 2_097_152, 3_333_333, -2, -1])
    assert rows == 4_194_303


def test_full_size_config5_whole_buffer(Q):
    """BASELINE.json configs[4] on the resident 2^32-sample buffer the bench uses (32 GiB): two lowpasses, 4-point windows."""
    rate = 400_000_000
    st = [("lowpass", 20_000_000, 8, 40), ("lowpass", 500_000, 32, 40)]
    synth = (0x5EED0005, [(Q.tone_step(0.1e6, rate), 160, 1_000_000), (Q.tone_step(90e6, rate), 3000, 0)], 40)
    rows = _full_size_sparkfft(Q, Q.CF32, rate, 2**32, st, 4, 2, (0.001, 0.01), synth, (4 * 32 + 40) * 8 + 40,
                               [0, 1, 2, 4_194_303, 4_194_304, -2, -1])
    assert rows > 8_000_000


@pytest.mark.parametrize("fmt,stages,sink", [
    (O.CS16, [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)], ("spark", 128, 128, (0.5, 50.0))),   # config 4: FUSE = 1
    (O.CU8, [("shift", -3_000_000), ("lowpass", 1_500_000, 16, 404)], ("write", 0x1000)),                    # no STFT in the kernel
    (O.CS8, [("shift", 1_000_000), ("shift", 250_000), ("lowpass", 900_000, 32, 800)], ("spark", 64, 64, (0.05, 2.0))),
    (O.CF32, [("shift", 1_300_000), ("lowpass", 900_000, 32, 400)], ("spark", 64, 16, (0.05, 2.0))),        # config 1's shape: snapshots
])
def test_overlap_carried_between_consecutive_tiles(Q, fmt, stages, sink):
    """Long run-time-length filters over integer captures (fk_fir CARRY): a CTA walks a contiguous run of tiles and
    moves the L - D overlapping, already mixed samples across in registers instead of decoding them again.  With one
    CTA per SM and a capture of a few hundred tiles every CTA carries (runs of 2 - 3 tiles); every output against the
    oracle and against the general executor, bit for bit -- and against the same chain cut differently (two more
    CTAs per SM: other tiles are run-first)."""
    D = stages[-1][2]
    tile_samples = (512 if D == 16 else 256) * D
    n = 330 * tile_samples + 12_345
    raw, _ = synth_raw(fmt, n, rate=100e6)
    fused = gpu_chain(raw, fmt, 100_000_000, stages).set_option("fir_cta_cap", 1)
    other = gpu_chain(raw, fmt, 100_000_000, stages).set_option("fir_cta_cap", 3)
    general = gpu_chain(raw, fmt, 100_000_000, stages).set_option("use_fast", 0)
    got, rc = _run(fused, sink)
    alt, rc1 = _run(other, sink)
    ref, rc2 = _run(general, sink)
    with kept_only():
        want, rc3 = _run_oracle(oracle_chain(raw, fmt, 100_000_000, stages), sink)
    assert rc == rc1 == rc2 == rc3
    if sink[0] == "write":
        assert_bit_equal(got, want, "carried tiles vs oracle")
        assert_bit_equal(alt, want, "other run lengths vs oracle")
        assert_bit_equal(ref, want, "general executor vs oracle")
    else:
        assert np.array_equal(got[0], want[0]) and np.array_equal(alt[0], want[0]) and np.array_equal(ref[0], want[0])
        assert_bit_equal(got[1], want[1], "magnitudes, carried tiles vs oracle")
