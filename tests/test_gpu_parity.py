"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs.  Bar (BASELINE.json north_star): bit-exact integer decode and sparkfft bucket indices;
cf32 samples and FFT magnitudes within 1e-5 relative.  In EXACT mode the GPU reproduces the oracle's
operation order, so these tests demand bit equality for samples and magnitudes too."""
import numpy as np
import pytest

import oracle_lib as O
from helpers import assert_bit_equal, gpu_chain, kept_only, oracle_chain, rel_err, synth_raw

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star tolerance for cf32 samples and FFT magnitudes (relative, max-norm)


@pytest.fixture(scope="module")
def Q():
    import quadrs_b200

    return quadrs_b200


FORMATS = [O.CF32, O.CS8, O.CU8, O.CS16]


# ---------------------------------------------------------------- decode (a2, a3)
@pytest.mark.parametrize("fmt", FORMATS)
def test_decode_bit_exact(Q, fmt):
    if fmt == O.CS16:
        v = np.arange(-32768, 32768, dtype=np.int16)
        raw = np.stack([v, v[::-1]], axis=1).reshape(-1).view(np.uint8)
    elif fmt == O.CF32:
        words = np.random.default_rng(1).integers(0, 2**32, size=4096, dtype=np.uint32)  # any bit pattern, NaNs too
        raw = words.view(np.uint8)
    else:
        raw = np.stack([np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8)[::-1]], axis=1).reshape(-1)
    want = O.decode(fmt, raw)
    got = Q.Samples.from_bytes(raw, fmt, 1000).read_at(0, len(want))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"decode fmt {fmt}"  # raw bits, NaN payloads too


def test_file_source_len_floor_and_short_reads(Q, tmp_path):
    raw, _ = synth_raw(O.CS16, 1000)
    raw = np.concatenate([raw, np.zeros(3, dtype=np.uint8)])  # trailing partial pair
    path = tmp_path / "cap.sr48k.cs16"
    raw.tofile(path)
    g = Q.Samples.from_file(path, Q.CS16, 48_000)
    o = O.Samples.from_file(path, O.CS16, 48_000)
    assert g.len() == o.len() == 1000
    assert_bit_equal(g.read_at(990, 64), o.read_at(990, 64), "short read at eof")
    with pytest.raises(Q.QdError) as e:
        g.read_at(1000, 1)
    assert e.value.code == O.E_OFFSET_EOF
    with pytest.raises(Q.QdError) as e:
        g.read_exact_at(990, 64)
    assert e.value.code == O.E_SHORT_READ


# ---------------------------------------------------------------- shift (a4)
@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("freq", [280_000, -7_000_000, 9_999_999])
def test_shift_bit_exact(Q, fmt, freq):
    raw, _ = synth_raw(fmt, 50_000)
    st = [("shift", freq)]
    want = oracle_chain(raw, fmt, 20_000_000, st).read_at(123, 40_000)
    got = gpu_chain(raw, fmt, 20_000_000, st).read_at(123, 40_000)
    assert_bit_equal(got, want, f"shift fmt {fmt} f {freq}")


@pytest.mark.parametrize("fmt,base", [(O.CS16, 2**33 - 70_000), (O.CF32, 2**34 - 50_000), (O.CS8, 2**30 - 40_000)])
def test_shift_phase_at_capture_scale_offsets(Q, fmt, base):
    # phase = fl64(n * ratio) at n ~ 2^30..2^34: the shard holds only a window of the capture
    total = base + 40_000
    raw, _ = synth_raw(fmt, 40_000, first=base)
    st = [("shift", 49_999_999)]  # just under Nyquist at 100 MS/s: the worst case for phase ulp
    want = oracle_chain(raw, fmt, 100_000_000, st, base, total).read_at(base + 100, 30_000)
    got = gpu_chain(raw, fmt, 100_000_000, st, base, total).read_at(base + 100, 30_000)
    assert_bit_equal(got, want, "shift at large n")


# ---------------------------------------------------------------- lowpass (a5, a6)
def test_taps_match_oracle(Q):
    for freq, sr, size in [(200_000, 21_000_000, 400), (1_000_000, 20_000_000, 40), (2_000_000, 100_000_000, 800)]:
        g = Q.Samples.from_bytes(np.zeros(8 * 4000, dtype=np.uint8), Q.CF32, sr).lowpass(freq, 8, size)
        assert_bit_equal(g.taps(0), O.taps(freq, sr, size), "taps")


LOWPASS_CASES = [
    # fmt, stages, (off, n) reads
    (O.CS8, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], [(0, 4096), (3, 100), (500, 1)]),
    (O.CU8, [("lowpass", 300_000, 3, 10)], [(0, 500), (17, 64)]),
    (O.CS16, [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)], [(0, 128), (200, 128), (5, 700)]),
    (O.CF32, [("shift", 280_000), ("lowpass", 200_000, 32, 400)], [(0, 64), (16, 64), (100, 333)]),
    (O.CF32, [("lowpass", 2_000_000, 8, 40), ("lowpass", 500_000, 4, 40)], [(0, 64), (7, 200)]),
    (O.CS8, [("lowpass", 2_000_000, 4, 24), ("shift", -600_000), ("lowpass", 300_000, 2, 7)], [(0, 300), (11, 50)]),
    (O.CS8, [("shift", 100_000), ("shift", -2_000_000), ("lowpass", 1_000_000, 5, 33)], [(2, 77)]),
]


@pytest.mark.parametrize("fmt,stages,reads", LOWPASS_CASES)
def test_read_at_bit_exact_including_truncated_tail(Q, fmt, stages, reads):
    raw, _ = synth_raw(fmt, 60_000)
    o = oracle_chain(raw, fmt, 20_000_000, stages)
    g = gpu_chain(raw, fmt, 20_000_000, stages)
    assert g.len() == o.len() and g.sample_rate() == o.sample_rate()
    with kept_only():
        for off, n in reads:
            assert_bit_equal(g.read_at(off, n), o.read_at(off, n), f"read_at({off},{n}) {stages}")


def test_read_at_matches_literal_convolve(Q):
    # against the literal full-rate complex_convolve (filter.rs:107-124), not the kept-only shortcut
    raw, _ = synth_raw(O.CS8, 9000)
    st = [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]
    want = oracle_chain(raw, O.CS8, 20_000_000, st).read_at(5, 600)
    assert_bit_equal(gpu_chain(raw, O.CS8, 20_000_000, st).read_at(5, 600), want, "literal convolve")


@pytest.mark.parametrize("fmt", FORMATS)
def test_read_at_end_of_capture(Q, fmt):
    n = 5003
    raw, _ = synth_raw(fmt, n)
    st = [("shift", 1_000_000), ("lowpass", 1_000_000, 8, 40)]
    o, g = oracle_chain(raw, fmt, 20_000_000, st), gpu_chain(raw, fmt, 20_000_000, st)
    ln = o.len()
    assert g.len() == ln == 1 + (n - 40) // 8
    with kept_only():
        for off, cnt in [(ln - 30, 64), (ln - 2, 5), (ln - 1, 4)]:  # the last index always yields 0 (Q8)
            assert_bit_equal(g.read_at(off, cnt), o.read_at(off, cnt), f"eof read {off}")
    assert len(g.read_at(ln - 1, 4)) == 0


def test_lowpass_panics_map_to_codes(Q):
    raw, _ = synth_raw(O.CS8, 100)
    g = gpu_chain(raw, O.CS8, 48_000, [("lowpass", 2000, 8, 40)])
    with pytest.raises(Q.QdError) as e:
        g.read_at(9, 4)  # inner returns 28 < 40 samples (filter.rs:76)
    assert e.value.code == O.E_SHORT_INPUT
    g = gpu_chain(raw[:60], O.CS8, 48_000, [("lowpass", 2000, 8, 40)])
    with pytest.raises(Q.QdError) as e:
        g.len()  # filter.rs:46
    assert e.value.code == O.E_SHORT_INPUT


# ---------------------------------------------------------------- write (a11)
@pytest.mark.parametrize("fmt,stages", [
    (O.CS8, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]),  # config 2 shape
    (O.CS16, [("lowpass", 1_000_000, 4, 16)]),
    (O.CU8, [("shift", -300_000)]),
    (O.CF32, []),
])
def test_write_matches_oracle_chunk_for_chunk(Q, fmt, stages):
    n = 3 * 0x1000 * 8 + 4321
    raw, _ = synth_raw(fmt, n)
    o, g = oracle_chain(raw, fmt, 20_000_000, stages), gpu_chain(raw, fmt, 20_000_000, stages)
    with kept_only():
        want, want_rc = o.write_mem()
    got, got_rc = g.write_mem()
    assert got_rc == want_rc  # E_WRITE_SHORT after a lowpass (lib.rs:203), OK otherwise
    assert_bit_equal(got, want, "write")
    with kept_only():
        w1, _ = o.write_mem(first_chunk=2, max_chunks=1)
    g1, _ = g.write_mem(first_chunk=2, max_chunks=1)
    assert_bit_equal(g1, w1, "write chunk 2")


def test_write_file_roundtrip(Q, tmp_path):
    raw, _ = synth_raw(O.CS8, 40_000)
    st = [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]
    g = gpu_chain(raw, O.CS8, 20_000_000, st)
    name = g.write_file(str(tmp_path / "out"))
    assert name.endswith("out.sr2500000.cf32")
    data = np.fromfile(name, dtype=np.complex64)
    with kept_only():
        want, _ = oracle_chain(raw, O.CS8, 20_000_000, st).write_mem()
    assert_bit_equal(data, want, "written file")
    with pytest.raises(Q.QdError) as e:
        g.write_file(str(tmp_path / "out"))
    assert e.value.code == O.E_EXISTS
    g.write_file(str(tmp_path / "out"), overwrite=True)
    with pytest.raises(Q.QdError) as e:
        g.write_file("-")
    assert e.value.code == O.E_UNIMPLEMENTED
    # the written name re-parses as a `from` source (args.rs:328-333)
    back = Q.Samples.from_file(name, Q.CF32, 2_500_000)
    assert back.len() == len(want)


# ---------------------------------------------------------------- sparkfft (a7)
def test_readme_ook_known_answer_on_gpu(Q, golden_dir):
    import itertools

    g = Q.Samples.from_file(golden_dir / "cupboard-superdec.sr400.cf32", Q.CF32, 400)
    idx, mag = g.spark_fft(4, 2, (0.001, 0.01), want_mag=True)
    o = O.Samples.from_file(golden_dir / "cupboard-superdec.sr400.cf32", O.CF32, 400)
    widx, wmag = o.spark_fft(4, 2, (0.001, 0.01))
    assert idx.shape == (995, 4) and np.array_equal(idx, widx)
    assert_bit_equal(mag, wmag, "cupboard magnitudes")
    bits = "".join("." if (r == 0).all() else "X" for r in idx)
    runs = [(k, len(list(gr))) for k, gr in itertools.groupby(bits)]
    want = [(".", 8), ("X", 8), (".", 16), ("X", 17), (".", 15), ("X", 16)]  # README.md:135-140
    assert any(runs[i : i + 6] == want for i in range(len(runs)))
    assert g.spark_fft_text(4, 2, (0.001, 0.01)) == o.spark_fft_text(4, 2, (0.001, 0.01))


def test_config1_fsk_fixture_full(Q, golden_dir):
    # BASELINE.json configs[0]: shift 280000 | lowpass -power 200 -decimate 32 200000 | sparkfft -width 64 -stride 16
    path = golden_dir / "fsk-example.sr21M.fc32"
    g = Q.Samples.from_file(path, Q.CF32, 21_000_000).shift(280_000).lowpass(200_000, 32, 400)
    o = O.Samples.from_file(path, O.CF32, 21_000_000).shift(280_000).lowpass(200_000, 32, 400)
    with kept_only():
        widx, wmag = o.spark_fft(64, 16)
    idx, mag = g.spark_fft(64, 16, want_mag=True)
    assert idx.shape == (380, 64)
    assert np.array_equal(idx, widx)
    assert_bit_equal(mag, wmag, "config 1 magnitudes")
    gold = np.load(golden_dir / "config1_idx.npy")
    assert np.array_equal(idx, gold)


SPARK_CASES = [
    (O.CS16, [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)], 128, 128, (0.5, 50.0), 60_000),  # config 4 shape
    (O.CF32, [("lowpass", 20_000_000, 8, 40), ("lowpass", 500_000, 32, 40)], 4, 2, (0.001, 0.01), 40_000),  # config 5
    (O.CU8, [], 4096, 1024, (2.0, 500.0), 30_000),  # config 3 shape
    (O.CS8, [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], 64, 16, (0.05, 2.0), 30_000),
    (O.CS8, [], 8, 8, None, 2_000),
    (O.CF32, [("shift", -3_000_000)], 32, 5, (0.01, 3.0), 3_000),
    (O.CS8, [("lowpass", 1_000_000, 2, 6)], 2, 1, (0.01, 1.0), 300),
    (O.CS8, [], 1, 3, (0.01, 0.5), 100),
    (O.CU8, [("lowpass", 100_000, 8, 40)], 512, 100, (0.5, 20.0), 60_000),
]


@pytest.mark.parametrize("fmt,stages,W,S,rng,n", SPARK_CASES)
def test_sparkfft_bucket_indices_bit_exact(Q, fmt, stages, W, S, rng, n):
    raw, _ = synth_raw(fmt, n, rate=100e6)
    o, g = oracle_chain(raw, fmt, 100_000_000, stages), gpu_chain(raw, fmt, 100_000_000, stages)
    assert g.spark_rows(W, S) == o.spark_rows(W, S)
    with kept_only():
        widx, wmag = o.spark_fft(W, S, rng)
    idx, mag = g.spark_fft(W, S, rng, want_mag=True)
    assert idx.shape == widx.shape and idx.shape[0] > 0
    assert np.array_equal(idx, widx), f"{(idx != widx).sum()} bucket indices differ"
    for r in range(0, len(mag), max(1, len(mag) // 50)):  # per-window max-norm relative error
        assert rel_err(mag[r], wmag[r]) <= TOL
    assert_bit_equal(mag, wmag, "magnitudes")
    # a sub-range of rows reproduces the same rows
    r0 = idx.shape[0] // 3
    sub, _ = g.spark_fft(W, S, rng, first_row=r0, max_rows=5)
    assert np.array_equal(sub, widx[r0 : r0 + 5])


def test_sparkfft_at_capture_scale_offsets(Q):
    # config 4 near the end of a 2^33-sample capture: phase, truncation and EOF all use absolute indices
    total = 2**33
    keep = 128 * 16 * 40 + 800
    base = total - keep
    raw, _ = synth_raw(O.CS16, keep, first=base, rate=100e6)
    st = [("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)]
    o = oracle_chain(raw, O.CS16, 100_000_000, st, base, total)
    g = gpu_chain(raw, O.CS16, 100_000_000, st, base, total)
    rows = o.spark_rows(128, 128)
    assert g.spark_rows(128, 128) == rows == 4_194_303
    with kept_only():
        widx, wmag = o.spark_fft(128, 128, (0.5, 50.0), first_row=rows - 30, max_rows=30)
    idx, mag = g.spark_fft(128, 128, (0.5, 50.0), first_row=rows - 30, max_rows=30, want_mag=True)
    assert np.array_equal(idx, widx)
    assert_bit_equal(mag, wmag, "config 4 tail magnitudes")
    with pytest.raises(Q.QdError) as e:
        g.spark_fft(128, 128, (0.5, 50.0), first_row=0, max_rows=2)
    assert e.value.code == Q._lib.E_NOT_RESIDENT


def test_sparkfft_errors(Q):
    raw, _ = synth_raw(O.CS8, 3000)
    g = gpu_chain(raw, O.CS8, 1000, [])
    with pytest.raises(Q.QdError) as e:
        g.spark_fft(48, 48)
    assert e.value.code == O.E_FFT_WIDTH
    with pytest.raises(Q.QdError) as e:
        g.spark_fft(64, 0, max_rows=1)
    assert e.value.code == O.E_ZERO_STRIDE
    short = gpu_chain(raw[:20], O.CS8, 1000, [])
    with pytest.raises(Q.QdError) as e:
        short.spark_fft(64, 64, max_rows=1)  # len - width wraps; the first read_exact_at fails (fft.rs:28-30)
    assert e.value.code == O.E_SHORT_READ
    # graph[7] panic (fft.rs:59): with this range the f32 quotient reaches 7.0 for the magnitude just below max
    lo, hi, x = 0.8631789088249207, 1.9465599060058594, 1.9465597867965698
    assert O.glyph_index(x, lo, hi) == -1
    sig = np.zeros(16, dtype=np.complex64)
    sig[5] = x  # width-1 windows: the bin magnitude is |sample| exactly
    o = O.Samples.from_bytes(sig.view(np.uint8), O.CF32, 1000)
    with pytest.raises(O.OracleError) as eo:
        o.spark_fft(1, 1, (lo, hi))
    assert eo.value.code == O.E_GLYPH_RANGE
    with pytest.raises(Q.QdError) as e:
        Q.Samples.from_bytes(sig.view(np.uint8), Q.CF32, 1000).spark_fft(1, 1, (lo, hi))
    assert e.value.code == O.E_GLYPH_RANGE


# ---------------------------------------------------------------- freq_levels (a8), take_fft (a9), gen (a10)
def test_freq_levels_matches_oracle(Q):
    n = 40_000
    t = np.arange(n)
    f = np.where((t // 512) % 2 == 0, 0.11, -0.07)
    sig = (0.4 * np.exp(2j * np.pi * np.cumsum(f))).astype(np.complex64)
    noise = np.random.default_rng(3).standard_normal(2 * n).astype(np.float32).view(np.complex64) * np.float32(0.05)
    raw = (sig + noise).astype(np.complex64).view(np.uint8)
    for st, W, S in [([], 128, 128), ([("lowpass", 300_000, 4, 24)], 64, 20)]:
        o, g = oracle_chain(raw, O.CF32, 1_000_000, st), gpu_chain(raw, O.CF32, 1_000_000, st)
        with kept_only():
            want, wt = o.freq_levels(W, S)
        got, gt = g.freq_levels(W, S)
        assert gt == wt and np.array_equal(got, want)
        assert 0 < got.mean() < 1
    with pytest.raises(Q.QdError) as e:
        g.freq_levels(64, 64, levels=3)
    assert e.value.code == O.E_LEVELS


@pytest.mark.parametrize("W,bh", [(64, False), (256, True), (4096, True), (4, False), (12, True), (100, False), (513, True),
                                  (1000, False)])  # any width: FftPlanner takes them all (ffts.rs:25)
def test_take_fft_matches_oracle(Q, W, bh):
    raw, _ = synth_raw(O.CS8, 80_000)
    st = [("shift", 1_000_000), ("lowpass", 2_000_000, 4, 24)]
    o, g = oracle_chain(raw, O.CS8, 20_000_000, st), gpu_chain(raw, O.CS8, 20_000_000, st)
    with kept_only():
        want = o.take_fft(W, 37, blackman_harris=bh)
        want_s = o.take_fft(W, 11, slice_=(100, 9000), blackman_harris=bh)
    assert_bit_equal(g.take_fft(W, 37, blackman_harris=bh), want, "take_fft")
    assert_bit_equal(g.take_fft(W, 11, slice_=(100, 9000), blackman_harris=bh), want_s, "take_fft slice")
    for sl, code in [((100, 100), O.E_SLICE), ((100, 10**9), O.E_SLICE), ((100, 105), O.E_VISIBLE)]:
        with pytest.raises(Q.QdError) as e:
            g.take_fft(W, 10, slice_=sl)
        assert e.value.code == code


def test_gen_source_matches_oracle(Q):
    o = O.Samples.gen([1000, -2500, 123_456], 2_400_000, 0.01)
    g = Q.Samples.gen([1000, -2500, 123_456], 2_400_000, 0.01)
    assert g.len() == o.len() == 24_000
    assert_bit_equal(g.read_at(23_990, 40), o.read_at(23_990, 40), "gen past len")  # gen.rs:36,46
    assert_bit_equal(g.read_at(0, 20_000), o.read_at(0, 20_000), "gen")
    widx, wmag = o.shift(-100_000).lowpass(200_000, 8, 40).spark_fft(64, 32, (0.05, 3.0))
    idx, mag = g.shift(-100_000).lowpass(200_000, 8, 40).spark_fft(64, 32, (0.05, 3.0), want_mag=True)
    assert np.array_equal(idx, widx)
    assert_bit_equal(mag, wmag, "gen chain magnitudes")
    data, rc = g.write_mem()
    assert rc == 0 and len(data) == 6 * 0x1000  # whole chunks: Gen ignores len()


# ---------------------------------------------------------------- synthetic generator twin
@pytest.mark.parametrize("fmt", FORMATS)
def test_device_synth_matches_cpu_twin(Q, fmt):
    import torch

    n = 100_000
    first = 2**33 + 12345
    raw, p = synth_raw(fmt, n, first=first)
    pb = O.FORMAT_BYTES[fmt]
    buf = torch.empty(n * pb, dtype=torch.uint8, device="cuda")
    qp = Q.make_synth(p.seed, [(p.tone_step[i], p.tone_amp[i], p.key_period[i]) for i in range(p.n_tones)], p.noise_amp)
    Q.synth_fill_device(qp, fmt, first, n, buf.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(buf.cpu().numpy(), raw)


def test_device_resident_source_and_device_outputs(Q):
    import torch

    n = 300_000
    raw, _ = synth_raw(O.CS8, n)
    st = [("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)]
    d_raw = torch.from_numpy(raw).cuda()
    g = Q.Samples.from_device(d_raw.data_ptr(), raw.size, Q.CS8, 20_000_000, keep=(d_raw,))
    g = g.shift(1_500_000).lowpass(1_000_000, 8, 40)
    out = torch.zeros(2 * (n // 8 + 8), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own (non-blocking) stream
    got_n, rc = g.write_into(0x1000, 0, 10**6, out.data_ptr(), out.numel() // 2, Q._lib.SPACE_DEVICE)
    g.synchronize()
    with kept_only():
        want, want_rc = oracle_chain(raw, O.CS8, 20_000_000, st).write_mem()
    assert rc == want_rc and got_n == len(want)
    assert_bit_equal(out.cpu().numpy()[: 2 * got_n].view(np.complex64), want, "device write")
    rows = g.spark_rows(64, 64)
    d_idx = torch.zeros(rows * 64, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()  # the fill runs on torch's stream, the chain on its own (non-blocking) stream
    assert g.spark_fft_device(64, 64, (0.05, 2.0), 0, rows, d_idx.data_ptr()) == rows
    g.synchronize()
    with kept_only():
        widx, _ = oracle_chain(raw, O.CS8, 20_000_000, st).spark_fft(64, 64, (0.05, 2.0))
    assert np.array_equal(d_idx.cpu().numpy().reshape(rows, 64), widx)


# ---------------------------------------------------------------- fast STFT kernel, every width
@pytest.mark.parametrize("logw", range(0, 13))
def test_stft_every_width_both_epilogues(Q, logw):
    """fk_stft (qd_stft.cu): widths 1..4096, windows cut from raw bytes and from a filtered stream, glyph
    index via the f64 thresholds (no magnitudes requested) and via hypot (magnitudes requested)."""
    W = 1 << logw
    S = max(1, W // 3)
    n = max(20_000, 6 * W)
    raw, _ = synth_raw(O.CS8, n)
    rng = (0.3 * np.sqrt(W) / 8, 4.0 * np.sqrt(W))
    for stages in ([], [("lowpass", 4_000_000, 2, 4)]):
        o, g = oracle_chain(raw, O.CS8, 20_000_000, stages), gpu_chain(raw, O.CS8, 20_000_000, stages)
        with kept_only():
            widx, wmag = o.spark_fft(W, S, rng)
        idx_only, _ = g.spark_fft(W, S, rng, want_mag=False)
        idx, mag = g.spark_fft(W, S, rng, want_mag=True)
        assert widx.shape[0] > 3
        assert np.array_equal(idx_only, widx), f"threshold path W={W}: {(idx_only != widx).sum()} differ"
        assert np.array_equal(idx, widx), f"hypot path W={W}"
        assert_bit_equal(mag, wmag, f"magnitudes W={W}")
        assert len(np.unique(widx)) >= 3  # the range really splits the bins into several glyphs
        ref, _ = g.set_option("use_fast", 0).spark_fft(W, S, rng, want_mag=False)
        assert np.array_equal(ref, widx)


def test_stft_thresholds_on_special_values(Q):
    # magnitudes exactly on glyph boundaries, zeros, denormals, huge values, inf and NaN (fft.rs:53-60)
    def run(chain, W, S, rng_):
        try:
            idx, _ = chain.spark_fft(W, S, rng_)
            return 0, idx
        except (O.OracleError, Q.QdError) as e:
            return e.code, None

    # a range whose top glyph boundary does not reach graph[7] (the default 0.08:1 does: see below)
    lo, hi = next((a, b) for a, b in [(0.08, 1.0), (0.1, 1.5), (0.05, 2.0), (0.2, 3.0)]
                  if O.glyph_index(float(np.nextafter(np.float32(b), np.float32(0))), a, b) >= 0)
    d = np.float32((np.float32(hi) - np.float32(lo)) / np.float32(7))
    edges = [np.float32(lo) + np.float32(k) * d for k in range(8)]
    vals = []
    for e in edges:
        vals += [np.nextafter(e, np.float32(0)), e, np.nextafter(e, np.float32(4))]
    vals += [0.0, 1e-42, 1e-30, np.nextafter(np.float32(hi), np.float32(0)), hi, 3e38, np.inf, -np.inf, np.nan]
    vals = np.array(vals, dtype=np.float32)
    sig = np.zeros(2 * len(vals), dtype=np.complex64)
    sig.real[0::2] = vals          # real part only
    sig.imag[1::2] = vals          # imaginary part only
    sig = np.concatenate([sig, np.array([3 + 4j, 0.06 + 0.08j, complex(np.inf, np.nan), complex(np.nan, 1)], dtype=np.complex64),
                          np.zeros(2, dtype=np.complex64)])
    raw = sig.view(np.uint8)
    o = O.Samples.from_bytes(raw, O.CF32, 1000)
    g = Q.Samples.from_bytes(raw, Q.CF32, 1000)
    rc_o, widx = run(o, 1, 1, (lo, hi))                        # width 1: each magnitude is |sample|
    rc_g, idx = run(g, 1, 1, (lo, hi))
    assert rc_o == rc_g == 0
    assert np.array_equal(idx, widx), (idx.ravel(), widx.ravel())
    # the default range (which hits the reference's graph[7] panic just below 1.0), an inverted one, a tiny one
    for rng_ in (None, (0.5, 0.2), (0.0, 1e-3)):
        rc_o, widx = run(o, 2, 1, rng_)
        rc_g, idx = run(g, 2, 1, rng_)
        assert rc_o == rc_g, (rng_, rc_o, rc_g)
        if rc_o == 0:
            assert np.array_equal(idx, widx), rng_


# ---------------------------------------------------------------- glyphs through the linear form (qd_stft_epilogue.cuh glyph4)
LIN_RANGES = [(0.05, 2.0), (0.001, 0.01), (0.5, 50.0), (100.0, 100.5), (0.0, 1.0), (3.0, 4.0e6), (1e-20, 1e-18), (0.08, 1.0)]


def _boundary_magnitudes(lo, hi, seed):
    """f32 magnitudes on, next to and between the glyph boundaries of lo:hi (fft.rs:45-60)."""
    lo32, hi32 = np.float32(lo), np.float32(hi)
    d = np.float32((hi32 - lo32) / np.float32(7))
    vals = []
    for k in range(9):
        e = np.float32(lo32 + np.float32(k) * d) if k < 8 else hi32
        for near in (e, np.float32(float(e) * (1 - 3e-6)), np.float32(float(e) * (1 + 3e-6)), np.float32(float(e) * (1 - 4e-5)), np.float32(float(e) * (1 + 4e-5))):
            v = np.float32(near)
            run = [v]
            up, dn = v, v
            for _ in range(24):
                up = np.nextafter(up, np.float32(np.inf))
                dn = np.nextafter(dn, np.float32(0))
                run += [up, dn]
            vals += run
    r = np.random.default_rng(seed)
    vals += list(np.exp(r.uniform(np.log(max(lo, 1e-30) / 30), np.log(hi * 30), 600)).astype(np.float32))
    vals += list(r.uniform(lo, hi, 600).astype(np.float32))
    vals += [0.0, 1e-42, 1e-30, 1.9e19, 3e38, np.inf, np.nan]
    return np.array(vals, dtype=np.float32)


@pytest.mark.parametrize("rng_", LIN_RANGES)
@pytest.mark.parametrize("W", [4, 16, 64, 256, 4096])
def test_glyph_linear_form_agrees_on_and_around_every_boundary(Q, W, rng_):
    """A window that is v at sample 0 and zero elsewhere transforms to v in every bin, exactly (only additions of
    zeros), so the bins can be placed on, one ulp beside and a few 1e-6 beside every glyph boundary: the linear
    form must hand exactly those to the thresholds and agree with the oracle everywhere else.  v is split over
    re and im at several angles, so re^2 + im^2 rounds differently in f32 (kernel) and f64 (thresholds)."""
    mags = _boundary_magnitudes(*rng_, seed=W)
    if W == 4096:
        mags = mags[:: 7]

    def run(chain):
        try:
            return 0, chain.spark_fft(W, W, rng_)[0]
        except (O.OracleError, Q.QdError) as e:
            return e.code, None

    # first with every magnitude (a range whose top boundary reaches graph[7] panics: same status), then without
    # the few ulps below max where the reference panics
    for keep_panic_zone in (True, False):
        m = mags if keep_panic_zone else mags[~((mags > rng_[1] * (1 - 1e-5)) & (mags < rng_[1]))]
        ang = np.random.default_rng(W + 1).choice([0.0, np.pi / 2, 0.3, 0.7853981, 1.2, 2.9], size=len(m))
        sig = np.zeros((len(m), W), dtype=np.complex64)
        with np.errstate(invalid="ignore", over="ignore"):
            sig[:, 0] = (m * np.cos(ang)).astype(np.float32) + 1j * (m * np.sin(ang)).astype(np.float32)
        sig[ang == 0.0, 0] = m[ang == 0.0]  # the magnitude itself, bit for bit
        raw = sig.reshape(-1).view(np.uint8)
        rc_o, widx = run(O.Samples.from_bytes(raw, O.CF32, 1000))
        rc_g, idx = run(Q.Samples.from_bytes(raw, Q.CF32, 1000))
        rc_t, tidx = run(Q.Samples.from_bytes(raw, Q.CF32, 1000).set_option("glyph_lin", 0))
        assert rc_o == rc_g == rc_t, (keep_panic_zone, rc_o, rc_g, rc_t)
        if rc_o == 0:
            bad = np.argwhere(idx != widx)
            assert len(bad) == 0, (rng_, W, len(bad), [(float(m[r]), int(idx[r, c]), int(widx[r, c])) for r, c in bad[:5]])
            assert np.array_equal(tidx, widx)
            assert len(np.unique(widx)) >= 5
            break
    else:
        raise AssertionError(f"range {rng_} panics even without the zone below max")
