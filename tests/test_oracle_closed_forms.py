"""Checks the CPU oracle against closed forms derived from the reference source."""
import struct

import numpy as np
import pytest

import oracle_lib as O

f32 = np.float32


def _rand_bytes(seed, n):
    return np.random.default_rng(seed).integers(0, 256, size=n, dtype=np.uint8)


# ---- A.1 decode (src/lib.rs:241-255) ----
def test_decode_cs8_all_values():
    raw = np.arange(256, dtype=np.uint8).repeat(2)
    out = O.decode(O.CS8, raw)
    want = (np.arange(256).astype(np.uint8).view(np.int8).astype(f32) / f32(127.0)).astype(f32)
    assert np.array_equal(out.real, want) and np.array_equal(out.imag, want)
    assert out.real[128] == f32(-128.0) / f32(127.0)  # below -1.0, as the reference produces


def test_decode_cu8_all_values():
    raw = np.arange(256, dtype=np.uint8).repeat(2)
    out = O.decode(O.CU8, raw)
    want = (np.arange(256).astype(f32) / f32(255.0) - f32(127.5)).astype(f32)
    assert np.array_equal(out.real, want)
    assert out.real.min() == f32(-127.5) and out.real.max() == f32(-126.5)


def test_decode_cs16_all_values_and_iq_order():
    v = np.arange(-32768, 32768, dtype=np.int16)
    raw = np.stack([v, v[::-1]], axis=1).reshape(-1).view(np.uint8)
    out = O.decode(O.CS16, raw)
    want = (v.astype(f32) / f32(65535.0) - f32(32767.5)).astype(f32)
    assert np.array_equal(out.real, want) and np.array_equal(out.imag, want[::-1])
    assert len(np.unique(out.real)) <= 514  # ulp 2^-9 at 32768: ~512 levels survive


def test_decode_cf32_is_a_bit_copy_including_nan_payloads():
    words = np.array([0x7FC00001, 0xFFC12345, 0x00000001, 0x80000000, 0x7F800000, 0x3F800000], dtype=np.uint32)
    out = O.decode(O.CF32, words.view(np.uint8))
    assert np.array_equal(out.view(np.uint32), words)


# ---- A.2 file source (src/samples.rs:63-94) ----
def test_file_len_floor_and_short_read():
    raw = _rand_bytes(1, 4 * 10 + 3)  # 10 cs16 samples and a trailing partial pair
    s = O.Samples.from_bytes(raw, O.CS16, 1000)
    assert s.len() == 10
    assert len(s.read_at(7, 8)) == 3
    with pytest.raises(O.OracleError) as e:
        s.read_at(10, 1)
    assert e.value.code == O.E_OFFSET_EOF
    with pytest.raises(O.OracleError) as e:
        s.read_exact_at(7, 8)
    assert e.value.code == O.E_SHORT_READ


# ---- A.3 shift (src/shift.rs) ----
def test_shift_ratio_and_asserts():
    assert O.shift_ratio(280_000, 21_000_000) == (np.pi * 2.0) * 280_000.0 / 21_000_000.0
    s = O.Samples.from_bytes(_rand_bytes(2, 64), O.CS8, 1000)
    with pytest.raises(O.OracleError) as e:
        s.shift(500)
    assert e.value.code == O.E_SHIFT_NYQUIST
    s = O.Samples.from_bytes(_rand_bytes(2, 64), O.CS8, 1001)
    with pytest.raises(O.OracleError):
        s.shift(-500)  # |f| < (1001/2)=500 fails
    s.shift(-499)


def test_shift_matches_numpy_f32_ops():
    raw = _rand_bytes(3, 2 * 500)
    base = O.decode(O.CS8, raw)
    s = O.Samples.from_bytes(raw, O.CS8, 48_000).shift(-7_000)
    got = s.read_at(100, 300)
    ratio = O.shift_ratio(-7_000, 48_000)
    n = np.arange(100, 400, dtype=np.uint64).astype(np.float64)
    place = n * ratio
    c, sn = np.cos(place).astype(f32), np.sin(place).astype(f32)
    a, b = base.real[100:400], base.imag[100:400]
    re = (a * c).astype(f32) - (b * sn).astype(f32)
    im = (b * c).astype(f32) + (a * sn).astype(f32)
    # numpy's cos/sin may differ from glibc in the last f64 bit; after rounding to f32 they agree here
    assert np.array_equal(got.real, re) and np.array_equal(got.imag, im)


# ---- A.4 lowpass (src/filter.rs) ----
def test_taps_config1_statistics():
    t = O.taps(200_000, 21_000_000, 400)
    seq = f32(0)
    for v in t:
        seq = f32(seq + v)
    assert abs(float(seq) - 1.0) < 2e-7
    assert t.max() == pytest.approx(0.019039, rel=1e-4)
    assert abs(t[0]) < 1e-9
    assert np.abs(t - t[::-1]).max() > 0  # not exactly symmetric: taps must not be folded


def test_taps_formula_f32():
    L, freq, sr = 40, 1_000_000, 20_000_000
    t = O.taps(freq, sr, L)
    PI = f32(np.pi)
    cutoff = f32(freq / sr)
    i = np.arange(L).astype(f32)
    x = (f32(2.0) * cutoff) * (i - (f32(L) - f32(1)) / f32(2))
    xp = (x * PI).astype(f32)
    wave = (np.sin(xp.astype(np.float64)).astype(f32) / xp).astype(f32)
    a1 = ((f32(2.0) * PI) * i / (f32(L) - f32(1))).astype(f32)
    a2 = ((f32(4.0) * PI) * i / (f32(L) - f32(1))).astype(f32)
    win = ((f32(0.42) - f32(0.5) * np.cos(a1.astype(np.float64)).astype(f32)).astype(f32)
           + f32(0.08) * np.cos(a2.astype(np.float64)).astype(f32)).astype(f32)
    tt = (wave * win).astype(f32)
    ssum = f32(0)
    for v in tt:
        ssum = f32(ssum + v)
    want = (tt / ssum).astype(f32)
    # sinf/cosf (glibc, correctly rounded in practice) vs f64 numpy rounded to f32: allow 1 ulp
    assert np.abs(t - want).max() <= np.spacing(np.abs(want).max())


def _closed_form(raw_cf32, taps, D, n, off, total):
    """y[k] = sum_{j<J(k)} raw[k*D + L - L/2 + j] * f[j], sequential f32 mul-then-add."""
    L = len(taps)
    start = off * D
    valid = min(n * D + L, total - start)
    m = (valid - L) // D
    out = np.zeros(m, dtype=np.complex64)
    i0 = L - L // 2
    for k in range(m):
        J = min(L, valid - k * D - i0)
        re, im = f32(0), f32(0)
        for j in range(J):
            x = raw_cf32[start + k * D + i0 + j]
            re = f32(re + f32(x.real * taps[j]))
            im = f32(im + f32(x.imag * taps[j]))
        out[k] = re + 1j * im
    return out


@pytest.mark.parametrize("fmt,L,D,n,off", [(O.CS8, 40, 8, 64, 3), (O.CU8, 10, 3, 7, 0), (O.CS16, 16, 4, 20, 5),
                                           (O.CF32, 12, 5, 9, 2), (O.CS8, 7, 2, 11, 1)])
def test_lowpass_closed_form_and_truncation(fmt, L, D, n, off):
    total = 2000
    raw = _rand_bytes(10 + L, total * O.FORMAT_BYTES[fmt])
    if fmt == O.CF32:
        raw = (np.random.default_rng(5).standard_normal(2 * total).astype(f32)).view(np.uint8)
    base = O.decode(fmt, raw)
    t = O.taps(1000, 48_000, L)
    want = _closed_form(base, t, D, n, off, total)
    for kept in (False, True):
        O.set_kept_only(kept)
        try:
            got = O.Samples.from_bytes(raw, fmt, 48_000).lowpass(1000, D, L).read_at(off, n)
        finally:
            O.set_kept_only(False)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), kept
    assert len(want) == n


def test_truncated_positions_config_shapes():
    # positions with (n-k)*D + L/2 < L use a zero-truncated filter (filter.rs:68-80,107-124)
    def truncated(L, D, n):
        return [k for k in range(n) if (n - k) * D + L // 2 < L]

    assert truncated(400, 32, 64) == list(range(58, 64))       # config 1
    assert truncated(800, 16, 128) == list(range(104, 128))    # config 4
    assert truncated(40, 8, 4096) == [4094, 4095]              # config 2 via write
    assert truncated(40, 32, 4) == []                          # config 5 outer stage


def test_read_size_dependence_is_reproduced():
    raw = _rand_bytes(77, 2 * 5000)
    s = O.Samples.from_bytes(raw, O.CS8, 48_000).lowpass(2000, 8, 40)
    a = s.read_at(10, 16)
    b = s.read_at(10, 64)[:16]
    assert np.array_equal(a[:14], b[:14]) and not np.array_equal(a[14:], b[14:])


def test_lowpass_len_over_reports_by_one_and_write_panics_after_data():
    # SURVEY A.4 Q8; lib.rs:203
    for total, L, D in [(1000, 40, 8), (1001, 40, 8), (4136, 40, 8), (999, 10, 3)]:
        raw = _rand_bytes(total, 2 * total)
        s = O.Samples.from_bytes(raw, O.CS8, 48_000).lowpass(2000, D, L)
        ln = s.len()
        assert ln == 1 + (total - L) // D
        assert len(s.read_at(ln - 1, 16)) == 0
        data, rc = s.write_mem()
        assert rc == O.E_WRITE_SHORT and len(data) == ln - 1
        assert s.sample_rate() == 48_000 // D


def test_write_chunks_use_per_chunk_truncation():
    total = 8 * 0x1000 * 2 + 500
    raw = _rand_bytes(9, 2 * total)
    s = O.Samples.from_bytes(raw, O.CS8, 20_000_000).shift(1_500_000).lowpass(1_000_000, 8, 40)
    O.set_kept_only(True)
    try:
        data, rc = s.write_mem()
        c1, _ = s.write_mem(first_chunk=1, max_chunks=1)
        big = s.read_at(0, 3 * 0x1000)
    finally:
        O.set_kept_only(False)
    assert rc == O.E_WRITE_SHORT
    assert np.array_equal(data[0x1000:0x2000], c1)
    # chunked output differs from one big read exactly at the last two positions of each chunk
    diff = np.nonzero(data[: 2 * 0x1000] != big[: 2 * 0x1000])[0].tolist()
    assert diff == [4094, 4095, 8190, 8191]


def test_short_input_panics_map_to_codes():
    raw = _rand_bytes(4, 2 * 30)
    s = O.Samples.from_bytes(raw, O.CS8, 48_000).lowpass(2000, 8, 40)
    with pytest.raises(O.OracleError) as e:
        s.len()
    assert e.value.code == O.E_SHORT_INPUT
    raw = _rand_bytes(4, 2 * 100)
    s = O.Samples.from_bytes(raw, O.CS8, 48_000).lowpass(2000, 8, 40)
    with pytest.raises(O.OracleError) as e:
        s.read_at(9, 4)  # inner returns 28 < 40 samples
    assert e.value.code == O.E_SHORT_INPUT


# ---- FFT (own definition; rustfft absent) ----
@pytest.mark.parametrize("n", [1, 2, 4, 8, 16, 32, 64, 128, 512, 4096])
def test_fft_bounded_by_complex128_dft(n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    got = O.fft(x)
    want = np.fft.fft(x.astype(np.complex128))
    err = np.abs(got - want).max() / np.abs(want).max()
    assert err < 4e-7 * max(1, np.log2(n)), err
    if n <= 128:
        assert np.abs(O.dft_c128(x) - want).max() < 1e-10 * max(1.0, np.abs(want).max())


def test_fft_impulse_and_tone_are_exactly_placed():
    x = np.zeros(64, dtype=np.complex64)
    x[0] = 1
    assert np.array_equal(O.fft(x), np.ones(64, dtype=np.complex64))
    k = 5
    tone = np.exp(2j * np.pi * k * np.arange(64) / 64).astype(np.complex64)
    assert np.argmax(np.abs(O.fft(tone))) == k  # forward kernel e^{-2 pi i jk/N}


def test_fft_width_must_be_power_of_two():
    s = O.Samples.from_bytes(_rand_bytes(1, 2 * 1000), O.CS8, 1000)
    with pytest.raises(O.OracleError) as e:
        s.spark_fft(48, 48)
    assert e.value.code == O.E_FFT_WIDTH


# ---- A.5 sparkfft (src/fft.rs:12-69) ----
def test_glyph_index_thresholds():
    assert O.glyph_index(0.07999, 0.08, 1.0) == 0
    assert O.glyph_index(0.08, 0.08, 1.0) == 1
    assert O.glyph_index(1.0, 0.08, 1.0) == 8
    assert O.glyph_index(float("nan"), 0.08, 1.0) == 1  # NaN: both compares false, `as usize` -> 0
    assert O.glyph_index(float("inf"), 0.08, 1.0) == 8
    lo, hi = f32(0.08), f32(1.0)
    d = f32((hi - lo) / f32(7))
    xs = np.linspace(0.08, 0.99999, 20001).astype(f32)
    got = np.array([O.glyph_index(float(x), 0.08, 1.0) for x in xs])
    want = 1 + np.floor(((xs - lo) / d).astype(f32)).astype(int)
    assert np.array_equal(got[want <= 7], want[want <= 7])
    assert (np.diff(got[got >= 0]) >= 0).all()


def test_spark_rows_is_ceil_and_order_is_fftshifted():
    n = 1000
    t = np.arange(n)
    sig = (0.5 * np.exp(2j * np.pi * 0.25 * t)).astype(np.complex64)  # +fs/4
    s = O.Samples.from_bytes(sig.view(np.uint8), O.CF32, 1000)
    assert s.spark_rows(64, 64) == -(-(n - 64) // 64)
    assert s.spark_rows(64, 10) == -(-(n - 64) // 10)
    idx, mag = s.spark_fft(64, 64)
    assert (np.argmax(mag, axis=1) == 32 + 16).all()  # bin 16 sits at display column 48
    assert (idx[:, 48] == 8).all()                    # |X| = 32 >= max


def test_spark_fft_len_not_greater_than_width():
    s = O.Samples.from_bytes(np.zeros(8 * 64, dtype=np.uint8), O.CF32, 1000)
    assert s.spark_rows(64, 64) == 0
    with pytest.raises(O.OracleError) as e:  # len - width == 0: loop does not run at all
        s2 = O.Samples.from_bytes(np.zeros(8 * 10, dtype=np.uint8), O.CF32, 1000)
        s2.spark_fft(64, 64, max_rows=1)
    assert e.value.code == O.E_SHORT_READ  # wrapped limit, first read_exact_at fails (fft.rs:28-30)


# ---- A.6 freq_levels, A.7 take_fft, A.8 gen ----
def test_freq_levels_two_tone():
    n = 2048
    t = np.arange(n)
    f = np.where((t // 256) % 2 == 0, 0.125, -0.125)  # alternate upper / lower half of the spectrum
    sig = (0.5 * np.exp(2j * np.pi * np.cumsum(f))).astype(np.complex64)
    s = O.Samples.from_bytes(sig.view(np.uint8), O.CF32, 1000)
    vals, total = s.freq_levels(64, 64)
    assert total == (n - 64) // 64  # floor, unlike sparkfft's ceil
    # positive frequency -> bins 0..W/2 (natural order) -> first >= second -> 1
    assert vals[0] == 1 and vals[4] == 0 and vals[8] == 1
    with pytest.raises(O.OracleError) as e:
        s.freq_levels(64, 64, levels=3)
    assert e.value.code == O.E_LEVELS


def test_take_fft_rows_and_window():
    n = 5000
    rng = np.random.default_rng(0)
    sig = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    s = O.Samples.from_bytes(sig.view(np.uint8), O.CF32, 1000)
    out = s.take_fft(64, 10, slice_=(100, 3000))
    step = (3000 - 100) / 10
    for i in (0, 3, 9):
        at = 100 + int(np.floor(step * i + 0.5))
        want = np.abs(np.fft.fftshift(np.fft.fft(sig[at : at + 64].astype(np.complex128))))
        assert np.allclose(out[i], want, rtol=2e-5, atol=1e-5)
    w = O.blackman_harris(64)
    assert w[0] == pytest.approx(0.35875 - 0.48829 + 0.14128 - 0.01168, abs=1e-6) and w.argmax() in (31, 32)
    outw = s.take_fft(64, 4, blackman_harris=True)
    at = int(np.floor((n - 64) / 4 * 1 + 0.5))
    want = np.abs(np.fft.fftshift(np.fft.fft(sig[at : at + 64].astype(np.complex128) * w)))
    assert np.allclose(outw[1], want, rtol=2e-5, atol=1e-5)
    with pytest.raises(O.OracleError) as e:
        s.take_fft(64, 10, slice_=(100, 100))
    assert e.value.code == O.E_SLICE
    with pytest.raises(O.OracleError) as e:
        s.take_fft(64, 10, slice_=(100, 5000))
    assert e.value.code == O.E_SLICE
    with pytest.raises(O.OracleError) as e:
        s.take_fft(64, 50, slice_=(100, 150))
    assert e.value.code == O.E_VISIBLE


def test_gen_matches_formula_and_ignores_len():
    g = O.Samples.gen([1000, -2500], 48_000, 0.01)
    assert g.len() == 480 and g.sample_rate() == 48_000
    got = g.read_at(470, 40)  # read_at fills the whole buffer past len() (gen.rs:36,46)
    assert len(got) == 40
    n = np.arange(470, 510).astype(np.float64)
    base = n * (np.pi * 2.0) / 48_000.0
    re = np.zeros(40, dtype=f32)
    im = np.zeros(40, dtype=f32)
    for fr in (1000.0, -2500.0):
        re = (re + np.cos(fr * base).astype(f32)).astype(f32)
        im = (im + np.sin(fr * base).astype(f32)).astype(f32)
    assert np.abs(got.real - re).max() <= 1.2e-7 and np.abs(got.imag - im).max() <= 1.2e-7
    for bad in ([], ):
        with pytest.raises(O.OracleError) as e:
            O.Samples.gen(bad, 48_000, 1.0)
        assert e.value.code == O.E_GEN_ARGS
    with pytest.raises(O.OracleError):
        O.Samples.gen([1], 0, 1.0)
    with pytest.raises(O.OracleError):
        O.Samples.gen([1], 10, 0.0)


def test_write_file_roundtrip(tmp_path):
    g = O.Samples.gen([100], 8000, 0.5)
    name = g.write_file(str(tmp_path / "tone"))
    assert name.endswith("tone.sr8000.cf32")  # lib.rs:194; re-parsable by `from` (args.rs:328-333)
    data = np.fromfile(name, dtype=np.complex64)
    # Gen::read_at ignores len() (gen.rs:36,46), so `gen | write` emits whole 0x1000 chunks
    assert g.len() == 4000 and len(data) == 4096 and np.array_equal(data, g.read_at(0, 4096))
    with pytest.raises(O.OracleError) as e:
        g.write_file(str(tmp_path / "tone"))
    assert e.value.code == O.E_EXISTS
    g.write_file(str(tmp_path / "tone"), overwrite=True)
    with pytest.raises(O.OracleError) as e:
        g.write_file("-")
    assert e.value.code == O.E_UNIMPLEMENTED


# ---- synthetic generator twin ----
def test_synth_is_index_keyed_and_in_range():
    p = O.make_synth(0x5EED0002, [(O.tone_step(1.5e6, 20e6), 40, 0), (O.tone_step(-3e6, 20e6), 25, 1000)], 6)
    for fmt in (O.CS8, O.CU8, O.CS16, O.CF32):
        a = O.synth_fill(p, fmt, 0, 5000)
        b = O.synth_fill(p, fmt, 1234, 1000)
        pb = O.FORMAT_BYTES[fmt]
        assert np.array_equal(a[1234 * pb : 2234 * pb], b)
    big = O.synth_fill(p, O.CS8, 2**33 + 5, 64)
    assert big.view(np.int8).std() > 5
    x = O.decode(O.CS8, O.synth_fill(p, O.CS8, 0, 4096))
    spec = np.abs(np.fft.fft(x))
    assert np.argmax(spec) == round(1.5e6 / 20e6 * 4096)
