"""ctypes binding for oracle/_build/libquadrs_oracle.so.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB_PATH = ORACLE_DIR / "_build" / "libquadrs_oracle.so"

CF32, CS8, CU8, CS16 = 0, 1, 2, 3
FORMAT_BYTES = {CF32: 8, CS8: 2, CU8: 2, CS16: 4}

OK = 0
E_INVALID_ARG, E_SHIFT_NYQUIST, E_ZERO_RATE, E_OFFSET_EOF, E_SHORT_INPUT = 1, 2, 3, 4, 5
E_SHORT_READ, E_FFT_WIDTH, E_GLYPH_RANGE, E_LEVELS, E_SLICE, E_VISIBLE = 6, 7, 8, 9, 10, 11
E_GEN_ARGS, E_WRITE_SHORT, E_IO = 12, 13, 14
E_UNIMPLEMENTED, E_EXISTS, E_NOMEM, E_ZERO_STRIDE = 17, 18, 19, 20


class OracleError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def build() -> Path:
    src_mtime = max(p.stat().st_mtime for p in ORACLE_DIR.glob("*.[ch]"))
    if not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < src_mtime:
        subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, capture_output=True)
    return LIB_PATH


class Synth(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("n_tones", C.c_uint32),
        ("tone_step", C.c_uint32 * 8),
        ("tone_amp", C.c_int32 * 8),
        ("key_period", C.c_uint32 * 8),
        ("noise_amp", C.c_int32),
    ]


class Job(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("n_bytes", C.c_uint64),
        ("format", C.c_int),
        ("sample_rate", C.c_uint64),
        ("n_stages", C.c_uint32),
        ("stage_kind", C.c_int32 * 8),
        ("stage_freq", C.c_int64 * 8),
        ("stage_decimate", C.c_uint64 * 8),
        ("stage_size", C.c_uint64 * 8),
        ("sink", C.c_int),
        ("width", C.c_uint64),
        ("stride", C.c_uint64),
        ("has_range", C.c_int),
        ("min", C.c_float),
        ("max", C.c_float),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(str(build()))
    vp, u64, sz, i32, f32 = C.c_void_p, C.c_uint64, C.c_size_t, C.c_int, C.c_float
    L.qo_last_error.restype = C.c_char_p
    L.qo_from_mem.restype = vp
    L.qo_from_mem.argtypes = [vp, u64, i32, u64]
    L.qo_from_mem_window.restype = vp
    L.qo_from_mem_window.argtypes = [vp, u64, i32, u64, u64, u64]
    L.qo_from_file.restype = vp
    L.qo_from_file.argtypes = [C.c_char_p, i32, u64]
    L.qo_gen.argtypes = [C.POINTER(C.c_int64), sz, u64, C.c_double, C.POINTER(vp)]
    L.qo_shift.argtypes = [vp, C.c_int64, C.POINTER(vp)]
    L.qo_lowpass.argtypes = [vp, u64, u64, sz, C.POINTER(vp)]
    L.qo_free.argtypes = [vp]
    L.qo_free.restype = None
    L.qo_len.argtypes = [vp, C.POINTER(u64)]
    L.qo_sample_rate.argtypes = [vp]
    L.qo_sample_rate.restype = u64
    L.qo_read_at.argtypes = [vp, u64, vp, sz, C.POINTER(sz)]
    L.qo_read_exact_at.argtypes = [vp, u64, vp, sz]
    L.qo_set_kept_only_convolve.argtypes = [i32]
    L.qo_set_kept_only_convolve.restype = None
    L.qo_spark_fft.argtypes = [vp, sz, u64, i32, f32, i32, f32, u64, u64, vp, vp, C.POINTER(u64)]
    L.qo_spark_rows.argtypes = [vp, sz, u64, C.POINTER(u64)]
    L.qo_spark_fft_text.argtypes = [vp, sz, u64, i32, f32, i32, f32, vp, sz, C.POINTER(sz)]
    L.qo_freq_levels.argtypes = [vp, sz, u64, sz, u64, u64, vp, C.POINTER(u64)]
    L.qo_take_fft.argtypes = [vp, i32, u64, u64, sz, i32, sz, vp]
    L.qo_write_mem.argtypes = [vp, sz, u64, u64, vp, u64, C.POINTER(u64)]
    L.qo_write_file.argtypes = [vp, C.c_char_p, i32, C.c_char_p, sz]
    L.qo_decode.argtypes = [i32, vp, sz, vp]
    L.qo_decode.restype = None
    L.qo_taps.argtypes = [u64, u64, sz, vp]
    L.qo_blackman_harris.argtypes = [sz, vp]
    L.qo_blackman_harris.restype = None
    L.qo_fft.argtypes = [vp, sz]
    L.qo_dft_c128.argtypes = [vp, sz, vp]
    L.qo_dft_c128.restype = None
    L.qo_shift_ratio.argtypes = [C.c_int64, u64]
    L.qo_shift_ratio.restype = C.c_double
    L.qo_glyph_index.argtypes = [f32, f32, f32]
    L.qo_format_row.argtypes = [vp, sz, vp]
    L.qo_format_row.restype = sz
    L.qo_synth_fill.argtypes = [C.POINTER(Synth), i32, u64, u64, vp]
    L.qo_synth_fill.restype = None
    L.qo_timed_run.argtypes = [C.POINTER(Job), u64, u64, i32, C.POINTER(u64)]
    L.qo_timed_run.restype = C.c_double
    _lib = L
    return L


def _check(rc: int):
    if rc != OK:
        raise OracleError(rc, lib().qo_last_error().decode("utf-8", "replace"))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Samples:
    """A node of the oracle's lazy pull graph (trait Samples, src/samples.rs:11-28)."""

    def __init__(self, handle, keep=()):
        self._h = handle
        self._keep = keep  # numpy buffers the C side borrows

    # -- construction (Operation::exec arms, src/lib.rs:89-121) --
    @staticmethod
    def from_bytes(data, fmt: int, sample_rate: int) -> "Samples":
        arr = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1)
        arr = np.ascontiguousarray(arr)
        h = lib().qo_from_mem(_ptr(arr), arr.size, fmt, sample_rate)
        if not h:
            raise OracleError(E_INVALID_ARG, "qo_from_mem failed")
        return Samples(h, (arr,))

    @staticmethod
    def from_window(data: np.ndarray, fmt: int, sample_rate: int, base_sample: int, total_samples: int) -> "Samples":
        arr = np.ascontiguousarray(data.view(np.uint8).reshape(-1))
        h = lib().qo_from_mem_window(_ptr(arr), arr.size, fmt, sample_rate, base_sample, total_samples)
        if not h:
            raise OracleError(E_INVALID_ARG, "qo_from_mem_window failed")
        return Samples(h, (arr,))

    @staticmethod
    def from_file(path, fmt: int, sample_rate: int) -> "Samples":
        h = lib().qo_from_file(os.fsencode(str(path)), fmt, sample_rate)
        if not h:
            raise OracleError(E_IO, lib().qo_last_error().decode())
        return Samples(h)

    @staticmethod
    def gen(cos_hz, sample_rate: int, seconds: float = 1.0) -> "Samples":
        arr = (C.c_int64 * max(1, len(cos_hz)))(*cos_hz)
        out = C.c_void_p()
        _check(lib().qo_gen(arr, len(cos_hz), sample_rate, seconds, C.byref(out)))
        return Samples(out.value)

    def _take(self):
        h, self._h = self._h, None
        return h

    def shift(self, frequency: int) -> "Samples":
        out = C.c_void_p()
        rc = lib().qo_shift(self._h, frequency, C.byref(out))
        _check(rc)
        keep = self._keep
        self._take()
        return Samples(out.value, keep)

    def lowpass(self, frequency: int, decimate: int = 8, size: int = 40) -> "Samples":
        out = C.c_void_p()
        _check(lib().qo_lowpass(self._h, frequency, decimate, size, C.byref(out)))
        keep = self._keep
        self._take()
        return Samples(out.value, keep)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().qo_free(self._h)
            self._h = None

    # -- trait Samples --
    def len(self) -> int:
        v = C.c_uint64()
        _check(lib().qo_len(self._h, C.byref(v)))
        return v.value

    def sample_rate(self) -> int:
        return lib().qo_sample_rate(self._h)

    def read_at(self, off: int, n: int) -> np.ndarray:
        buf = np.zeros(n, dtype=np.complex64)
        got = C.c_size_t()
        _check(lib().qo_read_at(self._h, off, _ptr(buf), n, C.byref(got)))
        return buf[: got.value]

    def read_exact_at(self, off: int, n: int) -> np.ndarray:
        buf = np.zeros(n, dtype=np.complex64)
        _check(lib().qo_read_exact_at(self._h, off, _ptr(buf), n))
        return buf

    # -- sinks --
    def spark_rows(self, width: int, stride: int) -> int:
        v = C.c_uint64()
        _check(lib().qo_spark_rows(self._h, width, stride, C.byref(v)))
        return v.value

    def spark_fft(self, width=128, stride=None, rng=None, first_row=0, max_rows=None, want_mag=True):
        stride = width if stride is None else stride
        if max_rows is None:
            max_rows = max(0, self.spark_rows(width, stride) - first_row)
        idx = np.zeros((max_rows, width), dtype=np.uint8)
        mag = np.zeros((max_rows, width), dtype=np.float32) if want_mag else None
        rows = C.c_uint64()
        has = 1 if rng is not None else 0
        lo, hi = rng if rng is not None else (0.0, 0.0)
        rc = lib().qo_spark_fft(self._h, width, stride, has, lo, has, hi, first_row, max_rows, _ptr(idx),
                                _ptr(mag) if want_mag else None, C.byref(rows))
        _check(rc)
        r = rows.value
        return idx[:r], (mag[:r] if want_mag else None)

    def spark_fft_text(self, width=128, stride=None, rng=None) -> str:
        stride = width if stride is None else stride
        rows = self.spark_rows(width, stride)
        cap = 64 + (rows + 1) * (3 * width + 8)
        buf = C.create_string_buffer(cap)
        n = C.c_size_t()
        has = 1 if rng is not None else 0
        lo, hi = rng if rng is not None else (0.0, 0.0)
        _check(lib().qo_spark_fft_text(self._h, width, stride, has, lo, has, hi, buf, cap, C.byref(n)))
        return buf.raw[: n.value].decode("utf-8")

    def freq_levels(self, width=128, stride=None, levels=2, first=0, max_n=None):
        stride = width if stride is None else stride
        total = C.c_uint64()
        if max_n is None:
            probe = np.zeros(1, dtype=np.uint8)
            _check(lib().qo_freq_levels(self._h, width, stride, levels, 0, 0, _ptr(probe), C.byref(total)))
            max_n = max(0, total.value - first)
        vals = np.zeros(max(1, max_n), dtype=np.uint8)
        _check(lib().qo_freq_levels(self._h, width, stride, levels, first, max_n, _ptr(vals), C.byref(total)))
        return vals[: min(max_n, max(0, total.value - first))], total.value

    def take_fft(self, width: int, output_len: int, slice_=None, blackman_harris=False) -> np.ndarray:
        out = np.zeros((output_len, width), dtype=np.float32)
        has = 1 if slice_ is not None else 0
        a, b = slice_ if slice_ is not None else (0, 0)
        _check(lib().qo_take_fft(self._h, has, a, b, width, 1 if blackman_harris else 0, output_len, _ptr(out)))
        return out

    def write_mem(self, chunk=0x1000, first_chunk=0, max_chunks=None, allow_short=True):
        """do_write's pull loop.  Returns (samples, status) -- status is E_WRITE_SHORT when
        the reference's assert at lib.rs:203 would fire after the data is out."""
        if max_chunks is None:
            max_chunks = (self.len() + chunk - 1) // chunk + 1
        cap = max_chunks * chunk
        out = np.zeros(cap, dtype=np.complex64)
        n = C.c_uint64()
        rc = lib().qo_write_mem(self._h, chunk, first_chunk, max_chunks, _ptr(out), cap, C.byref(n))
        if rc != OK and not (allow_short and rc == E_WRITE_SHORT):
            _check(rc)
        return out[: n.value], rc

    def write_file(self, prefix: str, overwrite=False) -> str:
        name = C.create_string_buffer(4096)
        _check(lib().qo_write_file(self._h, os.fsencode(prefix), 1 if overwrite else 0, name, 4096))
        return os.fsdecode(name.value)


def set_kept_only(on: bool):
    lib().qo_set_kept_only_convolve(1 if on else 0)


def decode(fmt: int, raw: np.ndarray) -> np.ndarray:
    raw = np.ascontiguousarray(raw.view(np.uint8).reshape(-1))
    n = raw.size // FORMAT_BYTES[fmt]
    out = np.zeros(n, dtype=np.complex64)
    lib().qo_decode(fmt, _ptr(raw), n, _ptr(out))
    return out


def taps(frequency: int, sample_rate: int, size: int) -> np.ndarray:
    out = np.zeros(size, dtype=np.float32)
    _check(lib().qo_taps(frequency, sample_rate, size, _ptr(out)))
    return out


def blackman_harris(n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.float32)
    lib().qo_blackman_harris(n, _ptr(out))
    return out


def fft(x: np.ndarray) -> np.ndarray:
    buf = np.ascontiguousarray(x, dtype=np.complex64).copy()
    _check(lib().qo_fft(_ptr(buf), buf.size))
    return buf


def dft_c128(x: np.ndarray) -> np.ndarray:
    buf = np.ascontiguousarray(x, dtype=np.complex64)
    out = np.zeros(buf.size, dtype=np.complex128)
    lib().qo_dft_c128(_ptr(buf), buf.size, _ptr(out))
    return out


def shift_ratio(frequency: int, sample_rate: int) -> float:
    return lib().qo_shift_ratio(frequency, sample_rate)


def glyph_index(norm: float, lo: float, hi: float) -> int:
    return lib().qo_glyph_index(norm, lo, hi)


def format_row(idx: np.ndarray) -> str:
    idx = np.ascontiguousarray(idx, dtype=np.uint8)
    buf = C.create_string_buffer(3 * idx.size + 16)
    n = lib().qo_format_row(_ptr(idx), idx.size, buf)
    return buf.raw[:n].decode("utf-8")


def make_synth(seed: int, tones, noise_amp: int = 0) -> Synth:
    """tones: list of (step_u32, amp, key_period)."""
    p = Synth()
    p.seed = seed
    p.n_tones = len(tones)
    for i, (step, amp, key) in enumerate(tones):
        p.tone_step[i] = step & 0xFFFFFFFF
        p.tone_amp[i] = amp
        p.key_period[i] = key
    p.noise_amp = noise_amp
    return p


def tone_step(freq_hz: float, sample_rate: float) -> int:
    return int(round(freq_hz / sample_rate * 2**32)) & 0xFFFFFFFF


def synth_fill(p: Synth, fmt: int, first: int, n: int) -> np.ndarray:
    out = np.zeros(n * FORMAT_BYTES[fmt], dtype=np.uint8)
    lib().qo_synth_fill(C.byref(p), fmt, first, n, _ptr(out))
    return out


def timed_run(raw: np.ndarray, fmt: int, sample_rate: int, stages, sink: str, first_unit: int, n_units: int,
              n_threads: int, width=0, stride=0, rng=None):
    """stages: list of ('shift', f) / ('lowpass', f, decimate, size).  Returns (seconds, checksum)."""
    raw = np.ascontiguousarray(raw.view(np.uint8).reshape(-1))
    j = Job()
    j.data = raw.ctypes.data
    j.n_bytes = raw.size
    j.format = fmt
    j.sample_rate = sample_rate
    j.n_stages = len(stages)
    for i, st in enumerate(stages):
        if st[0] == "shift":
            j.stage_kind[i], j.stage_freq[i] = 1, st[1]
        else:
            j.stage_kind[i], j.stage_freq[i], j.stage_decimate[i], j.stage_size[i] = 2, st[1], st[2], st[3]
    j.sink = 0 if sink == "write" else 1
    j.width, j.stride = width, stride
    j.has_range = 1 if rng is not None else 0
    if rng is not None:
        j.min, j.max = rng
    lib().qo_set_kept_only_convolve(0)
    cs = C.c_uint64()
    secs = lib().qo_timed_run(C.byref(j), first_unit, n_units, n_threads, C.byref(cs))
    if secs < 0:
        raise OracleError(int(-secs), "qo_timed_run failed")
    return secs, cs.value
