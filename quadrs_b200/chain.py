"""Host-side mirror of the reference's operator interface over the C ABI.

`Samples` plays the role of `Box<dyn Samples>` (src/samples.rs:11-28): `len`, `sample_rate`,
`read_at`, `read_exact_at`.  `from_file` / `gen` / `.shift` / `.lowpass` are the graph-building arms
of `Operation::exec` (src/lib.rs:89-121); `spark_fft`, `freq_levels`, `take_fft`, `do_write` are the
sinks (src/fft.rs, src/ffts.rs, src/lib.rs:178-213).  Same argument meaning, same error behaviour
(reference panics surface as `QdError` with a distinct code).  All arithmetic runs in
libquadrs_gpu.so on the GPU; nothing here computes samples.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from ._lib import QdError  # noqa: F401  (re-export)

CF32, CS8, CU8, CS16 = L.FMT_CF32, L.FMT_CS8, L.FMT_CU8, L.FMT_CS16
EXACT, FAST = L.PRECISION_EXACT, L.PRECISION_FAST

_EXT_FORMATS = {  # guess_from_extension, src/args.rs:392-402
    "cf32": CF32, "fc32": CF32, "cs8": CS8, "sc8": CS8, "c8": CS8, "cu8": CU8, "su8": CU8,
    "cs16": CS16, "sc16": CS16, "c16": CS16,
}


def format_from_extension(ext: str) -> Optional[int]:
    return _EXT_FORMATS.get(ext)


def _ptr(a) -> Optional[int]:
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return int(a)


class Samples:
    """A node of the lazy graph; immutable, cheap to extend with `.shift()` / `.lowpass()`.

    `device` is a CUDA device index, or a sequence of them: the chain is then sharded by sample range over those
    GPUs inside this one process (qd_chain_create_sharded), every sink call fanning its rows / chunks out and
    gathering the results into the one host buffer -- bit-identical to the one-device chain."""

    def __init__(self, source: L.Source, stages: Sequence[L.Stage] = (), device: int = 0, keep=(),
                 precision: int = EXACT, stream: Optional[int] = None):
        self._source = source
        self._stages = tuple(stages)
        self._device = device
        self._keep = keep
        self._precision = precision
        self._stream = stream
        self._h = None
        self._create()

    # ---- construction: Operation::From / Gen / Shift / LowPass (src/lib.rs:89-121) ----
    @staticmethod
    def from_file(path, fmt: int, sample_rate: int, device: int = 0) -> "Samples":
        src = L.Source()
        src.kind, src.format, src.sample_rate = L.SRC_FILE, fmt, sample_rate
        p = os.fsencode(str(path))
        src.path = p
        return Samples(src, device=device, keep=(p,))

    @staticmethod
    def from_bytes(data, fmt: int, sample_rate: int, device: int = 0, base_sample: int = 0,
                   total_samples: int = 0) -> "Samples":
        """Raw capture bytes in host memory (what SampleFile would pread)."""
        arr = data if isinstance(data, np.ndarray) else np.frombuffer(data, dtype=np.uint8)
        arr = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        src = L.Source()
        src.kind, src.format, src.sample_rate = L.SRC_HOST_MEM, fmt, sample_rate
        src.data, src.n_bytes = arr.ctypes.data, arr.size
        src.base_sample, src.total_samples = base_sample, total_samples
        return Samples(src, device=device, keep=(arr,))

    @staticmethod
    def from_host_ptr(ptr: int, n_bytes: int, fmt: int, sample_rate: int, device: int = 0, base_sample: int = 0,
                      total_samples: int = 0, keep=()) -> "Samples":
        src = L.Source()
        src.kind, src.format, src.sample_rate = L.SRC_HOST_MEM, fmt, sample_rate
        src.data, src.n_bytes = ptr, n_bytes
        src.base_sample, src.total_samples = base_sample, total_samples
        return Samples(src, device=device, keep=keep)

    @staticmethod
    def from_device(ptr: int, n_bytes: int, fmt: int, sample_rate: int, device: int = 0, base_sample: int = 0,
                    total_samples: int = 0, keep=()) -> "Samples":
        """Raw capture bytes already resident in HBM (a torch tensor's data_ptr(), for instance)."""
        src = L.Source()
        src.kind, src.format, src.sample_rate = L.SRC_DEVICE_MEM, fmt, sample_rate
        src.data, src.n_bytes = ptr, n_bytes
        src.base_sample, src.total_samples = base_sample, total_samples
        return Samples(src, device=device, keep=keep)

    @staticmethod
    def gen(cos: Sequence[int], sample_rate: int, seconds: float = 1.0, device: int = 0) -> "Samples":
        arr = (C.c_int64 * max(1, len(cos)))(*cos)
        src = L.Source()
        src.kind, src.sample_rate = L.SRC_GEN, sample_rate
        src.gen_seconds, src.gen_cos, src.gen_n_cos = seconds, arr, len(cos)
        return Samples(src, device=device, keep=(arr,))

    def _extend(self, st: L.Stage) -> "Samples":
        return Samples(self._source, self._stages + (st,), self._device, self._keep, self._precision, self._stream)

    def shift(self, frequency: int) -> "Samples":
        st = L.Stage()
        st.kind, st.frequency = L.STAGE_SHIFT, frequency
        return self._extend(st)

    def lowpass(self, frequency: int, decimate: int = 8, size: int = 40) -> "Samples":
        """size = number of taps (args.rs:161-166: `-power P` means size 2*P, default 40)."""
        st = L.Stage()
        st.kind, st.frequency, st.decimate, st.size = L.STAGE_LOWPASS, frequency, decimate, size
        return self._extend(st)

    def on_devices(self, devices: Sequence[int]) -> "Samples":
        """The same graph sharded over `devices` (one host process; see the class docstring)."""
        return Samples(self._source, self._stages, list(devices), self._keep, self._precision, None)

    def n_devices(self) -> int:
        v = C.c_size_t()
        L.check(L.lib().qd_chain_n_devices(self._h, C.byref(v)))
        return v.value

    def with_precision(self, precision: int) -> "Samples":
        return Samples(self._source, self._stages, self._device, self._keep, precision, self._stream)

    def with_stream(self, cuda_stream: Optional[int]) -> "Samples":
        return Samples(self._source, self._stages, self._device, self._keep, self._precision, cuda_stream)

    def _create(self):
        lib = L.lib()
        n = len(self._stages)
        arr = (L.Stage * max(1, n))(*self._stages)
        h = C.c_void_p()
        if isinstance(self._device, (list, tuple)):
            devs = (C.c_int * len(self._device))(*self._device)
            L.check(lib.qd_chain_create_sharded(C.byref(self._source), arr, n, devs, len(self._device), C.byref(h)))
        else:
            L.check(lib.qd_chain_create(C.byref(self._source), arr, n, self._device, C.byref(h)))
        self._h = h
        if self._precision != EXACT:
            L.check(lib.qd_chain_set_precision(h, self._precision))
        if self._stream is not None:
            L.check(lib.qd_chain_set_stream(h, self._stream))

    def close(self):
        if self._h is not None:
            L.lib().qd_chain_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int) -> "Samples":
        """Tuning knob on THIS node's chain handle ("use_fast", "segment_bytes", "scratch_budget")."""
        L.check(L.lib().qd_chain_set_option(self._h, key.encode(), value))
        return self

    def synchronize(self):
        L.check(L.lib().qd_chain_synchronize(self._h))

    def profile(self, enable: bool = True):
        L.check(L.lib().qd_chain_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """-> (bracketed regions, summed device ms, dominant kernel name); synchronises."""
        n, ms, name = C.c_uint64(), C.c_double(), C.create_string_buffer(256)
        L.check(L.lib().qd_chain_profile_read(self._h, C.byref(n), C.byref(ms), name, 256))
        return n.value, ms.value, name.value.decode()

    # ---- trait Samples (src/samples.rs:11-28) ----
    def len(self) -> int:
        v = C.c_uint64()
        L.check(L.lib().qd_chain_len(self._h, C.byref(v)))
        return v.value

    def sample_rate(self) -> int:
        v = C.c_uint64()
        L.check(L.lib().qd_chain_sample_rate(self._h, C.byref(v)))
        return v.value

    def taps(self, stage: int) -> np.ndarray:
        n = C.c_size_t()
        L.check(L.lib().qd_chain_taps(self._h, stage, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.float32)
        L.check(L.lib().qd_chain_taps(self._h, stage, out.ctypes.data, n.value, C.byref(n)))
        return out

    def read_at(self, off: int, n: int) -> np.ndarray:
        buf = np.zeros(n, dtype=np.complex64)
        got = C.c_size_t()
        L.check(L.lib().qd_chain_read_at(self._h, off, buf.ctypes.data, n, L.SPACE_HOST, C.byref(got)))
        return buf[: got.value]

    def read_exact_at(self, off: int, n: int) -> np.ndarray:
        buf = np.zeros(n, dtype=np.complex64)
        L.check(L.lib().qd_chain_read_exact_at(self._h, off, buf.ctypes.data, n, L.SPACE_HOST))
        return buf

    # ---- sinks ----
    def spark_rows(self, width: int = 128, stride: Optional[int] = None) -> int:
        stride = width if stride is None else stride
        v = C.c_uint64()
        L.check(L.lib().qd_sparkfft_rows(self._h, width, stride, C.byref(v)))
        return v.value

    def spark_fft(self, width: int = 128, stride: Optional[int] = None, rng: Optional[Tuple[float, float]] = None,
                  first_row: int = 0, max_rows: Optional[int] = None, want_mag: bool = False):
        """spark_fft (src/fft.rs:12-69) -> (idx[rows, width] u8, mag[rows, width] f32 | None)."""
        stride = width if stride is None else stride
        if max_rows is None:
            # at least one row: when len <= width the reference still attempts (and fails) its first read
            max_rows = max(0, self.spark_rows(width, stride) - first_row) or (1 if first_row == 0 and self.len() < width else 0)
        idx = np.zeros((max_rows, width), dtype=np.uint8)
        mag = np.zeros((max_rows, width), dtype=np.float32) if want_mag else None
        rows = C.c_uint64()
        lo, hi = rng if rng is not None else (0.0, 0.0)
        rc = L.lib().qd_sparkfft(self._h, width, stride, 1 if rng is not None else 0, lo, hi, first_row, max_rows,
                                 idx.ctypes.data, _ptr(mag), L.SPACE_HOST, C.byref(rows))
        r = rows.value
        if rc == L.E_GLYPH_RANGE:  # fft.rs:59 panics inside the first row that holds such a bin: nothing after it exists
            bad = np.nonzero((idx[:r] == 9).any(axis=1))[0]
            if bad.size:
                r = int(bad[0]) + 1
        try:
            L.check(rc)
        except QdError as e:  # the rows the reference had printed before it stopped
            e.partial = (idx[:r], (mag[:r] if want_mag else None))
            raise
        return idx[:r], (mag[:r] if want_mag else None)

    def spark_fft_device(self, width: int, stride: int, rng, first_row: int, n_rows: int, idx_ptr: int,
                         mag_ptr: Optional[int] = None) -> int:
        """Same, writing into device buffers (no host copy); returns rows produced."""
        rows = C.c_uint64()
        lo, hi = rng if rng is not None else (0.0, 0.0)
        L.check(L.lib().qd_sparkfft(self._h, width, stride, 1 if rng is not None else 0, lo, hi, first_row, n_rows,
                                    idx_ptr, mag_ptr, L.SPACE_DEVICE, C.byref(rows)))
        return rows.value

    def spark_fft_into(self, width: int, stride: int, rng, first_row: int, n_rows: int, idx_host_ptr: int,
                       mag_host_ptr: Optional[int] = None) -> int:
        rows = C.c_uint64()
        lo, hi = rng if rng is not None else (0.0, 0.0)
        L.check(L.lib().qd_sparkfft(self._h, width, stride, 1 if rng is not None else 0, lo, hi, first_row, n_rows,
                                    idx_host_ptr, mag_host_ptr, L.SPACE_HOST, C.byref(rows)))
        return rows.value

    def spark_fft_text(self, width: int = 128, stride: Optional[int] = None, rng=None) -> str:
        """Exact stdout of spark_fft: header (fft.rs:19) then one row per window (fft.rs:63)."""
        idx, _ = self.spark_fft(width, stride, rng)
        lines = [f"sparkfft sample_rate={self.sample_rate()}"]
        lines += [format_row(r) for r in idx]
        return "\n".join(lines) + "\n"

    def freq_levels(self, width: int = 128, stride: Optional[int] = None, levels: int = 2, first: int = 0,
                    max_n: Optional[int] = None):
        """freq_levels (src/fft.rs:77-101) -> (vals u8[], total)."""
        stride = width if stride is None else stride
        total = C.c_uint64()
        if max_n is None:
            L.check(L.lib().qd_freq_levels(self._h, width, stride, levels, 0, 0, None, L.SPACE_HOST, C.byref(total)))
            max_n = max(0, total.value - first)
        vals = np.zeros(max(1, max_n), dtype=np.uint8)
        L.check(L.lib().qd_freq_levels(self._h, width, stride, levels, first, max_n, vals.ctypes.data, L.SPACE_HOST,
                                       C.byref(total)))
        return vals[: min(max_n, max(0, total.value - first))], total.value

    def take_fft(self, width: int, output_len: int, slice_: Optional[Tuple[int, int]] = None,
                 blackman_harris: bool = False) -> np.ndarray:
        """take_fft (src/ffts.rs:18-85) -> magnitudes [output_len, width]."""
        out = np.zeros((output_len, width), dtype=np.float32)
        a, b = slice_ if slice_ is not None else (0, 0)
        L.check(L.lib().qd_take_fft(self._h, 1 if slice_ is not None else 0, a, b, width, 1 if blackman_harris else 0,
                                    output_len, out.ctypes.data, L.SPACE_HOST))
        return out

    def write_mem(self, chunk: int = 0x1000, first_chunk: int = 0, max_chunks: Optional[int] = None):
        """do_write's pull loop (src/lib.rs:199-210) into memory -> (samples, status).
        status is E_WRITE_SHORT where the reference panics at lib.rs:203 after delivering the data."""
        if max_chunks is None:
            max_chunks = (self.len() + chunk - 1) // chunk + 1
        cap = max_chunks * chunk
        out = np.zeros(cap, dtype=np.complex64)
        n = C.c_uint64()
        rc = L.lib().qd_write_cf32(self._h, chunk, first_chunk, max_chunks, out.ctypes.data, cap, L.SPACE_HOST,
                                   C.byref(n))
        L.check(rc, allow=(L.E_WRITE_SHORT,))
        return out[: n.value], rc

    def write_into(self, chunk: int, first_chunk: int, n_chunks: int, out_ptr: int, cap: int, space: int) -> Tuple[int, int]:
        n = C.c_uint64()
        rc = L.lib().qd_write_cf32(self._h, chunk, first_chunk, n_chunks, out_ptr, cap, space, C.byref(n))
        L.check(rc, allow=(L.E_WRITE_SHORT,))
        return n.value, rc

    def write_file(self, prefix: str, overwrite: bool = False) -> str:
        """do_write (src/lib.rs:178-213): '{prefix}.sr{rate}.cf32'."""
        name = C.create_string_buffer(4096)
        rc = L.lib().qd_write_file(self._h, os.fsencode(prefix), 1 if overwrite else 0, name, 4096)
        L.check(rc, allow=(L.E_WRITE_SHORT,))
        return os.fsdecode(name.value)


def format_row(idx: np.ndarray) -> str:
    idx = np.ascontiguousarray(idx, dtype=np.uint8)
    cap = 3 * idx.size + 16
    buf = C.create_string_buffer(cap)
    n = L.lib().qd_format_row(idx.ctypes.data, idx.size, buf, cap)
    return buf.raw[:n].decode("utf-8")


# ---- free functions with the reference's names ----
def spark_fft(samples: Samples, fft_width: int, stride: int, min: Optional[float] = None, max: Optional[float] = None):
    """fft::spark_fft (src/fft.rs:12-18).  A lone min or max takes the other's default (fft.rs:22-23)."""
    rng = None if (min is None and max is None) else (0.08 if min is None else min, 1.0 if max is None else max)
    return samples.spark_fft(fft_width, stride, rng)


def freq_levels(samples: Samples, fft_width: int, stride: int, levels: int):
    return samples.freq_levels(fft_width, stride, levels)[0]


def take_fft(samples: Samples, slice_, width: int, blackman_harris: bool, output_len: int) -> np.ndarray:
    return samples.take_fft(width, output_len, slice_, blackman_harris)


def do_write(samples: Samples, overwrite: bool, prefix: str) -> str:
    return samples.write_file(prefix, overwrite)
