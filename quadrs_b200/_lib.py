"""ctypes loader for libquadrs_gpu.so (the C ABI of include/quadrs_gpu.h).

The product path fails loudly when the CUDA extension is missing or no B200 is present: there is no
CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libquadrs_gpu.so"
CSRC_DIR = PKG_DIR / "csrc"

# qd_status
OK = 0
(E_INVALID_ARG, E_SHIFT_NYQUIST, E_ZERO_RATE, E_OFFSET_EOF, E_SHORT_INPUT, E_SHORT_READ, E_FFT_WIDTH, E_GLYPH_RANGE,
 E_LEVELS, E_SLICE, E_VISIBLE, E_GEN_ARGS, E_WRITE_SHORT, E_IO, E_CUDA, E_NOT_RESIDENT, E_UNIMPLEMENTED, E_EXISTS,
 E_NOMEM, E_ZERO_STRIDE) = range(1, 21)

FMT_CF32, FMT_CS8, FMT_CU8, FMT_CS16 = 0, 1, 2, 3
SRC_HOST_MEM, SRC_DEVICE_MEM, SRC_FILE, SRC_GEN = 0, 1, 2, 3
STAGE_SHIFT, STAGE_LOWPASS = 1, 2
SPACE_HOST, SPACE_DEVICE = 0, 1
PRECISION_EXACT, PRECISION_FAST = 0, 1

PAIR_BYTES = {FMT_CF32: 8, FMT_CS8: 2, FMT_CU8: 2, FMT_CS16: 4}

# every symbol include/quadrs_gpu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "qd_last_error", "qd_abi_version", "qd_device_count", "qd_kernel_launches", "qd_status_name",
    "qd_chain_create", "qd_chain_create_sharded", "qd_chain_n_devices", "qd_chain_destroy", "qd_chain_set_stream", "qd_chain_set_precision", "qd_chain_synchronize", "qd_chain_set_option",
    "qd_chain_profile", "qd_chain_profile_read", "qd_chain_len", "qd_chain_sample_rate", "qd_chain_taps", "qd_chain_read_at", "qd_chain_read_exact_at",
    "qd_sparkfft_rows", "qd_sparkfft", "qd_format_row", "qd_freq_levels", "qd_take_fft", "qd_write_cf32",
    "qd_write_file", "qd_shard_plan", "qd_synth_fill",
]


class QdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{status_name(code)} ({code}): {msg}")
        self.code = code
        self.msg = msg


class Source(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("format", C.c_int32), ("sample_rate", C.c_uint64),
        ("data", C.c_void_p), ("n_bytes", C.c_uint64), ("path", C.c_char_p),
        ("base_sample", C.c_uint64), ("total_samples", C.c_uint64),
        ("gen_seconds", C.c_double), ("gen_cos", C.POINTER(C.c_int64)), ("gen_n_cos", C.c_uint64),
    ]


class Stage(C.Structure):
    _fields_ = [("kind", C.c_int32), ("reserved", C.c_int32), ("frequency", C.c_int64), ("decimate", C.c_uint64),
                ("size", C.c_uint64)]


class Shard(C.Structure):
    _fields_ = [("first_unit", C.c_uint64), ("n_units", C.c_uint64), ("first_sample", C.c_uint64),
                ("n_samples", C.c_uint64)]


class Synth(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_tones", C.c_uint32), ("tone_step", C.c_uint32 * 8),
                ("tone_amp", C.c_int32 * 8), ("key_period", C.c_uint32 * 8), ("noise_amp", C.c_int32)]


def build(force: bool = False) -> Path:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = list(CSRC_DIR.glob("*.cu")) + list(CSRC_DIR.glob("*.cuh")) + list(CSRC_DIR.glob("*.h")) + \
        list(CSRC_DIR.glob("*.cpp")) + [PKG_DIR.parent / "include" / "quadrs_gpu.h", CSRC_DIR / "Makefile"]
    newest = max(p.stat().st_mtime for p in srcs)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest:
        r = subprocess.run(["make", "-C", str(CSRC_DIR), "-j8"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libquadrs_gpu.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    return LIB_PATH


_lib = None


def lib():
    """Loads the library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(quadrs_b200 has no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, u64, sz, i32, f32 = C.c_void_p, C.c_uint64, C.c_size_t, C.c_int, C.c_float
    L.qd_last_error.restype = C.c_char_p
    L.qd_status_name.restype = C.c_char_p
    L.qd_status_name.argtypes = [i32]
    L.qd_kernel_launches.restype = u64
    L.qd_device_count.argtypes = [C.POINTER(i32)]
    L.qd_chain_create.argtypes = [C.POINTER(Source), C.POINTER(Stage), sz, i32, C.POINTER(vp)]
    L.qd_chain_create_sharded.argtypes = [C.POINTER(Source), C.POINTER(Stage), sz, C.POINTER(i32), sz, C.POINTER(vp)]
    L.qd_chain_n_devices.argtypes = [vp, C.POINTER(sz)]
    L.qd_chain_destroy.argtypes = [vp]
    L.qd_chain_destroy.restype = None
    L.qd_chain_set_stream.argtypes = [vp, vp]
    L.qd_chain_set_precision.argtypes = [vp, i32]
    L.qd_chain_synchronize.argtypes = [vp]
    L.qd_chain_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.qd_chain_profile.argtypes = [vp, i32]
    L.qd_chain_profile_read.argtypes = [vp, C.POINTER(u64), C.POINTER(C.c_double), C.c_char_p, sz]
    L.qd_chain_len.argtypes = [vp, C.POINTER(u64)]
    L.qd_chain_sample_rate.argtypes = [vp, C.POINTER(u64)]
    L.qd_chain_taps.argtypes = [vp, sz, vp, sz, C.POINTER(sz)]
    L.qd_chain_read_at.argtypes = [vp, u64, vp, sz, i32, C.POINTER(sz)]
    L.qd_chain_read_exact_at.argtypes = [vp, u64, vp, sz, i32]
    L.qd_sparkfft_rows.argtypes = [vp, sz, u64, C.POINTER(u64)]
    L.qd_sparkfft.argtypes = [vp, sz, u64, i32, f32, f32, u64, u64, vp, vp, i32, C.POINTER(u64)]
    L.qd_format_row.argtypes = [vp, sz, vp, sz]
    L.qd_format_row.restype = sz
    L.qd_freq_levels.argtypes = [vp, sz, u64, sz, u64, u64, vp, i32, C.POINTER(u64)]
    L.qd_take_fft.argtypes = [vp, i32, u64, u64, sz, i32, sz, vp, i32]
    L.qd_write_cf32.argtypes = [vp, sz, u64, u64, vp, u64, i32, C.POINTER(u64)]
    L.qd_write_file.argtypes = [vp, C.c_char_p, i32, C.c_char_p, sz]
    L.qd_shard_plan.argtypes = [C.POINTER(Source), C.POINTER(Stage), sz, i32, u64, u64, C.c_uint32, C.c_uint32,
                                C.POINTER(Shard)]
    L.qd_synth_fill.argtypes = [C.POINTER(Synth), i32, u64, u64, vp, i32, vp]
    _lib = L
    return L


def status_name(code: int) -> str:
    try:
        return lib().qd_status_name(code).decode()
    except Exception:
        return f"status {code}"


def check(rc: int, allow=()):
    if rc != OK and rc not in allow:
        raise QdError(rc, lib().qd_last_error().decode("utf-8", "replace"))
    return rc
