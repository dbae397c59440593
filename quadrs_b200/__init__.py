"""quadrs_b200: B200-native (sm_100a) implementation of quadrs's streaming IQ DSP chain.

The product is libquadrs_gpu.so (CUDA kernels + C ABI, include/quadrs_gpu.h).  This package is the
Python host-side mirror of the reference's interface over that ABI; it computes nothing itself and
raises if the library is missing or no B200 is visible.
"""
from . import _lib
from ._lib import QdError, build
from .chain import (CF32, CS8, CU8, CS16, EXACT, FAST, Samples, do_write, format_from_extension, format_row,
                    freq_levels, spark_fft, take_fft)
from .shard import plan_shards, shard_plan
from .synth import make_synth, synth_fill_device, tone_step

# FAST arithmetic (block-anchored f64 phase + FMA FIR) is validated against the oracle at 1e-5 for cs8/cf32
# (tests/test_gpu_fast.py::test_fast_mode_*); bench.py uses it for those formats with a cf32 sink.
FAST_READY = True

__all__ = [
    "QdError", "build", "CF32", "CS8", "CU8", "CS16", "EXACT", "FAST", "Samples", "do_write",
    "format_from_extension", "format_row", "freq_levels", "spark_fft", "take_fft", "plan_shards", "shard_plan",
    "make_synth", "synth_fill_device", "tone_step", "_lib",
]
