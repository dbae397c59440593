"""Synthetic IQ generator front-end (device-side fill; integer-only, keyed by absolute sample index)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

from . import _lib as L


def tone_step(freq_hz: float, sample_rate: float) -> int:
    return int(round(freq_hz / sample_rate * 2**32)) & 0xFFFFFFFF


def make_synth(seed: int, tones: Sequence[Tuple[int, int, int]], noise_amp: int = 0) -> L.Synth:
    """tones: (step_u32, amplitude, key_period)."""
    p = L.Synth()
    p.seed = seed
    p.n_tones = len(tones)
    for i, (step, amp, key) in enumerate(tones):
        p.tone_step[i] = step & 0xFFFFFFFF
        p.tone_amp[i] = amp
        p.key_period[i] = key
    p.noise_amp = noise_amp
    return p


def synth_fill_device(p: L.Synth, fmt: int, first_sample: int, n_samples: int, device_ptr: int, device: int = 0,
                      stream: Optional[int] = None) -> None:
    L.check(L.lib().qd_synth_fill(C.byref(p), fmt, first_sample, n_samples, device_ptr, device, stream))
