"""Sharding by sample range: each rank owns a contiguous range of sink units and holds
only the raw samples those units touch (filter-tap and FFT-window halo included).  Phase and
end-of-file arithmetic use absolute sample indices, so no collective and no hand-off is needed."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Sequence, Tuple

from . import _lib as L

SINK_WRITE, SINK_SPARKFFT, SINK_FREQ_LEVELS = 0, 1, 2


@dataclass(frozen=True)
class ShardPlan:
    first_unit: int
    n_units: int
    first_sample: int
    n_samples: int


def _stages(stages: Sequence[Tuple]) -> Tuple:
    arr = (L.Stage * max(1, len(stages)))()
    for i, st in enumerate(stages):
        if st[0] == "shift":
            arr[i].kind, arr[i].frequency = L.STAGE_SHIFT, st[1]
        elif st[0] == "lowpass":
            arr[i].kind, arr[i].frequency, arr[i].decimate, arr[i].size = L.STAGE_LOWPASS, st[1], st[2], st[3]
        else:
            raise ValueError(f"unknown stage {st!r}")
    return arr, len(stages)


def shard_plan(fmt: int, sample_rate: int, total_samples: int, stages: Sequence[Tuple], sink_kind: int,
               unit_len: int, stride: int, n_shards: int, shard: int) -> ShardPlan:
    """stages: ('shift', f) / ('lowpass', f, decimate, size).  Pure host arithmetic (no GPU needed)."""
    src = L.Source()
    src.kind, src.format, src.sample_rate = L.SRC_HOST_MEM, fmt, sample_rate
    src.n_bytes = total_samples * L.PAIR_BYTES[fmt]
    arr, n = _stages(stages)
    out = L.Shard()
    L.check(L.lib().qd_shard_plan(C.byref(src), arr, n, sink_kind, unit_len, stride, n_shards, shard, C.byref(out)))
    return ShardPlan(out.first_unit, out.n_units, out.first_sample, out.n_samples)


def plan_shards(fmt, sample_rate, total_samples, stages, sink_kind, unit_len, stride, n_shards) -> List[ShardPlan]:
    return [shard_plan(fmt, sample_rate, total_samples, stages, sink_kind, unit_len, stride, n_shards, r)
            for r in range(n_shards)]
