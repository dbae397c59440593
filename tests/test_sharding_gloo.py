"""world_size-2 (and 3) runs of the N>1 path on CPU with the gloo backend.

The multi-GPU design has no data-path collective: each rank plans its own
contiguous range of sink units with qd_shard_plan (pure host arithmetic in libquadrs_gpu.so), holds
only the raw samples those units touch, and evaluates them at their ABSOLUTE offsets.  Without a GPU
the per-rank evaluation here is done by the CPU oracle standing in for the kernels; what is under test
is the host logic both share: unit ranges, halos, absolute indexing, and the gather of outputs on
rank 0, which must reproduce the unsharded result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from helpers import kept_only, oracle_chain, synth_raw


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


CONFIGS = {
    "cfg2_write": dict(fmt=O.CS8, rate=20_000_000, n=300_000, stages=[("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)],
                       sink=0, unit=0x1000, stride=0x1000, rng=None),
    "cfg4_spark": dict(fmt=O.CS16, rate=100_000_000, n=200_000, stages=[("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)],
                       sink=1, unit=128, stride=128, rng=(0.5, 50.0)),
    "cfg1_overlap": dict(fmt=O.CF32, rate=21_000_000, n=60_000, stages=[("shift", 280_000), ("lowpass", 200_000, 32, 400)],
                         sink=1, unit=64, stride=16, rng=None),
    "cfg5_two_stage": dict(fmt=O.CF32, rate=400_000_000, n=120_000,
                           stages=[("lowpass", 20_000_000, 8, 40), ("lowpass", 500_000, 32, 40)],
                           sink=1, unit=4, stride=2, rng=(0.001, 0.01)),
}


def _worker(rank, world, port, name, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import quadrs_b200 as Q

        cfg = CONFIGS[name]
        fmt, rate, n, stages = cfg["fmt"], cfg["rate"], cfg["n"], cfg["stages"]
        plan = Q.shard_plan(fmt, rate, n, stages, cfg["sink"], cfg["unit"], cfg["stride"], world, rank)
        pb = O.FORMAT_BYTES[fmt]
        # every rank regenerates ONLY its own byte range (index-keyed generator), as each GPU does in place
        raw, _ = synth_raw(fmt, plan.n_samples, first=plan.first_sample, rate=rate) if plan.n_samples else (np.zeros(0, np.uint8), None)
        assert raw.size == plan.n_samples * pb
        with kept_only():
            if plan.n_units == 0:
                mine = np.zeros((0, cfg["unit"]), dtype=np.uint8) if cfg["sink"] else np.zeros(0, dtype=np.complex64)
            else:
                chain = oracle_chain(raw, fmt, rate, stages, plan.first_sample, n)
                if cfg["sink"] == 0:
                    mine, _ = chain.write_mem(chunk=cfg["unit"], first_chunk=plan.first_unit, max_chunks=plan.n_units)
                else:
                    mine, _ = chain.spark_fft(cfg["unit"], cfg["stride"], cfg["rng"], first_row=plan.first_unit,
                                              max_rows=plan.n_units, want_mag=False)
        # gather to the host writer on rank 0 (outputs only; inputs never move)
        payload = torch.from_numpy(np.ascontiguousarray(mine).view(np.uint8).reshape(-1).copy())
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([payload.numel()], dtype=torch.int64))
        cap = int(max(s.item() for s in sizes))
        padded = torch.zeros(cap, dtype=torch.uint8)
        padded[: payload.numel()] = payload
        gathered = [torch.zeros(cap, dtype=torch.uint8) for _ in range(world)] if rank == 0 else None
        dist.gather(padded, gathered, dst=0)
        units = torch.tensor([plan.first_unit, plan.n_units], dtype=torch.int64)
        all_units = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(all_units, units)
        if rank == 0:
            whole = np.concatenate([g[: int(s.item())].numpy() for g, s in zip(gathered, sizes)])
            np.save(out_path, whole)
            np.save(out_path + ".units.npy", torch.stack(all_units).numpy())
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", sorted(CONFIGS))
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_run_reproduces_unsharded(tmp_path, name, world):
    import quadrs_b200

    quadrs_b200.build()
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(world, _free_port(), name, out), nprocs=world, join=True)
    cfg = CONFIGS[name]
    raw, _ = synth_raw(cfg["fmt"], cfg["n"], rate=cfg["rate"])
    with kept_only():
        chain = oracle_chain(raw, cfg["fmt"], cfg["rate"], cfg["stages"])
        if cfg["sink"] == 0:
            want, _ = chain.write_mem(chunk=cfg["unit"])
        else:
            want, _ = chain.spark_fft(cfg["unit"], cfg["stride"], cfg["rng"], want_mag=False)
    got = np.load(out)
    assert np.array_equal(got, np.ascontiguousarray(want).view(np.uint8).reshape(-1))
    units = np.load(out + ".units.npy")
    assert units[0, 0] == 0 and all(units[i, 0] + units[i, 1] == units[i + 1, 0] for i in range(world - 1))
