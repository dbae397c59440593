#!/usr/bin/env python
"""bench.py -- throughput of the quadrs IQ DSP hot path on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload all|cfg1|cfg2|cfg2s|cfg3|cfg4|cfg5] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input.  The headline workload is
BASELINE.json configs[3], the metric's literal chain on its largest single-GPU configuration: synthetic cs16 at
100 MS/s, 2^33 samples, shift | lowpass -power 400 -decimate 16 | sparkfft -width 128, sharded over the N GPUs by
contiguous row ranges with halo (strong scaling: the capture stays 2^33 samples; absolute sample indices drive
phase and truncation; no data-path collective).  With `--workload all` (the default) the same JSON line carries a
`configs` map with the other BASELINE configurations measured the same way (config 1's chain at capture scale,
config 2, config 2's input through sparkfft, config 3 at 2^30, config 5 as a resident 2^32-sample buffer
processed at 4 absolute offsets), each a fixed per-GPU size (weak scaling).

Rank 0 prints ONE JSON line.  `value` = input Msamples/s with the input resident in HBM, timed with CUDA events
on the stream the kernels run on (max over ranks).  `e2e` = the same metric through the public C ABI with HOST
buffers: pinned host input -> H2D -> kernels -> D2H of the result inside the timed region; with N > 1 it is ONE
process (rank 0) driving all N devices through qd_chain_create_sharded, the way the reference's single-process
caller would.  `roofline` = algorithmic bytes / device time against the measured HBM peak of MEASURED_PEAKS.json,
with the FP32 co-bound of the exact-order FIR / FFT stated.  `cpu_baseline` = the CPU oracle (a C port of the
reference algorithm; Rust cannot be built here) timed on a bounded sample of the same workload; its output
checksum over those units is compared with the GPU's (`parity_checked`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

CS8, CU8, CS16, CF32 = 1, 2, 3, 0
PAIR = {CF32: 8, CS8: 2, CU8: 2, CS16: 4}
FMT_NAME = {CF32: "cf32", CS8: "cs8", CU8: "cu8", CS16: "cs16"}

# name -> one workload.  sink: ("write", chunk) or ("sparkfft", W, S, (lo, hi)).  samples: per GPU (weak) or of the
# whole capture (strong).  passes: the resident buffer is processed at this many absolute offsets per step.
WORKLOADS = {
    # BASELINE.json configs[3] -- the headline
    "cfg4": dict(
        title="synthetic cs16 100 MS/s, 2^33 samples: shift 7000000 | lowpass -power 400 -decimate 16 2000000 | sparkfft -width 128 -range 0.5:50, sharded over the GPUs",
        fmt=CS16, rate=100_000_000, samples=2**33, scaling="strong",
        stages=[("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)], sink=("sparkfft", 128, 128, (0.5, 50.0)),
        tones=[(7.3e6, 9000, 0), (6.2e6, 6000, 50_000), (-20e6, 4000, 0)], noise=1200, seed=0x5EED0004,
        cpu_units=1536, ref_units_per_thread=96),
    # BASELINE.json configs[0] (the reference's own example chain) at capture scale: overlapping windows
    "cfg1": dict(
        title="synthetic cf32 21 MS/s, 2^27 samples/GPU: shift 280000 | lowpass -power 200 -decimate 32 200000 | sparkfft -width 64 -stride 16 -range 0.01:3",
        fmt=CF32, rate=21_000_000, samples=2**27, scaling="weak",
        stages=[("shift", 280_000), ("lowpass", 200_000, 32, 400)], sink=("sparkfft", 64, 16, (0.01, 3.0)),
        tones=[(-250e3, 6000, 2000), (-310e3, 6000, 2000), (3e6, 9000, 0)], noise=300, seed=0x5EED0001,
        cpu_units=1024, ref_units_per_thread=128),
    # BASELINE.json configs[1]
    "cfg2": dict(
        title="synthetic cs8 20 MS/s, 2^30 samples/GPU: decode + shift 1500000 + lowpass -power 20 -decimate 8 1000000 | write",
        fmt=CS8, rate=20_000_000, samples=2**30, scaling="weak",
        stages=[("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], sink=("write", 0x1000),
        tones=[(1.6e6, 45, 0), (-4.1e6, 30, 0), (0.3e6, 20, 3000)], noise=6, seed=0x5EED0002,
        cpu_units=2048, ref_units_per_thread=512),
    # configs[1]'s input through the metric's literal chain: shift + lowpass + sparkfft with overlapping windows
    "cfg2s": dict(
        title="synthetic cs8 20 MS/s, 2^30 samples/GPU: shift 1500000 | lowpass -power 20 -decimate 8 1000000 | sparkfft -width 64 -stride 16 -range 0.01:3",
        fmt=CS8, rate=20_000_000, samples=2**30, scaling="weak",
        stages=[("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], sink=("sparkfft", 64, 16, (0.01, 3.0)),
        tones=[(1.6e6, 45, 0), (-4.1e6, 30, 0), (0.3e6, 20, 3000)], noise=6, seed=0x5EED0002,
        cpu_units=8192, ref_units_per_thread=512),
    # BASELINE.json configs[2]
    "cfg3": dict(
        title="synthetic cu8 2.4 MS/s multi-tone, 2^30 samples/GPU: sparkfft -width 4096 -stride 1024 -range 2:500",
        fmt=CU8, rate=2_400_000, samples=2**30, scaling="weak",
        stages=[], sink=("sparkfft", 4096, 1024, (2.0, 500.0)),
        tones=[(-800e3, 40, 0), (-123_456, 30, 0), (300e3, 25, 0), (1_000_001, 20, 0)], noise=4, seed=0x5EED0003,
        cpu_units=2048, ref_units_per_thread=256),
    # BASELINE.json configs[4]: 2^34 samples as a resident 2^32-sample buffer processed at 4 absolute offsets
    "cfg5": dict(
        title="synthetic cf32 400 MS/s, 2^34 samples/GPU (resident 2^32-sample buffer at 4 absolute offsets): lowpass -decimate 8 20000000 | lowpass -decimate 32 500000 | sparkfft -width 4 -stride 2 -range 0.001:0.01",
        fmt=CF32, rate=400_000_000, samples=2**32, passes=4, scaling="weak",
        stages=[("lowpass", 20_000_000, 8, 40), ("lowpass", 500_000, 32, 40)], sink=("sparkfft", 4, 2, (0.001, 0.01)),
        tones=[(0.1e6, 160, 1_000_000), (90e6, 3000, 0)], noise=40, seed=0x5EED0005,
        cpu_units=4096, ref_units_per_thread=256),
    # ---- kernel experiments (not part of the default line): the long filter alone, and the mixer alone ----
    "x_fir16": dict(
        title="experiment: cf32, lowpass -power 400 -decimate 16 | sparkfft -width 128 (config 4's filter without its decode + shift)",
        fmt=CF32, rate=100_000_000, samples=2**28, scaling="weak",
        stages=[("lowpass", 2_000_000, 16, 800)], sink=("sparkfft", 128, 128, (0.5, 50.0)),
        tones=[(7.3e6, 9000, 0), (6.2e6, 6000, 50_000)], noise=1200, seed=0x5EED0014, cpu_units=256, ref_units_per_thread=16),
    "x_mix16": dict(
        title="experiment: cs16, shift | lowpass -power 8 -decimate 16 | sparkfft -width 128 (config 4's decode + shift with a short filter)",
        fmt=CS16, rate=100_000_000, samples=2**28, scaling="weak",
        stages=[("shift", 7_000_000), ("lowpass", 2_000_000, 16, 16)], sink=("sparkfft", 128, 128, (0.5, 50.0)),
        tones=[(7.3e6, 9000, 0), (6.2e6, 6000, 50_000)], noise=1200, seed=0x5EED0024, cpu_units=256, ref_units_per_thread=16),
}
HEADLINE = "cfg4"
MAP_ORDER = ["cfg1", "cfg2", "cfg2s", "cfg3", "cfg5"]

METRIC = "input Msamples/s through shift+lowpass+sparkfft"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="override the workload's sample count")
    ap.add_argument("--precision", default="auto", choices=["auto", "exact", "fast"])
    ap.add_argument("--segment-mb", type=int, default=0, help="host-path segment size in MiB (default: the library's)")
    ap.add_argument("--opt", action="append", default=[], help="chain option key=value (experiments)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def sink_kind(w):
    return 0 if w["sink"][0] == "write" else 1


def unit_geometry(w):
    if w["sink"][0] == "write":
        return w["sink"][1], w["sink"][1]
    return w["sink"][1], w["sink"][2]


def chain_geometry(w):
    """-> (raw samples per top-level sample, raw samples one unit's read touches)"""
    unit_len, _ = unit_geometry(w)
    mult, need = 1, unit_len
    for st in reversed(w["stages"]):
        if st[0] == "lowpass":
            need = need * st[2] + st[3]
            mult *= st[2]
    return mult, need


def workload_config(name, w, samples_total_note):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": w["title"], "name": name, "format": FMT_NAME[w["fmt"]], "sample_rate": w["rate"],
            "capture_samples": samples_total_note, "scaling": w["scaling"],
            "l2": "inputs far larger than L2 (126 MB); no flush needed"}


def make_oracle_synth(O, w):
    return O.make_synth(w["seed"], [(O.tone_step(f, w["rate"]), a, k) for f, a, k in w["tones"]], w["noise"])


def checksum_np(np, arr, w):
    """The oracle's qo_timed_run checksum of the same units, from the GPU's output bytes."""
    if w["sink"][0] == "write":
        v = arr.view(np.uint32).reshape(-1, w["sink"][1], 2).astype(np.uint64)
        wt = np.arange(1, w["sink"][1] + 1, dtype=np.uint64)
        with np.errstate(over="ignore"):
            return int(((v[:, :, 0] + (v[:, :, 1] << np.uint64(1))) * wt[None, :]).sum(dtype=np.uint64))
    W = w["sink"][1]
    v = arr.reshape(-1, W).astype(np.uint64)
    return int((v * np.arange(1, W + 1, dtype=np.uint64)[None, :]).sum(dtype=np.uint64))


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_cpus(gpu_index: int):
    """Run this rank on the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the e2e
    path are allocated on that NUMA node (what a deployment does with numactl).  Returns the CPU count or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[gpu_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, ~1 ms period, from a thread:
    ctypes releases the GIL while the library call runs).  Falls back to nvidia-smi -lms when NVML is missing."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []  # (sm_mhz, reasons_mask)
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        self.mode = None
        self.proc = None
        self.lines = []

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.mode = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.mode = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((float(sm), int(mask)))
            except Exception:
                break
            time.sleep(0.001)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        if self.mode == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=1)
            sm = sorted(x[0] for x in self.samples)
            mask = 0
            for _, m in self.samples:
                mask |= m
            reasons = sorted(name for bit, name in self.REASONS.items() if mask & bit)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle (port of the reference algorithm) on all host threads
# ------------------------------------------------------------------------------------------------
def run_reference(args, name, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    import oracle_lib as O

    threads = os.cpu_count() or 1
    unit_len, stride = unit_geometry(w)
    mult, need = chain_geometry(w)
    units = min(threads * w["ref_units_per_thread"], 16384)  # a bounded sample of the workload per step
    n_in = (units - 1) * stride * mult + need + 64
    raw = O.synth_fill(make_oracle_synth(O, w), w["fmt"], 0, n_in)
    sink = w["sink"]
    kw = dict(width=sink[1], stride=sink[2], rng=sink[3]) if sink[0] == "sparkfft" else {}
    samples_per_step = units * stride * mult

    def step():
        secs, _ = O.timed_run(raw, w["fmt"], w["rate"], w["stages"], "write" if sink[0] == "write" else "sparkfft",
                              0, units, threads, **kw)
        return secs

    for _ in range(args.warmup):
        step()
    total = sum(step() for _ in range(args.steps))
    ms = 1e3 * total / max(1, args.steps)
    value = samples_per_step / (ms * 1e-3) / 1e6
    sample = (f"{units} sink units = {samples_per_step} input samples per step (the first units of the same synthetic "
              f"capture), {threads} threads over disjoint unit ranges")
    total_samples = (args.samples or w["samples"]) * w.get("passes", 1) * (1 if w["scaling"] == "strong" else args.gpus)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, w, total_samples),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "CPU oracle: C port of the reference algorithm (lazy per-window pull, full-rate "
                                 "complex_convolve, per-sample f64 sin/cos); the Rust reference cannot be built here"},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class Ctx:
    pass


def fp32_cobound(w, precision_fast, sm_count, sm_mhz):
    """FP32 co-bound: packed f32x2 instructions per input sample that the reference's arithmetic
    needs at the very least -- FIR: one FFMA2 per complex MAC in FAST mode, FMUL2 + FFMA2 in the reference's exact
    mul-then-add order; FFT (our radix-4 DIT with individually rounded operations): per point and radix-4 level
    3 twiddle products of 3 packed instructions and 8 packed adds for 4 points, i.e. 4.25, plus ~4 for the
    magnitude/threshold epilogue -- against the FMA pipe's 64 packed lanes per clock per SM."""
    import math

    macs, rate_div = 0.0, 1
    for st in w["stages"]:
        if st[0] == "lowpass":
            rate_div *= st[2]
            macs += st[3] / rate_div
    instr = macs * (1 if precision_fast else 2)
    fft = 0.0
    if w["sink"][0] == "sparkfft":
        W, S = w["sink"][1], w["sink"][2]
        per_point = 4.25 * math.log(W, 4) + 4.0
        fft = per_point * W / (S * rate_div)
    total = instr + fft
    if total <= 0:
        return None
    peak_instr = sm_count * 64 * sm_mhz * 1e6
    return {"fir_complex_macs_per_sample": macs, "fir_packed_instr_per_sample": instr,
            "fft_packed_instr_per_sample": fft, "ceiling_msamples_per_s": peak_instr / total / 1e6}


def run_workload(cx, name, w, headline):
    """Runs one workload on this rank; rank 0 gets the result dict (others get None)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    Q, args = cx.Q, cx.args
    rank, world, local, dev, stream = cx.rank, cx.world, cx.local, cx.dev, cx.stream
    lib = Q._lib.lib()
    fmt, rate, pb = w["fmt"], w["rate"], PAIR[w["fmt"]]
    passes = w.get("passes", 1)
    samples = args.samples or w["samples"]
    strong = w["scaling"] == "strong"
    n_shards = world * passes
    total = samples * passes * (1 if strong else world)  # the logical capture
    unit_len, stride = unit_geometry(w)
    sk = sink_kind(w)
    mult, need = chain_geometry(w)
    out_per_unit = unit_len * 8 if sk == 0 else unit_len
    plans = [Q.shard_plan(fmt, rate, total, w["stages"], sk, unit_len, stride, n_shards, rank * passes + p)
             for p in range(passes)]
    n_in = max(p.n_samples for p in plans)
    n_units = sum(p.n_units for p in plans)

    # ---- synthetic input, generated in place on this GPU at the absolute sample range of its first pass ----
    d_in = torch.empty(n_in * pb + 64, dtype=torch.uint8, device=dev)
    synth = Q.make_synth(w["seed"], [(Q.tone_step(f, rate), a, k) for f, a, k in w["tones"]], w["noise"])
    Q.synth_fill_device(synth, fmt, plans[0].first_sample, n_in, d_in.data_ptr(), local, stream.cuda_stream)
    torch.cuda.synchronize()

    if args.precision == "exact":
        precision = Q.EXACT
    elif args.precision == "fast":
        precision = Q.FAST
    else:  # FAST only where it meets the 1e-5 bar: cs8 / cf32 with a cf32 sink (tests/test_gpu_fast.py)
        precision = Q.FAST if (fmt in (CS8, CF32) and sk == 0) else Q.EXACT
    if precision == Q.FAST and fmt in (CU8, CS16):
        precision = Q.EXACT

    def build_chain(src, prec, on_stream=True):
        s = src
        for st in w["stages"]:
            s = s.shift(st[1]) if st[0] == "shift" else s.lowpass(st[1], st[2], st[3])
        s = s.with_precision(prec)
        if on_stream:
            s = s.with_stream(stream.cuda_stream)
        if args.segment_mb:
            s.set_option("segment_bytes", args.segment_mb << 20)
        for kv in args.opt:
            k, v = kv.split("=")
            s.set_option(k, int(v))
        return s

    def device_chains(prec):
        return [build_chain(Q.Samples.from_device(d_in.data_ptr(), p.n_samples * pb, fmt, rate, local,
                                                  base_sample=p.first_sample, total_samples=total, keep=(d_in,)), prec)
                for p in plans]

    d_out = torch.empty(n_units * out_per_unit + 64, dtype=torch.uint8, device=dev)
    out_off = [0]
    for p in plans:
        out_off.append(out_off[-1] + p.n_units * out_per_unit)

    def run_one(chain, p, out_ptr, space):
        if sk == 0:
            n, _ = chain.write_into(unit_len, p.first_unit, p.n_units, out_ptr, p.n_units * unit_len, space)
            return n
        if space == Q._lib.SPACE_DEVICE:
            return chain.spark_fft_device(unit_len, stride, w["sink"][3], p.first_unit, p.n_units, out_ptr)
        return chain.spark_fft_into(unit_len, stride, w["sink"][3], p.first_unit, p.n_units, out_ptr)

    def run_device(chains):
        got = 0
        for i, (c, p) in enumerate(zip(chains, plans)):
            got += run_one(c, p, d_out.data_ptr() + out_off[i], Q._lib.SPACE_DEVICE)
        return got

    samples_per_step = n_units * stride * mult  # input samples consumed by this rank's units (halo excluded)
    steps = args.steps if headline else max(3, min(args.steps, 5))
    warm = max(args.warmup, 3)

    def timed(chains, k):
        for _ in range(warm):
            produced = run_device(chains)
        cx.barrier()
        for c in chains:
            c.profile(True)
        l0 = lib.qd_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cx.barrier()
        e0.record(stream)
        for _ in range(k):
            run_device(chains)
        e1.record(stream)
        cx.barrier()
        ms = e0.elapsed_time(e1) / k
        launches = lib.qd_kernel_launches() - l0
        kern_ms, kern_name = 0.0, ""
        for c in chains:  # the passes run back to back on one stream: their dominant regions add up
            _, t, nm = c.profile_read()
            kern_ms += t
            kern_name = nm or kern_name
            c.profile(False)
        return ms, launches, kern_ms / k, kern_name, produced

    chains = device_chains(precision)
    sampler = ClockSampler(local) if (rank == 0 and headline) else None
    if sampler:
        sampler.start()
    ms_dev, launches, kern_ms, kern_name, produced = timed(chains, steps)
    clocks = sampler.stop() if sampler else None

    exact_ms = None
    if precision == Q.FAST:  # the bit-exact arithmetic mode, timed the same way
        ex = device_chains(Q.EXACT)
        exact_ms = timed(ex, min(steps, 5))[0]
        del ex
        run_device(chains)  # leave the FAST result in d_out for the comparisons below
        torch.cuda.synchronize()

    # ---- parity: the CPU oracle on the first units of rank 0's own input, checksum against the GPU's output ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle_lib as O

        units = min(w["cpu_units"], plans[0].n_units)
        n = min((units - 1) * stride * mult + need + 64, plans[0].n_samples)
        raw = d_in[: n * pb].cpu().numpy()
        kw = dict(width=w["sink"][1], stride=w["sink"][2], rng=w["sink"][3]) if sk == 1 else {}
        secs, want_sum = O.timed_run(raw, fmt, rate, w["stages"], "write" if sk == 0 else "sparkfft", 0, units, 1, **kw)
        exact_out = d_out
        if precision == Q.FAST:  # bit-exactness is EXACT's property; FAST is within 1e-5 of it (tests/test_gpu_fast.py)
            exact_out = torch.empty_like(d_out)
            ex = device_chains(Q.EXACT)
            run_one(ex[0], plans[0], exact_out.data_ptr(), Q._lib.SPACE_DEVICE)
            ex[0].synchronize()
        got_sum = checksum_np(np, exact_out[: units * out_per_unit].cpu().numpy(), w)
        cpu = {"value": units * stride * mult / secs / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
               "sample": f"first {units} sink units ({units * stride * mult} input samples) of the same input, {secs:.1f} s, "
                         "1 thread (the reference hot path is single-threaded)",
               "parity_checked": bool(got_sum == want_sum),
               "parity": f"position-weighted checksum of the EXACT-mode GPU output over those {units} units "
                         f"{'==' if got_sum == want_sum else '!='} the oracle's"}
        del exact_out

    # ---- end-to-end through host buffers: ONE process (rank 0) drives every device ----
    e2e = None
    if not args.no_e2e and (headline or world == 1):
        e2e = run_e2e(cx, name, w, plans, total, precision, build_chain, d_in, d_out, out_per_unit, n_units)

    # ---- max / sum over ranks ----
    if world > 1:
        t = torch.tensor([ms_dev, float(samples_per_step), float(launches), exact_ms or 0.0, kern_ms],
                         dtype=torch.float64, device=dev)
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_dev, exact_ms, kern_ms = tmax[0].item(), (tmax[3].item() or None), tmax[4].item()
        total_samples_step, launches_all = tsum[1].item(), int(tsum[2].item())
    else:
        total_samples_step, launches_all = float(samples_per_step), int(launches)

    del chains, d_in, d_out
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    peak, peak_src = cx.peak
    # algorithmic bytes per GPU and step: input read once + final output written once (cf32 for write, u8 per bin for sparkfft)
    alg_bytes = sum(p.n_samples for p in plans) * pb + n_units * out_per_unit
    prec_name = "fast" if precision == Q.FAST else "exact"
    traffic, traffic_src = None, None
    ent = cx.traffic.get(f"{name}:{prec_name}")
    if ent:  # DRAM bytes of the step's kernels, from the ncu capture under profiles/, scaled to this run's samples
        traffic, traffic_src = ent["dram_bytes_per_sample"] * samples_per_step, ent["source"]
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    cob = fp32_cobound(w, precision == Q.FAST, cx.sm_count, sm_mhz)
    value = total_samples_step / (ms_dev * 1e-3) / 1e6
    achieved = alg_bytes / (ms_dev * 1e-3) / 1e9  # on the whole step's device time (every kernel of the chain)
    hbm_ceiling = peak * 1e9 / (alg_bytes / samples_per_step) / 1e6  # Msamples/s per GPU at 100 % of the measured HBM peak
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src,
            "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
            "kernel": kern_name, "kernel_ms_per_step": kern_ms,
            "kernel_frac": (alg_bytes / (kern_ms * 1e-3) / 1e9 / peak) if kern_ms > 0 else None,
            "step_ms": ms_dev, "algorithmic_bytes_per_step": alg_bytes, "peak_source": peak_src,
            "hbm_ceiling_msamples_per_s": hbm_ceiling}
    if "fk_tcfir" in (kern_name or ""):
        # the filter ran on the tensor cores: per 64-sample row one 128 (K) x N (columns) slice of a tcgen05.mma tile, N =
        # 2 tap halves x (re, im) x the outputs a row takes part in, padded to 8 columns per half
        lp = [st for st in w["stages"] if st[0] == "lowpass"][0]
        D, L = lp[2], lp[3]
        nout = (63 - (L - L // 2)) // D + ((L - L // 2) + L - 1) // D + 1
        n_cols = 2 * ((2 * nout + 7) // 8 * 8)
        flop = 2.0 * 128 * n_cols / 64
        tp, tp_src = getattr(cx, "tensor_peak", (None, None))
        ach = value / world * 1e6 * flop / 1e12
        roof["tensor_cobound"] = {"flop_per_sample": flop, "mma_n": n_cols, "achieved_tflops": ach, "peak_tflops": tp,
                                  "peak_source": tp_src, "frac": (ach / tp) if tp else None,
                                  "note": "f16 x f16 -> f32 tcgen05.mma; the HBM roofline binds, not the tensor cores"}
        cob = None  # the packed-FP32 floor describes the CUDA-core filter
    if cob:
        per_gpu = value / world
        cob["min_hbm_fp32_ceiling_msamples_per_s"] = min(hbm_ceiling, cob["ceiling_msamples_per_s"])
        cob["frac_of_min_ceiling"] = per_gpu / cob["min_hbm_fp32_ceiling_msamples_per_s"]
        roof["fp32_cobound"] = cob
    res = {"value": value, "unit": "Msamples/s", "ms_per_step": ms_dev, "steps": steps, "precision": prec_name,
           "samples_per_gpu": samples * passes, "units_per_gpu": n_units, "gpu_launches": launches_all,
           "scaling": w["scaling"], "roofline": roof, "clocks": clocks,
           "config": workload_config(name, w, total)}
    if exact_ms:
        res["exact_mode"] = {"value": total_samples_step / (exact_ms * 1e-3) / 1e6, "unit": "Msamples/s",
                             "ms_per_step": exact_ms,
                             "note": "same workload in EXACT arithmetic (bit-identical to the CPU oracle); FAST is within "
                                     "1e-5 of it (tests/test_gpu_fast.py::test_fast_mode_full_size_config2_against_exact)"}
    if e2e:
        res["e2e"] = e2e
    if cpu:
        res["cpu_baseline"] = cpu
        res["parity_checked"] = cpu["parity_checked"]
    return res


def _run_e2e_rank0(cx, name, w, plans, total, precision, build_chain, d_in, d_out, out_per_unit, n_units_rank):
    """Pinned host capture -> H2D -> kernels -> D2H of the result, inside the timed region, through the C ABI.  With
    several GPUs rank 0 alone drives all of them with one sharded chain (the other ranks wait at a CPU barrier)."""
    import torch

    Q, args = cx.Q, cx.args
    rank, world, local, dev, stream = cx.rank, cx.world, cx.local, cx.dev, cx.stream
    fmt, rate, pb = w["fmt"], w["rate"], PAIR[w["fmt"]]
    passes = w.get("passes", 1)
    unit_len, stride = unit_geometry(w)
    sk = sink_kind(w)
    mult, _ = chain_geometry(w)
    res = None
    if True:
        if world == 1:
            # the capture of every pass is this rank's resident range
            n_host = max(p.n_samples for p in plans)
            h_in = cx.pinned_in(n_host * pb)
            h_in.copy_(d_in[: n_host * pb])
            jobs = [(p.first_sample, p.n_samples, p.first_unit, p.n_units) for p in plans]
            devices = None
        else:
            # the whole capture in host memory (generated in pieces on this GPU), all units, all devices
            whole = Q.shard_plan(fmt, rate, total, w["stages"], sk, unit_len, stride, 1, 0)
            n_host = whole.n_samples
            h_in = cx.pinned_in(n_host * pb)
            synth = Q.make_synth(w["seed"], [(Q.tone_step(f, rate), a, k) for f, a, k in w["tones"]], w["noise"])
            piece = 1 << 30
            tmp = torch.empty(piece * pb, dtype=torch.uint8, device=dev)
            for s0 in range(0, n_host, piece):
                n = min(piece, n_host - s0)
                Q.synth_fill_device(synth, fmt, whole.first_sample + s0, n, tmp.data_ptr(), local, stream.cuda_stream)
                stream.synchronize()
                h_in[s0 * pb : (s0 + n) * pb].copy_(tmp[: n * pb])
            del tmp
            jobs = [(whole.first_sample, whole.n_samples, whole.first_unit, whole.n_units)]
            devices = list(range(world))
        units_all = sum(j[3] for j in jobs)
        h_out = cx.pinned_out(units_all * out_per_unit)
        chains = []
        for first_sample, n_samples, _, _ in jobs:
            src = Q.Samples.from_host_ptr(h_in.data_ptr(), n_samples * pb, fmt, rate, local if devices is None else devices,
                                          base_sample=first_sample, total_samples=total, keep=(h_in,))
            chains.append(build_chain(src, precision, on_stream=devices is None))

        def step():
            off = 0
            for c, (_, _, first_unit, nu) in zip(chains, jobs):
                if sk == 0:
                    c.write_into(unit_len, first_unit, nu, h_out.data_ptr() + off, nu * unit_len, Q._lib.SPACE_HOST)
                else:
                    c.spark_fft_into(unit_len, stride, w["sink"][3], first_unit, nu, h_out.data_ptr() + off)
                off += nu * out_per_unit
            for c in chains:
                c.synchronize()

        k = 3 if passes == 1 else 2
        step()
        if passes == 1:
            step()
        t0 = time.perf_counter()
        for _ in range(k):
            step()
        ms = 1e3 * (time.perf_counter() - t0) / k
        # the host-path result is the device-path result (rank 0's own units)
        nb = n_units_rank * out_per_unit
        same = bool(torch.equal(h_out[:nb], d_out[:nb].cpu()))
        samples_step = units_all * stride * mult
        res = {"value": samples_step / (ms * 1e-3) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": sum(j[1] for j in jobs) * pb, "d2h_bytes_per_step": units_all * out_per_unit,
               "ms_per_step": ms, "steps": k, "same_as_device_path": same,
               "driver": "one host process" + (f", one sharded chain over {world} devices (qd_chain_create_sharded)"
                                                if devices else "")}
        del chains
    return res


def run_e2e(cx, name, w, plans, total, precision, build_chain, d_in, d_out, out_per_unit, n_units_rank):
    """Rank 0 measures; every rank meets at the CPU barrier afterwards, whatever happened (a failed pinned
    allocation on a small host must not cost the headline or hang the other ranks)."""
    res = None
    try:
        if cx.rank == 0:
            res = _run_e2e_rank0(cx, name, w, plans, total, precision, build_chain, d_in, d_out, out_per_unit, n_units_rank)
    except Exception as e:  # noqa: BLE001
        res = {"error": f"{type(e).__name__}: {e}"}
    finally:
        cx.cpu_barrier()
    return res


def host_copy_peak(cx):
    """Aggregate pinned-host -> device copy bandwidth with every GPU copying at once and no kernel running: the
    denominator of the end-to-end number at N GPUs (rank 0 drives all devices, as the e2e leg does)."""
    import torch

    if cx.rank != 0:
        cx.cpu_barrier()
        return None
    try:
        return _host_copy_peak_rank0(cx)
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}
    finally:
        cx.cpu_barrier()


def _host_copy_peak_rank0(cx):
    import torch

    n = 1 << 30
    h = cx.pinned_in(n * cx.world)
    bufs, streams = [], []
    for d in range(cx.world):
        with torch.cuda.device(d):
            bufs.append(torch.empty(n, dtype=torch.uint8, device=f"cuda:{d}"))
            streams.append(torch.cuda.Stream(device=d))

    def once():
        for d in range(cx.world):
            with torch.cuda.stream(streams[d]):
                bufs[d].copy_(h[d * n : (d + 1) * n], non_blocking=True)
        for s in streams:
            s.synchronize()

    once()
    t0 = time.perf_counter()
    for _ in range(3):
        once()
    dt = (time.perf_counter() - t0) / 3
    del bufs
    return {"h2d_gb_per_s": cx.world * n / dt / 1e9, "devices": cx.world,
            "how": "1 GiB per device from one pinned buffer, all devices at once, 3 repeats, wall clock"}


def run_b200(args):
    import torch
    import torch.distributed as dist

    import quadrs_b200 as Q

    cx = Ctx()
    cx.Q, cx.args = Q, args
    cx.rank = int(os.environ.get("RANK", "0"))
    cx.world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; quadrs_b200 has no CPU fallback")
    torch.cuda.set_device(cx.local)
    cx.dev = torch.device("cuda", cx.local)
    # no CPU binding here: rank 0's library worker threads bind themselves to each GPU's local CPUs (qd_multi.cu)
    gloo = None
    if cx.world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=cx.dev)
        gloo = dist.new_group(backend="gloo")  # CPU-side waits while rank 0 drives every device (e2e leg)

    def barrier():
        if cx.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_barrier():
        if cx.world > 1:
            dist.barrier(group=gloo)

    cx.barrier, cx.cpu_barrier = barrier, cpu_barrier
    cx.stream = torch.cuda.Stream(device=cx.dev)  # the library's kernels, our CUDA events and the generator all run on it
    torch.cuda.set_stream(cx.stream)
    cx.sm_count = torch.cuda.get_device_properties(cx.dev).multi_processor_count
    pin = {}

    def pinned(kind, nbytes):  # one pinned buffer per direction, grown on demand (pinning tens of GiB is slow)
        t = pin.get(kind)
        if t is None or t.numel() < nbytes:
            pin[kind] = None
            t = torch.empty(nbytes + (16 << 20), dtype=torch.uint8, pin_memory=True)  # slack: later workloads reuse it
            pin[kind] = t
        return t[:nbytes]

    cx.pinned_in = lambda n: pinned("in", n)
    cx.pinned_out = lambda n: pinned("out", n + 64)
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        cx.peak = (json.loads(peaks_path.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)")
        cx.tensor_peak = (json.loads(peaks_path.read_text()).get("bf16_tflops_sustained"), "MEASURED_PEAKS.json bf16_tflops_sustained")
    else:
        cx.peak = (6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)")
    tpath = ROOT / "profiles" / "traffic.json"
    cx.traffic = json.loads(tpath.read_text()) if tpath.exists() else {}

    head_name = HEADLINE if args.workload == "all" else args.workload
    t_start = time.time()
    head = run_workload(cx, head_name, WORKLOADS[head_name], True)
    others = {}
    if args.workload == "all":
        for nm in MAP_ORDER:
            try:
                r = run_workload(cx, nm, WORKLOADS[nm], False)
            except Exception as e:  # a failing side workload must not cost the headline
                r = {"error": f"{type(e).__name__}: {e}"}
            if cx.rank == 0:
                others[nm] = r
    copy_peak = host_copy_peak(cx) if (cx.world > 1 and not args.no_e2e) else None

    if cx.rank == 0:
        w = WORKLOADS[head_name]
        line = {
            "metric": METRIC, "value": head["value"], "unit": "Msamples/s", "n_gpus": cx.world, "steps": head["steps"],
            "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": w["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": head["config"],
            "run": {"precision": head["precision"], "samples_per_gpu": head["samples_per_gpu"],
                    "units_per_gpu": head["units_per_gpu"],
                    "sharding": f"{cx.world} contiguous unit ranges with halo, absolute indices, no collective",
                    "wall_s": round(time.time() - t_start, 1)},
            "clocks": head["clocks"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
        }
        for k in ("exact_mode", "e2e", "cpu_baseline", "parity_checked"):
            if k in head:
                line[k] = head[k]
        if copy_peak:
            line["host_copy_peak"] = copy_peak
            if "e2e" in line and "ms_per_step" in line["e2e"] and "h2d_gb_per_s" in copy_peak:
                bps = line["e2e"]["h2d_bytes_per_step"] / (line["e2e"]["ms_per_step"] * 1e-3) / 1e9
                line["e2e"]["h2d_gb_per_s"] = bps
                line["e2e"]["frac_of_host_copy_peak"] = bps / copy_peak["h2d_gb_per_s"]
        if others:
            for r in others.values():
                if r:
                    r.pop("clocks", None)
            line["configs"] = others
        print(json.dumps(line), flush=True)
    if cx.world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        name = HEADLINE if args.workload == "all" else args.workload
        run_reference(args, name, WORKLOADS[name])
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
