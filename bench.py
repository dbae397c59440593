#!/usr/bin/env python
"""bench.py -- throughput of the quadrs IQ DSP hot path on B200, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input.  The default workload is
BASELINE.json configs[1]: synthetic cs8 (HackRF) at 20 MS/s, 2^30 samples per GPU, decode + shift +
lowpass -power 20 -decimate 8, delivered as do_write's 0x1000-sample chunks.  With N GPUs the logical
capture is N * 2^30 samples and rank r owns the r-th contiguous range of write chunks plus its halo
(weak scaling; absolute sample indices drive phase and truncation; no data-path collective).

Rank 0 prints ONE JSON line.  `value` = input Msamples/s with the input resident in HBM, timed with
CUDA events on the stream the kernels run on (max over ranks).  `e2e` = the same metric through the
public API with HOST buffers: pinned host input -> H2D -> kernels -> D2H of the result, all inside the
timed region.  `roofline` = algorithmic bytes of the dominant kernel / its device time, against the
measured HBM peak of MEASURED_PEAKS.json.  `cpu_baseline` = the CPU oracle (a port of the reference
algorithm; Rust cannot be built here) timed on a bounded sample of the same workload.
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

# name -> description of one workload.  sink: ("write", chunk) or ("sparkfft", W, S, (lo, hi))
WORKLOADS = {
    # BASELINE.json configs[1]
    "cfg2": dict(
        title="synthetic cs8 20 MS/s, 2^30 samples/GPU: decode + shift 1500000 + lowpass -power 20 -decimate 8 1000000 | write",
        fmt=CS8, rate=20_000_000, samples=2**30,
        stages=[("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], sink=("write", 0x1000),
        tones=[(1.6e6, 45, 0), (-4.1e6, 30, 0), (0.3e6, 20, 3000)], noise=6, seed=0x5EED0002,
        out_bytes_per_unit=0x1000 * 8, cpu_units=8192, ref_units_per_thread=512),
    # BASELINE.json configs[3] (per-GPU shard of 2^30 samples by default; --samples overrides)
    "cfg4": dict(
        title="synthetic cs16 100 MS/s: shift 7000000 | lowpass -power 400 -decimate 16 2000000 | sparkfft -width 128 -range 0.5:50",
        fmt=CS16, rate=100_000_000, samples=2**30,
        stages=[("shift", 7_000_000), ("lowpass", 2_000_000, 16, 800)], sink=("sparkfft", 128, 128, (0.5, 50.0)),
        tones=[(7.3e6, 9000, 0), (6.2e6, 6000, 50_000), (-20e6, 4000, 0)], noise=1200, seed=0x5EED0004,
        out_bytes_per_unit=128, cpu_units=2048, ref_units_per_thread=96),
    # BASELINE.json configs[4] shape
    "cfg5": dict(
        title="synthetic cf32 400 MS/s: lowpass -decimate 8 20000000 | lowpass -decimate 32 500000 | sparkfft -width 4 -stride 2 -range 0.001:0.01",
        fmt=CF32, rate=400_000_000, samples=2**29,
        stages=[("lowpass", 20_000_000, 8, 40), ("lowpass", 500_000, 32, 40)], sink=("sparkfft", 4, 2, (0.001, 0.01)),
        tones=[(0.1e6, 160, 1_000_000), (90e6, 3000, 0)], noise=40, seed=0x5EED0005,
        out_bytes_per_unit=4, cpu_units=4096, ref_units_per_thread=256),
    # configs[1]'s input through the metric's literal chain: shift + lowpass + sparkfft with overlapping windows
    "cfg2s": dict(
        title="synthetic cs8 20 MS/s: shift 1500000 | lowpass -power 20 -decimate 8 1000000 | sparkfft -width 64 -stride 16 -range 0.01:3",
        fmt=CS8, rate=20_000_000, samples=2**30,
        stages=[("shift", 1_500_000), ("lowpass", 1_000_000, 8, 40)], sink=("sparkfft", 64, 16, (0.01, 3.0)),
        tones=[(1.6e6, 45, 0), (-4.1e6, 30, 0), (0.3e6, 20, 3000)], noise=6, seed=0x5EED0002,
        out_bytes_per_unit=64, cpu_units=8192, ref_units_per_thread=512),
    # BASELINE.json configs[0] (the reference's own example chain) at capture scale: overlapping windows
    "cfg1": dict(
        title="synthetic cf32 21 MS/s: shift 280000 | lowpass -power 200 -decimate 32 200000 | sparkfft -width 64 -stride 16 -range 0.01:3",
        fmt=CF32, rate=21_000_000, samples=2**27,
        stages=[("shift", 280_000), ("lowpass", 200_000, 32, 400)], sink=("sparkfft", 64, 16, (0.01, 3.0)),
        tones=[(-250e3, 6000, 2000), (-310e3, 6000, 2000), (3e6, 9000, 0)], noise=300, seed=0x5EED0001,
        out_bytes_per_unit=64, cpu_units=2048, ref_units_per_thread=128),
    # BASELINE.json configs[2] shape
    "cfg3": dict(
        title="synthetic cu8 2.4 MS/s multi-tone: sparkfft -width 4096 -stride 1024 -range 2:500",
        fmt=CU8, rate=2_400_000, samples=2**28,
        stages=[], sink=("sparkfft", 4096, 1024, (2.0, 500.0)),
        tones=[(-800e3, 40, 0), (-123_456, 30, 0), (300e3, 25, 0), (1_000_001, 20, 0)], noise=4, seed=0x5EED0003,
        out_bytes_per_unit=4096, cpu_units=4096, ref_units_per_thread=256),
}

METRIC = "input Msamples/s through shift+lowpass+sparkfft"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--samples", type=int, default=0, help="samples per GPU (default: the workload's)")
    ap.add_argument("--precision", default="auto", choices=["auto", "exact", "fast"])
    ap.add_argument("--segment-mb", type=int, default=0, help="host-path segment size in MiB (default: the library's)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def sink_kind(w):
    return 0 if w["sink"][0] == "write" else 1


def unit_geometry(w):
    if w["sink"][0] == "write":
        return w["sink"][1], w["sink"][1]
    return w["sink"][1], w["sink"][2]


def make_oracle_synth(O, w):
    scale = 1
    return O.make_synth(w["seed"], [(O.tone_step(f, w["rate"]), a * scale, k) for f, a, k in w["tones"]], w["noise"])


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
def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    import numpy as np

    import oracle_lib as O

    threads = os.cpu_count() or 1
    unit_len, stride = unit_geometry(w)
    mult, need = 1, unit_len
    for st in reversed(w["stages"]):
        if st[0] == "lowpass":
            need = need * st[2] + st[3]
            mult *= st[2]
    units = min(threads * w["ref_units_per_thread"], 16384)  # bounded: at most 2^29 input samples per step for cfg2
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
    sample = f"{units} sink units = {samples_per_step} input samples per step, {threads} threads over disjoint unit ranges"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["title"], "name": args.workload,
                   "note": "CPU oracle: C port of the reference algorithm (lazy per-chunk pull, full-rate "
                           "complex_convolve, per-sample f64 sin/cos); the Rust reference cannot be built here"},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, w):
    import numpy as np
    import torch
    import torch.distributed as dist

    import quadrs_b200 as Q

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; quadrs_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_cpus(local)  # pinned host buffers (e2e) are first-touched on the GPU's NUMA node
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fmt, rate, pb = w["fmt"], w["rate"], PAIR[w["fmt"]]
    per_gpu = args.samples or w["samples"]
    total = per_gpu * world
    unit_len, stride = unit_geometry(w)
    sk = sink_kind(w)
    plan = Q.shard_plan(fmt, rate, total, w["stages"], sk, unit_len, stride, world, rank)
    n_units, first_unit = plan.n_units, plan.first_unit
    n_in = plan.n_samples

    # ---- synthetic input, generated in place on this GPU at its absolute sample range ----
    d_in = torch.empty(n_in * pb, dtype=torch.uint8, device=dev)
    synth = Q.make_synth(w["seed"], [(Q.tone_step(f, rate), a, k) for f, a, k in w["tones"]], w["noise"])
    # a non-default stream: the library's kernels, our CUDA events and the generator all run on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    Q.synth_fill_device(synth, fmt, plan.first_sample, n_in, d_in.data_ptr(), local, stream.cuda_stream)
    torch.cuda.synchronize()

    if args.precision == "exact":
        precision = Q.EXACT
    elif args.precision == "fast":
        precision = Q.FAST
    else:  # FAST only where it meets the 1e-5 bar: cs8 / cf32 with a cf32 sink (tests/test_gpu_fast.py)
        precision = Q.FAST if (fmt in (CS8, CF32) and sk == 0 and hasattr(Q, "FAST_READY")) else Q.EXACT

    def build_chain(src, prec=None):
        s = src
        for st in w["stages"]:
            s = s.shift(st[1]) if st[0] == "shift" else s.lowpass(st[1], st[2], st[3])
        s = s.with_precision(precision if prec is None else prec).with_stream(stream.cuda_stream)
        if args.segment_mb:
            s.set_option("segment_bytes", args.segment_mb << 20)
        return s

    dev_chain = build_chain(Q.Samples.from_device(d_in.data_ptr(), n_in * pb, fmt, rate, local,
                                                  base_sample=plan.first_sample, total_samples=total, keep=(d_in,)))
    out_bytes = n_units * w["out_bytes_per_unit"]
    d_out = torch.empty(out_bytes + 64, dtype=torch.uint8, device=dev)

    def run(chain, out_ptr, space):
        if sk == 0:
            n, _ = chain.write_into(unit_len, first_unit, n_units, out_ptr, n_units * unit_len, space)
            return n
        lo_hi = w["sink"][3]
        if space == Q._lib.SPACE_DEVICE:
            return chain.spark_fft_device(unit_len, stride, lo_hi, first_unit, n_units, out_ptr)
        return chain.spark_fft_into(unit_len, stride, lo_hi, first_unit, n_units, out_ptr)

    mult = 1
    for st in w["stages"]:
        if st[0] == "lowpass":
            mult *= st[2]
    samples_per_step = n_units * stride * mult  # input samples consumed by this rank's units (halo excluded)

    # ---- device-resident timing (`value`) ----
    lib = Q._lib.lib()
    for _ in range(max(args.warmup, 3)):
        produced = run(dev_chain, d_out.data_ptr(), Q._lib.SPACE_DEVICE)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_chain.profile(True)
    launches0 = lib.qd_kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        run(dev_chain, d_out.data_ptr(), Q._lib.SPACE_DEVICE)
    ev1.record(stream)
    barrier()
    ms_dev = ev0.elapsed_time(ev1) / args.steps
    launches = lib.qd_kernel_launches() - launches0
    regions, kern_ms, kern_name = dev_chain.profile_read()
    dev_chain.profile(False)
    clocks = sampler.stop() if rank == 0 else None

    # the bit-exact arithmetic mode, timed the same way, when the headline ran in FAST mode
    exact_ms = None
    if precision == Q.FAST:
        exact_chain = build_chain(Q.Samples.from_device(d_in.data_ptr(), n_in * pb, fmt, rate, local,
                                                        base_sample=plan.first_sample, total_samples=total,
                                                        keep=(d_in,)), Q.EXACT)
        for _ in range(3):
            run(exact_chain, d_out.data_ptr(), Q._lib.SPACE_DEVICE)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(min(args.steps, 5)):
            run(exact_chain, d_out.data_ptr(), Q._lib.SPACE_DEVICE)
        e1.record(stream)
        barrier()
        exact_ms = e0.elapsed_time(e1) / min(args.steps, 5)
        run(dev_chain, d_out.data_ptr(), Q._lib.SPACE_DEVICE)  # leave the FAST result in d_out for the e2e comparison
        barrier()

    # ---- end-to-end timing through host buffers (`e2e`) ----
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty(n_in * pb, dtype=torch.uint8, pin_memory=True)
        h_in.copy_(d_in)
        h_out = torch.empty(out_bytes + 64, dtype=torch.uint8, pin_memory=True)
        host_chain = build_chain(Q.Samples.from_host_ptr(h_in.data_ptr(), n_in * pb, fmt, rate, local,
                                                         base_sample=plan.first_sample, total_samples=total,
                                                         keep=(h_in,)))
        e2e_steps = max(1, min(args.steps, 5))
        for _ in range(2):
            run(host_chain, h_out.data_ptr(), Q._lib.SPACE_HOST)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            run(host_chain, h_out.data_ptr(), Q._lib.SPACE_HOST)
            host_chain.synchronize()
        barrier()
        ms_e2e = 1e3 * (time.perf_counter() - t0) / e2e_steps
        # the host-path result is the device-path result
        same = bool(torch.equal(h_out[:out_bytes], d_out[:out_bytes].cpu()))
        e2e = {"ms": ms_e2e, "same_as_device_path": same, "steps": e2e_steps}
        del h_in, h_out, host_chain

    # ---- max over ranks ----
    if world > 1:
        t = torch.tensor([ms_dev, e2e["ms"] if e2e else 0.0, float(samples_per_step), float(launches),
                          exact_ms or 0.0], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms_dev, ms_e2e_max = tmax[0].item(), tmax[1].item()
        exact_ms = tmax[4].item() or None
        total_samples_step = tsum[2].item()
        launches_all = int(tsum[3].item())
    else:
        ms_e2e_max = e2e["ms"] if e2e else 0.0
        total_samples_step = float(samples_per_step)
        launches_all = int(launches)

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = json.loads(peaks_path.read_text())["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (of measured)"
        else:
            peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
        # algorithmic bytes: input read once + final output written once (cf32 for write, u8 per bin for sparkfft)
        alg_bytes = n_in * pb + (produced * 8 if sk == 0 else n_units * unit_len)
        traffic, traffic_src = None, None
        tpath = ROOT / "profiles" / "traffic.json"
        if tpath.exists():
            ent = json.loads(tpath.read_text()).get(f"{args.workload}:{'fast' if precision == Q.FAST else 'exact'}:{per_gpu}")
            if ent:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        # FP32 co-bound of the FIR: complex MACs per input sample, each one packed FFMA2 in FAST
        # mode and two packed instructions (FMUL2 + FFMA2) in the reference's exact mul-then-add order; the FMA
        # pipe retires 64 packed lanes per clock per SM
        macs, rate_div = 0.0, 1
        for st in w["stages"]:
            if st[0] == "lowpass":
                rate_div *= st[2]
                macs += st[3] / rate_div
        fir = None
        if macs:
            instr = macs * (1 if precision == Q.FAST else 2)
            sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
            peak_instr = torch.cuda.get_device_properties(dev).multi_processor_count * 64 * sm_mhz * 1e6
            fir = {"complex_macs_per_sample": macs, "packed_instr_per_sample": instr,
                   "ceiling_msamples_per_s": peak_instr / instr / 1e6}
        kern_avg_ms = kern_ms / max(1, args.steps)  # device ms per step of the dominant kernel (CUDA events around its launches)
        achieved = alg_bytes / (kern_avg_ms * 1e-3) / 1e9 if kern_avg_ms > 0 else None
        value = total_samples_step / (ms_dev * 1e-3) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["title"], "name": args.workload, "samples_per_gpu": per_gpu,
                       "capture_samples": total, "units_per_gpu": n_units, "input_bytes_per_gpu": n_in * pb,
                       "precision": "fast" if precision == Q.FAST else "exact",
                       "l2": "inputs larger than L2 (no flush needed)" if n_in * pb > 256 * 2**20 else "input smaller than 2x L2",
                       "sharding": f"{world} contiguous unit ranges with halo, absolute indices, no collective",
                       "rank_cpu_affinity": affinity},
            "clocks": clocks,
            "gpu_launches": launches_all,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": kern_name,
                         "kernel_ms_per_step": kern_avg_ms, "algorithmic_bytes_per_step": alg_bytes,
                         "peak_source": peak_src, "fir_fp32_cobound": fir},
        }
        if fir:
            fir["frac_of_ceiling"] = (value / world) / fir["ceiling_msamples_per_s"]
        if exact_ms:
            line["exact_mode"] = {"value": total_samples_step / (exact_ms * 1e-3) / 1e6, "unit": "Msamples/s",
                                  "ms_per_step": exact_ms,
                                  "note": "same workload in EXACT arithmetic (bit-identical to the CPU oracle); the "
                                          "headline FAST mode is within 1e-5 of it on this workload "
                                          "(tests/test_gpu_fast.py::test_fast_mode_full_size_config2_against_exact)"}
        if e2e:
            line["e2e"] = {"value": total_samples_step / (ms_e2e_max * 1e-3) / 1e6, "unit": "Msamples/s",
                           "h2d_bytes_per_step": n_in * pb, "d2h_bytes_per_step": out_bytes,
                           "ms_per_step": ms_e2e_max, "same_as_device_path": e2e["same_as_device_path"]}
        if not args.no_cpu_baseline and world == 1:  # the CPU baseline is reported at N = 1 only
            line["cpu_baseline"] = cpu_baseline(w, d_in, plan, pb)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(w, d_in, plan, pb):
    """The oracle (a port of the reference algorithm), 1 thread as the reference runs, on the first
    cpu_units sink units of rank 0's own input bytes."""
    import oracle_lib as O

    unit_len, stride = unit_geometry(w)
    mult, need = 1, unit_len
    for st in reversed(w["stages"]):
        if st[0] == "lowpass":
            need = need * st[2] + st[3]
            mult *= st[2]
    units = min(w["cpu_units"], plan.n_units)
    n = min((units - 1) * stride * mult + need + 64, plan.n_samples)
    raw = d_in[: n * pb].cpu().numpy()
    sink = w["sink"]
    kw = dict(width=sink[1], stride=sink[2], rng=sink[3]) if sink[0] == "sparkfft" else {}
    # rank 0 of a sharded run starts at sample 0, so unit indices are the shard's own
    secs, _ = O.timed_run(raw, w["fmt"], w["rate"], w["stages"], "write" if sink[0] == "write" else "sparkfft",
                          0, units, 1, **kw)
    samples = units * stride * mult
    return {"value": samples / secs / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
            "sample": f"first {units} sink units ({samples} input samples) of the same input, {secs:.1f} s, 1 thread "
                      "(the reference hot path is single-threaded)"}


def main():
    args = parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_b200(args, w)


if __name__ == "__main__":
    main()
