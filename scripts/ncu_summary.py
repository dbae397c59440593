#!/usr/bin/env python
"""Summarise an .ncu-rep (captured under gpurun with `ncu --set full`) into the text kept under profiles/.

    python scripts/ncu_summary.py gpurun_out/x.ncu-rep [--title "..."] [--stalls] > profiles/x_ncu_summary.txt

Per kernel launch: duration, DRAM bytes, pipe utilisation, occupancy limits, instruction counts; with --stalls
the stall-reason shares of the not-issued warp samples from the source page (needs -lineinfo builds).
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__cycles_elapsed.avg.per_second",
]


def raw_page(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    return hdr, units, rows[2:]


def stalls(path):
    """-> per kernel: ({stall reason: not-issued samples}, [(samples, instructions executed, sass)] hottest first)"""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    res = []
    hdr, cur, hot = None, None, None
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            hdr = None
            continue
        if row[0] == "Address":
            hdr = row
            cur, hot = defaultdict(float), []
            res.append((cur, hot))
            continue
        if hdr is None or len(row) < len(hdr):
            continue
        d = dict(zip(hdr, row))
        for name, val in d.items():
            if name.startswith("stall_") and name.endswith("(Not Issued)"):
                try:
                    cur[name[6:-13]] += float(val)
                except ValueError:
                    pass
        try:
            hot.append((int(d["# Samples"]), int(d["Instructions Executed"]), d["Source"].strip()))
        except (KeyError, ValueError):
            pass
    return res


def main():
    path = sys.argv[1]
    title = None
    if "--title" in sys.argv:
        title = sys.argv[sys.argv.index("--title") + 1]
    hdr, units, rows = raw_page(path)
    col = {h: i for i, h in enumerate(hdr)}
    if title:
        print(title)
    for r in rows:
        print(f"Kernel Name = {r[col['Kernel Name']]}")
        for m in METRICS:
            if m in col:
                print(f"{m} = {r[col[m]]} {units[col[m]]}")
        try:
            rd = float(r[col["dram__bytes_read.sum"]])
            wr = float(r[col["dram__bytes_write.sum"]])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            t = rd * scale[units[col["dram__bytes_read.sum"]]] + wr * scale[units[col["dram__bytes_write.sum"]]]
            print(f"DRAM traffic per launch = {t / 1e9:.4f} GB")
        except Exception:
            pass
        print()
    if "--stalls" in sys.argv:
        for i, (s, hot) in enumerate(stalls(path)):
            tot = sum(s.values())
            if tot <= 0:
                continue
            top = sorted(s.items(), key=lambda kv: -kv[1])[:8]
            print(f"launch {i}: stall reasons (not-issued warp samples): " + ", ".join(f"{k} {100 * v / tot:.0f} %" for k, v in top))
            n_inst = sum(h[1] for h in hot)
            by_op = defaultdict(int)
            for smp, ex, src in hot:
                by_op[src.split()[0].split(".")[0] if not src.startswith("@") else src.split()[1].split(".")[0]] += ex
            print(f"launch {i}: warp instructions by opcode: " + ", ".join(f"{k} {100 * v / max(1, n_inst):.1f} %" for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:14]))
            if "--hot" in sys.argv:
                allsmp = sum(h[0] for h in hot)
                for smp, ex, src in sorted(hot, key=lambda h: -h[0])[:25]:
                    print(f"    {100 * smp / max(1, allsmp):5.2f} % samples  {ex:>10} exec  {src}")


if __name__ == "__main__":
    main()
