#!/usr/bin/env python
"""profiles/traffic.json from the ncu summaries of an evidence run (scripts/gpu_evidence.sh):

    python scripts/make_traffic.py r2t r2v        # later tags override earlier ones, workload by workload

For every workload: DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels of ONE pass of the chain
over the captured sample count, per input sample.  bench.py scales that to its own sample count.
"""
import json
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tags = sys.argv[1:] or ["r2z"]
out = {}
files = [(tag, f) for tag in tags for f in sorted((ROOT / "profiles").glob(f"{tag}_full_*_summary.txt"))]
for tag, f in files:
    m = re.match(rf"{tag}_full_(\w+?)_(exact|fast)_summary", f.stem)
    if not m:
        continue
    wl, prec = m.groups()
    text = f.read_text()
    cmd = text.split("\n", 1)[0]
    ms = re.search(r"--samples (\d+)", cmd)
    samples = int(ms.group(1))
    kernels = {}
    for blk in text.split("Kernel Name = ")[1:]:
        name = blk.split("\n", 1)[0].strip()
        mt = re.search(r"DRAM traffic per launch = ([0-9.]+) GB", blk)
        md = re.search(r"gpu__time_duration.sum = ([0-9.]+) (us|ms)", blk)
        if mt and name not in kernels:  # one launch of each kernel = one pass of the chain
            kernels[name] = {"dram_bytes": float(mt.group(1)) * 1e9,
                             "duration_us": float(md.group(1)) * (1000.0 if md.group(2) == "ms" else 1.0)}
    total = sum(k["dram_bytes"] for k in kernels.values())
    out[f"{wl}:{prec}"] = {"dram_bytes_per_sample": total / samples, "capture_samples": samples, "kernels": kernels,
                            "source": f"profiles/{f.name} (dram__bytes_read.sum + dram__bytes_write.sum of one pass, ncu --set full), "
                                      f"scaled from a {samples}-sample capture"}
(ROOT / "profiles" / "traffic.json").write_text(json.dumps(out, indent=1) + "\n")
for k, v in out.items():
    print(k, round(v["dram_bytes_per_sample"], 4), "B/sample", list(v["kernels"]))
