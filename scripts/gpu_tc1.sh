#!/bin/bash
# first contact of fk_tcfir with a GPU: its tests under a timeout (a hang must not cost the box), then config 2 with
# and without the tensor-core kernel
set -u
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_tcfir.py -m gpu -q -x -s 2>&1 | tail -40 ) > gpurun_out/tc1_pytest.log 2>&1
tail -30 gpurun_out/tc1_pytest.log
for tc in 1 0; do
timeout 200 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --opt use_tc=$tc > gpurun_out/tc1_bench_tc$tc.json 2> gpurun_out/tc1_bench_tc$tc.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/tc1_bench_tc$tc.json") if l.startswith("{")][-1])
    print("use_tc=$tc", round(d["value"]), "Msamples/s", round(d["ms_per_step"],4), "ms", d["roofline"]["kernel"], round(d["roofline"]["frac"],4), d.get("parity_checked"))
except Exception as e: print("bench parse failed", e); print(open("gpurun_out/tc1_bench_tc$tc.err").read()[-1500:])
PY
done
