#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 300 python -m pytest tests/test_gpu_tcfir.py -m gpu -q -x 2>&1 | tail -1; done
for tc in 1; do
timeout 200 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --opt use_tc=$tc 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('use_tc=$tc', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4))"
done
