#!/bin/bash
# e2e (host buffers, H2D + kernels + D2H) of the default bench workload against the host-path segment size
for m in 16 32 64 128 256; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --segment-mb $m 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print($m, d['e2e']['value'], d['e2e']['ms_per_step'])"
done
