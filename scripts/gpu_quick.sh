#!/bin/bash
# quick check: GPU tests (stop at first failure) + a few short bench lines; WL="name:samples ..." TAG=...
set -u
mkdir -p gpurun_out
O=gpurun_out; TAG=${TAG:-q}
if [ "${NOTEST:-0}" = 0 ]; then ( time python -m pytest tests -m gpu -q -x ) > $O/${TAG}_pytest.log 2>&1; tail -4 $O/${TAG}_pytest.log; fi
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for item in ${WL:-cfg4:268435456}; do
  wl=${item%%:*}; S=${item##*:}
  python bench.py --workload $wl --samples $S $B ${OPTS:-} > $O/${TAG}_bench_$wl.json 2> $O/${TAG}_bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/${TAG}_bench_$wl.json").read().strip().splitlines()[-1])
    ex=d.get("exact_mode",{}).get("value")
    print("$wl", d["run"]["precision"], round(d["value"]), "ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "launches", d["gpu_launches"], "exact", ex and round(ex))
except Exception as e: print("$wl", "failed", e); print(open("$O/${TAG}_bench_$wl.err").read()[-1500:])
PY
done
