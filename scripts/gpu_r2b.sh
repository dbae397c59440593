#!/bin/bash
# round 2, call B: the new long-filter FIR loop (uniform tap loads, 4 outputs per thread, unstaged integer tiles)
set -u
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -q -x ) > $O/r2b_pytest.log 2>&1
tail -5 $O/r2b_pytest.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for wl in cfg4 cfg1 cfg2s cfg5; do
  S=$((2**30)); [ $wl = cfg1 ] && S=$((2**27)); [ $wl = cfg5 ] && S=$((2**30))
  python bench.py --workload $wl --samples $S $B > $O/r2b_bench_$wl.json 2> $O/r2b_bench_$wl.err
  python - <<PY
import json
d=json.loads(open("$O/r2b_bench_$wl.json").read().strip().splitlines()[-1])
print("$wl", round(d["value"]), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4))
PY
done
wl=cfg4; S=$((2**28)); B="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py --workload $wl --samples $S $B > $O/r2b_plain2_$wl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fk_fir' -s 2 -c 1 -f -o $O/r2b_full_$wl \
    python bench.py --workload $wl --samples $S $B > $O/r2b_ncuf_$wl.log 2>&1
python scripts/ncu_summary.py $O/r2b_full_$wl.ncu-rep --stalls --hot > $O/r2b_full_${wl}_summary.txt 2>&1
ncu -i $O/r2b_full_$wl.ncu-rep --page source --csv --print-source sass 2>/dev/null | cut -d, -f1-6,31-64 | gzip -9 > $O/r2b_full_${wl}_source.csv.gz
rm -f $O/r2b_full_$wl.ncu-rep
