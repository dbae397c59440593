#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
( time python -m pytest tests -m gpu -q -x ) > $O/r2d_pytest.log 2>&1
tail -5 $O/r2d_pytest.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for wl in cfg4 cfg1 cfg2s x_fir16 x_mix16; do
  S=$((2**30)); [ $wl = cfg1 ] && S=$((2**27)); [ $wl = x_fir16 ] && S=$((2**28)); [ $wl = x_mix16 ] && S=$((2**28))
  python bench.py --workload $wl --samples $S $B > $O/r2d_bench_$wl.json 2> $O/r2d_bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2d_bench_$wl.json").read().strip().splitlines()[-1])
    print("$wl", round(d["value"]), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["gpu_launches"])
except Exception as e: print("$wl", "failed", e)
PY
done
for wl in x_fir16; do
S=$((2**28)); B="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py --workload $wl --samples $S $B > $O/r2d_plain2_$wl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fk_fir' -s 2 -c 1 -f -o $O/r2d_full_$wl \
    python bench.py --workload $wl --samples $S $B > $O/r2d_ncuf_$wl.log 2>&1
python scripts/ncu_summary.py $O/r2d_full_$wl.ncu-rep --stalls --hot > $O/r2d_full_${wl}_summary.txt 2>&1
ncu -i $O/r2d_full_$wl.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > $O/r2d_full_${wl}_source.csv.gz
rm -f $O/r2d_full_$wl.ncu-rep
done
