#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
for cap in 1 2 3; do
for wl in x_fir16 x_mix16 cfg4; do
  S=$((2**28))
  python bench.py --workload $wl --samples $S $B --opt fir_cta_cap=$cap > $O/r2e_bench_${wl}_$cap.json 2> $O/r2e_bench_${wl}_$cap.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2e_bench_${wl}_$cap.json").read().strip().splitlines()[-1])
    print("$wl cap $cap", round(d["value"]), round(d["ms_per_step"],3))
except Exception as e: print("$wl", "failed", e)
PY
done
done
wl=x_fir16; S=$((2**28)); B="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline --opt fir_cta_cap=1"
python bench.py --workload $wl --samples $S $B > $O/r2e_plain2_$wl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'fk_fir' -s 2 -c 1 -f -o $O/r2e_full_$wl \
    python bench.py --workload $wl --samples $S $B > $O/r2e_ncuf_$wl.log 2>&1
python scripts/ncu_summary.py $O/r2e_full_$wl.ncu-rep --stalls --hot > $O/r2e_full_${wl}_summary.txt 2>&1
ncu -i $O/r2e_full_$wl.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > $O/r2e_full_${wl}_source.csv.gz
rm -f $O/r2e_full_$wl.ncu-rep
