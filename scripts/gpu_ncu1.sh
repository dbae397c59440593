#!/bin/bash
# one full ncu capture of fk_fir for a workload: WL=name S=samples TAG=... [OPTS]
set -u
mkdir -p gpurun_out
O=gpurun_out; wl=${WL}; S=${S}; TAG=${TAG}
B="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline ${OPTS:-}"
python bench.py --workload $wl --samples $S $B > $O/${TAG}_plain_$wl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KRE:-fk_fir} -s ${SKIP:-2} -c ${CNT:-1} -f -o $O/${TAG}_full_$wl \
    python bench.py --workload $wl --samples $S $B > $O/${TAG}_ncuf_$wl.log 2>&1
python scripts/ncu_summary.py $O/${TAG}_full_$wl.ncu-rep --stalls --hot > $O/${TAG}_full_${wl}_summary.txt 2>&1
ncu -i $O/${TAG}_full_$wl.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > $O/${TAG}_full_${wl}_source.csv.gz
rm -f $O/${TAG}_full_$wl.ncu-rep
tail -3 $O/${TAG}_full_${wl}_summary.txt | cut -c1-300
