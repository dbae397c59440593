#!/bin/bash
# Evidence run: GPU tests, the default bench line (both arms), and for every workload a launch list plus one
# `ncu --set full` capture of its kernels at a reduced size.  Summaries are made on the box (the reports are too large
# to travel).  TAG=r2z by default.
set -u
mkdir -p gpurun_out
O=gpurun_out; TAG=${TAG:-r2z}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
( time python -m pytest tests -m gpu -q --durations=8 ) > $O/${TAG}_pytest.log 2>&1; tail -3 $O/${TAG}_pytest.log
( time python bench.py --steps 20 --warmup 5 ) > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -c 300 $O/${TAG}_bench.err
( time python bench.py --impl reference --steps 5 --warmup 1 ) > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err
B="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
for item in ${ITEMS:-cfg4:268435456:exact cfg1:67108864:exact cfg2:268435456:fast cfg2:268435456:exact cfg2s:268435456:exact cfg3:268435456:exact cfg5:268435456:exact}; do
  wl=$(echo $item | cut -d: -f1); S=$(echo $item | cut -d: -f2); P=$(echo $item | cut -d: -f3)
  A="--workload $wl --samples $S --precision $P $B"
  python bench.py $A > $O/${TAG}_plain_${wl}_$P.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/${TAG}_launches_${wl}_$P.csv python bench.py $A > /dev/null 2>&1
  # the third warm-up step onwards: one step's worth of our kernels
  K=2; [ $wl = cfg4 ] && K=1; [ $wl = cfg3 ] && K=1; [ $wl = cfg2 ] && K=1; [ $wl = cfg5 ] && K=3
  SK=$((2*K)); [ $wl = cfg5 ] && SK=24
  python bench.py $A > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'fk_fir|fk_tcfir|fk_tail|fk_stft' -s $SK -c $K -f -o $O/${TAG}_full_${wl}_$P python bench.py $A > $O/${TAG}_ncuf_${wl}_$P.log 2>&1
  python scripts/ncu_summary.py $O/${TAG}_full_${wl}_$P.ncu-rep --stalls --hot --title "ncu --set full --clock-control none --import-source on: python bench.py $A" > $O/${TAG}_full_${wl}_${P}_summary.txt 2>&1
  rm -f $O/${TAG}_full_${wl}_$P.ncu-rep
done
ls -la $O | grep ${TAG} | wc -l
