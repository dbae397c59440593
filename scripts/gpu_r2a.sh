#!/bin/bash
# round 2, call A: the GPU test-suite, the full bench line, launch lists and full ncu captures of the kernels that
# had no evidence in round 1 (cfg4 / cfg1 / cfg5 FIR, fk_tail, small STFTs).  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/r2a_gpu.txt 2>&1
free -g >> $O/r2a_gpu.txt; nproc >> $O/r2a_gpu.txt
( time python -m pytest tests -m gpu -q --durations=15 ) > $O/r2a_pytest.log 2>&1
tail -5 $O/r2a_pytest.log
( time python bench.py ) > $O/r2a_bench.json 2> $O/r2a_bench.err
tail -c 600 $O/r2a_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r2a_bench_ref.json 2>> $O/r2a_bench.err
B="--steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
for wl in cfg4 cfg1 cfg5 cfg2s cfg3 cfg2; do
  S=$((2**28)); [ $wl = cfg1 ] && S=$((2**26)); [ $wl = cfg5 ] && S=$((2**28))
  python bench.py --workload $wl --samples $S $B > $O/r2a_plain_$wl.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2a_launches_$wl.csv \
      python bench.py --workload $wl --samples $S $B > $O/r2a_ncul_$wl.log 2>&1
done
for wl in cfg4 cfg1 cfg5; do
  S=$((2**28)); [ $wl = cfg1 ] && S=$((2**26))
  SK=4; CN=2; [ $wl = cfg1 ] && SK=6 && CN=3; [ $wl = cfg5 ] && SK=12 && CN=3
  python bench.py --workload $wl --samples $S $B > $O/r2a_plain2_$wl.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'fk_fir|fk_tail|fk_stft' -s $SK -c $CN -f -o $O/r2a_full_$wl \
      python bench.py --workload $wl --samples $S $B > $O/r2a_ncuf_$wl.log 2>&1
  # summaries are made here: the reports themselves are too large to travel back (64 MiB limit)
  python scripts/ncu_summary.py $O/r2a_full_$wl.ncu-rep --stalls --hot > $O/r2a_full_${wl}_summary.txt 2>&1
  ncu -i $O/r2a_full_$wl.ncu-rep --page source --csv --print-source sass 2>/dev/null | cut -d, -f1-6,31-64 | gzip -9 > $O/r2a_full_${wl}_source.csv.gz
  rm -f $O/r2a_full_$wl.ncu-rep
done
ls -la $O | tail -30
