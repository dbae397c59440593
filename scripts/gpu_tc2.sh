#!/bin/bash
# ncu capture of fk_tcfir on config 2 at 2^28 samples (one launch), report brought back for reading here
set -u
mkdir -p gpurun_out
A="--workload cfg2 --samples 268435456 --precision fast --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py $A > gpurun_out/tc2_plain.log 2>&1 || { tail -5 gpurun_out/tc2_plain.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fk_tcfir' -s 1 -c 1 -f -o gpurun_out/tc2_full python bench.py $A > gpurun_out/tc2_ncu.log 2>&1
ls -la gpurun_out/tc2_full.ncu-rep
python scripts/ncu_summary.py gpurun_out/tc2_full.ncu-rep --stalls --hot --title "ncu --set full: python bench.py $A" > gpurun_out/tc2_summary.txt 2>&1
head -50 gpurun_out/tc2_summary.txt
