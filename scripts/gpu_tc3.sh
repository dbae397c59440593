#!/bin/bash
# the whole FAST suite with the tensor-core kernel as the default, launch list of config 2, ncu capture
set -u
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fast.py tests/test_gpu_tcfir.py tests/test_gpu_random.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -25 ) > gpurun_out/tc3_pytest.log 2>&1
tail -12 gpurun_out/tc3_pytest.log
A="--workload cfg2 --samples 268435456 --precision fast --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
python bench.py $A > gpurun_out/tc3_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/tc3_launches.csv python bench.py $A > /dev/null 2>&1
grep -E "fk_|gk_" gpurun_out/tc3_launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fk_tcfir' -s 1 -c 1 -f -o gpurun_out/tc3_full python bench.py $A > gpurun_out/tc3_ncu.log 2>&1
python scripts/ncu_summary.py gpurun_out/tc3_full.ncu-rep --stalls --hot --title "ncu --set full: python bench.py $A" > gpurun_out/tc3_summary.txt 2>&1
head -34 gpurun_out/tc3_summary.txt
