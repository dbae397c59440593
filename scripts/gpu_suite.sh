#!/bin/bash
# the whole GPU suite, then the launch list of config 2 FAST
set -u
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q -x --durations=10 ) > gpurun_out/suite_pytest.log 2>&1; tail -25 gpurun_out/suite_pytest.log
A="--workload cfg2 --samples 1073741824 --precision fast --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/suite_launches.csv python bench.py $A > /dev/null 2>&1
grep -E "fk_" gpurun_out/suite_launches.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head
