#!/bin/bash
# quick A/B lines after a kernel change: GPU suite, then the workloads the change touches
set -u
mkdir -p gpurun_out
TAG=r2u WL="${WL:-cfg3:268435456 cfg4:268435456 cfg2s:268435456}" bash scripts/gpu_quick.sh
[ -n "${WLX:-}" ] && TAG=r2u_ex NOTEST=1 OPTS="--precision exact" WL="$WLX" bash scripts/gpu_quick.sh
