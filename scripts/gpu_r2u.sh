#!/bin/bash
# r2u: quick A/B lines after a kernel change: GPU suite, then the sparkfft workloads and config 2 in EXACT
set -u
mkdir -p gpurun_out
TAG=r2u WL="cfg3:268435456 cfg2s:268435456 cfg1:134217728 cfg4:268435456" bash scripts/gpu_quick.sh
TAG=r2u_ex NOTEST=1 OPTS="--precision exact" WL="cfg2:268435456" bash scripts/gpu_quick.sh
