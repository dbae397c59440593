#!/bin/bash
# same-box A/B of the tile-overlap carry (fk_fir CARRY): config 4 and config 1, carry on / off, twice each
set -u
mkdir -p gpurun_out
for rep in 1 2; do
  TAG=r2x_on$rep NOTEST=1 WL="cfg4:1073741824 cfg1:134217728" bash scripts/gpu_quick.sh
  TAG=r2x_off$rep NOTEST=1 OPTS="--opt fir_carry=0" WL="cfg4:1073741824 cfg1:134217728" bash scripts/gpu_quick.sh
done
