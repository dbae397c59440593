#!/bin/bash
# r2u: linear glyph epilogue, compile-time padded offsets, thread-ordered twiddles in fk_stft: GPU suite, A/B bench lines
set -u
mkdir -p gpurun_out
TAG=r2u WL="cfg3:268435456 cfg2s:268435456 cfg1:67108864" bash scripts/gpu_quick.sh
TAG=r2u_b3 NOTEST=1 OPTS="--opt stft_minb=3" WL="cfg3:268435456" bash scripts/gpu_quick.sh
TAG=r2u_b2 NOTEST=1 OPTS="--opt stft_minb=2" WL="cfg3:268435456" bash scripts/gpu_quick.sh
