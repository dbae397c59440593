#!/bin/bash
# r2u: linear glyph epilogue, compile-time padded offsets, thread-ordered twiddles, conflict-free pitches in fk_stft
set -u
mkdir -p gpurun_out
TAG=r2u WL="cfg3:268435456 cfg2s:268435456 cfg1:134217728 cfg4:268435456" bash scripts/gpu_quick.sh
