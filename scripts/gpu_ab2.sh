set -u; mkdir -p gpurun_out
python -m pytest tests/test_gpu_fast.py tests/test_gpu_parity.py tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -2
run() { python bench.py --workload $1 --samples $2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --opt fuse_stft=$3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$1 fuse=$3', round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'])"; }
for f in 0 1 2; do run cfg5 1073741824 $f; done
for f in 0 2; do run cfg2s 1073741824 $f; run cfg1 134217728 $f; done
