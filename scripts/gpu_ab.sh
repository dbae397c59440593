set -u; mkdir -p gpurun_out
python -m pytest tests/test_gpu_fast.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for rep in 1 2; do for f in 1 0; do
python bench.py --workload cfg4 --samples 1073741824 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --opt fuse_stft=$f 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('fuse=$f', round(d['value']), round(d['ms_per_step'],3), d['gpu_launches'])"
done; done
