#!/bin/bash
# multi-GPU: the sharded-chain tests on real devices, then the bench line under torchrun as the driver launches it
set -u
mkdir -p gpurun_out
O=gpurun_out; N=${N:-2}; TAG=${TAG:-m}
nvidia-smi -L > $O/${TAG}_gpus.txt; nvidia-smi topo -m >> $O/${TAG}_gpus.txt 2>&1; free -g >> $O/${TAG}_gpus.txt; nproc >> $O/${TAG}_gpus.txt
( time python -m pytest tests/test_gpu_multi.py -m gpu -q ) > $O/${TAG}_pytest_multi.log 2>&1; tail -4 $O/${TAG}_pytest_multi.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 ${BENCH_ARGS:-} ) > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err
tail -c 400 $O/${TAG}_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1])
    print("N=$N value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e"), "copy_peak", d.get("host_copy_peak"))
    for k,v in d.get("configs",{}).items(): print(" ", k, v.get("value") and round(v["value"]), v.get("error"))
except Exception as e: print("bench parse failed", e)
PY
