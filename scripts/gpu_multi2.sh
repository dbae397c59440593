#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out; N=${N:-2}; TAG=${TAG:-m}
( time python -m pytest tests/test_gpu_multi.py tests/test_cli.py -m gpu -q ) > $O/${TAG}_pytest_multi.log 2>&1; tail -3 $O/${TAG}_pytest_multi.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 ) > $O/${TAG}_bench_${N}gpu.json 2> $O/${TAG}_bench_${N}gpu.err
tail -c 300 $O/${TAG}_bench_${N}gpu.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 ) > $O/${TAG}_bench_ref_${N}gpu.json 2>> $O/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open("$O/${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1])
print("N=$N value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", d.get("e2e",{}).get("value"), d.get("e2e",{}).get("same_as_device_path"), "copy_peak", d.get("host_copy_peak",{}).get("h2d_gb_per_s"), "traffic/alg", d["roofline"].get("traffic_over_algorithmic"))
for k,v in d.get("configs",{}).items(): print(" ", k, v.get("value") and round(v["value"]), v.get("error"))
r=[l for l in open("$O/${TAG}_bench_ref_${N}gpu.json") if l.startswith("{")]
print("ref lines", len(r), json.loads(r[-1])["value"] if r else None)
PY
