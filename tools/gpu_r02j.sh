#!/bin/bash
# r02j (8 GPUs): the final build at 8 ranks — the contract bench line (its own 8-rank parity check first)
set -u
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29608 bench.py --gpus 8 --steps 20 --warmup 5 \
    > $OUT/bench_8gpu_$TAG.json 2> $OUT/bench_8gpu_$TAG.err
echo rc=$?
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_8gpu_$TAG.json").read().strip().splitlines()[-1])
    print("N=8", {k: d.get(k) for k in ("value", "ms_per_step")}, d.get("parity"), d["clocks"], d["timing_passes"])
    for k in d.get("kernels", []):
        print("    ", k["kernel"], k["launches_per_step"], round(k["ms_per_launch"], 4), round(k["frac_hbm"] or 0, 3))
except Exception as e:
    print("bench line unreadable:", e)
PY
tail -c 400 $OUT/bench_8gpu_$TAG.err
