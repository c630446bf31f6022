#!/bin/bash
# r02d (8 GPUs): parity + strong scaling of the contract bench line at 2, 8, 4 GPUs, oracle-based N-rank checks, C5 at 8 GPUs
set -u
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
bench() { # N
  timeout 300 $TR --nproc-per-node $1 --master-port $((29600 + $1)) bench.py --gpus $1 --steps 20 --warmup 5 > $OUT/bench_${1}gpu_$TAG.json 2> $OUT/bench_${1}gpu_$TAG.err
  local rc=$?
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${1}gpu_$TAG.json").read().strip().splitlines()[-1])
    print("N=$1 rc=$rc", {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d.get("parity"), d.get("timing_passes"))
    for k in d.get("kernels", []):
        print("    ", k["kernel"], k["launches_per_step"], round(k["ms_per_launch"], 4), round(k["frac_hbm"] or 0, 3))
except Exception as e:
    print("N=$1 rc=$rc bench line unreadable:", e)
PY
  return $rc
}
# (N=2 ran in its own 2-GPU call, r02c: parity 5e-15, 14.0 ms/step)
echo "== oracle-based N-rank checks, 8 ranks"
timeout 400 $TR --nproc-per-node 8 --master-port 29571 tools/multi_gpu_check.py > $OUT/multi_gpu_check_8gpu_$TAG.jsonl 2> $OUT/multi_gpu_check_8gpu_$TAG.err
echo "rc=$?"; cat $OUT/multi_gpu_check_8gpu_$TAG.jsonl; tail -c 300 $OUT/multi_gpu_check_8gpu_$TAG.err
echo "== bench N=8"; bench 8
echo "== bench N=4"; bench 4
echo "== config C5 at 8 GPUs"
timeout 400 $TR --nproc-per-node 8 --master-port 29573 tools/bench_amr.py synthetic 9 10 10 1 > $OUT/bench_amr_c5_8gpu_$TAG.json 2> $OUT/bench_amr_c5_8gpu_$TAG.err
echo "rc=$?"; tail -c 600 $OUT/bench_amr_c5_8gpu_$TAG.json; tail -c 300 $OUT/bench_amr_c5_8gpu_$TAG.err
