#!/bin/bash
# Multi-GPU call of round 2:   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_round2_multi.sh N r02m'
# 1. N-rank parity checks (uniform path incl. the kzr halo of the deferred x-update, tags/dump, bodies, the distributed
#    general-rows Poisson solve, multi-level steps with replicated operators and with the mesh distributed);  2. the contract bench line at N GPUs;
# 3. config C5 (synthetic 3-level mesh, 32 M cells) at N GPUs.
set -u
N=${1:-2}
TAG=${2:-r02m}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== 1. parity under torchrun, $N ranks"
timeout 400 $TR --master-port 29571 tools/multi_gpu_check.py > $OUT/multi_gpu_check_${N}gpu_$TAG.jsonl 2> $OUT/multi_gpu_check_${N}gpu_$TAG.err
echo "rc=$?"; cat $OUT/multi_gpu_check_${N}gpu_$TAG.jsonl; tail -c 400 $OUT/multi_gpu_check_${N}gpu_$TAG.err
echo "== 2. bench.py at $N GPUs"
timeout 300 $TR --master-port 29572 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err
echo "rc=$?"; tail -c 300 $OUT/bench_${N}gpu_$TAG.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${N}gpu_$TAG.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "value", "ms_per_step")}, d["poisson_iteration"]["ms_per_iteration"], d["clocks"])
except Exception as e:
    print("bench line unreadable:", e)
PY
echo "== 3. config C5 at $N GPUs (mesh distributed by block ranges, cup2d_amr_create_ranks)"
timeout 400 $TR --master-port 29573 tools/bench_amr.py synthetic 9 10 10 1 > $OUT/bench_amr_c5_${N}gpu_$TAG.json 2> $OUT/bench_amr_c5_${N}gpu_$TAG.err
echo "rc=$?"; tail -c 600 $OUT/bench_amr_c5_${N}gpu_$TAG.json; tail -c 300 $OUT/bench_amr_c5_${N}gpu_$TAG.err
