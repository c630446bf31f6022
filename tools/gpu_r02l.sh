#!/bin/bash
# r02l (1 GPU): final build after the SpMV change: parity suite, contract bench line, launch list, ncu of the Krylov kernels
set -u
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=3) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -6 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_1gpu_$TAG.json 2> $OUT/bench_1gpu_$TAG.err; echo rc=$?
python - <<PY
import json
d = json.loads(open("$OUT/bench_1gpu_$TAG.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["clocks"], d["poisson_iteration"]["ms_per_iteration"], d["roofline"]["frac"])
for k in d.get("kernels", []):
    print("    ", k["kernel"], k["launches_per_step"], round(k["ms_per_launch"], 4), round(k["frac_hbm"] or 0, 3))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_list_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_pupdate|k_spmv|k_r_update|k_final' -s 10 -c 5 -o $OUT/krylov_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_kry_$TAG.log 2>&1
ls $OUT | grep $TAG
