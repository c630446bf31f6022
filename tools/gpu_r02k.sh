#!/bin/bash
# r02k (1 GPU): row-mapped stencil kernels with one warp-wide load for the S/N edge rows, branch-free row access and in-chunk
# W/E ghosts (rows.cuh: CUP2D_ROWS_COOP): parity suite, bench A/B against the round-1 form, ncu of the Krylov + pressure kernels.
set -u
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
echo "== 1. pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=3) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -6 $OUT/pytest_gpu_$TAG.log
echo "== 2. bench default / nocoop"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo rc=$?
CUP2D_B200_LIB=$PWD/cup2d_b200/libcup2d_b200_nocoop.so timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $OUT/bench_nocoop_$TAG.json 2> $OUT/bench_nocoop_$TAG.err
python - <<PY | tee $OUT/variants_$TAG.jsonl
import json
for v in ("", "nocoop"):
    f = "$OUT/bench_" + (v + "_" if v else "") + "$TAG.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = {k["kernel"]: round(k["ms_per_launch"], 4) for k in d["kernels"]}
        print(json.dumps({"variant": v or "default (cooperative edge rows, in-chunk W/E ghosts)", "ms_per_step": d["ms_per_step"], "value": d["value"],
                          "poisson_iteration_ms": d["poisson_iteration"]["ms_per_iteration"], "kernels_ms": ks, "clocks": d["clocks"]}))
    except Exception as e:
        print(json.dumps({"variant": v or "default", "error": str(e)}))
PY
echo "== 3. ncu"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_pupdate|k_spmv|k_r_update|k_final' -s 10 -c 5 -o $OUT/krylov_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_kry_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'pressure_rhs|pressure_correct|k_init' -s 3 -c 3 -o $OUT/press_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_prs_$TAG.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_list_$TAG.log 2>&1
ls $OUT | grep $TAG | head -20
