#!/bin/bash
# r02h (1 GPU): row-mapped stencil kernels reading the neighbour blocks of their own chunk from the scratch (rows.cuh):
# parity suite, contract bench line, A/B against the build without it, full ncu capture of the Krylov and pressure kernels.
set -u
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
echo "== 1. pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=5) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -8 $OUT/pytest_gpu_$TAG.log
echo "== 2. bench (default build)"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo rc=$?
tail -c 300 $OUT/bench_$TAG.err
for v in noinchunk; do
  L=$PWD/cup2d_b200/libcup2d_b200_$v.so
  CUP2D_B200_LIB=$L timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
done
python - <<PY | tee $OUT/variants_$TAG.jsonl
import json
for v in ("", "noinchunk"):
    f = "$OUT/bench_" + (v + "_" if v else "") + "$TAG.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        ks = {k["kernel"]: round(k["ms_per_launch"], 4) for k in d["kernels"]}
        print(json.dumps({"variant": v or "default (in-chunk neighbours from the scratch)", "ms_per_step": d["ms_per_step"], "value": d["value"],
                          "poisson_iteration_ms": d["poisson_iteration"]["ms_per_iteration"], "kernels_ms": ks, "clocks": d["clocks"]}))
    except Exception as e:
        print(json.dumps({"variant": v or "default", "error": str(e)}))
PY
echo "== 3. reference arm"
(time timeout 900 python bench.py --impl reference --steps 5 --warmup 3) > $OUT/bench_reference_$TAG.json 2> $OUT/bench_reference_$TAG.err
cut -c1-300 $OUT/bench_reference_$TAG.json | head -2; tail -3 $OUT/bench_reference_$TAG.err
echo "== 4. ncu: launch list, Krylov + pressure kernels"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_list_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_pupdate|k_spmv|k_r_update|k_final' -s 10 -c 5 -o $OUT/krylov_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_kry_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'pressure_rhs|pressure_correct|umax' -s 3 -c 3 -o $OUT/press_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_prs_$TAG.log 2>&1
ls -la $OUT | tail -8
