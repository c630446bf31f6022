#!/bin/bash
# r02e (1 GPU): the rewritten advect stage (per-row bulk copies straight into the interleaved plane, third-difference WENO
# algebra, warp-local passes): parity suite, contract bench line, A/B of the loader / residency variants, launch list,
# full ncu capture of the advect kernels.
set -u
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
echo "== 1. pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=5) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -8 $OUT/pytest_gpu_$TAG.log
echo "== 2. bench (default build)"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo rc=$?
tail -c 300 $OUT/bench_$TAG.err
echo "== 3. variants (parity of the advect tests, then a short bench)"
for v in ctas5 ldgsts ldgsts5; do
  L=$PWD/cup2d_b200/libcup2d_b200_$v.so
  [ -f $L ] || { echo "$v not built"; continue; }
  CUP2D_B200_LIB=$L timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "advect_stage_vs_oracle or operators_vs_reference_golden or rk2_and_dt" 2>&1 | tail -1
  CUP2D_B200_LIB=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
done
python - <<PY | tee $OUT/variants_$TAG.jsonl
import json
for v in ("", "ctas5", "ldgsts", "ldgsts5"):
    f = "$OUT/bench_" + (v + "_" if v else "") + "$TAG.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        adv = [k for k in d["kernels"] if k["kernel"].startswith("advect")][0]
        print(json.dumps({"variant": v or "default", "advect_ms": adv["ms_per_launch"], "advect_frac_hbm": adv["frac_hbm"], "ms_per_step": d["ms_per_step"],
                          "value": d["value"], "clocks": d["clocks"]}))
    except Exception as e:
        print(json.dumps({"variant": v or "default", "error": str(e)}))
PY
echo "== 4. ncu launch list + full capture of the advect kernels (default build)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_list_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:advect_stage -s 2 -c 2 -o $OUT/advect_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_adv_$TAG.log 2>&1
tail -3 $OUT/ncu_adv_$TAG.log
ls -la $OUT | tail -12
