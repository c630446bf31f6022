#!/bin/bash
# r02g (1 GPU): advect stage, A/B of {uniform-sign upwind core on/off} x {pass-specialised code / one copy}; cp.async loader.
set -u
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
K="advect or operators_vs_reference_golden or rk2_and_dt or time_steps_vs or amr_fast or amr_advect_diffuse or synthetic_three_level"
echo "== 1. parity (default build)"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_amr.py -m gpu -x -q -k "$K" 2>&1 | tail -2
echo "== 2. bench"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo rc=$?
tail -c 300 $OUT/bench_$TAG.err
VARS="fast_onecopy nofast nofast_onecopy"
for v in $VARS; do
  L=$PWD/cup2d_b200/libcup2d_b200_$v.so
  [ -f $L ] || { echo "$v not built"; continue; }
  CUP2D_B200_LIB=$L timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "advect or operators_vs_reference_golden or rk2_and_dt" 2>&1 | tail -1
  CUP2D_B200_LIB=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
done
python - <<PY | tee $OUT/variants_$TAG.jsonl
import json
for v in ("", "fast_onecopy", "nofast", "nofast_onecopy"):
    f = "$OUT/bench_" + (v + "_" if v else "") + "$TAG.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        adv = [k for k in d["kernels"] if k["kernel"].startswith("advect")][0]
        print(json.dumps({"variant": v or "default (fast path, specialised passes)", "advect_ms": adv["ms_per_launch"], "advect_frac_hbm": adv["frac_hbm"],
                          "ms_per_step": d["ms_per_step"], "value": d["value"], "clocks": d["clocks"]}))
    except Exception as e:
        print(json.dumps({"variant": v or "default", "error": str(e)}))
PY
echo "== 3. full ncu capture of the advect kernels (default build)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:advect_stage -s 2 -c 2 -o $OUT/advect_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_adv_$TAG.log 2>&1
tail -2 $OUT/ncu_adv_$TAG.log
