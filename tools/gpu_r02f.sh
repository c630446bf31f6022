#!/bin/bash
# r02f (2 GPUs): advect stage with the uniform-sign upwind core and the cp.async loader — parity suite (incl. the 2-GPU test),
# contract bench line, A/B (bulk-copy loader / no fast path / 5 CTAs per SM), the same at 2 GPUs (parity + timing: the new
# advect kernel on partial tiles with halo slots, Krylov pushes without the per-lane fence), N-rank oracle checks,
# config C3 (tolerance-driven solves at 4096^2), small grids on one GPU (how much of the multi-GPU loss is kernel size),
# launch list + full ncu capture of the advect kernels.
set -u
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== 1. pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=5) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -8 $OUT/pytest_gpu_$TAG.log
echo "== 2. bench (default build), 1 GPU"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo rc=$?
tail -c 300 $OUT/bench_$TAG.err
echo "== 3. variants"
for v in tma nofast ctas5; do
  L=$PWD/cup2d_b200/libcup2d_b200_$v.so
  [ -f $L ] || { echo "$v not built"; continue; }
  CUP2D_B200_LIB=$L timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "advect or operators_vs_reference_golden or rk2_and_dt" 2>&1 | tail -1
  CUP2D_B200_LIB=$L timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/bench_${v}_$TAG.json 2> $OUT/bench_${v}_$TAG.err
done
python - <<PY | tee $OUT/variants_$TAG.jsonl
import json
for v in ("", "tma", "nofast", "ctas5"):
    f = "$OUT/bench_" + (v + "_" if v else "") + "$TAG.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        adv = [k for k in d["kernels"] if k["kernel"].startswith("advect")][0]
        print(json.dumps({"variant": v or "default", "advect_ms": adv["ms_per_launch"], "advect_frac_hbm": adv["frac_hbm"], "ms_per_step": d["ms_per_step"],
                          "value": d["value"], "clocks": d["clocks"]}))
    except Exception as e:
        print(json.dumps({"variant": v or "default", "error": str(e)}))
PY
echo "== 4. 2 GPUs: bench (parity inside) + oracle-based checks"
timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_2gpu_$TAG.json 2> $OUT/bench_2gpu_$TAG.err; echo rc=$?
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_2gpu_$TAG.json").read().strip().splitlines()[-1])
    print("N=2", {k: d.get(k) for k in ("value", "ms_per_step")}, d.get("parity"))
    for k in d.get("kernels", []):
        print("    ", k["kernel"], k["launches_per_step"], round(k["ms_per_launch"], 4), round(k["frac_hbm"] or 0, 3))
except Exception as e:
    print("bench line unreadable:", e)
PY
tail -c 300 $OUT/bench_2gpu_$TAG.err
timeout 400 $TR --nproc-per-node 2 --master-port 29571 tools/multi_gpu_check.py > $OUT/multi_gpu_check_2gpu_$TAG.jsonl 2> $OUT/multi_gpu_check_2gpu_$TAG.err
echo "rc=$?"; cat $OUT/multi_gpu_check_2gpu_$TAG.jsonl; tail -c 300 $OUT/multi_gpu_check_2gpu_$TAG.err
echo "== 5. config C3: 4096^2, solves stop at 1e-6"
timeout 300 python tools/bench_c3.py 9 6 1e-6 > $OUT/bench_c3_$TAG.json 2> $OUT/bench_c3_$TAG.err; echo rc=$?
cut -c1-700 $OUT/bench_c3_$TAG.json; tail -c 300 $OUT/bench_c3_$TAG.err
echo "== 6. small grids on one GPU (level 9 = 1/4, level 8 = 1/16 of the 8192^2 cells)"
for lv in 9 8; do
  timeout 200 python bench.py --level $lv --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > $OUT/bench_level${lv}_$TAG.json 2> $OUT/bench_level${lv}_$TAG.err
  python -c "
import json; d=json.loads(open('$OUT/bench_level${lv}_$TAG.json').read().strip().splitlines()[-1]); print('level $lv', d['ms_per_step'], d['value'], d['poisson_iteration']['ms_per_iteration'])"
done
echo "== 7. ncu launch list + full capture of the advect kernels (default build)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_list_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:advect_stage -s 2 -c 2 -o $OUT/advect_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_adv_$TAG.log 2>&1
tail -2 $OUT/ncu_adv_$TAG.log
ls -la $OUT | tail -12
