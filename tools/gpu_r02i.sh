#!/bin/bash
# r02i (2 GPUs): the final build of the round — parity suite (incl. the 2-GPU test), contract bench line (1 GPU, then the
# reference arm, then 2 GPUs), config C5 on 1 and 2 GPUs (multi-level advect kernel with the upwind core), config C3,
# ncu launch list of the bench command and full captures of the advect, Krylov and pressure kernels.
set -u
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== 1. pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=5) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -8 $OUT/pytest_gpu_$TAG.log
echo "== 2. bench, 1 GPU"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_1gpu_$TAG.json 2> $OUT/bench_1gpu_$TAG.err; echo rc=$?
tail -c 300 $OUT/bench_1gpu_$TAG.err
echo "== 3. reference arm"
(time timeout 900 python bench.py --impl reference --steps 5 --warmup 3) > $OUT/bench_reference_$TAG.json 2> $OUT/bench_reference_$TAG.err
cut -c1-200 $OUT/bench_reference_$TAG.json | head -1
echo "== 4. bench, 2 GPUs"
timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/bench_2gpu_$TAG.json 2> $OUT/bench_2gpu_$TAG.err; echo rc=$?
python - <<PY
import json
for n in (1, 2):
    try:
        d = json.loads(open(f"$OUT/bench_{n}gpu_$TAG.json").read().strip().splitlines()[-1])
        print(f"N={n}", {k: d.get(k) for k in ("value", "ms_per_step")}, d["e2e"]["value"], d.get("parity"), d["clocks"])
        for k in d.get("kernels", []):
            print("    ", k["kernel"], k["launches_per_step"], round(k["ms_per_launch"], 4), round(k["frac_hbm"] or 0, 3))
    except Exception as e:
        print(f"N={n} bench line unreadable:", e)
PY
tail -c 300 $OUT/bench_2gpu_$TAG.err
echo "== 5. config C5 (synthetic 3-level mesh, 16384^2 effective) on 1 and 2 GPUs; run.sh mesh"
timeout 400 python tools/bench_amr.py synthetic 9 10 10 1 > $OUT/bench_amr_c5_1gpu_$TAG.json 2> $OUT/bench_amr_c5_1gpu_$TAG.err
tail -c 500 $OUT/bench_amr_c5_1gpu_$TAG.json; tail -c 200 $OUT/bench_amr_c5_1gpu_$TAG.err
timeout 400 $TR --nproc-per-node 2 --master-port 29573 tools/bench_amr.py synthetic 9 10 10 1 > $OUT/bench_amr_c5_2gpu_$TAG.json 2> $OUT/bench_amr_c5_2gpu_$TAG.err
tail -c 500 $OUT/bench_amr_c5_2gpu_$TAG.json; tail -c 200 $OUT/bench_amr_c5_2gpu_$TAG.err
echo "== 6. config C3"
timeout 300 python tools/bench_c3.py 9 6 1e-6 > $OUT/bench_c3_$TAG.json 2> $OUT/bench_c3_$TAG.err; cut -c1-600 $OUT/bench_c3_$TAG.json
echo "== 7. ncu: launch list of the bench command, full captures"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_list_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:advect_stage -s 2 -c 2 -o $OUT/advect_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_adv_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_pupdate|k_spmv|k_r_update|k_final' -s 10 -c 5 -o $OUT/krylov_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_kry_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'pressure_rhs|pressure_correct|umax' -s 3 -c 3 -o $OUT/press_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-graph > $OUT/ncu_prs_$TAG.log 2>&1
ls -la $OUT | tail -10
