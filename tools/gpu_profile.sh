#!/bin/bash
# Run on the GPU box (via gpurun): bench line + ncu launch list + ncu full captures of the top kernels.
# Outputs land in gpurun_out/ (scratch); summaries worth judging are copied to profiles/ by hand.
set -u
TAG=${1:-r01i}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 600 $OUT/bench_$TAG.err
python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2>> $OUT/bench_$TAG.err
# every launch with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_$TAG.log 2>&1
# full captures: the advect stage (both variants), and the five Krylov kernels
ncu --set full --clock-control none --import-source on -k regex:advect_stage -s 2 -c 2 -o $OUT/advect_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_adv_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_pupdate|k_spmv|k_r_update|k_final' -s 10 -c 5 -o $OUT/krylov_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_kry_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'pressure_rhs|pressure_correct|umax' -s 3 -c 3 -o $OUT/press_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_prs_$TAG.log 2>&1
timeout 300 python tools/ref_gpu_compare_amr.py 14 2>/dev/null | tail -1 > $OUT/amr_compare_$TAG.json
(time timeout 600 python -m pytest tests -m gpu -x -q) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -5 $OUT/pytest_gpu_$TAG.log
ls -la $OUT | tail -20
