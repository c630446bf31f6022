#!/bin/bash
# RECORD of the first GPU call of round 2 (r02a).  The three advect variants it A/B-tested (CUP2D_ADV_WARP_ROWS,
# CUP2D_WENO_CUBIC_RCP, CUP2D_WENO_LAZY_BETAS) no longer exist as switches: warp-local rows and the cubic reciprocal step
# are the default since the rewrite of the advect stage (r02e), the lazy indicators are subsumed by the upwind core.
# First GPU-box call of round 2, ≈ 20 min of box time on one GPU:
#   gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh r02a'
# 1. the validated parity suite; 2. the never-run multi-level device path (every test, no -x) + compute-sanitizer on its
# smallest cases; 3. the contract bench line; 4. A/B of the advect variants; 5. multi-level bench (baseline / fast) + ncu.
# Everything lands in gpurun_out/ with the tag in the name; nothing here changes clocks.
set -u
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
export OMP_NUM_THREADS=8

echo "== 1. validated suite"
(time timeout 500 python -m pytest tests -m gpu -x -q --durations=8) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -5 $OUT/pytest_gpu_$TAG.log

echo "== 2. multi-level device path, first contact"
(time timeout 900 python -m pytest tests/test_gpu_amr.py tests/test_gpu_parity.py -m gpu -q --durations=0 -p no:cacheprovider \
    -k "test_gpu_amr or host_pipeline or tiny_values") \
    > $OUT/pytest_amr_$TAG.log 2>&1
tail -25 $OUT/pytest_amr_$TAG.log
for tool in memcheck racecheck; do
    timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 \
        python -m pytest tests/test_gpu_amr.py -m gpu -q -x -k "advect or fast or full_step or bodies or adapt_tags or amr_dump" > $OUT/sanitizer_${tool}_amr_$TAG.log 2>&1
    echo "sanitizer $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT/sanitizer_${tool}_amr_$TAG.log | tail -3
done

echo "== 3. contract bench line"
timeout 300 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 300 $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "clocks")}, d["e2e"]["value"], d["poisson_iteration"])
except Exception as e:
    print("bench line unreadable:", e)
PY

echo "== 4. advect variants (same bench, kernel table only)"
ab() { # name, defines
    make -s -C cup2d_b200/csrc variant EXTRA="$2" > $OUT/variant_$1_$TAG.log 2>&1 || { echo "variant $1 failed to build"; return; }
    cp cup2d_b200/libcup2d_b200_variant.so cup2d_b200/libcup2d_b200_$1.so
    CUP2D_B200_LIB=$PWD/cup2d_b200/libcup2d_b200_$1.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "advect or rk2 or time_steps or full_step or operators or poisson" \
        > $OUT/variant_$1_pytest_$TAG.log 2>&1
    echo "variant $1 parity rc=$? $(tail -1 $OUT/variant_$1_pytest_$TAG.log)"
    CUP2D_B200_LIB=$PWD/cup2d_b200/libcup2d_b200_$1.so timeout 150 python bench.py --no-e2e --no-cpu-baseline --steps 10 \
        > $OUT/bench_variant_$1_$TAG.json 2>> $OUT/variant_$1_$TAG.log
}
ab base ""
ab warprows "-DCUP2D_ADV_WARP_ROWS=1"
ab cubic "-DCUP2D_WENO_CUBIC_RCP=1"
ab lazy "-DCUP2D_WENO_LAZY_BETAS=1"
ab all3 "-DCUP2D_ADV_WARP_ROWS=1 -DCUP2D_WENO_CUBIC_RCP=1 -DCUP2D_WENO_LAZY_BETAS=1"
ab spmv4 "-DSPMV_CTAS=4"        # SpMV at 4 CTAs/SM (64 registers, ~90 B of spills) instead of 3
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_variant_*_$TAG.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        adv = [k for k in d["kernels"] if k["kernel"].startswith("advect")][0]
        spmv = [k["ms_per_launch"] for k in d["kernels"] if k["kernel"].startswith("k_spmv")]
        print(f, "advect ms", round(adv["ms_per_launch"], 4), "spmv ms", [round(x, 4) for x in spmv], "step ms", round(d["ms_per_step"], 3),
              d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "unreadable:", e)
PY

echo "== 5. multi-level steps: baseline vs fast kernels, then a launch list and one full capture of the fast path"
for fast in 0 1; do
    timeout 300 python tools/bench_amr.py 9 10 10 $fast > $OUT/bench_amr_fast${fast}_$TAG.json 2> $OUT/bench_amr_fast${fast}_$TAG.err
    tail -c 600 $OUT/bench_amr_fast${fast}_$TAG.json; tail -c 300 $OUT/bench_amr_fast${fast}_$TAG.err
done
# config C5: synthetic 3-level mesh, 16384^2 effective (501 376 blocks, 32 M cells), fast kernels
timeout 400 python tools/bench_amr.py synthetic 9 10 10 1 > $OUT/bench_amr_c5_$TAG.json 2> $OUT/bench_amr_c5_$TAG.err
tail -c 600 $OUT/bench_amr_c5_$TAG.json; tail -c 300 $OUT/bench_amr_c5_$TAG.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_amr_$TAG.csv \
    python tools/bench_amr.py 9 2 10 1 > $OUT/ncu_list_amr_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'amr_.*fast|amr_fluxcorr' -s 20 -c 8 -o $OUT/amr_fast_$TAG -f \
    python tools/bench_amr.py 9 1 4 1 > $OUT/ncu_amr_$TAG.log 2>&1
ls -la $OUT | tail -30

echo "== 6. config C1 (the reference's run.sh case, levelMax 8, 40 steps, 20 Krylov iterations each): wall time of the whole program"
# ref_harness_gpu = the unmodified reference with its own cuda.cu solver; the other two run its hot path on cup2d_amr
for exe in ref_harness_gpu ref_harness_amrloop ref_harness_amrresident; do
    [ -x oracle/_ref/$exe ] || { echo "$exe not built"; continue; }
    s=$(date +%s%N)
    OMP_NUM_THREADS=$(nproc) CUP2D_B200_AMR_FAST=1 CUP2D_B200_MAX_ITER=20 timeout 600 oracle/_ref/$exe asteps 8 40 20 /tmp/c1_$exe.bin > /dev/null 2>&1
    echo "$exe rc=$? wall $(( ($(date +%s%N) - s) / 1000000 )) ms" | tee -a $OUT/c1_walltime_$TAG.txt
done
python - <<PY | tee -a $OUT/c1_walltime_$TAG.txt
import numpy as np, os
def last(path):
    a, i, rec = np.fromfile(path), 0, None
    while i < len(a):
        nb = int(a[i + 1]); rec = (nb, a[i + 2 + 3 * nb:i + 2 + 131 * nb]); i += 2 + 195 * nb
    return rec
try:
    ref = last("/tmp/c1_ref_harness_gpu.bin")
    for exe in ("ref_harness_amrloop", "ref_harness_amrresident"):
        got = last(f"/tmp/c1_{exe}.bin")
        print(exe, "blocks", got[0], "vs", ref[0], "vel rel diff after 40 steps",
              float(np.abs(got[1] - ref[1]).max() / np.abs(ref[1]).max()) if got[0] == ref[0] else "mesh differs")
except Exception as e:
    print("C1 comparison unavailable:", e)
PY
