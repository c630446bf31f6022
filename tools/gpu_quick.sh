#!/bin/bash
# Short GPU-box pass: parity suite, bench line, ncu launch list and a full capture of the Krylov kernels.
set -u
TAG=${1:-r01k}
OUT=gpurun_out
mkdir -p $OUT
(time timeout 400 python -m pytest tests -m gpu -x -q --durations=8) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -14 $OUT/pytest_gpu_$TAG.log
timeout 150 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 400 $OUT/bench_$TAG.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "clocks")}, d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["poisson_iteration"])
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_pupdate|k_spmv|k_r_update|k_final' -s 10 -c 5 -o $OUT/krylov_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_kry_$TAG.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/ncu_list_$TAG.log 2>&1
ls -la $OUT | tail -8
