#!/bin/bash
# r02b (1 GPU): the graph-replayed step — parity suite incl. the new graph tests, the contract bench line with the step
# captured in a CUDA graph, the same with kernel-by-kernel launches, and the reference arm.
set -u
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
echo "== 1. pytest -m gpu"
(time timeout 900 python -m pytest tests -m gpu -x -q --durations=8) > $OUT/pytest_gpu_$TAG.log 2>&1
tail -6 $OUT/pytest_gpu_$TAG.log
echo "== 2. bench (graph)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo rc=$?
tail -c 400 $OUT/bench_$TAG.err
echo "== 3. bench (no graph)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-graph > $OUT/bench_nograph_$TAG.json 2> $OUT/bench_nograph_$TAG.err; echo rc=$?
python - <<PY
import json
for f in ("$OUT/bench_$TAG.json", "$OUT/bench_nograph_$TAG.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, (d.get("e2e") or {}).get("value"), d["timing_passes"])
        for k in d["kernels"]:
            print("   ", k["kernel"], k["launches_per_step"], round(k["ms_per_launch"], 4), round(k["frac_hbm"] or 0, 3))
    except Exception as e:
        print(f, "unreadable:", e)
PY
echo "== 4. reference arm, twice"
for i in 1 2; do
  (time timeout 900 python bench.py --impl reference --steps 5 --warmup 3) > $OUT/bench_reference_${i}_$TAG.json 2> $OUT/bench_reference_${i}_$TAG.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_reference_${i}_$TAG.json").read().strip().splitlines()[0])
    c = d["cpu_baseline"]
    print("reference arm run $i:", d["value"], c["cores"], c["value_min_median_max"], c["with_reference_gpu_solver"])
except Exception as e:
    print("reference arm unreadable:", e)
PY
  tail -3 $OUT/bench_reference_${i}_$TAG.err
done
