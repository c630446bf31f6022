"""GPU box: A/B the run-time tuning knobs through bench.py (each combination in its own process)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
combos = [dict(CUP2D_ADV_SPLIT="0"), dict(CUP2D_ADV_SPLIT="1")]
out = []
for c in combos:
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-e2e", "--steps", "5"],
                       env=dict(os.environ, **c), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    try:
        d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        row = {"knobs": c, "ms_per_step": d["ms_per_step"], **{k["kernel"]: round(k["ms_per_launch"], 4) for k in d["kernels"]}}
    except Exception as e:
        row = {"knobs": c, "error": str(e)}
    print(json.dumps(row), flush=True)
    out.append(json.dumps(row))
open(os.path.join(ROOT, "gpurun_out", "ab_test.jsonl"), "w").write("\n".join(out) + "\n")
