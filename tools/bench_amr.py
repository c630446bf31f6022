"""Round-2 tool (needs a GPU; not part of bench.py's contract line): multi-level time steps on the mesh the unmodified
reference grows for its run.sh case (config C5 of SURVEY 8(d) at the chosen levelMax).

  python tools/bench_amr.py [levelMax=9] [steps=10] [poisson_iters=10] [fast=1]

The mesh comes from oracle/_ref/ref_harness amrlab (which travels to the GPU box prebuilt), the fields are its seeded ones,
bodies are left out (u_def = 0, chi = 0).  Prints one JSON line: blocks, cells, levels, ms per step with CUDA events on the
context's stream, Mcell-updates/s = cells * (2 + K) / time, and the host-side plan build time."""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 9
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    fast = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    import torch
    from cup2d_b200.amr import AmrSimulation
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "amr.bin")
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_harness"), "amrlab", str(lmax), "3", out], check=True,
                       stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="8"))
        a = np.fromfile(out)
    i, rec = 0, {}
    while i < len(a):
        tag, n = int(a[i]), int(a[i + 1])
        rec[tag] = a[i + 2:i + 2 + n]
        i += 2 + n
    nu, _, h0, bpdx, bpdy, _ = rec[10]
    blocks = rec[11].reshape(-1, 3).astype(np.int32)
    nb = len(blocks)
    t0 = time.time()
    sim = AmrSimulation(blocks, int(bpdx), int(bpdy), h0, nu)
    t_create = time.time() - t0
    sim.upload("vel", rec[12].reshape(nb, 8, 8, 2))
    sim.upload("pres", rec[13].reshape(nb, 8, 8, 1))
    sim.set_fast(bool(fast))
    for _ in range(3):
        sim.step(cfl=0.5, max_iter=K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.time()
    for _ in range(steps):
        sim.step(cfl=0.5, max_iter=K)
    torch.cuda.synchronize()
    wall = (time.time() - t0) / steps
    cells = nb * 64
    levels = sorted(set(blocks[:, 0].tolist()))
    print(json.dumps({"workload": f"run.sh mesh, levelMax {lmax}", "blocks": nb, "cells": cells, "levels": levels,
                      "fast": bool(fast), "poisson_iters": K, "ms_per_step_wall": wall * 1e3,
                      "Mcell_updates_per_s": cells * (2 + K) / wall / 1e6, "create_s": t_create}))
    sim.close()


if __name__ == "__main__":
    main()
