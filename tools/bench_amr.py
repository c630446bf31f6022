"""Round-2 tool (needs a GPU; not part of bench.py's contract line): multi-level time steps on the mesh the unmodified
reference grows for its run.sh case (config C5 of SURVEY 8(d) at the chosen levelMax).

  python tools/bench_amr.py [levelMax=9] [steps=10] [poisson_iters=10] [fast=1]
  python tools/bench_amr.py synthetic [base_level=9] [steps=10] [poisson_iters=10] [fast=1]     (config C5: 3 levels;
      under torchrun on N GPUs: the mesh distributed by block ranges, cup2d_amr_create_ranks)

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


def three_level_mesh(base_level, r1=0.22, r2=0.11, centre=(0.5, 0.5)):
    """Synthetic 3-level mesh for config C5 (3-level block-AMR, 16384^2 effective resolution at base_level 9):
    a unit square of 2^base_level blocks per side; the base blocks within r2 of `centre` are refined twice, those within r1
    (and, for the 2:1 rule, every 8-neighbour of a twice-refined one) once.  Refinement is uniform inside a base block, so level
    jumps occur between base blocks only and never exceed one.  Blocks follow the library's space-filling-curve order of the
    base blocks, children in Z order.  Returns int32 (level, i, j) triples."""
    from cup2d_b200 import lib as _l
    import ctypes as C
    n = 1 << base_level
    ij = np.empty(2 * n * n, dtype=np.int32)
    _l.check(_l.load_library().cup2d_block_order(1, 1, base_level, ij.ctypes.data_as(C.POINTER(C.c_int32))))
    ij = ij.reshape(-1, 2)
    x, y = (ij[:, 0] + 0.5) / n - centre[0], (ij[:, 1] + 0.5) / n - centre[1]
    d2 = x * x + y * y
    depth = np.zeros((n, n), dtype=np.int32)          # [j, i]
    depth[ij[:, 1], ij[:, 0]] = (d2 < r1 * r1).astype(np.int32) + (d2 < r2 * r2)
    two = np.pad(depth == 2, 1)
    near_two = np.zeros_like(depth, dtype=bool)
    for dj in range(3):
        for di in range(3):
            near_two |= two[dj:dj + n, di:di + n]
    depth = np.maximum(depth, near_two.astype(np.int32))
    dep = depth[ij[:, 1], ij[:, 0]]
    out = []
    for d in range(3):
        sel = ij[dep == d]
        m = 1 << d
        cj, ci = np.meshgrid(np.arange(m), np.arange(m), indexing="ij")     # Z order would interleave; row-major is fine for d <= 2
        t = np.empty((len(sel), m * m, 3), dtype=np.int32)
        t[:, :, 0] = base_level + d
        t[:, :, 1] = sel[:, 0:1] * m + ci.reshape(1, -1)
        t[:, :, 2] = sel[:, 1:2] * m + cj.reshape(1, -1)
        out.append((np.flatnonzero(dep == d), t))
    # stitch back in curve order of the base blocks
    counts = np.array([1, 4, 16])[dep]
    start = np.concatenate([[0], np.cumsum(counts)])
    blocks = np.empty((start[-1], 3), dtype=np.int32)
    for d, (pos, t) in enumerate(out):
        m2 = 1 << (2 * d)
        idx = (start[pos][:, None] + np.arange(m2)[None, :]).reshape(-1)
        blocks[idx] = t.reshape(-1, 3)
    return blocks


def seeded_fields(blocks, h0):
    """Taylor-Green velocity + a smooth pressure on the cell centres of a multi-level mesh"""
    h = h0 / (1 << blocks[:, 0]).astype(np.float64)
    ix = np.arange(8, dtype=np.float64) + 0.5
    X = (blocks[:, 1] * 8)[:, None, None] * h[:, None, None] + h[:, None, None] * ix[None, None, :] + 0 * ix[None, :, None]
    Y = (blocks[:, 2] * 8)[:, None, None] * h[:, None, None] + h[:, None, None] * ix[None, :, None] + 0 * ix[None, None, :]
    vel = np.stack([np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y), -np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y)], axis=-1)
    pres = (np.cos(2 * np.pi * X) * np.cos(2 * np.pi * Y))[..., None]
    return vel, pres


def run(sim, nb, blocks, steps, K, label, t_create, fast):
    import torch
    for _ in range(3):
        sim.step(cfl=0.5, max_iter=K)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps):
        sim.step(cfl=0.5, max_iter=K)
    torch.cuda.synchronize()
    wall = (time.time() - t0) / steps
    cells = nb * 64
    levels = sorted(set(blocks[:, 0].tolist()))
    print(json.dumps({"workload": label, "blocks": nb, "cells": cells, "levels": levels,
                      "fast": bool(fast), "poisson_iters": K, "ms_per_step_wall": wall * 1e3,
                      "Mcell_updates_per_s": cells * (2 + K) / wall / 1e6, "create_s": t_create}))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "synthetic":   # python tools/bench_amr.py synthetic [base_level=9] [steps] [K] [fast]
        base = int(sys.argv[2]) if len(sys.argv) > 2 else 9
        steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
        K = int(sys.argv[4]) if len(sys.argv) > 4 else 10
        fast = int(sys.argv[5]) if len(sys.argv) > 5 else 1
        from cup2d_b200.amr import AmrSimulation
        blocks = three_level_mesh(base)
        h0 = 1.0 / 8
        vel, pres = seeded_fields(blocks, h0)
        world, rank, lrank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
        t0 = time.time()
        if world > 1:   # under torchrun: the mesh distributed over the GPUs by block ranges (cup2d_amr_create_ranks)
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(lrank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))
            rb = [round(r * len(blocks) / world) for r in range(world + 1)]
            sim = AmrSimulation.distributed(blocks, 1, 1, h0, 1e-4, rank, rb, dist, device=lrank)
            vel, pres = vel[rb[rank]:rb[rank + 1]], pres[rb[rank]:rb[rank + 1]]
        else:
            sim = AmrSimulation(blocks, 1, 1, h0, 1e-4, device=lrank)
            sim.set_fast(bool(fast))
        sim.upload("vel", vel)
        sim.upload("pres", pres)
        sim.step(cfl=0.5, max_iter=1)          # first step builds the compact tables and the Poisson rows
        t_create = time.time() - t0
        label = f"synthetic 3-level mesh, base level {base} ({8 << (base + 2)}^2 effective), {world} GPU(s)"
        if rank == 0:
            run(sim, len(blocks), blocks, steps, K, label, t_create, fast)
        else:           # same calls, no report (every step synchronises the ranks inside the solve)
            for _ in range(3 + steps):
                sim.step(cfl=0.5, max_iter=K)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        sim.close()
        return
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
    run(sim, nb, blocks, steps, K, f"run.sh mesh, levelMax {lmax}", t_create, fast)
    sim.close()


if __name__ == "__main__":
    main()
