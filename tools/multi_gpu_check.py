"""Run under torchrun on N GPUs of one box: N-rank run vs the numpy oracle (small grid) and timing of
the N-rank Poisson iteration / step (large grid).  Rank 0 prints one JSON line per check."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist

import cup2d_b200
import cup2d_oracle as orc  # checker only

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dist.init_process_group("nccl", device_id=torch.device("cuda", lrank))


def gather_field(sim, name, dim):
    """every rank contributes its blocks; returns global arrays on all ranks"""
    flat = sim.download_blocks(name)
    parts = [None] * world
    dist.all_gather_object(parts, flat)
    allb = np.concatenate(parts)
    return cup2d_b200.from_blocks(allb, sim.order, sim.nbx, sim.nby, dim)


# ---- parity at 256^2: 2 steps, 10 iterations each, vs the oracle ----
L = 5
N = 8 << L
x = (np.arange(N) + 0.5) / N
X, Y = np.meshgrid(x, x)
rng = np.random.default_rng(5)
u = np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y) + 0.05 * rng.uniform(-1, 1, (N, N))
v = -np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y) + 0.05 * rng.uniform(-1, 1, (N, N))
p = np.cos(2 * np.pi * X) * np.cos(2 * np.pi * Y)
sim = cup2d_b200.Simulation(L, nu=1e-3, cfl=0.5, device=lrank, rank=rank, nranks=world)
sim.attach_peers(dist)
sim.upload("vel", u, v)
sim.upload("pres", p)
ru, rv, rp = u, v, p
worst = 0.0
for s in range(2):
    dt, it, err = sim.step(max_iter=10)
    ref = orc.step(ru, rv, rp, 1e-3, 0.5, kiter=10)
    gu, gv = gather_field(sim, "vel", 2)
    gp = gather_field(sim, "pres", 1)
    worst = max(worst, abs(dt - ref["dt"]) / ref["dt"], np.abs(gu - ref["u"]).max(), np.abs(gv - ref["v"]).max(),
                np.abs(gp - ref["p"]).max())
    ru, rv, rp = ref["u"], ref["v"], ref["p"]
if rank == 0:
    print(json.dumps({"check": "parity_256", "ranks": world, "Linf_u_v_p_dt": worst, "nhalo": int(sim.lib.cup2d_nblocks_halo(sim._h))}), flush=True)
assert worst < 1e-8, worst
sim.close()
dist.barrier()
dist.destroy_process_group()
