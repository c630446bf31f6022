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

# ---- regridding criterion (vorticity part), dump files and the penalisation phase on N ranks ----
gu, gv = gather_field(sim, "vel", 2)
linf = sim.vorticity_tag()
w = orc.vorticity(gu, gv, 1.0 / N)
want = orc.block_linf(w)[sim.local_order[:, 1], sim.local_order[:, 0]]
gw = gather_field(sim, "tmp", 1)
tag_err = max(float(np.abs(linf - want).max() / np.abs(want).max()), float(np.abs(gw - w).max() / np.abs(w).max()))
pref = "/tmp/cup2d_mgpu_dump"
if rank == 0:
    for ext in (".xyz.raw", ".attr.raw", ".xdmf2"):
        if os.path.exists(pref + ext):
            os.remove(pref + ext)
dist.barrier()
sim.dump(0.25, pref)
dist.barrier()
xyz, attr = orc.dump_arrays(gu, gv, sim.order, 1.0 / 8, L)
dump_ok = (open(pref + ".xyz.raw", "rb").read() == xyz.tobytes() and open(pref + ".attr.raw", "rb").read() == attr.tobytes()
           and open(pref + ".xdmf2").read() == orc.dump_xdmf(0.25, len(sim.order) * 64, "cup2d_mgpu_dump.xyz.raw", "cup2d_mgpu_dump.attr.raw"))
if rank == 0:
    print(json.dumps({"check": "tags_dump", "ranks": world, "tag_rel_err": tag_err, "dump_identical": bool(dump_ok)}), flush=True)
assert tag_err < 1e-12 and dump_ok
sim.close()

# ---- adapt()'s full criterion (vorticity + body proximity, diagonal neighbours across ranks) vs the reference golden
chi_ok = True
for name in ("L3_finest", "L3_coarser"):
    d = np.load(os.path.join(ROOT, "tests", "golden", f"tags_{name}.npz"))
    rtol_ = float(d["rtol"])
    sim = cup2d_b200.Simulation(int(d["L"]), device=lrank, rank=rank, nranks=world)
    sim.attach_peers(dist)
    sim.upload("vel", d["u"], d["v"])
    sim.upload("chi", d["chi"])
    linf = sim.adapt_tags(rtol_, int(d["offset"]))
    got = gather_field(sim, "tmp", 1)
    want = orc.block_linf(d["tags"])[sim.local_order[:, 1], sim.local_order[:, 0]]
    chi_ok = (chi_ok and np.array_equal(got == 2 * rtol_, d["tags"] == 2 * rtol_)
              and np.abs(got - d["tags"]).max() < 1e-12 * np.abs(d["tags"]).max()
              and np.abs(linf - want).max() < 1e-12 * np.abs(want).max())
    sim.close()
if rank == 0:
    print(json.dumps({"check": "adapt_tags_chi", "ranks": world, "same_blocks_flagged": bool(chi_ok)}), flush=True)
assert chi_ok

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import load_penal  # noqa: E402

PL, steps = load_penal(os.path.join(ROOT, "tests", "golden", "penal_L4.npz"))
sim = cup2d_b200.Simulation(PL, device=lrank, rank=rank, nranks=world)
sim.attach_peers(dist)
pen_err, pen_exact = 0.0, True
for st in steps:
    sim.upload("vel", st["u0"], st["v0"])
    sim.upload("chi", st["chi"])
    for k, sh in enumerate(st["shapes"]):
        mine = (sh["ids"] >= sim.gbegin) & (sh["ids"] < sim.gend)
        sim.shape_set(k, sh["ids"][mine] - sim.gbegin, sh["X"][mine], sh["udef"][mine])
    for k, sh in enumerate(st["shapes"]):
        Q = sim.shape_integrals(k, st["lam"], st["dt"], sh["cx"], sh["cy"])  # global sums on every rank
        pen_err = max(pen_err, float(np.abs(Q - sh["Q"]).max() / np.abs(sh["Q"]).max()))
    for k, sh in enumerate(st["shapes"]):
        sim.penalize(k, st["lam"], st["dt"], sh["cx"], sh["cy"], sh["u"], sh["v"], sh["omega"])
    sim.udef_assemble()
    gu, gv = gather_field(sim, "vel", 2)
    du, dv = gather_field(sim, "tmpV", 2)
    pen_exact = pen_exact and all(np.array_equal(a, b) for a, b in ((gu, st["u1"]), (gv, st["v1"]), (du, st["udu"]), (dv, st["udv"])))
if rank == 0:
    print(json.dumps({"check": "penalisation", "ranks": world, "integrals_rel_err": pen_err, "blend_udef_bit_exact": bool(pen_exact)}), flush=True)
assert pen_err < 1e-12 and pen_exact
sim.close()
# ---- the Poisson matrix of the reference's 7-level run.sh mesh (neighbour table + coarse-fine rows from the plan),
#      distributed over the ranks by block ranges, vs the same solve on one GPU (rank 0) ----
from cup2d_b200.amr import AmrPlan, DistributedPoisson
g = np.load(os.path.join(ROOT, "tests", "golden", "amrlab_lmax8.npz"))
blocks = np.ascontiguousarray(g["blocks"], dtype=np.int32)
nb = len(blocks)
plan = AmrPlan(blocks, int(g["bpdx"]), int(g["bpdy"]))
nbr, rows, rowptr, col, val = plan.poisson()
rb = [round(r * nb / world) for r in range(world + 1)]
rng = np.random.default_rng(7)
b, x0 = rng.uniform(-1, 1, (nb, 64)), rng.uniform(-0.1, 0.1, (nb, 64))
dp = DistributedPoisson(nbr, rows, rowptr, col, val, rb, rank, device=lrank)
dp.attach_peers(dist)
xr, it, err = dp.solve(b[rb[rank]:rb[rank + 1]], x0[rb[rank]:rb[rank + 1]], max_iter=12)
parts = [None] * world
dist.all_gather_object(parts, xr)
dist.barrier()
dp.close()
if rank == 0:
    one = DistributedPoisson(nbr, rows, rowptr, col, val, [0, nb], 0, device=lrank)
    one.attach_peers()
    xs, its, errs = one.solve(b, x0, max_iter=12)
    one.close()
    amr_err = float(np.abs(np.concatenate(parts) - xs).max() / np.abs(xs).max())
    print(json.dumps({"check": "amr_poisson_ranks", "ranks": world, "iters": it, "rel_err_vs_one_gpu": amr_err, "err": err, "err_one_gpu": errs}), flush=True)
    assert it == its == 12 and amr_err < 1e-9

# ---- multi-level time steps on several GPUs (cup2d_amr_set_ranks: operators replicated, Poisson solve distributed) vs the
#      same steps on one GPU ----
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_amr
from cup2d_b200.amr import AmrSimulation
mblocks = bench_amr.three_level_mesh(4, r1=0.3, r2=0.15, centre=(0.45, 0.55))
mnb, mh0 = len(mblocks), 1 / 8
mvel, mpres = bench_amr.seeded_fields(mblocks, mh0)


def amr_steps(sim):
    out = []
    sim.set_fast(True)
    sim.upload("vel", mvel)
    sim.upload("pres", mpres)
    for _ in range(2):
        info = sim.step(cfl=0.5, max_iter=10)
        out.append((info, sim.download("vel"), sim.download("pres")))
    return out


asim = AmrSimulation(mblocks, 1, 1, mh0, 1e-3, device=lrank)
asim.set_ranks(rank, [round(r * mnb / world) for r in range(world + 1)], dist)
got = amr_steps(asim)
dist.barrier()
asim.close()
allgot = [None] * world
dist.all_gather_object(allgot, [(g[0], float(np.abs(g[1]).sum()), float(np.abs(g[2]).sum())) for g in got])
if rank == 0:
    one = AmrSimulation(mblocks, 1, 1, mh0, 1e-3, device=lrank)
    ref = amr_steps(one)
    one.close()
    amr_step_err = max(max(float(np.abs(g[1] - r[1]).max() / np.abs(r[1]).max()), float(np.abs(g[2] - r[2]).max() / np.abs(r[2]).max()))
                       for g, r in zip(got, ref))
    same = all(a == allgot[0] for a in allgot)
    print(json.dumps({"check": "amr_step_ranks", "ranks": world, "blocks": mnb, "rel_err_vs_one_gpu": amr_step_err,
                      "ranks_identical": bool(same)}), flush=True)
    assert amr_step_err < 1e-9 and same

# ---- the same steps with the mesh itself DISTRIBUTED over the GPUs (cup2d_amr_create_ranks) ----
mrb = [round(r * mnb / world) for r in range(world + 1)]
dsim = AmrSimulation.distributed(mblocks, 1, 1, mh0, 1e-3, rank, mrb, dist, device=lrank)
sl = slice(mrb[rank], mrb[rank + 1])
dsim.upload("vel", mvel[sl])
dsim.upload("pres", mpres[sl])
dgot = []
for _ in range(2):
    info = dsim.step(cfl=0.5, max_iter=10)
    dgot.append((info, dsim.download("vel"), dsim.download("pres")))
dist.barrier()
dsim.close()
parts = [None] * world
dist.all_gather_object(parts, dgot)
if rank == 0:
    derr = 0.0
    for s_ in range(2):
        v = np.concatenate([parts[r][s_][1] for r in range(world)])
        p_ = np.concatenate([parts[r][s_][2] for r in range(world)])
        derr = max(derr, float(np.abs(v - ref[s_][1]).max() / np.abs(ref[s_][1]).max()), float(np.abs(p_ - ref[s_][2]).max() / np.abs(ref[s_][2]).max()))
    print(json.dumps({"check": "amr_distributed_ranks", "ranks": world, "blocks": mnb, "rel_err_vs_one_gpu": derr}), flush=True)
    assert derr < 1e-9
dist.barrier()
dist.destroy_process_group()
