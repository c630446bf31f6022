"""GPU-box sweep of the advect-stage tuning knobs (CUP2D_ADV_UNROLL x CUP2D_ADV_CONSTMEM x CUP2D_ADV_ONECOPY).
Each variant runs in its own process (the knobs are read once); reports ms per stage launch at 8192^2
for both stage kinds and the parity error against the numpy oracle at 256^2."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, os, json
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, cup2d_b200, cup2d_oracle as orc
def fields(N, seed):
    rng = np.random.default_rng(seed)
    x = (np.arange(N) + 0.5) / N
    X, Y = np.meshgrid(x, x)
    u = np.sin(2*np.pi*X)*np.cos(2*np.pi*Y) + 0.05*rng.uniform(-1, 1, (N, N))
    v = -np.cos(2*np.pi*X)*np.sin(2*np.pi*Y) + 0.05*rng.uniform(-1, 1, (N, N))
    return u, v
# parity at 256^2 (TG+noise) and 64^2 random
errs = []
for L, kind in ((5, "tg"), (3, "rand")):
    N = 8 << L
    u, v = fields(N, 3) if kind == "tg" else tuple(np.random.default_rng(9).uniform(-1, 1, (2, N, N)))
    sim = cup2d_b200.Simulation(L, nu=1e-3)
    sim.upload("vel", u, v)
    dt = 0.3 / N
    sim.advect_diffuse_rhs(dt)
    au, av = sim.download("tmpV")
    ru, rv = orc.advect_diffuse(u, v, 1.0/N, 1e-3, dt)
    errs.append(max(np.abs(au-ru).max(), np.abs(av-rv).max()) / max(np.abs(ru).max(), np.abs(rv).max()))
    sim.close()
L = int(os.environ.get("TUNE_LEVEL", "10")); N = 8 << L
sim = cup2d_b200.Simulation(L, nu=1e-4)
order = sim.local_order
bi = order[:, 0].astype(np.float64)[:, None, None]; bj = order[:, 1].astype(np.float64)[:, None, None]
X = (bi*8 + np.arange(8.)[None, None, :] + 0.5)/N; Y = (bj*8 + np.arange(8.)[None, :, None] + 0.5)/N
blk = np.empty((len(order), 8, 8, 2))
blk[..., 0] = np.sin(2*np.pi*X)*np.cos(2*np.pi*Y); blk[..., 1] = -np.cos(2*np.pi*X)*np.sin(2*np.pi*Y)
sim.upload_blocks("vel", blk.reshape(-1)); sim.upload_blocks("vold", blk.reshape(-1))
dt = 0.25 / N
res = {}
for name, args in (("stage1", ("vel", "vel", "tmpV", 0.5)), ("stage2", ("tmpV", "vold", "vel", 1.0))):
    for _ in range(3): sim.advect_diffuse_stage(*args, dt)
    sim.profile(True)
    for _ in range(10): sim.advect_diffuse_stage(*args, dt)
    p = sim.profile_read(); sim.profile(False)
    ms, n = p["advect_stage_kernel"]; res[name] = ms / n
    sim.upload_blocks("vel", blk.reshape(-1))
print(json.dumps({"unroll": os.environ.get("CUP2D_ADV_UNROLL"), "constmem": os.environ.get("CUP2D_ADV_CONSTMEM"), "onecopy": os.environ.get("CUP2D_ADV_ONECOPY"),
                  "ms_stage1": res["stage1"], "ms_stage2": res["stage2"], "rel_err": errs,
                  "Gcell_s_stage1": N*N/res["stage1"]/1e6, "Gcell_s_stage2": N*N/res["stage2"]/1e6}))
'''
out = []
for unr in (5, 10):
    for cm in (0, 1):
        for one in (0, 1):
            env = dict(os.environ, CUP2D_ADV_UNROLL=str(unr), CUP2D_ADV_CONSTMEM=str(cm), CUP2D_ADV_ONECOPY=str(one))
            r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"unroll": unr, "cm": cm, "one": one, "error": r.stderr[-400:]})
            print(line, flush=True)
            out.append(line)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "tune_advect.jsonl"), "w").write("\n".join(out) + "\n")
