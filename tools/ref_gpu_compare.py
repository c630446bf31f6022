"""GPU box only: pin the oracle chain against the REAL reference GPU solver.
  ref_harness_gpu  = unmodified main.cpp + unmodified cuda.cu   (1000 BiCGSTAB iterations per step, cuda.cu:438)
  ref_harness      = unmodified main.cpp + CPU restatement of cuda.cu (oracle/ref_spmat_cpu.cpp)
  ref_harness_b200 = unmodified main.cpp + dropin/local_spmat_adapter.cpp + libcup2d_b200.so (the drop-in)
  cup2d_b200       = this library end to end (its own time step), max_iter=1000
All three run the same 2 time steps from the same seeded field; prints the pairwise L-inf differences."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cup2d_b200

L = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N = 8 << L
nsteps = 2
x = (np.arange(N) + 0.5) / N
X, Y = np.meshgrid(x, x)
rng = np.random.default_rng(42)
u = np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y) + 0.05 * rng.uniform(-1, 1, (N, N))
v = -np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y) + 0.05 * rng.uniform(-1, 1, (N, N))
p = np.cos(2 * np.pi * X) * np.cos(2 * np.pi * Y)
z = np.zeros((N, N))
res = {}
with tempfile.TemporaryDirectory() as tmp:
    fin = os.path.join(tmp, "in.bin")
    np.concatenate([a.ravel() for a in (u, v, p, z, z, z)]).tofile(fin)
    for name in ("ref_harness_gpu", "ref_harness", "ref_harness_b200"):
        fout = os.path.join(tmp, name + ".bin")
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", name), "steps", str(L), "1e-3", "0.5", str(nsteps), "1000", fin, fout],
                       check=True, stderr=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="16"))
        raw = np.fromfile(fout).reshape(nsteps, 1 + 5 * N * N)
        res[name] = (raw[:, 0].copy(), raw[:, 1:].reshape(nsteps, 5, N, N))
sim = cup2d_b200.Simulation(L, nu=1e-3, cfl=0.5)
sim.upload("vel", u, v)
sim.upload("pres", p)
mine = []
for s in range(nsteps):
    dt, it, err = sim.step(max_iter=1000)
    gu, gv = sim.download("vel")
    mine.append((dt, gu, gv, sim.download("pres"), it, err))
out = {"L": L, "N": N}
for s in range(nsteps):
    g, c = res["ref_harness_gpu"][1][s], res["ref_harness"][1][s]
    out[f"step{s}"] = {
        "real_cuda_cu_vs_cpu_restatement": {"u": float(np.abs(g[0] - c[0]).max()), "v": float(np.abs(g[1] - c[1]).max()), "p": float(np.abs(g[2] - c[2]).max())},
        "real_cuda_cu_vs_reference_driver_with_cup2d_adapter": (lambda a: {"u": float(np.abs(g[0] - a[0]).max()), "v": float(np.abs(g[1] - a[1]).max()),
                                                                          "p": float(np.abs(g[2] - a[2]).max())})(res["ref_harness_b200"][1][s]),
        "real_cuda_cu_vs_cup2d_b200": {"u": float(np.abs(g[0] - mine[s][1]).max()), "v": float(np.abs(g[1] - mine[s][2]).max()),
                                        "p": float(np.abs(g[2] - mine[s][3]).max()), "dt": float(abs(res["ref_harness_gpu"][0][s] - mine[s][0]))},
        "cup2d_iters": mine[s][4], "cup2d_err": mine[s][5],
    }
print(json.dumps(out))
