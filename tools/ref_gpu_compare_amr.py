"""GPU box only: the reference's own run.sh case (2 fish, block-AMR, levels 5..7) driven by the UNMODIFIED
main.cpp, once with the reference's cuda.cu and once with dropin/local_spmat_adapter.cpp + libcup2d_b200.so
(general rows through the CSR side table).  Compares, step by step, the grid (level,i,j of every block), the
Poisson right-hand side and the returned solution."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def read_records(path):
    raw = open(path, "rb").read()
    off, recs = 0, []
    while off < len(raw):
        nrows, nblk = np.frombuffer(raw, dtype=np.int64, count=2, offset=off)
        off += 16
        dt = np.frombuffer(raw, dtype=np.float64, count=1, offset=off)[0]
        off += 8
        b = np.frombuffer(raw, dtype=np.float64, count=nrows, offset=off)
        off += 8 * nrows
        x = np.frombuffer(raw, dtype=np.float64, count=nrows, offset=off)
        off += 8 * nrows
        lij = np.frombuffer(raw, dtype=np.int32, count=3 * nblk, offset=off).reshape(-1, 3)
        off += 12 * nblk
        recs.append(dict(dt=dt, b=b, x=x, lij=lij))
    return recs


def compare(nsteps=12, level_max=8, names=("ref_harness_gpu", "ref_harness_b200"), timeout=None):
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        recs = []
        for i, n in enumerate(names):   # the same name twice = two runs of one program (run-to-run variation)
            f = os.path.join(tmp, f"{i}_{n}.bin")
            subprocess.run([os.path.join(ROOT, "oracle", "_ref", n), "amr", str(level_max), str(nsteps), "1000", f],
                           check=True, stderr=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="8"), timeout=timeout)
            recs.append(read_records(f))
    a, b = recs
    rows = []
    for s in range(min(len(a), len(b))):
        same_grid = a[s]["lij"].shape == b[s]["lij"].shape and bool((a[s]["lij"] == b[s]["lij"]).all())
        row = {"step": s, "blocks": int(len(a[s]["lij"])), "levels": sorted(set(a[s]["lij"][:, 0].tolist())),
               "same_grid": same_grid, "dt_diff": float(abs(a[s]["dt"] - b[s]["dt"]))}
        if same_grid:
            # The system is singular (pure Neumann): the constant mode of x is arbitrary and, with tolerance 0 and
            # 1000 iterations on a residual at round-off level, drifts freely in BOTH solvers.  The driver removes
            # the volume-weighted mean right after the solve (main.cpp:7126-7148), so that is what is compared.
            w = np.repeat(4.0 ** (-a[s]["lij"][:, 0].astype(np.float64)), 64)   # h^2 per cell
            xa = a[s]["x"] - (w * a[s]["x"]).sum() / w.sum()
            xb = b[s]["x"] - (w * b[s]["x"]).sum() / w.sum()
            row["b_Linf"] = float(np.abs(a[s]["b"] - b[s]["b"]).max())
            row["b_scale"] = float(np.abs(a[s]["b"]).max())
            row["x_Linf"] = float(np.abs(xa - xb).max())
            row["x_scale"] = float(np.abs(xa).max())
            row["x_mean_ref"] = float((w * a[s]["x"]).sum() / w.sum())
            row["x_mean_b200"] = float((w * b[s]["x"]).sum() / w.sum())
        rows.append(row)
    out["steps"] = rows
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    if len(sys.argv) > 2 and sys.argv[2] == "self":   # the reference against itself: how reproducible is the comparison's yardstick
        print(json.dumps({"pair": "ref_harness_gpu twice", **compare(n, names=("ref_harness_gpu", "ref_harness_gpu"))}))
    else:
        print(json.dumps(compare(n)))
