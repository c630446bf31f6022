"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (needs /root/reference, which does not exist on the GPU box):

    make -C oracle ref && python tests/golden/make_golden.py

Every fixture is the output of oracle/_ref/ref_harness (= /root/reference/main.cpp compiled as is;
see oracle/ref_harness.cpp for the modes) on seeded inputs.  Inputs are stored alongside outputs so
the tests never need the reference at run time.  All arrays are global row-major a[iy, ix].
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


def run(*args, bpd=(1, 1)):
    # one OpenMP thread: the threaded dot products of the CPU solver restatement are then summed in a fixed
    # order and the fixtures regenerate bit-identically
    env = dict(os.environ, CUP2D_REF_BPDX=str(bpd[0]), CUP2D_REF_BPDY=str(bpd[1]), OMP_NUM_THREADS="1")
    subprocess.run([HARNESS, *map(str, args)], check=True, stderr=subprocess.DEVNULL, env=env)


def taylor_green(N):
    x = (np.arange(N) + 0.5) / N
    X, Y = np.meshgrid(x, x)  # X[iy, ix]
    u = np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y)
    v = -np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y)
    p = np.cos(2 * np.pi * X) * np.cos(2 * np.pi * Y)
    return u, v, p


def make_inputs(kind, L, seed):
    N = 8 << L
    rng = np.random.default_rng(seed)
    if kind == "random":
        u, v, p = (rng.uniform(-1, 1, (N, N)) for _ in range(3))
        chi = rng.uniform(0, 1, (N, N))
        udu, udv = (rng.uniform(-1, 1, (N, N)) for _ in range(2))
    else:
        u, v, p = taylor_green(N)
        # small seeded perturbation so that no symmetry hides an indexing error
        u = u + 0.05 * rng.uniform(-1, 1, (N, N))
        v = v + 0.05 * rng.uniform(-1, 1, (N, N))
        x = (np.arange(N) + 0.5) / N
        X, Y = np.meshgrid(x, x)
        chi = np.exp(-((X - 0.4) ** 2 + (Y - 0.55) ** 2) / 0.02)
        udu, udv = 0.3 * np.sin(3 * X + Y), 0.2 * np.cos(2 * Y - X)
    return [np.ascontiguousarray(a, dtype=np.float64) for a in (u, v, p, chi, udu, udv)]


def gen_rect(tmp):
    """rectangular domains (bpdx != bpdy): the non-regular branch of the space-filling curve
    (main.cpp:6358-6376) and operators / steps on a 2x1 box (h0 = extent/max(bpdx,bpdy)/8)."""
    for bx, by, L in ((2, 1, 2), (3, 2, 1), (1, 2, 2)):
        out = os.path.join(tmp, "order.bin")
        run("order", L, out, bpd=(bx, by))
        np.save(os.path.join(HERE, f"order_{bx}x{by}_L{L}.npy"), np.fromfile(out, dtype=np.int32).reshape(-1, 2))
    bx, by, L = 2, 1, 2
    NX, NY = (8 << L) * bx, (8 << L) * by
    rng = np.random.default_rng(99)
    x = (np.arange(NX) + 0.5) / NX
    y = (np.arange(NY) + 0.5) / NY
    X, Y = np.meshgrid(x, y)
    u = np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y) + 0.05 * rng.uniform(-1, 1, (NY, NX))
    v = -np.cos(2 * np.pi * X) * np.sin(2 * np.pi * Y) + 0.05 * rng.uniform(-1, 1, (NY, NX))
    p = np.cos(2 * np.pi * X) * np.cos(2 * np.pi * Y)
    chi = np.exp(-((X - 0.4) ** 2 + (Y - 0.55) ** 2) / 0.02)
    udu, udv = 0.3 * np.sin(3 * X + Y), 0.2 * np.cos(2 * Y - X)
    ins = [np.ascontiguousarray(a) for a in (u, v, p, chi, udu, udv)]
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    np.concatenate([a.ravel() for a in ins]).tofile(fin)
    nu, dt = 1e-3, 2e-3
    run("ops", L, repr(nu), repr(dt), fin, fout, bpd=(bx, by))
    o = np.fromfile(fout).reshape(6, NY, NX)
    nsteps, kiter, cfl = 2, 10, 0.5
    run("steps", L, repr(nu), repr(cfl), nsteps, kiter, fin, fout, bpd=(bx, by))
    raw = np.fromfile(fout).reshape(nsteps, 1 + 5 * NX * NY)
    f = raw[:, 1:].reshape(nsteps, 5, NY, NX)
    np.savez_compressed(os.path.join(HERE, "rect_2x1_L2.npz"), nu=nu, dt=dt, L=L, bpdx=bx, bpdy=by, cfl=cfl, kiter=kiter,
                        u=ins[0], v=ins[1], p=ins[2], chi=ins[3], udef_u=ins[4], udef_v=ins[5],
                        adv_u=o[0], adv_v=o[1], rhs=o[2], rhs1=o[3], gradp_u=o[4], gradp_v=o[5],
                        step_dt=raw[:, 0].copy(), step_u=f[:, 0], step_v=f[:, 1], step_p=f[:, 2])
    print("rect 2x1", raw[:, 0])


def gen_order(tmp):
    for L in (0, 1, 2, 3, 4, 5):
        out = os.path.join(tmp, "order.bin")
        run("order", L, out)
        o = np.fromfile(out, dtype=np.int32).reshape(-1, 2)
        np.save(os.path.join(HERE, f"order_L{L}.npy"), o)
        print("order", L, o.shape)


def gen_ops(tmp, kind, L, seed, nu, dt):
    N = 8 << L
    ins = make_inputs(kind, L, seed)
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    np.concatenate([a.ravel() for a in ins]).tofile(fin)
    run("ops", L, repr(nu), repr(dt), fin, fout)
    o = np.fromfile(fout).reshape(6, N, N)
    np.savez_compressed(
        os.path.join(HERE, f"ops_L{L}_{kind}.npz"), nu=nu, dt=dt, L=L,
        u=ins[0], v=ins[1], p=ins[2], chi=ins[3], udef_u=ins[4], udef_v=ins[5],
        adv_u=o[0], adv_v=o[1], rhs=o[2], rhs1=o[3], gradp_u=o[4], gradp_v=o[5])
    print("ops", kind, L, float(np.abs(o).max()))


def gen_vort(tmp, kind, L, seed):
    N = 8 << L
    ins = make_inputs(kind, L, seed)
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    np.concatenate([a.ravel() for a in ins]).tofile(fin)
    run("vort", L, fin, fout)
    np.savez_compressed(os.path.join(HERE, f"vort_L{L}_{kind}.npz"), L=L, u=ins[0], v=ins[1],
                        vort=np.fromfile(fout).reshape(N, N))
    print("vort", kind, L)


def body_chi(N, blobs):
    """compactly supported chi with values below 0 and above 1 (GradChiOnTmp clamps before testing)"""
    x = (np.arange(N) + 0.5) / N
    X, Y = np.meshgrid(x, x)
    chi = np.full((N, N), -0.25)
    for (cx, cy, r) in blobs:
        chi = np.maximum(chi, 1.5 * (1.0 - np.hypot(X - cx, Y - cy) / r))
    return chi


def gen_tags(tmp, name, L, seed, rtol, extra, blobs):
    N = 8 << L
    ins = make_inputs("tg", L, seed)
    ins[3] = body_chi(N, blobs)
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    np.concatenate([a.ravel() for a in ins]).tofile(fin)
    run("tags", L, repr(rtol), extra, fin, fout)
    np.savez_compressed(os.path.join(HERE, f"tags_{name}.npz"), L=L, rtol=rtol, offset=4 if extra == 0 else 2,
                        u=ins[0], v=ins[1], chi=ins[3], tags=np.fromfile(fout).reshape(N, N))
    print("tags", name)


def gen_dump(tmp, L, seed, time):
    N = 8 << L
    ins = make_inputs("random", L, seed)
    fin, pref = os.path.join(tmp, "in.bin"), os.path.join(tmp, "vel.00000007")
    np.concatenate([a.ravel() for a in ins]).tofile(fin)
    run("dump", L, repr(time), fin, pref)
    np.savez_compressed(os.path.join(HERE, f"dump_L{L}.npz"), L=L, time=time, u=ins[0], v=ins[1],
                        xyz=np.fromfile(pref + ".xyz.raw", dtype=np.float32),
                        attr=np.fromfile(pref + ".attr.raw", dtype=np.float32),
                        xdmf=np.frombuffer(open(pref + ".xdmf2", "rb").read(), dtype=np.uint8))
    print("dump", L)


def parse_penal(path, L):
    """records of `ref_harness penal` (see oracle/ref_harness.cpp) -> list of per-step dicts"""
    N = 8 << L
    n2 = N * N
    a = np.fromfile(path)
    i, steps = 0, []
    while i < len(a):
        tag, n = int(a[i]), int(a[i + 1])
        p = a[i + 2:i + 2 + n]
        i += 2 + n
        if tag == 2:
            cur = {"u0": p[:n2].reshape(N, N), "v0": p[n2:].reshape(N, N), "shapes": []}
            steps.append(cur)
        elif tag == 1:
            nob = int(p[13])
            o = 14
            ids = p[o:o + nob].astype(np.int32)
            o += nob
            X = p[o:o + nob * 64].reshape(nob, 8, 8)
            o += nob * 64
            cur["lam"], cur["dt"] = p[2], p[3]
            cur["shapes"].append(dict(ids=ids, X=X, udef=p[o:o + nob * 128].reshape(nob, 8, 8, 2), cx=p[4], cy=p[5],
                                      Q=p[6:13].copy()))
        elif tag == 3:
            S = len(cur["shapes"])
            for k, sh in enumerate(cur["shapes"]):
                sh["u"], sh["v"], sh["omega"] = p[3 * k:3 * k + 3]
            cur["u1"], cur["v1"], cur["udu"], cur["udv"], cur["chi"] = p[3 * S:].reshape(5, N, N)
    return steps


def gen_penal(tmp, L, nsteps, keep):
    """penalisation phase of the reference with two fish close enough for their blocks to overlap (and, at the
    last step, to collide, which changes the rigid motion the host hands to the blend)"""
    fout = os.path.join(tmp, "penal.bin")
    env = dict(os.environ, OMP_NUM_THREADS="1",
               CUP2D_REF_SHAPES="angle=0 L=0.8 xpos=0.52 ypos=0.44\n angle=175 L=0.8 xpos=0.47 ypos=0.56")
    subprocess.run([HARNESS, "penal", str(L), str(nsteps), "5", fout], check=True, stderr=subprocess.DEVNULL,
                   stdout=subprocess.DEVNULL, env=env)
    steps = parse_penal(fout, L)
    out = {"L": L, "nshapes": 2, "steps": np.array(keep)}
    for s in keep:
        st = steps[s]
        for k in ("u0", "v0", "u1", "v1", "udu", "udv", "chi", "lam", "dt"):
            out[f"s{s}_{k}"] = st[k]
        for j, sh in enumerate(st["shapes"]):
            for k in ("ids", "X", "udef", "Q"):
                out[f"s{s}_sh{j}_{k}"] = sh[k]
            out[f"s{s}_sh{j}_rigid"] = np.array([sh["cx"], sh["cy"], sh["u"], sh["v"], sh["omega"]])
    np.savez_compressed(os.path.join(HERE, f"penal_L{L}.npz"), **out)
    print("penal", L, keep)


def load_penal(path):
    """inverse of gen_penal's flattening: list of per-step dicts (used by the tests)"""
    d = np.load(path)
    steps = []
    for s in d["steps"]:
        st = {k: d[f"s{s}_{k}"] for k in ("u0", "v0", "u1", "v1", "udu", "udv", "chi")}
        st["lam"], st["dt"] = float(d[f"s{s}_lam"]), float(d[f"s{s}_dt"])
        st["shapes"] = []
        for j in range(int(d["nshapes"])):
            cx, cy, u, v, om = d[f"s{s}_sh{j}_rigid"]
            st["shapes"].append(dict(ids=d[f"s{s}_sh{j}_ids"], X=d[f"s{s}_sh{j}_X"], udef=d[f"s{s}_sh{j}_udef"],
                                     Q=d[f"s{s}_sh{j}_Q"], cx=float(cx), cy=float(cy), u=float(u), v=float(v),
                                     omega=float(om)))
        steps.append(st)
    return int(d["L"]), steps


def gen_amrlab(tmp, level_max, nsteps):
    """ghost assembly + the four operators of the reference on its own multi-level run.sh mesh:
    mesh, seeded inputs, flux-corrected outputs for every block and the assembled labs of every second block"""
    fout = os.path.join(tmp, "amrlab.bin")
    fcoo = os.path.join(tmp, "coo.bin")
    subprocess.run([HARNESS, "amrlab", str(level_max), str(nsteps), fout], check=True, stderr=subprocess.DEVNULL,
                   stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1", CUP2D_REF_DUMP_COO=fcoo))
    raw = open(fcoo, "rb").read()   # the Poisson matrix of the same mesh, as the reference's assembly pushed it
    m_rows, nnz = np.frombuffer(raw, dtype=np.int64, count=2)
    coo_row = np.frombuffer(raw, dtype=np.int32, count=nnz, offset=16)
    coo_col = np.frombuffer(raw, dtype=np.int32, count=nnz, offset=16 + 4 * nnz)
    coo_val = np.frombuffer(raw, dtype=np.float64, count=nnz, offset=16 + 8 * nnz)
    a = np.fromfile(fout)
    i, rec = 0, {}
    while i < len(a):
        tag, n = int(a[i]), int(a[i + 1])
        rec[tag] = a[i + 2:i + 2 + n]
        i += 2 + n
    nu, dt, h0, bpdx, bpdy, lmax = rec[10]
    blocks = rec[11].reshape(-1, 3).astype(np.int32)
    nb = len(blocks)
    sub = np.arange(0, nb, 2)
    np.savez_compressed(
        os.path.join(HERE, f"amrlab_lmax{level_max}.npz"), nu=nu, dt=dt, h0=h0, bpdx=int(bpdx), bpdy=int(bpdy),
        blocks=blocks, vel=rec[12].reshape(nb, 8, 8, 2), pres=rec[13].reshape(nb, 8, 8, 1),
        chi=rec[14].reshape(nb, 8, 8, 1), udef=rec[15].reshape(nb, 8, 8, 2), lab_blocks=sub,
        lab_vel3=rec[20].reshape(nb, 14, 14, 2)[sub], lab_vel1=rec[21].reshape(nb, 10, 10, 2)[sub],
        lab_pres1=rec[22].reshape(nb, 10, 10, 1)[sub], adv=rec[30].reshape(nb, 8, 8, 2),
        rhs=rec[31].reshape(nb, 8, 8, 1), rhs1=rec[32].reshape(nb, 8, 8, 1), gradp=rec[33].reshape(nb, 8, 8, 2),
        coo_row=coo_row, coo_col=coo_col, coo_val=coo_val)
    assert m_rows == 64 * nb, "the dumped matrix belongs to another mesh"
    print("amrlab", nb, "blocks, levels", sorted(set(blocks[:, 0].tolist())))


def gen_amrtags(tmp, level_max, nsteps):
    """what adapt() looks at (main.cpp:4676-4678) on the reference's own multi-level run.sh mesh with its two fish, after
    nsteps steps: mesh, vel, chi, the chi lab of GradChiOnTmp ({-4,-4,5,5,tensorial}), the tagging field before and after
    the chi rule"""
    fout = os.path.join(tmp, "atags.bin")
    subprocess.run([HARNESS, "atags", str(level_max), str(nsteps), fout], check=True, stderr=subprocess.DEVNULL,
                   stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1"))
    a = np.fromfile(fout)
    i, rec = 0, {}
    while i < len(a):
        tag, n = int(a[i]), int(a[i + 1])
        rec[tag] = a[i + 2:i + 2 + n]
        i += 2 + n
    rtol, ctol, lmax, h0, bpdx, bpdy = rec[10]
    blocks = rec[11].reshape(-1, 3).astype(np.int32)
    nb = len(blocks)
    np.savez_compressed(
        os.path.join(HERE, f"amrtags_lmax{level_max}.npz"), rtol=rtol, ctol=ctol, level_max=int(lmax), h0=h0, bpdx=int(bpdx),
        bpdy=int(bpdy), blocks=blocks, vel=rec[12].reshape(nb, 8, 8, 2), chi=rec[14].reshape(nb, 8, 8, 1),
        lab_chi4=rec[23].reshape(nb, 16, 16), vort=rec[40].reshape(nb, 8, 8), tagfield=rec[41].reshape(nb, 8, 8))
    fired = (rec[40] != rec[41]).reshape(nb, 64).any(axis=1)
    print("amrtags", nb, "blocks; chi rule fired on", int(fired.sum()))


def gen_amrdump(tmp, level_max, nsteps):
    """the reference's dump() of its velocity on its own multi-level run.sh mesh after nsteps steps: mesh, vel, the three files"""
    pref = os.path.join(tmp, "vel.00000003")
    subprocess.run([HARNESS, "adump", str(level_max), str(nsteps), pref], check=True, stderr=subprocess.DEVNULL,
                   stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1"))
    a = np.fromfile(pref + ".bin")
    i, rec = 0, {}
    while i < len(a):
        tag, n = int(a[i]), int(a[i + 1])
        rec[tag] = a[i + 2:i + 2 + n]
        i += 2 + n
    time, h0, bpdx, bpdy = rec[10]
    blocks = rec[11].reshape(-1, 3).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, f"amrdump_lmax{level_max}.npz"), time=time, h0=h0, bpdx=int(bpdx), bpdy=int(bpdy),
                        blocks=blocks, vel=rec[12].reshape(len(blocks), 8, 8, 2),
                        xyz=np.fromfile(pref + ".xyz.raw", dtype=np.float32), attr=np.fromfile(pref + ".attr.raw", dtype=np.float32),
                        xdmf=np.frombuffer(open(pref + ".xdmf2", "rb").read(), dtype=np.uint8))
    print("amrdump", len(blocks), "blocks")


def gen_steps(tmp, kind, L, seed, nu, cfl, nsteps, kiter):
    N = 8 << L
    ins = make_inputs(kind, L, seed)
    fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
    np.concatenate([a.ravel() for a in ins]).tofile(fin)
    run("steps", L, repr(nu), repr(cfl), nsteps, kiter, fin, fout)
    raw = np.fromfile(fout).reshape(nsteps, 1 + 5 * N * N)
    dts = raw[:, 0].copy()
    f = raw[:, 1:].reshape(nsteps, 5, N, N)
    np.savez_compressed(
        os.path.join(HERE, f"steps_L{L}_{kind}_k{kiter}.npz"), nu=nu, cfl=cfl, L=L, kiter=kiter,
        u0=ins[0], v0=ins[1], p0=ins[2], dt=dts,
        u=f[:, 0], v=f[:, 1], p=f[:, 2], b=f[:, 3], x=f[:, 4])
    print("steps", kind, L, dts)


if __name__ == "__main__":
    if not os.path.exists(HARNESS):
        sys.exit("build oracle/_ref/ref_harness first: make -C oracle ref")
    with tempfile.TemporaryDirectory() as tmp:
        gen_rect(tmp)
        gen_order(tmp)
        gen_vort(tmp, "random", 2, 4242)
        gen_vort(tmp, "tg", 3, 4243)
        gen_tags(tmp, "L3_finest", 3, 4244, 5.0, 0, [(0.4, 0.55, 0.12), (0.02, 0.97, 0.05)])
        gen_tags(tmp, "L3_coarser", 3, 4245, 3.0, 1, [(0.7, 0.3, 0.1), (0.99, 0.01, 0.04), (0.26, 0.76, 0.015)])
        gen_dump(tmp, 2, 4246, 0.1875)
        gen_penal(tmp, 4, 4, [2, 3])
        gen_amrlab(tmp, 8, 3)
        gen_amrtags(tmp, 8, 5)
        gen_amrdump(tmp, 8, 4)
        gen_ops(tmp, "random", 2, 1234, 1e-3, 2.5e-3)
        gen_ops(tmp, "tg", 3, 4321, 1e-3, 1.2e-3)
        gen_steps(tmp, "tg", 2, 777, 1e-3, 0.5, 3, 12)
        gen_steps(tmp, "random", 2, 778, 1e-2, 0.4, 2, 8)
        gen_steps(tmp, "tg", 3, 779, 1e-3, 0.5, 2, 40)
        gen_steps(tmp, "tg", 3, 779, 1e-3, 0.5, 2, 15)
