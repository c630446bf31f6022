"""CPU tests of the host-side AMR ghost-stencil plan (cup2d_b200/csrc/amr_plan.cpp through the C ABI): the CSR tables,
applied to the seeded fields of the reference's own 7-level mesh, must reproduce the labs the unmodified reference
assembled (tests/golden/amrlab_lmax8.npz) and, fed through the operators, its flux-corrected results."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import cup2d_amr_oracle as amr
from cup2d_b200.amr import AmrPlan, LAB_SHAPES


@pytest.fixture(scope="module")
def setup(golden_dir):
    d = np.load(os.path.join(golden_dir, "amrlab_lmax8.npz"))
    plan = AmrPlan(d["blocks"], int(d["bpdx"]), int(d["bpdy"]))
    yield d, plan
    plan.close()


def apply(plan, which, field):
    """lab[k, iy, ix, d] = table . field;  unwritten cells -> NaN"""
    rowptr, sb, sc, w = plan.stencil(which)
    ny, nx, dim = LAB_SHAPES[which]
    nb = len(plan.blocks)
    cols = sb.astype(np.int64) * (64 * dim) + sc
    T = sp.csr_matrix((w, cols, rowptr), shape=(nb * ny * nx * dim, nb * 64 * dim))
    lab = (T @ field.reshape(-1)).reshape(nb, ny, nx, dim)
    empty = (np.diff(rowptr) == 0).reshape(nb, ny, nx, dim)
    lab[empty] = np.nan
    return lab


@pytest.mark.parametrize("which,name,fieldname", [(2, "lab_pres1", "pres"), (1, "lab_vel1", "vel"), (0, "lab_vel3", "vel")])
def test_ghost_stencil_tables_reproduce_reference_labs(setup, which, name, fieldname):
    d, plan = setup
    lab = apply(plan, which, d[fieldname])
    got = lab[d["lab_blocks"]]
    want = d[name]
    written = ~np.isnan(got)
    scale = np.abs(want).max()
    assert np.abs(got[written] - want[written]).max() < 1e-13 * scale
    g = 3 if which == 0 else 1
    if which == 0:
        assert written.all()
    else:  # everything but the four corner ghosts of the cross-shaped stencils
        assert written[:, g:-g, :, :].all() and written[:, :, g:-g, :].all()
    # interior rows are the identity
    assert np.array_equal(got[:, g:g + 8, g:g + 8, :], d[fieldname][d["lab_blocks"]])


def test_tables_are_small_and_mostly_copies(setup):
    d, plan = setup
    rowptr, sb, sc, w = plan.stencil(0)
    per_row = np.diff(rowptr)
    assert per_row.max() <= 64              # the widest ghost (Taylor from 3x3 coarse cells that are 2x2 averages)
    assert (per_row == 1).mean() > 0.6      # interior + same-level copies + wall ghosts


def test_flux_correction_faces_match_oracle(setup):
    d, plan = setup
    mesh = amr.Mesh(d["blocks"], int(d["bpdx"]), int(d["bpdy"]))
    faces = plan.faces()
    want = []
    for k, (level, I, J) in enumerate(mesh.blocks):
        st = amr.face_states(mesh, k)
        for f, (cx, cy) in enumerate(amr.FACE_CODES):
            if st[f] == "coarse":
                want.append((k, f, mesh.index[(level - 1, (I + cx) // 2, (J + cy) // 2)], f ^ 1, (J % 2) if cx else (I % 2)))
    assert sorted(map(tuple, faces.tolist())) == sorted(want) and len(want) > 50


def test_operators_on_table_labs_match_reference(setup):
    """the advect kernel of the oracle fed with labs from the TABLES (not from the oracle's own assembly) + the oracle's
    flux correction reproduces the reference's flux-corrected tmpV to rounding"""
    import cup2d_oracle as orc
    d, plan = setup
    mesh = amr.Mesh(d["blocks"], int(d["bpdx"]), int(d["bpdy"]))
    nu, dt, h0 = float(d["nu"]), float(d["dt"]), float(d["h0"])
    lab = apply(plan, 0, d["vel"])
    nb = len(mesh.blocks)
    adv = np.empty((nb, 8, 8, 2))
    faces = []
    for k in range(nb):
        m = lab[k]
        h = h0 / (1 << mesh.blocks[k][0])
        adv[k, :, :, 0], adv[k, :, :, 1] = orc.advect_diffuse_padded(m[:, :, 0], m[:, :, 1], h, nu, dt)
        st = amr.face_states(mesh, k)
        c = m[3:11, 3:11]
        fl = [nu * dt * (c[:, 0] - m[3:11, 2]), nu * dt * (c[:, 7] - m[3:11, 11]), nu * dt * (c[0, :] - m[2, 3:11]),
              nu * dt * (c[7, :] - m[11, 3:11])]
        faces.append([fl[f] if st[f] in ("coarse", "fine") else None for f in range(4)])
    amr.flux_correct(mesh, adv, faces, 2)
    assert np.abs(adv - d["adv"]).max() < 1e-12 * np.abs(d["adv"]).max()


def test_plan_rejects_bad_meshes():
    from cup2d_b200 import Cup2dError
    with pytest.raises(Cup2dError):
        AmrPlan([[0, 0, 0], [0, 0, 0]], 1, 1)          # duplicate block
    with pytest.raises(Cup2dError):
        AmrPlan([[1, 2, 0]], 1, 1)                     # outside the domain
    # level 0 next to level 2: more than one level apart
    blocks = [[0, 0, 0]] + [[2, 4 + i, j] for i in range(4) for j in range(4)]
    with pytest.raises(Cup2dError):
        AmrPlan(blocks, 2, 1)


def balanced_disc_mesh(l0, lmax, bpdx=1, bpdy=1):
    """synthetic 2:1-balanced mesh: level l0 everywhere, refined level by level toward a circle"""
    blocks = {(l0, i, j) for i in range(bpdx << l0) for j in range(bpdy << l0)}
    for L in range(l0, lmax):
        n = 1 << L
        want = {b for b in blocks if b[0] == L and abs(np.hypot((b[1] + 0.5) / n - 0.5 * bpdx, (b[2] + 0.5) / n - 0.5 * bpdy) - 0.3) < 1.5 / n}
        # ripple: a block may only be refined if all its neighbours are at its level or finer
        changed = True
        while changed:
            changed = False
            for (l, i, j) in list(want):
                for di in (-1, 0, 1):
                    for dj in (-1, 0, 1):
                        a, b = i + di, j + dj
                        if 0 <= a < (bpdx << l) and 0 <= b < (bpdy << l) and (l, a, b) not in blocks:
                            par = (l - 1, a >> 1, b >> 1)
                            if par in blocks:        # coarser neighbour: it has to be refined first -> skip this one
                                want.discard((l, i, j))
                                changed = True
        for (l, i, j) in want:
            blocks.discard((l, i, j))
            blocks.update({(l + 1, 2 * i + a, 2 * j + b) for a in (0, 1) for b in (0, 1)})
    return np.array(sorted(blocks), dtype=np.int32)


def test_compact_ghost_tables_equal_full_rows_and_scale():
    """the compact tables (ghost rows of irregular blocks, built on several host threads) are exactly the corresponding
    rows of the full tables, on a synthetic mesh of a few thousand blocks"""
    import time
    blocks = balanced_disc_mesh(4, 8)
    assert len(blocks) > 1500 and len(set(blocks[:, 0].tolist())) >= 4
    plan = AmrPlan(blocks, 1, 1)
    irr = plan.irregular()
    assert 0 < len(irr) < len(blocks)
    for which in (0, 2):
        ny, nx, dim = LAB_SHAPES[which]
        t0 = time.time()
        rp, dst, sb, sc, w = plan.ghosts(which)
        t_compact = time.time() - t0
        frp, fsb, fsc, fw = plan.stencil(which)
        ncell = ny * nx * dim
        q, rem = np.divmod(dst, ncell)
        full_rows = irr[q].astype(np.int64) * ncell + rem
        assert np.array_equal(np.diff(rp), frp[full_rows + 1] - frp[full_rows])
        take = np.concatenate([np.arange(frp[r], frp[r + 1]) for r in full_rows[:4000]])
        n = len(take)
        assert np.array_equal(sb[:n], fsb[take]) and np.array_equal(sc[:n], fsc[take]) and np.array_equal(w[:n], fw[take])
        assert t_compact < 20.0
    plan.close()


def test_pattern_dictionary_and_neighbour_table(setup):
    d, plan = setup
    plan.ghosts(0)
    st = plan.stats(0)
    assert st["fallbacks"] == 0 and 10 < st["patterns"] < len(plan.irregular())
    mesh = amr.Mesh(d["blocks"], int(d["bpdx"]), int(d["bpdy"]))
    nb = plan.neighbours()
    codes = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]
    irregular = []
    for k, (l, I, J) in enumerate(mesh.blocks):
        NX, NY = mesh.bpdx << l, mesh.bpdy << l
        for q, (cx, cy) in enumerate(codes):
            if not (0 <= I + cx < NX and 0 <= J + cy < NY):
                want = -1
            else:
                s = mesh.state(l, I + cx, J + cy)
                want = s if s >= 0 else (-2 if s == -2 else -3)
            assert nb[k, q] == want
        if (nb[k] < -1).any():
            irregular.append(k)
    assert irregular == plan.irregular().tolist()


def test_coarse_face_formulation_of_the_flux_correction(setup):
    """csrc/amr_ops.cu applies fillcases per COARSE face: own flux + (fine a + fine b) per position, added once, and once
    more for flat entries >= 9 of a vector face when both fine blocks exist.  Same algorithm in numpy, from the plan's
    face list and table-made labs, against the reference's flux-corrected advect result (bit for bit when the labs are
    the oracle's; 1e-12 with table labs)."""
    import cup2d_oracle as orc
    d, plan = setup
    mesh = amr.Mesh(d["blocks"], int(d["bpdx"]), int(d["bpdy"]))
    nu, dt, h0 = float(d["nu"]), float(d["dt"]), float(d["h0"])
    nb = len(mesh.blocks)
    lab = amr.Lab(mesh, d["vel"], (-3, -3, 4, 4, True), "vector")
    labs = np.stack([lab.load(k).copy() for k in range(nb)])
    adv = np.empty((nb, 8, 8, 2))
    for k in range(nb):
        adv[k, :, :, 0], adv[k, :, :, 1] = orc.advect_diffuse_padded(labs[k, :, :, 0], labs[k, :, :, 1],
                                                                     h0 / (1 << mesh.blocks[k][0]), nu, dt)

    def cells(face, t):
        return ((0 if face == 0 else 7, t, -1 if face == 0 else 8, t) if face < 2
                else (t, 0 if face == 2 else 7, t, -1 if face == 2 else 8))

    def flux(k, face, t, comp):
        ix, iy, gx, gy = cells(face, t)
        return (nu * dt) * (labs[k, iy + 3, ix + 3, comp] - labs[k, gy + 3, gx + 3, comp])

    byface = {}
    for fine, ff, kc, fc, half in plan.faces().tolist():
        byface.setdefault((kc, fc), [-1, -1])[half] = fine
    for xfaces in (True, False):
        for (kc, fc), fine in sorted(byface.items()):
            if (fc < 2) != xfaces:
                continue
            for t in range(8):
                for comp in range(2):
                    acc = flux(kc, fc, t, comp)
                    fb = fine[t >> 2]
                    if fb >= 0:
                        t2 = 2 * (t & 3)
                        acc += flux(fb, fc ^ 1, t2, comp) + flux(fb, fc ^ 1, t2 + 1, comp)
                    ix, iy, _, _ = cells(fc, t)
                    v = adv[kc, iy, ix, comp] + acc
                    if fine[0] >= 0 and fine[1] >= 0 and 2 * t + comp >= 9:
                        v += acc
                    adv[kc, iy, ix, comp] = v
    assert np.array_equal(adv, d["adv"])


def test_poisson_rows_equal_the_reference_assembly(setup):
    """stencil rows (from the face-neighbour table) + general rows (CSR) == the COO the reference's own assembly loop
    pushed for this mesh (main.cpp:7051-7113), value for value"""
    d, plan = setup
    nbr, rows, rowptr, col, val = plan.poisson()
    n = 64 * len(plan.blocks)
    ref = sp.coo_matrix((d["coo_val"], (d["coo_row"], d["coo_col"])), shape=(n, n)).tocsr()
    ref.sum_duplicates()
    ref.sort_indices()
    general = set(rows.tolist())
    assert 1000 < len(general) < n // 2
    # general rows: bitwise
    for q, r in enumerate(rows):
        a, b = ref.indptr[r], ref.indptr[r + 1]
        nz = ref.data[a:b] != 0
        mine_nz = val[rowptr[q]:rowptr[q + 1]] != 0
        assert np.array_equal(ref.indices[a:b][nz], col[rowptr[q]:rowptr[q + 1]][mine_nz])
        assert np.array_equal(ref.data[a:b][nz], val[rowptr[q]:rowptr[q + 1]][mine_nz])
    # all other rows: the 5-point stencil through the neighbour table
    for k in range(len(plan.blocks)):
        for iy in range(8):
            for ix in range(8):
                r = 64 * k + 8 * iy + ix
                if r in general:
                    continue
                cols = []
                for (x, y, face) in ((ix - 1, iy, 0), (ix + 1, iy, 1), (ix, iy - 1, 2), (ix, iy + 1, 3)):
                    if 0 <= x < 8 and 0 <= y < 8:
                        cols.append(64 * k + 8 * y + x)
                    elif nbr[k, face] >= 0:
                        cols.append(64 * int(nbr[k, face]) + 8 * (y % 8) + (x % 8))
                want = dict.fromkeys(cols, 1.0)
                want[r] = -float(len(cols))
                a, b = ref.indptr[r], ref.indptr[r + 1]
                assert dict(zip(ref.indices[a:b].tolist(), ref.data[a:b].tolist())) == want


def test_chi_lab_tables_of_the_tagging_rule(golden_dir):
    """stencil kind 3 = GradChiOnTmp's {-4,-4,5,5,tensorial} chi lab (main.cpp:4633): the tables applied to the reference's
    chi field give its lab to rounding and with the same sign pattern (the rule only asks where the lab is positive)"""
    from cup2d_b200.amr import AmrPlan
    d = np.load(os.path.join(golden_dir, "amrtags_lmax8.npz"))
    nb = len(d["blocks"])
    plan = AmrPlan(np.ascontiguousarray(d["blocks"], dtype=np.int32), int(d["bpdx"]), int(d["bpdy"]))
    rp, sb, sc, w = plan.stencil(3)
    assert len(rp) == nb * 256 + 1 and (rp[1:] > rp[:-1]).all()       # tensorial: every lab cell has sources
    vals = w * d["chi"].reshape(nb, 64)[sb, sc]
    lab = np.add.reduceat(vals, rp[:-1]).reshape(nb, 16, 16)
    assert np.abs(lab - d["lab_chi4"]).max() < 1e-14
    assert np.array_equal(lab > 0, d["lab_chi4"] > 0)


def test_plan_of_the_c5_bench_mesh_scales():
    """config C5 (3-level AMR, 16384^2 effective): the synthetic mesh of tools/bench_amr.py at full size — 501 376 blocks,
    32 M cells — is accepted (2:1 balanced), and what a device context needs from the plan (compact ghost tables of the three
    stencils, Poisson rows) is built in seconds and grows with the level interfaces (5 400 irregular blocks), not with the mesh"""
    import sys
    import time
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import bench_amr
    from cup2d_b200.amr import AmrPlan
    blocks = bench_amr.three_level_mesh(9)
    assert len(blocks) == 501376 and sorted(set(blocks[:, 0].tolist())) == [9, 10, 11]
    assert (blocks[:, 1] >= 0).all() and (blocks[:, 1] < (1 << blocks[:, 0])).all()
    t0 = time.time()
    plan = AmrPlan(blocks, 1, 1)
    irr = plan.irregular()
    nnz = [len(plan.ghosts(k)[-1]) for k in range(3)]
    nbr, rows, rowptr, col, val = plan.poisson()
    dt = time.time() - t0
    assert len(irr) == 5400 and nnz[0] < 1000 * len(irr) and len(rows) < 64 * len(irr)
    assert dt < 60, dt
