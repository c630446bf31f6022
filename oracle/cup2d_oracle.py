"""TEST INFRASTRUCTURE — CPU restatement (numpy) of the CUP2D hot path on a uniform grid.

This module is the *checker*.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline/reference leg may import it; the product path (cup2d_b200/) never does.

Parity status: PINNED.  Every function below is checked in tests/test_oracle_golden.py against
outputs of the unmodified reference (oracle/_ref/ref_harness = /root/reference/main.cpp compiled as
is, driven through oracle/shim/), committed as fixtures under tests/golden/ together with the
script that generated them (tests/golden/make_golden.py).  The reference ships no tests or golden
vectors of its own.

All citations are file:line in /root/reference.  Fields are held as GLOBAL 2-D arrays
a[iy, ix] (row-major, N = 8*2^L cells per side, h = extent/N); `to_blocks/from_blocks` convert
to the reference's memory layout (8x8 blocks concatenated in Hilbert order, vectors interleaved
u,v per cell: main.cpp:6517, 5467-5468).
"""
import numpy as np

BS = 8  # _BS_, Makefile:12


# --------------------------------------------------------------------------------------------
# Space-filling curve (main.cpp:342-446) — regular case bpdx == bpdy == 2^k
# --------------------------------------------------------------------------------------------
def hilbert_xy2d(b, x, y):
    """SpaceCurve::AxestoTranspose (main.cpp:347-359): index of block (x,y) on a 2^b x 2^b curve."""
    n = 1 << b
    d = 0
    s = n // 2
    while s > 0:
        rx = 1 if (x & s) > 0 else 0
        ry = 1 if (y & s) > 0 else 0
        d += s * s * ((3 * rx) ^ ry)
        # rot (main.cpp:374-384) with n = full size
        if ry == 0:
            if rx == 1:
                x = n - 1 - x
                y = n - 1 - y
            x, y = y, x
        s //= 2
    return d


def hilbert_d2xy(b, d):
    """SpaceCurve::TransposetoAxes (main.cpp:360-373)."""
    n = 1 << b
    x = y = 0
    t = d
    s = 1
    while s < n:
        rx = 1 & (t // 2)
        ry = 1 & (t ^ rx)
        if ry == 0:
            if rx == 1:
                x = s - 1 - x
                y = s - 1 - y
            x, y = y, x
        x += s * rx
        y += s * ry
        t //= 4
        s *= 2
    return x, y


def block_order(L):
    """(i, j) of every block of a uniform level-L grid (bpdx=bpdy=1) in reference `infos` order:
    blocks are sorted by the Hilbert key (main.cpp:1550-1562, 6519-6537); returns int array (4^L, 2)."""
    nb = 1 << L
    out = np.empty((nb * nb, 2), dtype=np.int32)
    for d in range(nb * nb):
        out[d] = hilbert_d2xy(L, d)
    return out


def to_blocks(a, order, dim=1):
    """global field(s) -> reference block layout, flat.  dim=2: a = (u, v) interleaved per cell."""
    nblk = len(order)
    out = np.empty((nblk, BS, BS, dim))
    comps = (a,) if dim == 1 else a
    for k, (i, j) in enumerate(order):
        for c in range(dim):
            out[k, :, :, c] = comps[c][j * BS:(j + 1) * BS, i * BS:(i + 1) * BS]
    return out.reshape(-1)


def from_blocks(flat, order, dim=1):
    nblk = len(order)
    order = np.asarray(order)
    nbx, nby = int(order[:, 0].max()) + 1, int(order[:, 1].max()) + 1
    blk = np.asarray(flat).reshape(nblk, BS, BS, dim)
    outs = [np.empty((nby * BS, nbx * BS)) for _ in range(dim)]
    for k, (i, j) in enumerate(order):
        for c in range(dim):
            outs[c][j * BS:(j + 1) * BS, i * BS:(i + 1) * BS] = blk[k, :, :, c]
    return outs[0] if dim == 1 else tuple(outs)


# --------------------------------------------------------------------------------------------
# Ghost cells at the domain edge (what BlockLab + _apply_bc produce on a uniform grid)
# --------------------------------------------------------------------------------------------
def pad_vector(u, v, g):
    """VectorLab::applyBCface (main.cpp:3131-3154): every ghost layer takes the wall-adjacent cell,
    normal component negated, tangential copied (free-slip).  Corner ghosts are not used by any
    hot-path stencil (all are cross-shaped)."""
    up = np.pad(u, g, mode="edge")
    vp = np.pad(v, g, mode="edge")
    up[:, :g] *= -1.0  # x faces: u is normal
    up[:, -g:] *= -1.0
    vp[:g, :] *= -1.0  # y faces: v is normal
    vp[-g:, :] *= -1.0
    return up, vp


def pad_scalar(p, g):
    """ScalarLab::Neumann2D (main.cpp:3210-3245): constant extrapolation."""
    return np.pad(p, g, mode="edge")


# --------------------------------------------------------------------------------------------
# WENO5 (main.cpp:162-208)
# --------------------------------------------------------------------------------------------
def _betas(um2, um1, u, up1, up2):
    b1 = 13.0 / 12.0 * ((um2 + u) - 2 * um1) ** 2 + 0.25 * ((um2 + 3 * u) - 4 * um1) ** 2
    b2 = 13.0 / 12.0 * ((um1 + up1) - 2 * u) ** 2 + 0.25 * (um1 - up1) ** 2
    b3 = 13.0 / 12.0 * ((u + up2) - 2 * up1) ** 2 + 0.25 * ((3 * u + up2) - 4 * up1) ** 2
    return b1, b2, b3


def weno5_plus(um2, um1, u, up1, up2):
    """main.cpp:162-181"""
    e = 1e-6
    b1, b2, b3 = _betas(um2, um1, u, up1, up2)
    g1, g2, g3 = 0.1, 0.6, 0.3
    what1 = g1 / (b1 + e) ** 2
    what2 = g2 / (b2 + e) ** 2
    what3 = g3 / (b3 + e) ** 2
    aux = 1.0 / ((what1 + what3) + what2)
    w1, w2, w3 = what1 * aux, what2 * aux, what3 * aux
    f1 = (11.0 / 6.0) * u + ((1.0 / 3.0) * um2 - (7.0 / 6.0) * um1)
    f2 = (5.0 / 6.0) * u + ((-1.0 / 6.0) * um1 + (1.0 / 3.0) * up1)
    f3 = (1.0 / 3.0) * u + ((+5.0 / 6.0) * up1 - (1.0 / 6.0) * up2)
    return (w1 * f1 + w3 * f3) + w2 * f2


def weno5_minus(um2, um1, u, up1, up2):
    """main.cpp:182-201"""
    e = 1e-6
    b1, b2, b3 = _betas(um2, um1, u, up1, up2)
    g1, g2, g3 = 0.3, 0.6, 0.1
    what1 = g1 / (b1 + e) ** 2
    what2 = g2 / (b2 + e) ** 2
    what3 = g3 / (b3 + e) ** 2
    aux = 1.0 / ((what1 + what3) + what2)
    w1, w2, w3 = what1 * aux, what2 * aux, what3 * aux
    f1 = (1.0 / 3.0) * u + ((-1.0 / 6.0) * um2 + (5.0 / 6.0) * um1)
    f2 = (5.0 / 6.0) * u + ((1.0 / 3.0) * um1 - (1.0 / 6.0) * up1)
    f3 = (11.0 / 6.0) * u + ((-7.0 / 6.0) * up1 + (1.0 / 3.0) * up2)
    return (w1 * f1 + w3 * f3) + w2 * f2


def derivative(U, um3, um2, um1, u, up1, up2, up3):
    """main.cpp:202-208"""
    plus = weno5_plus(um2, um1, u, up1, up2) - weno5_plus(um3, um2, um1, u, up1)
    minus = weno5_minus(um1, u, up1, up2, up3) - weno5_minus(um2, um1, u, up1, up2)
    return np.where(U > 0, plus, minus)


# --------------------------------------------------------------------------------------------
# Operators
# --------------------------------------------------------------------------------------------
def advect_diffuse_padded(up, vp, h, nu, dt):
    """KernelAdvectDiffuse::operator() (main.cpp:5441-5503) on fields that already carry their 3 ghost layers
    (shape (NY + 6, NX + 6)): the undivided RHS of the NY x NX interior."""
    g = 3
    NY, NX = up.shape[0] - 2 * g, up.shape[1] - 2 * g
    dfac = nu * dt
    afac = -dt * h

    def sx(a, k):  # a(ix+k, iy)
        return a[g:g + NY, g + k:g + k + NX]

    def sy(a, k):  # a(ix, iy+k)
        return a[g + k:g + k + NY, g:g + NX]

    u, v = sx(up, 0), sx(vp, 0)
    dudx = derivative(u, sx(up, -3), sx(up, -2), sx(up, -1), u, sx(up, 1), sx(up, 2), sx(up, 3))
    dudy = derivative(v, sy(up, -3), sy(up, -2), sy(up, -1), u, sy(up, 1), sy(up, 2), sy(up, 3))
    dvdx = derivative(u, sx(vp, -3), sx(vp, -2), sx(vp, -1), v, sx(vp, 1), sx(vp, 2), sx(vp, 3))
    dvdy = derivative(v, sy(vp, -3), sy(vp, -2), sy(vp, -1), v, sy(vp, 1), sy(vp, 2), sy(vp, 3))
    tu = afac * (u * dudx + v * dudy) + dfac * (sx(up, 1) + sx(up, -1) + sy(up, 1) + sy(up, -1) - 4 * u)
    tv = afac * (u * dvdx + v * dvdy) + dfac * (sx(vp, 1) + sx(vp, -1) + sy(vp, 1) + sy(vp, -1) - 4 * v)
    return tu, tv


def advect_diffuse(u, v, h, nu, dt):
    """KernelAdvectDiffuse on a uniform grid: free-slip wall ghosts, then the kernel."""
    up, vp = pad_vector(u, v, 3)
    return advect_diffuse_padded(up, vp, h, nu, dt)


def compute_dt(u, v, h, nu, cfl):
    """main.cpp:6579-6595"""
    umax = max(np.abs(u).max(), np.abs(v).max())
    dt_diff = 0.25 * h * h / (nu + 0.25 * h * umax)
    dt_adv = h / (umax + 1e-8)
    return min(dt_diff, cfl * dt_adv)


def rk2(u, v, h, nu, dt):
    """main.cpp:6607-6642: vold=vel; V=Vold+0.5*K(V)/h^2; V=Vold+K(V)/h^2."""
    uo, vo = u.copy(), v.copy()
    tu, tv = advect_diffuse(u, v, h, nu, dt)
    ih2 = 0.5 / (h * h)
    u1, v1 = uo + tu * ih2, vo + tv * ih2
    tu, tv = advect_diffuse(u1, v1, h, nu, dt)
    ih2 = 1.0 / (h * h)
    return uo + tu * ih2, vo + tv * ih2


def pressure_rhs(u, v, udu, udv, chi, h, dt):
    """pressure_rhs::operator() (main.cpp:6105-6139)"""
    NY, NX = u.shape
    up, vp = pad_vector(u, v, 1)
    dup, dvp = pad_vector(udu, udv, 1)
    fac = 0.5 * h / dt
    cy, cx = slice(1, 1 + NY), slice(1, 1 + NX)
    div_v = up[cy, 2:] - up[cy, :-2] + vp[2:, cx] - vp[:-2, cx]
    div_u = dup[cy, 2:] - dup[cy, :-2] + dvp[2:, cx] - dvp[:-2, cx]
    return fac * div_v - fac * chi * div_u


def laplacian_neumann(p):
    """5-point undivided Laplacian with constant-extrapolation ghosts = the Poisson matrix rows of
    main.cpp:7074-7107 (interior: +1,+1,-4,+1,+1; domain edge: missing neighbour omitted, diagonal
    = -(number of neighbours)) and pressure_rhs1's stencil (main.cpp:6209-6230)."""
    NY, NX = p.shape
    pp = pad_scalar(p, 1)
    cy, cx = slice(1, 1 + NY), slice(1, 1 + NX)
    return pp[cy, :-2] + pp[cy, 2:] + pp[:-2, cx] + pp[2:, cx] - 4 * p


def pressure_rhs1(tmp, pold):
    """main.cpp:6209-6230: tmp -= lap(pold)"""
    return tmp - laplacian_neumann(pold)


def vorticity(u, v, h):
    """KernelVorticity::operator() (main.cpp:3343-3366): tmp = (0.5/h) * ((u_S - u_N + v_E) - v_W); this is the
    field adapt() tags blocks with (L-inf per block against Rtol / Ctol, main.cpp:4676-4689)."""
    NY, NX = u.shape
    up, vp = pad_vector(u, v, 1)
    cy, cx = slice(1, 1 + NY), slice(1, 1 + NX)
    i2h = 0.5 / h
    return i2h * (((up[:-2, cx] - up[2:, cx]) + vp[cy, 2:]) - vp[cy, :-2])


def block_linf(a):
    """per-block max |a| on the 8x8 blocks, as an array [nby, nbx] (the quantity adapt() compares, main.cpp:4676-4680)"""
    NY, NX = a.shape
    return np.abs(a).reshape(NY // BS, BS, NX // BS, BS).max(axis=(1, 3))


def adapt_tags(u, v, chi, h, rtol, offset):
    """The field adapt() thresholds (main.cpp:4659-4660): KernelVorticity, then GradChiOnTmp (main.cpp:4631-4656):
    a block whose surroundings — the block grown by `offset` cells on every side, corners included; offset = 4 on the
    finest level, else 2 — hold any chi > 0 (after clamping to [0,1]) gets its four centre cells set to 2 Rtol.
    Neumann ghosts outside the domain replicate wall-adjacent cells that are inside the grown box anyway."""
    w = vorticity(u, v, h).copy()
    NY, NX = chi.shape
    pos = np.minimum(chi, 1.0)
    pos = np.maximum(pos, 0.0) > 0.0
    c = BS // 2
    for j in range(NY // BS):
        for i in range(NX // BS):
            y0, y1 = max(0, BS * j - offset), min(NY, BS * (j + 1) + offset)
            x0, x1 = max(0, BS * i - offset), min(NX, BS * (i + 1) + offset)
            if pos[y0:y1, x0:x1].any():
                w[BS * j + c - 1:BS * j + c + 1, BS * i + c - 1:BS * i + c + 1] = 2 * rtol
    return w


def dump_arrays(u, v, order, h0, level):
    """dump() (main.cpp:3425-3453): per block in `infos` order, per cell row-major: the 4 corners of the cell quad
    (u0,v0, u0,v1, u1,v1, u1,v0) and the attribute (u, v, 0), all narrowed to float32.  origin as main.cpp:695-696."""
    order = np.asarray(order)
    h = h0 / (1 << level)
    ox = (order[:, 0] * BS) * h0 / (1 << level)
    oy = (order[:, 1] * BS) * h0 / (1 << level)
    k = np.arange(BS, dtype=np.float64)
    u0 = ox[:, None, None] + (h * k)[None, None, :] + np.zeros((1, BS, 1))
    v0 = oy[:, None, None] + (h * k)[None, :, None] + np.zeros((1, 1, BS))
    u1, v1 = u0 + h, v0 + h
    xyz = np.stack([u0, v0, u0, v1, u1, v1, u1, v0], axis=-1).astype(np.float32)
    vel = to_blocks((u, v), order, 2).reshape(len(order), BS, BS, 2)
    attr = np.concatenate([vel, np.zeros((len(order), BS, BS, 1))], axis=-1).astype(np.float32)
    return xyz.reshape(-1), attr.reshape(-1)


def dump_xdmf(time, ncell, xyz_base, attr_base):
    """the .xdmf2 text dump() writes (main.cpp:3390-3423), byte for byte"""
    return ("<Xdmf\n"
            "    Version=\"2.0\">\n"
            "  <Domain>\n"
            "    <Grid>\n"
            "      <Time Value=\"%.16e\"/>\n"
            "      <Topology\n"
            "          Dimensions=\"%d\"\n"
            "          TopologyType=\"Quadrilateral\"/>\n"
            "     <Geometry\n"
            "         GeometryType=\"XY\">\n"
            "       <DataItem\n"
            "           Dimensions=\"%d 2\"\n"
            "           Format=\"Binary\">\n"
            "         %s\n"
            "       </DataItem>\n"
            "     </Geometry>\n"
            "       <Attribute\n"
            "           AttributeType=\"Vector\"\n"
            "           Name=\"vort\"\n"
            "           Center=\"Cell\">\n"
            "         <DataItem\n"
            "             Dimensions=\"3 %d\"\n"
            "             Format=\"Binary\">\n"
            "           %s\n"
            "         </DataItem>\n"
            "       </Attribute>\n"
            "    </Grid>\n"
            "  </Domain>\n"
            "</Xdmf>\n") % (time, ncell, 4 * ncell, xyz_base, ncell, attr_base)


# --------------------------------------------------------------------------------------------
# Penalisation phase.  A shape is the reference's per-shape Obstacle list
# (main.cpp:3283-3286, 4245-4263): ids[nob] = block positions in `infos` order, X[nob, 8, 8] = the shape's own chi,
# udef[nob, 8, 8, 2] = its deformation velocity.
# --------------------------------------------------------------------------------------------
def _cell_centres(order, ids, h0, level):
    """p = origin + h (i + 0.5) of the cells of blocks `ids` (main.cpp:695-696, 6667-6668) -> px[nob,1,8], py[nob,8,1]"""
    order = np.asarray(order)
    h = h0 / (1 << level)
    ox = (order[ids, 0] * BS) * h0 / (1 << level)
    oy = (order[ids, 1] * BS) * h0 / (1 << level)
    k = np.arange(BS) + 0.5
    return ox[:, None, None] + (h * k)[None, None, :], oy[:, None, None] + (h * k)[None, :, None]


def shape_integrals(u, v, order, h0, level, ids, X, udef, lam, dt, cx, cy):
    """main.cpp:6643-6681: {PM, PJ, PX, PY, UM, VM, AM}, summed block by block, cell by cell in the reference's
    order (exactly reproducible with one OpenMP thread)."""
    h = h0 / (1 << level)
    hsq = h * h
    lambdt = lam * dt
    vel = to_blocks((u, v), np.asarray(order)[ids], 2).reshape(len(ids), BS, BS, 2)
    px, py = _cell_centres(order, ids, h0, level)
    px = px - cx + np.zeros((1, BS, 1))
    py = py - cy + np.zeros((1, 1, BS))
    Xl = np.where(X >= 0.5, lambdt, 0.0)
    F = hsq * Xl / (1 + Xl)
    du, dv = vel[..., 0] - udef[..., 0], vel[..., 1] - udef[..., 1]
    terms = [F, F * (px * px + py * py), F * px, F * py, F * du, F * dv, F * (px * dv - py * du)]
    sel = (X > 0).reshape(-1)
    out = []
    for t in terms:
        acc = 0.0
        for x in t.reshape(-1)[sel]:
            acc += x
        out.append(acc)
    return np.array(out)


def rigid_motion(Q):
    """main.cpp:6690-6703: (u, v, omega) of the shape from the 3x3 system (the reference uses GSL's LU)"""
    PM, PJ, PX, PY, UM, VM, AM = Q
    A = np.array([[PM, 0, -PY], [0, PM, PX], [-PY, PX, PJ]])
    return np.linalg.solve(A, np.array([UM, VM, AM]))


def penalize(u, v, chi, order, h0, level, shapes, lam, dt):
    """main.cpp:6944-6979.  shapes: dicts with ids, X, udef, cx, cy, u, v, omega; applied in order."""
    order = np.asarray(order)
    u, v = u.copy(), v.copy()
    for sh in shapes:
        ids, X, udef = sh["ids"], sh["X"], sh["udef"]
        CHI = to_blocks(chi, order[ids], 1).reshape(len(ids), BS, BS)
        V = to_blocks((u, v), order[ids], 2).reshape(len(ids), BS, BS, 2)
        px, py = _cell_centres(order, ids, h0, level)
        px, py = px - sh["cx"], py - sh["cy"]
        alpha = np.where(X > 0.5, 1 / (1 + lam * dt), 1.0)
        US = sh["u"] - sh["omega"] * py + udef[..., 0]
        VS = sh["v"] + sh["omega"] * px + udef[..., 1]
        on = ~(CHI > X) & ~(X <= 0)
        nu_ = np.where(on, alpha * V[..., 0] + (1 - alpha) * US, V[..., 0])
        nv_ = np.where(on, alpha * V[..., 1] + (1 - alpha) * VS, V[..., 1])
        for k, (i, j) in enumerate(order[ids]):
            u[j * BS:(j + 1) * BS, i * BS:(i + 1) * BS] = nu_[k]
            v[j * BS:(j + 1) * BS, i * BS:(i + 1) * BS] = nv_[k]
    return u, v


def udef_assemble(chi, order, shapes):
    """main.cpp:6980-7002: tmpV = 0, then += udef of every shape where its chi is not below the chi field"""
    order = np.asarray(order)
    udu, udv = np.zeros_like(chi), np.zeros_like(chi)
    for sh in shapes:
        ids, X, udef = sh["ids"], sh["X"], sh["udef"]
        CHI = to_blocks(chi, order[ids], 1).reshape(len(ids), BS, BS)
        on = ~(X < CHI)
        for k, (i, j) in enumerate(order[ids]):
            udu[j * BS:(j + 1) * BS, i * BS:(i + 1) * BS] += np.where(on[k], udef[k, :, :, 0], 0.0)
            udv[j * BS:(j + 1) * BS, i * BS:(i + 1) * BS] += np.where(on[k], udef[k, :, :, 1], 0.0)
    return udu, udv


def grad_p(p, h, dt):
    """pressureCorrectionKernel::operator() (main.cpp:6021-6043)"""
    NY, NX = p.shape
    pp = pad_scalar(p, 1)
    pfac = -0.5 * dt * h
    cy, cx = slice(1, 1 + NY), slice(1, 1 + NX)
    return pfac * (pp[cy, 2:] - pp[cy, :-2]), pfac * (pp[2:, cx] - pp[:-2, cx])


# --------------------------------------------------------------------------------------------
# Poisson: block-Jacobi preconditioned BiCGSTAB (cuda.cu:403-548), matrix-free on a uniform grid
# --------------------------------------------------------------------------------------------
def build_P_inv():
    """main.cpp:46-57 (getA_local) + 6451-6488: P_inv = -(A_loc)^-1 via dense Cholesky."""
    n = BS * BS
    A = np.zeros((n, n))
    for I1 in range(n):
        j1, i1 = divmod(I1, BS)
        for I2 in range(n):
            j2, i2 = divmod(I2, BS)
            if I1 == I2:
                A[I1, I2] = 4.0
            elif abs(i1 - i2) + abs(j1 - j2) == 1:
                A[I1, I2] = -1.0
    Lc = np.linalg.cholesky(A)
    Linv = np.linalg.solve(Lc, np.eye(n))
    return -(Linv.T @ Linv)


_P_INV = None


def precond(v):
    """cuda.cu:484-486: z_blk = P_inv^T v_blk for every 8x8 block (within-block index 8*iy+ix)."""
    global _P_INV
    if _P_INV is None:
        _P_INV = build_P_inv()
    NY, NX = v.shape
    nby, nbx = NY // BS, NX // BS
    blk = v.reshape(nby, BS, nbx, BS).transpose(0, 2, 1, 3).reshape(nby, nbx, BS * BS)
    z = blk @ _P_INV.T
    return z.reshape(nby, nbx, BS, BS).transpose(0, 2, 1, 3).reshape(NY, NX)


def bicgstab(b, x0, max_error=0.0, max_rel_error=0.0, max_restarts=0, max_iter=1000,
             A=laplacian_neumann, M=precond):
    """BiCGSTABSolver::main (cuda.cu:403-548), same operation order, eps guards and stopping rule.
    Returns (x_opt, iterations, error_opt)."""
    eps = 1e-21
    alpha = beta = omega = rho_prev = rho_curr = 1.0
    x = x0.copy()
    r = b - A(x)                                   # cuda.cu:412-415
    error = error_init = error_opt = np.abs(r).max()  # 420-430
    x_opt = x.copy()
    rhat = r.copy()
    nu = np.zeros_like(b)
    p = np.zeros_like(b)
    restarts = 0
    k = 0
    while k < max_iter:                            # cuda.cu:438
        rho_curr = float(np.vdot(rhat, r))         # 440
        nr2 = float(np.vdot(r, r))
        nrh2 = float(np.vdot(rhat, rhat))
        breakdown = rho_curr * rho_curr < 1e-16 * nr2 * nrh2   # 452-454
        beta = (rho_curr / (rho_prev + eps)) * (alpha / (omega + eps))  # 315-318
        if breakdown and max_restarts > 0:         # 457-477
            restarts += 1
            if restarts >= max_restarts:
                break
            rhat = r.copy()
            rho_curr = float(np.vdot(rhat, rhat))
            nu[:] = 0
            p[:] = 0
            rho_prev = alpha = omega = 1.0
            beta = (rho_curr / (rho_prev + eps)) * (alpha / (omega + eps))
        p = beta * (p - omega * nu) + r            # 478-483
        z = M(p)                                   # 484
        nu = A(z)                                  # 487
        alpha = rho_curr / (float(np.vdot(rhat, nu)) + eps)  # 488-496
        x = x + alpha * z                          # 498
        r = r - alpha * nu                         # 502
        z = M(r)                                   # 503
        t = A(z)                                   # 506
        omega = float(np.vdot(t, r)) / (float(np.vdot(t, t)) + eps)  # 507-518
        x = x + omega * z                          # 520
        r = r - omega * t                          # 524
        error = np.abs(r).max()                    # 525-534
        k += 1
        if error < error_opt:                      # 535-541
            error_opt = error
            x_opt = x.copy()
            if error <= max_error or error / error_init <= max_rel_error:
                break
        rho_prev = rho_curr                        # 325-327
    return x_opt, k, error_opt


# --------------------------------------------------------------------------------------------
# One full time step of the hot path without bodies (main.cpp:6576-7187 minus the OUT-of-scope parts)
# --------------------------------------------------------------------------------------------
def step(u, v, pres, nu, cfl, extent=1.0, kiter=1000, tol=0.0, tol_rel=0.0, max_restarts=100,
         chi=None, udef=None, dt=None):
    """Returns dict(dt, u, v, p, b, x, iters)."""
    N = max(u.shape)  # h0 = extent / max(bpdx, bpdy) / 8 / 2^level  (main.cpp:6338)
    h = extent / N
    if dt is None:
        dt = compute_dt(u, v, h, nu, cfl)
    u, v = rk2(u, v, h, nu, dt)
    if chi is None:
        chi = np.zeros_like(u)
    if udef is None:
        udef = (np.zeros_like(u), np.zeros_like(u))
    tmp = pressure_rhs(u, v, udef[0], udef[1], chi, h, dt)   # main.cpp:7011
    pold = pres.copy()                                        # 7016-7021
    tmp = pressure_rhs1(tmp, pold)                            # 7026
    x, iters, err = bicgstab(tmp, np.zeros_like(tmp), tol, tol_rel, max_restarts, kiter)  # 7114-7118
    # main.cpp:7120-7173 (uniform h: the h^2 weights cancel)
    p = x - x.sum() / x.size
    p = p + (pold - p.sum() / p.size)
    gu, gv = grad_p(p, h, dt)                                 # 7178
    ih2 = 1.0 / h / h                                         # 7180-7187
    u = u + gu * ih2
    v = v + gv * ih2
    return dict(dt=dt, u=u, v=v, p=p, b=tmp, x=x, iters=iters, err=err)
