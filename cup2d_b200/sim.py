"""Host-side mirror of the reference driver for the hot path (uniform grid, no bodies).

Configuration names follow the reference CLI (main.cpp:6321-6337): bpdx, bpdy, level (= levelStart
with levelMax = level+1, i.e. a uniform grid), extent, nu, CFL.  Fields cross the boundary as host
numpy arrays in the reference block layout; `to_blocks/from_blocks` convert from/to global 2-D arrays.
"""
import ctypes as C
import os

import numpy as np

from . import lib as _l

BS = 8


def block_order(bpdx, bpdy, level):
    """(i,j) of all blocks in the reference's Hilbert id order (SpaceCurve, main.cpp:342-446)."""
    lib = _l.load_library()
    n = (bpdx << level) * (bpdy << level)
    out = np.empty((n, 2), dtype=np.int32)
    _l.check(lib.cup2d_block_order(bpdx, bpdy, level, out.ctypes.data_as(C.POINTER(C.c_int32))))
    return out


def to_blocks(comps, order, nbx):
    """tuple of global arrays a[iy,ix] (1 or 2 components) -> flat reference block layout."""
    comps = comps if isinstance(comps, (tuple, list)) else (comps,)
    dim = len(comps)
    nby = comps[0].shape[0] // BS
    out = np.empty((len(order), BS, BS, dim))
    for c, a in enumerate(comps):
        blk = a.reshape(nby, BS, nbx, BS).transpose(0, 2, 1, 3)  # [j, i, iy, ix]
        out[:, :, :, c] = blk[order[:, 1], order[:, 0]]
    return np.ascontiguousarray(out.reshape(-1))


def from_blocks(flat, order, nbx, nby, dim):
    blk = np.asarray(flat).reshape(len(order), BS, BS, dim)
    outs = []
    for c in range(dim):
        g = np.empty((nby, nbx, BS, BS))
        g[order[:, 1], order[:, 0]] = blk[:, :, :, c]
        outs.append(g.transpose(0, 2, 1, 3).reshape(nby * BS, nbx * BS))
    return outs[0] if dim == 1 else tuple(outs)


class Simulation:
    def __init__(self, level, bpdx=1, bpdy=1, extent=1.0, nu=1e-3, cfl=0.5, device=0, rank=0, nranks=1,
                 order=None):
        self.lib = _l.load_library()
        self.bpdx, self.bpdy, self.level = bpdx, bpdy, level
        self.nbx, self.nby = bpdx << level, bpdy << level
        self.NX, self.NY = self.nbx * BS, self.nby * BS
        self.h = extent / max(bpdx, bpdy) / BS / (1 << level)   # main.cpp:6338 (h0) / 2^level
        self.nu, self.cfl = nu, cfl
        self.rank, self.nranks = rank, nranks
        self.order = np.ascontiguousarray(order if order is not None else block_order(bpdx, bpdy, level),
                                          dtype=np.int32)
        nglob = len(self.order)
        # contiguous SFC ranges, remainder to the first ranks (main.cpp:6494-6504)
        base, rem = divmod(nglob, nranks)
        counts = [base + (1 if r < rem else 0) for r in range(nranks)]
        self.rank_begin = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.gbegin, self.gend = int(self.rank_begin[rank]), int(self.rank_begin[rank + 1])
        self.local_order = self.order[self.gbegin:self.gend]
        cfg = _l.Config(self.nbx, self.nby, nglob, self.order.ctypes.data_as(C.POINTER(C.c_int32)), rank,
                        nranks, self.rank_begin.ctypes.data_as(C.POINTER(C.c_int64)), self.h, nu, cfl,
                        device, 0)
        self._h = C.c_void_p()
        _l.check(self.lib.cup2d_create(C.byref(cfg), C.byref(self._h)))
        self.nloc = self.lib.cup2d_nblocks_local(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.cup2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- fields -----------------------------------------------------------------------------
    def upload_blocks(self, name, flat):
        flat = np.ascontiguousarray(flat, dtype=np.float64)
        f = _l.FIELDS[name]
        assert flat.size == self.nloc * 64 * _l.FIELD_DIM[f]
        _l.check(self.lib.cup2d_field_upload(self._h, f, flat.ctypes.data))
        _l.check(self.lib.cup2d_sync(self._h))

    def download_blocks(self, name, out=None):
        f = _l.FIELDS[name]
        if out is None:
            out = np.empty(self.nloc * 64 * _l.FIELD_DIM[f])
        _l.check(self.lib.cup2d_field_download(self._h, f, out.ctypes.data))
        return out

    def upload(self, name, *comps):
        """global arrays a[iy,ix] (this rank's blocks are cut out)."""
        self.upload_blocks(name, to_blocks(comps, self.local_order, self.nbx))

    def download(self, name):
        """-> global array(s); with several ranks only this rank's blocks are filled (others NaN)."""
        f = _l.FIELDS[name]
        dim = _l.FIELD_DIM[f]
        flat = self.download_blocks(name)
        if self.nranks == 1:
            return from_blocks(flat, self.local_order, self.nbx, self.nby, dim)
        blk = flat.reshape(self.nloc, BS, BS, dim)
        outs = []
        for c in range(dim):
            g = np.full((self.nby, self.nbx, BS, BS), np.nan)
            g[self.local_order[:, 1], self.local_order[:, 0]] = blk[:, :, :, c]
            outs.append(g.transpose(0, 2, 1, 3).reshape(self.NY, self.NX))
        return outs[0] if dim == 1 else tuple(outs)

    def device_ptr(self, name):
        return self.lib.cup2d_field_device_ptr(self._h, _l.FIELDS[name])

    # ---- operators --------------------------------------------------------------------------
    def compute_dt(self):
        um, dt = C.c_double(), C.c_double()
        _l.check(self.lib.cup2d_compute_dt(self._h, C.byref(um), C.byref(dt)))
        return um.value, dt.value

    def advect_diffuse_rhs(self, dt, src="vel", dst="tmpV"):
        _l.check(self.lib.cup2d_advect_diffuse_rhs(self._h, _l.FIELDS[src], _l.FIELDS[dst], dt))

    def advect_diffuse_stage(self, src, old, dst, coef, dt):
        _l.check(self.lib.cup2d_advect_diffuse_stage(self._h, _l.FIELDS[src], _l.FIELDS[old], _l.FIELDS[dst], coef, dt))

    def rk2(self, dt):
        _l.check(self.lib.cup2d_advect_diffuse_rk2(self._h, dt))

    def pressure_rhs(self, dt):
        _l.check(self.lib.cup2d_pressure_rhs(self._h, dt))

    def poisson_solve(self, tol_abs=0.0, tol_rel=0.0, max_restarts=0, max_iter=1000):
        it, err = C.c_int(), C.c_double()
        _l.check(self.lib.cup2d_poisson_solve(self._h, tol_abs, tol_rel, max_restarts, max_iter, C.byref(it), C.byref(err)))
        return it.value, err.value

    def vorticity_tag(self):
        """tmp = vorticity(vel); returns max|tmp| per local block (reference infos[] order)."""
        out = np.empty(self.nloc)
        _l.check(self.lib.cup2d_vorticity_tag(self._h, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def adapt_tags(self, rtol, chi_cells):
        """adapt()'s criterion (vorticity + body proximity); returns max|tmp| per local block."""
        out = np.empty(self.nloc)
        _l.check(self.lib.cup2d_adapt_tags(self._h, float(rtol), int(chi_cells), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def dump(self, time, path):
        """the reference's dump(): <path>.xyz.raw, .attr.raw, .xdmf2"""
        _l.check(self.lib.cup2d_dump(self._h, float(time), os.fsencode(path)))

    def shape_set(self, shape, ids, X, udef):
        """upload one shape's obstacle blocks: ids[nob] local block ids, X[nob,8,8], udef[nob,8,8,2]"""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        X = np.ascontiguousarray(X, dtype=np.float64)
        udef = np.ascontiguousarray(udef, dtype=np.float64)
        assert X.size == 64 * len(ids) and udef.size == 128 * len(ids)
        dp = C.POINTER(C.c_double)
        _l.check(self.lib.cup2d_shape_set(self._h, shape, len(ids), ids.ctypes.data_as(C.POINTER(C.c_int32)),
                                          X.ctypes.data_as(dp), udef.ctypes.data_as(dp)))

    def shape_integrals(self, shape, lam, dt, cx, cy):
        out = np.empty(7)
        _l.check(self.lib.cup2d_shape_integrals(self._h, shape, lam, dt, cx, cy, out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def penalize(self, shape, lam, dt, cx, cy, us, vs, omega):
        _l.check(self.lib.cup2d_penalize(self._h, shape, lam, dt, cx, cy, us, vs, omega))

    def udef_assemble(self):
        _l.check(self.lib.cup2d_udef_assemble(self._h))

    def pressure_correct(self, dt):
        _l.check(self.lib.cup2d_pressure_correct(self._h, dt))

    def step(self, dt=0.0, keep_udef=False, tol_abs=0.0, tol_rel=0.0, max_restarts=100, max_iter=1000):
        dto, it, err = C.c_double(), C.c_int(), C.c_double()
        _l.check(self.lib.cup2d_step(self._h, dt, int(keep_udef), tol_abs, tol_rel, max_restarts, max_iter,
                                     C.byref(dto), C.byref(it), C.byref(err)))
        return dto.value, it.value, err.value

    def step_enqueue(self, dt=0.0, keep_udef=False, tol_abs=0.0, tol_rel=0.0, max_restarts=100, max_iter=1000):
        """the same step without waiting for the device (one CUDA graph launch from the second step on);
        dt <= 0: dt control runs on the device inside the step"""
        _l.check(self.lib.cup2d_step_enqueue(self._h, dt, int(keep_udef), tol_abs, tol_rel, max_restarts, max_iter))

    def step_result(self):
        """waits for the enqueued steps; (dt, iterations, residual) of the last one"""
        dto, it, err = C.c_double(), C.c_int(), C.c_double()
        _l.check(self.lib.cup2d_step_result(self._h, C.byref(dto), C.byref(it), C.byref(err)))
        return dto.value, it.value, err.value

    def set_graph(self, on):
        _l.check(self.lib.cup2d_set_graph(self._h, int(on)))

    # ---- host-buffer pipeline (cup2d_pipe_*): independent steps with inputs and results in host memory ----
    PIPE_SLOTS = 4

    def pipe_upload(self, slot, vel_ptr, pres_ptr):
        _l.check(self.lib.cup2d_pipe_upload(self._h, slot, vel_ptr, pres_ptr))

    def pipe_step(self, slot, dt=0.0, tol_abs=0.0, tol_rel=0.0, max_restarts=100, max_iter=1000):
        dto, it, err = C.c_double(), C.c_int(), C.c_double()
        _l.check(self.lib.cup2d_pipe_step(self._h, slot, dt, tol_abs, tol_rel, max_restarts, max_iter,
                                          C.byref(dto), C.byref(it), C.byref(err)))
        return dto.value, it.value, err.value

    def pipe_download(self, slot, vel_ptr, pres_ptr):
        _l.check(self.lib.cup2d_pipe_download(self._h, slot, vel_ptr, pres_ptr))

    def pipe_wait(self, slot):
        _l.check(self.lib.cup2d_pipe_wait(self._h, slot))

    def pipelined_steps(self, jobs, **step_args):
        """Run independent steps whose inputs and results are host buffers, overlapping upload(n+1), step(n) and
        download(n-1).  `jobs` yields (vel_in_ptr, pres_in_ptr, vel_out_ptr, pres_out_ptr) addresses of page-locked host
        buffers; returns the per-step (dt, iterations, error).  An output buffer may be reused after PIPE_SLOTS - 1
        further jobs; all of them are complete when this returns."""
        jobs = iter(jobs)
        out, n, nxt = [], 0, next(jobs, None)
        if nxt is not None:
            self.pipe_upload(0, nxt[0], nxt[1])
        while nxt is not None:
            cur, slot = nxt, n % self.PIPE_SLOTS
            nxt = next(jobs, None)
            if nxt is not None:  # the next step's inputs travel while this step computes
                self.pipe_upload((n + 1) % self.PIPE_SLOTS, nxt[0], nxt[1])
            out.append(self.pipe_step(slot, **step_args))
            self.pipe_download(slot, cur[2], cur[3])
            n += 1
        for slot in range(min(n, self.PIPE_SLOTS)):
            self.pipe_wait(slot)
        return out

    def sync(self):
        _l.check(self.lib.cup2d_sync(self._h))

    def launch_count(self):
        return self.lib.cup2d_launch_count(self._h)

    def profile(self, on):
        _l.check(self.lib.cup2d_profile_enable(self._h, int(on)))

    def profile_read(self):
        """-> {kernel class: (total ms, launches)} since profile(True)."""
        n = 16
        names = C.create_string_buffer(32 * n)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        k = self.lib.cup2d_profile_read(self._h, n, names, ms, cnt)
        if k < 0:
            _l.check(k)
        return {names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode(): (ms[i], cnt[i]) for i in range(k)}

    @property
    def stream(self):
        return self.lib.cup2d_stream(self._h)

    # ---- multi-GPU plumbing (torch.distributed only moves the opaque IPC blobs) ---------------
    def attach_peers(self, dist=None):
        n = self.lib.cup2d_peer_blob_size()
        blob = (C.c_ubyte * n)()
        _l.check(self.lib.cup2d_peer_export(self._h, blob))
        if self.nranks == 1:
            allb = bytes(blob)
        else:
            gathered = [None] * self.nranks
            dist.all_gather_object(gathered, bytes(blob))
            allb = b"".join(gathered)
        buf = (C.c_ubyte * len(allb)).from_buffer_copy(allb)
        _l.check(self.lib.cup2d_peer_attach(self._h, buf))
        if dist is not None and self.nranks > 1:
            dist.barrier()
