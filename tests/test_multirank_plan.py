"""CPU tests (gloo, world_size 2 and 4) of the host-side multi-rank logic of libcup2d_b200.so:
SFC-range partition, halo plan (who pulls which block from whom), neighbour table and advect tiles.
The data movement itself is emulated with torch.distributed point-to-point on CPU tensors, standing in
for the NVLink peer pulls the GPU kernels do (cup2d_b200/csrc/halo.cu)."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cup2d_b200
from cup2d_b200 import lib as L


def make_plan(level, rank, nranks, bpdx=1, bpdy=1):
    lib = cup2d_b200.load_library()
    order = cup2d_b200.block_order(bpdx, bpdy, level)
    nbx, nby = bpdx << level, bpdy << level
    nglob = len(order)
    base, rem = divmod(nglob, nranks)
    rb = np.concatenate([[0], np.cumsum([base + (1 if r < rem else 0) for r in range(nranks)])]).astype(np.int64)
    cfg = L.Config(nbx, nby, nglob, order.ctypes.data_as(C.POINTER(C.c_int32)), rank, nranks,
                   rb.ctypes.data_as(C.POINTER(C.c_int64)), 1.0 / (8 * nbx), 1e-3, 0.5, -1, 0)
    h = C.c_void_p()
    L.check(lib.cup2d_plan_create(C.byref(cfg), C.byref(h)))

    def table(which):
        n = lib.cup2d_plan_table(h, which, None)
        out = np.empty(n, dtype=np.int32)
        if n:
            lib.cup2d_plan_table(h, which, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    plan = dict(order=order, rank_begin=rb, nloc=int(lib.cup2d_nblocks_local(h)),
                halo_gid=table(0), halo_owner=table(1), halo_src=table(2), nbr=table(3).reshape(-1, 4),
                tiles=table(4).reshape(-1, 32), torg=table(5).reshape(-1, 2), nbx=nbx, nby=nby)
    lib.cup2d_destroy(h)
    return plan


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def worker(rank, world, port, level, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = make_plan(level, rank, world)
        order, rb, nloc = p["order"], p["rank_begin"], p["nloc"]
        gb = int(rb[rank])
        # block payload = a function of the GLOBAL block id, so every value can be checked locally
        def payload(gids):
            return torch.tensor(np.asarray(gids, dtype=np.float64)[:, None] * 10.0 + np.arange(4.0)[None, :])
        field = torch.empty(nloc + len(p["halo_gid"]), 4, dtype=torch.float64)
        field[:nloc] = payload(np.arange(gb, gb + nloc))
        field[nloc:] = float("nan")
        # everybody publishes what it wants to pull: (owner, slot on owner) per halo slot
        wants = [None] * world
        dist.all_gather_object(wants, (p["halo_owner"].tolist(), p["halo_src"].tolist()))
        reqs = []
        for peer in range(world):
            if peer == rank:
                continue
            # serve the peer's pulls from my local blocks
            slots = [s for o, s in zip(*wants[peer]) if o == rank]
            if slots:
                assert max(slots) < nloc and min(slots) >= 0
                reqs.append(dist.isend(field[:nloc][slots].contiguous(), peer))
        recv = {}
        for peer in range(world):
            if peer == rank:
                continue
            mine = [k for k, o in enumerate(p["halo_owner"]) if o == peer]
            if mine:
                buf = torch.empty(len(mine), 4, dtype=torch.float64)
                reqs.append(dist.irecv(buf, peer))
                recv[peer] = (mine, buf)
        for r in reqs:
            r.wait()
        for peer, (mine, buf) in recv.items():
            field[nloc + torch.tensor(mine)] = buf
        # 1. halo slots now hold exactly the blocks the plan says they hold
        assert torch.equal(field[nloc:], payload(p["halo_gid"]))
        # 2. every local block sees its four face neighbours through the neighbour table
        gid_of = -np.ones((p["nby"], p["nbx"]), dtype=np.int64)
        gid_of[order[:, 1], order[:, 0]] = np.arange(len(order))
        d = [(-1, 0), (1, 0), (0, -1), (0, 1)]
        for k in range(nloc):
            i, j = order[gb + k]
            for n, (di, dj) in enumerate(d):
                ii, jj = i + di, j + dj
                slot = p["nbr"][k, n]
                if ii < 0 or ii >= p["nbx"] or jj < 0 or jj >= p["nby"]:
                    assert slot == -1
                else:
                    assert slot >= 0 and field[slot, 0].item() == gid_of[jj, ii] * 10.0
        # 3. tiles: every local block appears in exactly one tile interior, ring slots are neighbours
        seen = np.zeros(nloc, dtype=int)
        for t, ts in enumerate(p["tiles"]):
            bi0, bj0 = p["torg"][t]
            for b in range(16):
                s = ts[b]
                if 0 <= s < nloc:
                    seen[s] += 1
                    assert tuple(order[gb + s]) == (bi0 + b % 4, bj0 + b // 4)
        assert (seen == 1).all()
        # 4. a two-value all-reduce in rank order (what peer_allreduce does over NVLink mailboxes)
        t = torch.tensor([float(rank + 1), float(nloc)], dtype=torch.float64)
        dist.all_reduce(t)
        assert t[0].item() == world * (world + 1) / 2 and t[1].item() == len(order)
        q.put((rank, "ok", len(p["halo_gid"])))
    except Exception as e:  # surface the failure in the parent
        q.put((rank, repr(e), -1))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,level", [(2, 3), (2, 4), (4, 3), (3, 2), (8, 3)])  # 3: uneven SFC ranges; 8: the node size
def test_halo_plan_two_and_four_ranks_gloo(world, level):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, level, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    assert all(r[1] == "ok" for r in res), res
    assert all(r[2] > 0 for r in res)


def test_plan_single_rank_has_no_halo_and_walls_everywhere_needed():
    p = make_plan(2, 0, 1)
    assert len(p["halo_gid"]) == 0 and p["nloc"] == 16
    order = p["order"]
    for k, (i, j) in enumerate(order):
        assert (p["nbr"][k, 0] == -1) == (i == 0) and (p["nbr"][k, 1] == -1) == (i == 3)
        assert (p["nbr"][k, 2] == -1) == (j == 0) and (p["nbr"][k, 3] == -1) == (j == 3)


def test_plan_context_refuses_device_calls():
    lib = cup2d_b200.load_library()
    order = cup2d_b200.block_order(1, 1, 1)
    rb = np.array([0, 4], dtype=np.int64)
    cfg = L.Config(2, 2, 4, order.ctypes.data_as(C.POINTER(C.c_int32)), 0, 1,
                   rb.ctypes.data_as(C.POINTER(C.c_int64)), 1.0 / 16, 1e-3, 0.5, -1, 0)
    h = C.c_void_p()
    L.check(lib.cup2d_plan_create(C.byref(cfg), C.byref(h)))
    assert lib.cup2d_sync(h) != 0 and b"plan-only" in lib.cup2d_last_error()
    lib.cup2d_destroy(h)
