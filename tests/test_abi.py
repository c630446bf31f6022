"""CPU-side checks of the C-ABI library: it loads, exports every symbol the header declares, mirrors
the reference's block ordering, and refuses to run without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

import cup2d_b200
from cup2d_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = cup2d_b200.load_library()
    hdr = open(os.path.join(ROOT, "include", "cup2d_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cup2d_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cup2d_b200.h but not exported"
    assert sorted(L.SYMBOLS) == declared
    assert lib.cup2d_version() >= 100


@pytest.mark.parametrize("lvl", range(6))
def test_block_order_matches_reference_golden(golden_dir, lvl):
    g = np.load(os.path.join(golden_dir, f"order_L{lvl}.npy"))
    assert np.array_equal(cup2d_b200.block_order(1, 1, lvl), g)


def test_block_order_rectangular_is_a_permutation_with_face_adjacency():
    # 2x1 base blocks (run.sh geometry): every block exactly once; consecutive blocks inside one base
    # block are face neighbours (Hilbert property)
    o = cup2d_b200.block_order(2, 1, 3)
    assert len({(i, j) for i, j in o}) == 16 * 8
    d = np.abs(np.diff(o, axis=0)).sum(axis=1)
    assert (d == 1).sum() >= len(o) - 2


def test_blocks_roundtrip_layout():
    lvl = 2
    N = 8 << lvl
    rng = np.random.default_rng(0)
    u, v = rng.normal(size=(N, N)), rng.normal(size=(N, N))
    order = cup2d_b200.block_order(1, 1, lvl)
    flat = cup2d_b200.to_blocks((u, v), order, 1 << lvl)
    i, j = order[5]
    assert flat[5 * 128 + 2 * (8 * 3 + 2) + 1] == v[j * 8 + 3, i * 8 + 2]  # main.cpp:5467-5468
    u2, v2 = cup2d_b200.from_blocks(flat, order, 1 << lvl, 1 << lvl, 2)
    assert np.array_equal(u, u2) and np.array_equal(v, v2)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cup2d_b200.Cup2dError, match="no CPU fallback"):
        cup2d_b200.Simulation(2)


@pytest.mark.parametrize("bx,by,lvl", [(2, 1, 2), (3, 2, 1), (1, 2, 2)])
def test_block_order_rectangular_matches_reference_golden(golden_dir, bx, by, lvl):
    """non-regular space-filling curve (bounding square not filled, main.cpp:6358-6376)"""
    g = np.load(os.path.join(golden_dir, f"order_{bx}x{by}_L{lvl}.npy"))
    assert np.array_equal(cup2d_b200.block_order(bx, by, lvl), g)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the driver's reference arm) on a tiny sample: one JSON line with the contract keys.
    Uses oracle/_ref/ref_harness (CPU): skipped where the reference was not compiled."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "ref_harness")):
        pytest.skip("oracle/_ref/ref_harness not built")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--cpu-level", "4",
                          "--steps", "3", "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                         timeout=300, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""), check=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mcell-updates/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config"):
        assert k in line


def test_multi_rank_constructors_reject_bad_partitions_before_touching_a_device():
    """argument errors of the rank-aware constructors are reported as CUP2D_EINVAL with a message, GPU or not"""
    import ctypes as C
    from cup2d_b200 import lib as L
    lib = L.load_library()
    I32, I64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    nbr = np.full(8, -1, dtype=np.int32)
    h = C.c_void_p()
    for rank, nranks, rb in ((2, 2, [0, 1, 2]), (0, 2, [0, 0, 2]), (0, 2, [1, 1, 2]), (0, 9, list(range(10)))):
        rb_a = np.array(rb, dtype=np.int64)
        rc = lib.cup2d_poisson_create_general_ranks(int(rb[-1]), rank, nranks, rb_a.ctypes.data_as(I64), nbr.ctypes.data_as(I32), 0, None, None,
                                                    None, None, 0, C.byref(h))
        assert rc == -1 and lib.cup2d_last_error(), (rank, nranks, rb)      # CUP2D_EINVAL
    blocks = np.array([[0, 0, 0]], dtype=np.int32)
    a = C.c_void_p()
    for rank, nranks, rb in ((1, 1, [0, 1]), (0, 1, [0, 2]), (0, 2, [0, 1, 1])):
        rb_a = np.array(rb, dtype=np.int64)
        rc = lib.cup2d_amr_create_ranks(1, blocks.ctypes.data_as(I32), 1, 1, 0.125, 1e-3, rank, nranks, rb_a.ctypes.data_as(I64), 0, C.byref(a))
        assert rc != 0 and not a.value, (rank, nranks, rb)
    # a partition that is not increasing ({0, 100, 50} with 50 blocks) is refused before any table is indexed by it
    four = np.array([[1, i, j] for j in range(2) for i in range(2)], dtype=np.int32)
    for nb, rank, nranks, rb in ((4, 0, 2, [0, 8, 4]), (4, 1, 2, [0, 8, 4]), (4, 0, 3, [0, 2, 2, 4])):
        rb_a = np.array(rb, dtype=np.int64)
        rc = lib.cup2d_amr_create_ranks(nb, four.ctypes.data_as(I32), 1, 1, 0.125, 1e-3, rank, nranks, rb_a.ctypes.data_as(I64), 0, C.byref(a))
        assert rc == -1 and lib.cup2d_last_error() and not a.value, (rank, nranks, rb)
