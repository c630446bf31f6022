"""The GPU parity tests, run on the CPU against a g++ build of the PRODUCT sources (tests/host_emu/build.py build_full: every
.cu of cup2d_b200/csrc with kernel launches turned into one OS thread per CUDA thread — real barriers, warp shuffles, shared
memory, atomics —, CUDA runtime calls mapped to malloc/memcpy and the handful of PTX statements (mbarrier / bulk copy /
reciprocal seed / system-scope loads and stores) replaced by their host meaning).

This checks the LOGIC of the kernels and of the host driver without a GPU — indexing, tables, reductions, the Krylov control
flow — not their execution on hardware (no launch configuration, no real asynchrony, no performance).  It is test
infrastructure: built on demand under tests/, never shipped, and not a CPU fallback of the product (cup2d_b200 still refuses to
run without its CUDA library; the emulated library is only ever loaded through the explicit CUP2D_B200_LIB override below).

A fast subset runs here; the whole suite: tools/run_emulated_gpu_tests.sh (about 5 minutes)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

SUBSET = ("operators_vs_reference_golden or vorticity_tagging or adapt_tags or dump_files or penalisation_phase or shape_calls "
          "or steps_L2_random_k8 or rectangular_domain or amr_fast or amr_advect_diffuse or amr_pressure_gradient")


@pytest.fixture(scope="module")
def emulated_library():
    sys.path.insert(0, os.path.join(HERE, "host_emu"))
    import build
    return build.build_full()


def test_gpu_parity_subset_on_the_emulated_library(emulated_library):
    env = dict(os.environ, CUP2D_B200_LIB=emulated_library, CUP2D_TEST_UNVALIDATED="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_gpu_parity.py"),
                        os.path.join(HERE, "test_gpu_amr.py"), "-m", "gpu", "-q", "-x", "-k", SUBSET, "-p", "no:cacheprovider"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=900, cwd=ROOT)
    tail = r.stdout[-1500:]
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail
    n = int(tail.split(" passed")[0].split()[-1])
    assert n >= 12, tail
