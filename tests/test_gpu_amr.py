"""GPU tests of the first multi-level device path (csrc/amr_ops.cu) against the reference's own flux-corrected operator
outputs on its 7-level run.sh mesh (tests/golden/amrlab_lmax8.npz).

This path was written after round 1's GPU budget was spent and has NOT been run on hardware yet: the tests are skipped
unless CUP2D_TEST_UNVALIDATED=1, so that the suite reports only what has actually been validated."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("CUP2D_TEST_UNVALIDATED") != "1",
                                 reason="csrc/amr_ops.cu has not been run on hardware yet (set CUP2D_TEST_UNVALIDATED=1)")]


@pytest.fixture(scope="module")
def case(golden_dir):
    from cup2d_b200.amr import AmrSimulation
    d = np.load(os.path.join(golden_dir, "amrlab_lmax8.npz"))
    sim = AmrSimulation(d["blocks"], int(d["bpdx"]), int(d["bpdy"]), float(d["h0"]), float(d["nu"]))
    yield d, sim
    sim.close()


def rel(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


def test_amr_advect_diffuse_with_flux_correction(case):
    d, sim = case
    sim.upload("vel", d["vel"])
    sim.advect_diffuse_rhs(float(d["dt"]))
    assert rel(sim.download("tmpV"), d["adv"]) < 1e-12


def test_amr_pressure_rhs_and_laplacian(case):
    d, sim = case
    sim.upload("vel", d["vel"])
    sim.upload("tmpV", d["udef"])
    sim.upload("chi", d["chi"])
    sim.upload("pold", d["pres"])
    sim.pressure_rhs(float(d["dt"]), with_laplacian=False)
    assert rel(sim.download("tmp"), d["rhs"]) < 1e-12
    sim.upload("tmpV", d["udef"])
    sim.pressure_rhs(float(d["dt"]), with_laplacian=True)
    assert rel(sim.download("tmp"), d["rhs1"]) < 1e-12


def test_amr_pressure_gradient(case):
    d, sim = case
    sim.upload("pres", d["pres"])
    sim.pressure_gradient(float(d["dt"]))
    assert rel(sim.download("tmpV"), d["gradp"]) < 1e-12
