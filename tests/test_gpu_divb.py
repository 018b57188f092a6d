"""On-device magnetic-divergence check (chl_mhd_max_divergence <- mhd::checkMagneticDivergence,
src/mhd/magnetic_divergence.cu:33-112) against the CPU oracle's, and the run loop's use of it."""
import numpy as np
import pytest

import cholla_b200 as cb
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for
from oracle import mhd

pytestmark = pytest.mark.gpu


def test_divergence_matches_oracle_and_detects_a_monopole():
    n, g = (24, 20, 16), n_ghost_for(cb.PPMC, mhd=True)
    full = ic.mhd_smooth(n, g, seed=2)
    G = Grid3D(n, (1.0, 1.0, 1.0), recon=cb.PPMC, solver=cb.HLLD, gamma=5.0 / 3.0, bc=[1] * 6, strict=True, mhd=True)
    H = G.H
    G.upload(full.reshape(-1))
    G.Set_Boundary_Conditions_Grid()
    G.advance(3)
    got = G.max_magnetic_divergence()
    want = mhd.max_divergence(G.download(), H.nx, H.ny, H.nz, g, H.dx, H.dy, H.dz)
    assert got == pytest.approx(want, rel=1e-12, abs=1e-300) and got <= 1e-11   # constrained transport holds
    bad = G.download().reshape(8, H.nz, H.ny, H.nx)
    bad[5, H.nz // 2, H.ny // 2, H.nx // 2] += 1.0e-3   # a monopole
    G.upload(bad.reshape(-1))
    d = G.max_magnetic_divergence()
    assert d == pytest.approx(1.0e-3 / H.dx, rel=1e-9) and d > 1e-9
    G.close()


def test_reference_unit_test_fiducial(golden_dir):
    """tMHDGrid3DcheckMagneticDivergence (magnetic_divergence_tests.cu:27-70): a (96 + 2 * 4)^3 grid of std::mt19937(1)
    values, dx = dy = dz = 3; the reference's known answer, the CPU oracle's value and the device reduction agree.  The
    maximum runs over every cell whose low faces exist, ghost cells included, like the reference's kernel."""
    import json
    import os

    from test_oracle_golden import nearly_equal
    from test_oracle_mhd_golden import divergence_test_grid

    with open(os.path.join(golden_dir, "mhd_scalars.json")) as fh:
        t = json.load(fh)["magnetic_divergence"]
    U, n = divergence_test_grid(t)
    nr, dx = t["grid_size"], t["dx"]
    assert n_ghost_for(cb.PLMC, mhd=True) == t["n_ghost"]
    G = Grid3D((nr, nr, nr), (dx * nr, dx * nr, dx * nr), recon=cb.PLMC, solver=cb.HLLD, gamma=5.0 / 3.0, bc=[1] * 6,
               strict=True, mhd=True)
    assert (G.H.nx, G.H.ny, G.H.nz) == (n, n, n) and G.H.dx == dx
    G.upload(U)
    got = G.max_magnetic_divergence()
    G.close()
    fid = t["fiducial_numerator"] / t["fiducial_denominator"]
    assert nearly_equal(got, fid) and got == pytest.approx(fid, rel=1e-13), got
    assert got == pytest.approx(mhd.check_magnetic_divergence(U, n, n, n, dx, dx, dx), rel=1e-14)
