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
