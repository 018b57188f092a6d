"""A few steps of every context added in build round 2 on a small ragged grid, to be run under compute-sanitizer
(memcheck / racecheck): dual energy + scalar with PPMC / PPMP / PLMC, static gravity, potential gravity (five-point), CIE
cooling, 1-D and 2-D dual energy, MHD with the per-launch CTA shapes."""
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cholla_b200 as cb  # noqa: E402
from cholla_b200 import initial_conditions as ic  # noqa: E402
from cholla_b200.grid import Grid3D, n_ghost_for  # noqa: E402


def run(n, recon, solver, ns, de, grav=None, pot=False, cooling=None, strict=True, steps=2):
    G = Grid3D(n, (1.0, 1.0, 1.0), bounds_min=(-0.53, -0.49, -0.51), recon=recon, solver=solver, gamma=5.0 / 3.0,
               bc=[1, 1, 1, 1, 1, 1] if n[2] > 1 else ([1, 1, 1, 1, 0, 0] if n[1] > 1 else [1, 1, 0, 0, 0, 0]), strict=strict,
               n_scalars=ns, dual_energy=de, cooling=cooling)
    G.Set_Initial_Conditions(ic.with_passive(ic.random_smooth(n, seed=1), ns, de, 5.0 / 3.0))
    if grav is not None:
        G.Set_Static_Gravity(grav)
    if pot:
        H = G.H
        G.Set_Gravity_Potential(np.random.default_rng(0).random((H.nz, H.ny, H.nx)), five_point=True)
    G.Set_Boundary_Conditions_Grid()
    G.advance(steps)
    ok = np.isfinite(G.download()).all()
    G.close()
    return ok


def main():
    n = (37, 19, 21)
    res = {}
    for strict in (True, False):
        res["ppmc_hllc_de_s1", strict] = run(n, cb.PPMC, cb.HLLC, 1, True, strict=strict)
        res["ppmp_hllc_de_cool", strict] = run(n, cb.PPMP, cb.HLLC, 0, True, cooling="cie", strict=strict)
        res["plmc_roe_s1_grav", strict] = run(n, cb.PLMC, cb.ROE, 1, False, grav=1, strict=strict)
        res["ppmc_hll_de_pot", strict] = run(n, cb.PPMC, cb.HLL, 0, True, pot=True, strict=strict)
    res["2d_de"] = run((45, 23, 1), cb.PPMC, cb.HLLC, 0, True, grav=2)
    res["1d_de_s1"] = run((64, 1, 1), cb.PLMC, cb.HLLC, 1, True)
    G = Grid3D((34, 22, 18), (1.0, 1.0, 1.0), recon=cb.PPMC, solver=cb.HLLD, gamma=5.0 / 3.0, mhd=True)
    G.upload(ic.mhd_smooth((34, 22, 18), n_ghost_for(cb.PPMC, True), seed=2).reshape(-1))
    G.Set_Boundary_Conditions_Grid()
    G.advance(2)
    res["mhd"] = bool(np.isfinite(G.download()).all())
    G.close()
    print(res)
    assert all(res.values())


if __name__ == "__main__":
    main()
