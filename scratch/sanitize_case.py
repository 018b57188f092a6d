"""Tiny hydro (PPMC+HLLC, strict + product) and MHD (PPMC+HLLD) steps for compute-sanitizer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cholla_b200 as cb
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for
for strict in (True, False):
    G = Grid3D((36, 20, 12), (1.0, 1.0, 1.0), recon=cb.PPMC, solver=cb.HLLC, gamma=5.0 / 3.0, bc=[1, 1, 2, 2, 3, 3], strict=strict)
    G.Set_Initial_Conditions(ic.random_smooth((36, 20, 12), seed=2)); G.Set_Boundary_Conditions_Grid(); G.advance(2); G.close()
G = Grid3D((40, 1, 1), (1.0, 1.0, 1.0), recon=cb.PLMC, solver=cb.ROE, gamma=1.4, bc=[3, 3, 0, 0, 0, 0], strict=False)
G.Set_Initial_Conditions(ic.riemann((40, 1, 1), gamma=1.4)); G.Set_Boundary_Conditions_Grid(); G.advance(2); G.close()
g = n_ghost_for(cb.PPMC, mhd=True)
for strict in (True, False):
    G = Grid3D((20, 16, 12), (1.0, 1.0, 1.0), recon=cb.PPMC, solver=cb.HLLD, gamma=5.0 / 3.0, bc=[1] * 6, strict=strict, mhd=True)
    G.upload(ic.mhd_smooth((20, 16, 12), g, seed=1)); G.Set_Boundary_Conditions_Grid(); G.advance(2); G.close()
print("sanitize case done")
