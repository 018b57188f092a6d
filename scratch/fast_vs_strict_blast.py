#!/usr/bin/env python3
"""How far does product arithmetic drift from the reference's bits on the headline workload?  Blast wave N^3, PPMC+HLLC,
periodic: relative L1 per field between strict (== reference, bit for bit) and product mode after k steps."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cholla_b200 as cb
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n, gamma = (N, N, N), 1.66666667
f = ic.spherical_overpressure(n, gamma=gamma)
G = {}
for strict in (True, False):
    g = Grid3D(n, (1.0, 1.0, 1.0), recon=cb.PPMC, solver=cb.HLLC, gamma=gamma, bc=[1] * 6, strict=strict)
    g.Set_Initial_Conditions(f); g.Set_Boundary_Conditions_Grid()
    G[strict] = g
done = 0
for k in (1, 5, 13, 50, 200):
    for g in G.values(): g.advance(k - done)
    done = k
    a, b = G[True].real_cells(), G[False].real_cells()
    grp = [[0], [1, 2, 3], [4]]
    err = [float(sum(np.abs(a[i] - b[i]).sum() for i in q) / sum(np.abs(a[i]).sum() for i in q)) for q in grp]
    print(json.dumps(dict(N=N, steps=k, rel_l1_density=err[0], rel_l1_momentum=err[1], rel_l1_energy=err[2],
                          max_abs=float(np.abs(a - b).max()))))
