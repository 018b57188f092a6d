import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import oracle, cholla_b200 as cb
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for
from helpers import ghosted, real_view

def one(n, recon, solver, flags, steps=1):
    fields = ic.random_smooth(n, seed=1)
    g = n_ghost_for(recon)
    flat, dims = ghosted(fields, g, flags)
    G = Grid3D(n, (1.0, 1.0, 1.0), recon=recon, solver=solver, gamma=5/3, bc=flags, strict=True)
    G.upload(flat)
    dti = G.Calc_Inverse_Timestep(); G.set_dt(dti)
    G.Execute_Hydro_Integrator()
    got = G.download(); G.close()
    ref = flat.copy()
    oracle.vl_step(ref, *dims, g, G.H.dx, G.H.dy, G.H.dz, G.H.dt, 5/3, recon, solver)
    a, b = real_view(got, dims, g), real_view(ref, dims, g)
    ia, ib = a.view(np.int64), b.view(np.int64)
    d = np.abs(ia - ib)
    bad = np.argwhere(d > 0)
    print(n, "recon", recon, "solver", solver, "max ulp", d.max(), "nbad", len(bad), "of", d.size)
    if len(bad):
        for f in range(5):
            bf = bad[bad[:, 0] == f]
            if len(bf):
                print("  field", f, "n", len(bf), "z", np.unique(bf[:, 1])[:12], "y", np.unique(bf[:, 2])[:12], "x", np.unique(bf[:, 3])[:40])

for n, flags in [((37, 1, 1), [1,1,0,0,0,0]), ((37, 19, 1), [1,1,1,1,0,0]), ((37, 19, 21), [1]*6)]:
    for recon in (0, 1, 2):
        one(n, recon, 0, flags)
