#!/usr/bin/env python3
"""BASELINE config 4: Orszag-Tang vortex, N^3, PPMC + HLLD + CT, fp64, one GPU: this library (strict, product) vs the
reference's own -DMHD kernels (oracle/_ref).  usage: bench_mhd.py [N] [steps]   (each arm in its own process: the
reference sizes global scratch at first use)"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ARM = os.environ.get("MHD_ARM")
if ARM is None:
    for arm in ("strict", "fast", "reference"):
        p = subprocess.run([sys.executable, __file__, str(N), str(STEPS)], env=dict(os.environ, MHD_ARM=arm),
                           capture_output=True, text=True)
        print((p.stdout.strip().splitlines() or ["%s failed: %s" % (arm, p.stderr[-400:])])[-1])
    sys.exit(0)
import numpy as np, torch
import cholla_b200 as cb
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for
gamma, flags = 5.0 / 3.0, [1] * 6
g = n_ghost_for(cb.PPMC, mhd=True)
n = (N, N, N)
full = ic.orszag_tang(n, g, gamma)
dims = (N + 2 * g,) * 3
if ARM == "reference":
    from oracle import reference
    t = torch.from_numpy(np.ascontiguousarray(full.reshape(-1))).cuda()
    d = 1.0 / N
    reference.advance(t, dims, g, d, d, d, gamma, 2, 2, flags, 2)
    torch.cuda.synchronize()
    _, _, ms = reference.advance(t, dims, g, d, d, d, gamma, 2, 2, flags, STEPS, timed=True)
    print(json.dumps(dict(arm=ARM, N=N, ms_per_step=ms / STEPS, gcells=N ** 3 * STEPS / ms / 1e6)))
else:
    G = Grid3D(n, (1.0, 1.0, 1.0), recon=cb.PPMC, solver=cb.HLLD, gamma=gamma, bc=flags, strict=(ARM == "strict"), mhd=True)
    G.upload(full)
    G.Set_Boundary_Conditions_Grid()
    G.advance(3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    G.advance(STEPS)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps(dict(arm=ARM, N=N, ms_per_step=ms / STEPS, gcells=N ** 3 * STEPS / ms / 1e6)))
