"""Full-size parity (BASELINE.json configs 2 and 4 at their quoted 256^3): strict arithmetic against the reference's OWN
kernels (oracle/_ref, unmodified sources compiled for sm_100) on the same bytes, bit for bit, dt history included.
The CPU oracle is left out at this size (16.8 M cells x 0.8 M cell-updates/s per core); it is pinned to the same
reference kernels at small sizes by tests/test_gpu_reference.py / test_gpu_mhd.py.

Reference system tests these mirror: src/system_tests/hydro_system_tests.cpp:36-68 (Sod, fiducial snapshot) and :99-138
(sound wave), src/system_tests/mhd_system_tests.cpp:703-787 (Orszag-Tang and friends against fiducial snapshots; the
cholla-tests-data snapshots are LFS pointers here, so the fiducial IS the reference's own kernels run side by side).

Each case runs in a fresh subprocess: the reference sizes its global scratch at the first call (VL_3D_cuda.cu:58-133)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys, hashlib
import numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from oracle import reference
import cholla_b200 as cb
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for

case, S, steps = %(case)r, %(size)d, %(steps)d
n = (S, S, S)
mhd = case == "orszag_tang"
recon, solver = cb.PPMC, (cb.HLLD if mhd else cb.HLLC)
g = n_ghost_for(recon, mhd=mhd)
if case == "sound_wave":     # examples/3D/sound_wave.txt
    lengths, lo, flags, gamma = (4.0, 4.0, 4.0), (0.0, 0.0, 0.0), [1] * 6, 1.666666666666667
    fields = ic.sound_wave(n, lengths=lengths, lo=lo, gamma=gamma)
elif case == "sod":          # examples/3D/sod256.txt: transmissive faces, diaphragm at x = 0.5
    lengths, lo, flags, gamma = (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), [3] * 6, 1.4
    fields = ic.riemann(n, lengths=lengths, lo=lo, gamma=gamma)
else:                        # examples/3D/orszag_tang_vortex.txt
    lengths, lo, flags, gamma = (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), [1] * 6, 1.666666666666667
    fields = None
G = Grid3D(n, lengths, bounds_min=lo, recon=recon, solver=solver, gamma=gamma, bc=flags, strict=True, mhd=mhd)
H = G.H
if mhd:
    G.upload(ic.orszag_tang(n, g, gamma=gamma).reshape(-1))
else:
    G.Set_Initial_Conditions(fields)
G.Set_Boundary_Conditions_Grid()           # both arms start from the same bytes, ghosts included
ref = G.C_device.clone()
dts = G.advance(steps)
mine = G.C_device.clone()
bad = G.bad_cells()
G.close(); del G
torch.cuda.empty_cache()
dims = (H.nx, H.ny, H.nz)
dts_ref, _, _ = reference.advance(ref, dims, g, H.dx, H.dy, H.dz, gamma, recon, solver, flags, steps)
torch.cuda.synchronize()
nf = H.n_fields
a = mine.view(nf, H.nz, H.ny, H.nx)[:, g:-g, g:-g, g:-g]
b = ref.view(nf, H.nz, H.ny, H.nx)[:, g:-g, g:-g, g:-g]
equal = bool(torch.equal(a, b))
l1 = float(((a - b).abs().sum() / b.abs().sum()).item())
print("RESULT " + json.dumps(dict(case=case, cells=S ** 3, steps=steps, bit_equal=equal, rel_l1=l1,
                                  dt_equal=bool(dts == dts_ref), bad_cells=bad, whole_equal=bool(torch.equal(mine, ref)),
                                  sha=hashlib.sha256(a.contiguous().cpu().numpy().tobytes()).hexdigest()[:16])))
"""


def run_child(**kw):
    kw["root"] = ROOT
    p = subprocess.run([sys.executable, "-c", CHILD % kw], capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def _need(recon, solver):
    from oracle import reference
    if not reference.available(recon, solver):
        pytest.skip("oracle/_ref library for this combination was not built")


@pytest.mark.parametrize("case,steps", [("sound_wave", 8), ("sod", 8)])
def test_config2_256_cubed_bit_parity_with_reference_kernels(case, steps):
    _need(2, 0)
    r = run_child(case=case, size=256, steps=steps)
    assert r["bit_equal"] and r["dt_equal"] and r["whole_equal"], r
    assert r["bad_cells"] == 0, r


def test_config4_orszag_tang_256_cubed_bit_parity_with_reference_kernels():
    _need(2, 2)
    r = run_child(case="orszag_tang", size=256, steps=6)
    assert r["bit_equal"] and r["dt_equal"], r
