"""The cooling source step of a -DCOOLING_GPU build ("the cooling source step that follows the
integrator", grid3D.cpp:512-522 -> Cooling_Update, src/cooling/cooling_cuda.cu:20-196 with the CIE curve :293-317).
Arms: the reference's own -DCOOLING_GPU kernels (oracle/_ref/*_cool), the CPU oracle, the B200-native kernels.  log10 / pow
come from the device library, so bit equality is asserted between the two GPU arms; the oracle (CPU libm) agrees to
round-off."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import json, sys
import numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import oracle
from oracle import reference
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for
from helpers import ghosted, real_view, rel_l1, max_ulp

recon, n, flags, gamma, steps, de = %(recon)d, %(n)r, %(flags)r, %(gamma)r, %(steps)d, %(de)d
variant = ("_de" if de else "") + "_cool"
# an interstellar gas in the reference's code units (kpc, kyr, M_sun): n_H ~ 0.05 - 5 cm^-3, 10^4 - 10^7.5 K
MP, KB = 1.672622e-24, 1.380658e-16
DU = 1.98847e33 / 3.08567758e21 ** 3
PU = DU * (3.08567758e21 / 3.15569e10) ** 2
shape = (n[2], n[1], n[0])
Z, Y, X = np.meshgrid(*[(np.arange(m) + 0.5) / m for m in shape], indexing="ij")
two_pi = 2.0 * np.pi
# smooth, periodic: densities 0.1 - 1 cm^-3, temperatures 10^4.5 - 10^6.5 K (cooling times from far below to far above dt)
nH = 10.0 ** (-0.5 + 0.5 * np.sin(two_pi * (X + 0.3 * Y)) * np.cos(two_pi * Z))
T = 10.0 ** (5.5 + 1.0 * np.cos(two_pi * (X - Y + 0.5 * Z)))
d = nH * 0.6 * MP / DU
P = nH * KB * T / PU
cs = np.sqrt(gamma * P / d)
v = [0.3 * cs * np.sin(two_pi * (X + Y + Z) + a) if n[a] > 1 else np.zeros(shape) for a in range(3)]
E = P / (gamma - 1.0) + 0.5 * d * (v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
fields = ic.with_passive(np.stack([d, d * v[0], d * v[1], d * v[2], E]), 0, bool(de), gamma)
g = n_ghost_for(recon)
flat, dims = ghosted(fields, g, flags)
lengths = (0.5, 0.5, 0.5)
dd = [lengths[i] / n[i] for i in range(3)]
out = {}
t = torch.from_numpy(flat.copy()).cuda()
dts_ref, _, _ = reference.advance(t, dims, g, dd[0], dd[1], dd[2], gamma, recon, 0, flags, steps, variant=variant)
ref = t.cpu().numpy()
out["flags"] = reference.lib(recon, 0, variant=variant).ref_build_flags().decode()
orc = flat.copy()
dts_orc = oracle.advance(orc, *dims, g, dd[0], dd[1], dd[2], gamma, recon, 0, flags, steps, dual_energy=bool(de), cooling=True)
adiabatic = flat.copy()
oracle.advance(adiabatic, *dims, g, dd[0], dd[1], dd[2], gamma, recon, 0, flags, steps, dual_energy=bool(de))
out["cooling_effect_l1"] = float(rel_l1(real_view(orc, dims, g), real_view(adiabatic, dims, g))[4])
out["oracle"] = dict(l1=float(rel_l1(real_view(orc, dims, g), real_view(ref, dims, g)).max()))
for strict in (True, False):
    G = Grid3D(n, lengths, recon=recon, solver=0, gamma=gamma, bc=flags, strict=strict, dual_energy=bool(de), cooling="cie")
    G.upload(flat)
    dts = G.advance(steps)
    got = G.download(); G.close()
    out["strict" if strict else "fast"] = dict(ulp=max_ulp(real_view(got, dims, g), real_view(ref, dims, g)),
                                               l1=float(rel_l1(real_view(got, dims, g), real_view(ref, dims, g)).max()),
                                               dt_equal=bool(dts == dts_ref))
print("RESULT " + json.dumps(out))
"""


def run_child(**kw):
    kw["root"] = ROOT
    p = subprocess.run([sys.executable, "-c", CHILD % kw], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


@pytest.mark.parametrize("recon,de,n,flags", [(2, 0, (29, 21, 19), [1, 1, 1, 1, 1, 1]), (1, 1, (29, 21, 19), [1, 1, 1, 1, 1, 1]),
                                              (1, 1, (45, 23, 1), [3, 3, 3, 3, 0, 0]), (2, 0, (64, 1, 1), [3, 3, 0, 0, 0, 0]),
                                              # builds/make.type.cooling as shipped: PPMP + HLLC + VL + DE + COOLING_GPU
                                              (4, 1, (29, 21, 19), [1, 1, 1, 1, 1, 1]), (4, 1, (45, 23, 1), [3, 3, 3, 3, 0, 0])])
def test_cie_cooling_matches_the_reference_build(recon, de, n, flags):
    from oracle import reference
    variant = ("_de" if de else "") + "_cool"
    if not reference.available(recon, 0, variant=variant):
        pytest.skip("oracle/_ref library for this combination was not built")
    r = run_child(recon=recon, n=n, flags=flags, gamma=5.0 / 3.0, steps=3, de=de)
    assert "-DCOOLING_GPU" in r["flags"].split(), json.dumps(r)
    assert r["cooling_effect_l1"] > 1e-3, json.dumps(r)      # the gas did cool
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], json.dumps(r)
    # CPU libm against the device library inside a sub-cycling loop with a threshold: round-off, occasionally one sub-step
    assert r["oracle"]["l1"] <= 1e-8 and r["fast"]["l1"] <= 1e-8, json.dumps(r)
