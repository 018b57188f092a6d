"""GPU parity tests of the MHD path (PLMC/PPMC + HLLD + constrained transport, BASELINE config 4) through the C ABI.

Three arms on the same bytes:
  * the reference's OWN kernels (oracle/_ref/libcholla_ref_{plmc,ppmc}_hlld.so: unmodified sources, -DMHD -DHLLD,
    compiled for sm_100) -- strict mode must reproduce them bit for bit on every real cell, dt history included;
  * the CPU oracle (oracle/mhd_oracle.c) -- 1e-12 relative L1 (rsqrt / rhypot have no correctly rounded libm twin);
  * product arithmetic (FMA) -- 1e-12 relative L1 on smooth data.
Each reference comparison runs in a fresh subprocess: the reference sizes its global scratch at the first call."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-12

CHILD = r"""
import json, sys
import numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from oracle import mhd, reference
import cholla_b200 as cb
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for

recon, n, flags, gamma, steps, icname, use_ref = %(recon)d, %(n)r, %(flags)r, %(gamma)r, %(steps)d, %(ic)r, %(use_ref)d
g = n_ghost_for(recon, mhd=True)
full = getattr(ic, icname)(n, g) if icname != "mhd_smooth" else ic.mhd_smooth(n, g, seed=3)
dims = (n[0] + 2 * g, n[1] + 2 * g, n[2] + 2 * g)
flat = np.ascontiguousarray(full.reshape(-1))
mhd.set_boundaries(flat, *dims, g, flags)          # every arm starts from the same bytes
dx, dy, dz = [1.0 / v for v in n]
def real(a):
    return a.reshape(8, dims[2], dims[1], dims[0])[:, g:-g, g:-g, g:-g]
def rel_l1(a, b):
    # relative L1 of density, momentum, energy and magnetic field (the north-star's four quantities): the three
    # components of a vector share one norm, so a component that is identically zero in the problem (Orszag-Tang: m_z,
    # B_z) is measured against the vector's size rather than against its own round-off
    groups = [[0], [1, 2, 3], [4], [5, 6, 7]]
    return max(sum(float(np.abs(a[f] - b[f]).sum()) for f in grp) / sum(float(np.abs(b[f]).sum()) for f in grp)
               for grp in groups)
def ulp(a, b):
    return int(np.abs(np.ascontiguousarray(a).view(np.int64) - np.ascontiguousarray(b).view(np.int64)).max())
out = {}
orc = flat.copy()
dts_orc = mhd.advance(orc, *dims, g, dx, dy, dz, gamma, recon, flags, steps)
out["divb_oracle"] = mhd.max_divergence(orc, *dims, g, dx, dy, dz)
ref = None
if use_ref:
    t = torch.from_numpy(flat.copy()).cuda()
    dts_ref, _, _ = reference.advance(t, dims, g, dx, dy, dz, gamma, recon, 2, flags, steps)
    ref = t.cpu().numpy()
    out["oracle_vs_ref_l1"] = rel_l1(real(orc), real(ref))
    out["oracle_vs_ref_dt"] = max(abs(a - b) / b for a, b in zip(dts_orc, dts_ref))
for strict in (True, False):
    G = Grid3D(n, (1.0, 1.0, 1.0), recon=recon, solver=cb.HLLD, gamma=gamma, bc=flags, strict=strict, mhd=True)
    assert (G.H.nx, G.H.ny, G.H.nz) == dims and G.H.n_fields == 8
    G.upload(flat)
    dts = G.advance(steps)
    got = G.download(); G.close()
    r = dict(l1_oracle=rel_l1(real(got), real(orc)), nan=int(np.isnan(real(got)).sum()),
             divb=mhd.max_divergence(got, *dims, g, dx, dy, dz),
             dt_oracle=max(abs(a - b) / b for a, b in zip(dts, dts_orc)))
    if ref is not None:
        r.update(ulp_ref=ulp(real(got), real(ref)), l1_ref=rel_l1(real(got), real(ref)), dt_equal_ref=bool(dts == dts_ref),
                 whole_equal_ref=bool(np.array_equal(got, ref)))
    out["strict" if strict else "fast"] = r
print("RESULT " + json.dumps(out))
"""


def run_child(**kw):
    kw["root"] = ROOT
    p = subprocess.run([sys.executable, "-c", CHILD % kw], capture_output=True, text=True, timeout=1200)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def have_ref(recon):
    from oracle import reference
    return reference.available(recon, 2)


@pytest.mark.parametrize("recon", [2, 1])
def test_mhd_smooth_vs_reference_kernels_and_oracle(recon):
    use_ref = have_ref(recon)
    r = run_child(recon=recon, n=(24, 20, 16), flags=[1] * 6, gamma=5.0 / 3.0, steps=3, ic="mhd_smooth",
                  use_ref=int(use_ref))
    assert r["strict"]["nan"] == 0 and r["fast"]["nan"] == 0, r
    assert r["strict"]["l1_oracle"] <= TOL and r["fast"]["l1_oracle"] <= TOL, r
    assert r["strict"]["dt_oracle"] <= 1e-13 and r["fast"]["dt_oracle"] <= 1e-12, r
    assert r["strict"]["divb"] <= 1e-11 and r["fast"]["divb"] <= 1e-11, r
    if use_ref:
        assert r["oracle_vs_ref_l1"] <= TOL, r
        assert r["strict"]["ulp_ref"] == 0 and r["strict"]["dt_equal_ref"], r   # bit-exact drop-in
        assert r["fast"]["l1_ref"] <= TOL, r


def test_mhd_orszag_tang_32_ppmc_config4():
    """BASELINE config 4 at reduced size: Orszag-Tang vortex, PPMC + HLLD + CT, 6 steps."""
    use_ref = have_ref(2)
    r = run_child(recon=2, n=(32, 32, 32), flags=[1] * 6, gamma=1.666666666666667, steps=6, ic="orszag_tang",
                  use_ref=int(use_ref))
    assert r["strict"]["nan"] == 0, r
    assert r["strict"]["l1_oracle"] <= TOL and r["fast"]["l1_oracle"] <= TOL, r
    assert r["strict"]["divb"] <= 1e-11, r
    if use_ref:
        assert r["strict"]["ulp_ref"] == 0 and r["strict"]["dt_equal_ref"], r


def test_mhd_blast_transmissive_strict_bits():
    """Discontinuous data, outflow boundaries (no diode in an MHD build): only bit-reproduction can match."""
    if not have_ref(1):
        pytest.skip("oracle/_ref MHD library not built")
    r = run_child(recon=1, n=(24, 24, 20), flags=[3] * 6, gamma=5.0 / 3.0, steps=4, ic="mhd_blast", use_ref=1)
    assert r["strict"]["ulp_ref"] == 0 and r["strict"]["dt_equal_ref"], r
    assert r["strict"]["whole_equal_ref"], r
