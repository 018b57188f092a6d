"""Static gravity (-DSTATIC_GRAV): the source terms of Update_Conserved_Variables_{1,2,3}D
(src/hydro/hydro_cuda.cu:59-68, 141-153, 301-317) with the accelerations of calc_g_{1,2,3}D (src/gravity/static_grav.h).
Arms: the reference's own -DSTATIC_GRAV kernels (oracle/_ref/*_sg), the CPU oracle, the B200-native kernels.  The analytic
fields call log / pow / atan2, whose device-library bits no CPU libm reproduces, so bit equality is asserted between the
two GPU arms (and against the oracle for the exact Rayleigh-Taylor field); the oracle's field agrees to round-off."""
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

recon, solver, n, flags, gamma, steps = %(recon)d, %(solver)d, %(n)r, %(flags)r, %(gamma)r, %(steps)d
custom_grav, lengths, lo, de, scale = %(custom_grav)d, %(lengths)r, %(lo)r, %(de)d, %(scale)r
variant = ("_de" if de else "") + "_sg"
base = ic.random_smooth(n, seed=6)
base[1:4] *= scale      # velocities comparable with the free-fall speed over a step of the galactic fields
base[4] = (0.5 + np.random.default_rng(1).random(base[0].shape)) * scale ** 2 / (gamma - 1.0) + 0.5 * (base[1:4] ** 2).sum(0) / base[0]
fields = ic.with_passive(base, 0, bool(de), gamma)
g = n_ghost_for(recon)
flat, dims = ghosted(fields, g, flags)
d = [lengths[i] / n[i] for i in range(3)]
out = {}
# the reference's kernels
t = torch.from_numpy(flat.copy()).cuda()
dts_ref, _, _ = reference.advance(t, dims, g, d[0], d[1], d[2], gamma, recon, solver, flags, steps, variant=variant,
                                  custom_grav=custom_grav, bounds=lo)
ref = t.cpu().numpy()
out["flags"] = reference.lib(recon, solver, variant=variant).ref_build_flags().decode()
# the same build with custom_grav = 0: gravity must have changed the answer for the test to mean anything
# CPU oracle
field = oracle.static_gravity_field(*dims, g, custom_grav, (0, 0, 0), d[0], d[1], d[2], lo)
oracle.set_gravity(field)
orc = flat.copy()
dts_orc = oracle.advance(orc, *dims, g, d[0], d[1], d[2], gamma, recon, solver, flags, steps, dual_energy=bool(de))
oracle.set_gravity(None)
plain = flat.copy()
oracle.advance(plain, *dims, g, d[0], d[1], d[2], gamma, recon, solver, flags, steps, dual_energy=bool(de))
out["gravity_effect_l1"] = float(rel_l1(real_view(orc, dims, g), real_view(plain, dims, g)).max())
out["oracle"] = dict(ulp=max_ulp(real_view(orc, dims, g), real_view(ref, dims, g)), dt_equal=bool(dts_orc == dts_ref),
                     l1=float(rel_l1(real_view(orc, dims, g), real_view(ref, dims, g)).max()))
for strict in (True, False):
    G = Grid3D(n, lengths, bounds_min=lo, recon=recon, solver=solver, gamma=gamma, bc=flags, strict=strict,
               dual_energy=bool(de))
    G.Set_Static_Gravity(custom_grav)
    if strict:
        gf = G.gravity_field()
        den = np.abs(field).max()
        out["field_vs_oracle"] = float(np.abs(gf - field).max() / den) if den > 0 else 0.0
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
    kw.setdefault("de", 0)
    kw.setdefault("scale", 1.0)
    p = subprocess.run([sys.executable, "-c", CHILD % kw], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def _need(recon, solver, variant):
    from oracle import reference
    if not reference.available(recon, solver, variant=variant):
        pytest.skip("oracle/_ref library for this combination was not built")


def test_rayleigh_taylor_field_2d_is_exact_on_every_arm():
    """custom_grav = 2 in 2-D: g = (0, -1), no transcendental arithmetic -- reference == oracle == strict kernels."""
    _need(1, 0, "_sg")
    r = run_child(recon=1, solver=0, n=(45, 23, 1), flags=[1, 1, 2, 2, 0, 0], gamma=1.4, steps=4, custom_grav=2,
                  lengths=(1.0, 1.0, 1.0), lo=(0.0, 0.0, 0.0))
    assert "-DSTATIC_GRAV" in r["flags"].split(), json.dumps(r)
    assert r["gravity_effect_l1"] > 1e-4, json.dumps(r)
    assert r["oracle"]["ulp"] == 0 and r["oracle"]["dt_equal"], json.dumps(r)
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], json.dumps(r)
    assert r["fast"]["l1"] <= 1e-12, json.dumps(r)


@pytest.mark.parametrize("recon,n,flags,custom_grav,lengths,lo,scale", [
    # (lower corners chosen so that no cell centre sits on the axis, where x / r_disk is 0 / 0 in the reference too)
    (2, (29, 21, 19), [3, 3, 3, 3, 3, 3], 1, (40.0, 40.0, 20.0), (-20.3, -19.9, -10.1), 0.2),   # Milky Way disk + halo
    (1, (29, 21, 19), [3, 3, 3, 3, 3, 3], 2, (4.0, 4.0, 4.0), (-2.03, -1.99, -2.01), 0.2),      # M82
    (1, (45, 23, 1), [3, 3, 3, 3, 0, 0], 1, (1.0, 1.0, 1.0), (-0.5, -0.5, 0.0), 1.0),           # Gresho
    (1, (45, 23, 1), [3, 3, 3, 3, 0, 0], 4, (40.0, 40.0, 1.0), (-20.3, -19.9, 0.0), 0.2),       # Kuzmin disk + NFW halo
    (1, (64, 1, 1), [3, 3, 0, 0, 0, 0], 1, (8.0, 1.0, 1.0), (-4.0, 0.0, 0.0), 0.2),             # 1-D column
])
def test_analytic_fields_match_the_reference_build(recon, n, flags, custom_grav, lengths, lo, scale):
    _need(recon, 0, "_sg")
    r = run_child(recon=recon, solver=0, n=n, flags=flags, gamma=5.0 / 3.0, steps=3, custom_grav=custom_grav,
                  lengths=lengths, lo=lo, scale=scale)
    assert r["gravity_effect_l1"] > 1e-9, json.dumps(r)
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], json.dumps(r)   # same device library, same order: same bits
    assert r["field_vs_oracle"] <= 1e-13, json.dumps(r)                          # CPU libm: round-off only
    assert r["oracle"]["l1"] <= 1e-12 and r["fast"]["l1"] <= 1e-11, json.dumps(r)


def test_gravity_with_dual_energy():
    """The source terms sit inside the update kernel, i.e. before Select_Internal_Energy / Sync_Energies."""
    _need(1, 0, "_de_sg")
    r = run_child(recon=1, solver=0, n=(29, 21, 19), flags=[3, 3, 3, 3, 3, 3], gamma=5.0 / 3.0, steps=3, custom_grav=1,
                  lengths=(40.0, 40.0, 20.0), lo=(-20.3, -19.9, -10.1), scale=0.2, de=1)
    assert "-DDE" in r["flags"].split() and "-DSTATIC_GRAV" in r["flags"].split(), json.dumps(r)
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], json.dumps(r)
    assert r["oracle"]["l1"] <= 1e-12, json.dumps(r)


POT_CHILD = r"""
import json, sys
import numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import oracle
from oracle import reference
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for
from helpers import ghosted, real_view, rel_l1, max_ulp

recon, n, flags, gamma, steps, five, dropin = %(recon)d, %(n)r, %(flags)r, %(gamma)r, %(steps)d, %(five)d, %(dropin)d
variant = "_grav5" if five else "_grav3"
fields = ic.random_smooth(n, seed=8)
g = n_ghost_for(recon)
flat, dims = ghosted(fields, g, flags)
d = [1.0 / n[i] for i in range(3)]
# a smooth potential on the whole local array, ghost cells included (what the Poisson solver + its boundary fill leave)
c = [((np.arange(dims[i]) - g) + 0.5) * d[i] for i in range(3)]
Z, Y, X = np.meshgrid(c[2], c[1], c[0], indexing="ij")
pot = -3.0 / np.sqrt((X - 0.45) ** 2 + (Y - 0.55) ** 2 + (Z - 0.5) ** 2 + 0.05) + 0.7 * np.sin(2 * np.pi * X) * np.cos(2 * np.pi * Y)
pot = np.ascontiguousarray(pot.reshape(-1))
out = {}
pt = torch.from_numpy(pot).cuda()
t = torch.from_numpy(flat.copy()).cuda()
dts_ref, _, _ = reference.advance(t, dims, g, d[0], d[1], d[2], gamma, recon, 0, flags, steps, variant=variant, potential=pt)
ref = t.cpu().numpy()
out["flags"] = reference.lib(recon, 0, variant=variant).ref_build_flags().decode()
if dropin:
    t2 = torch.from_numpy(flat.copy()).cuda()
    dts_d, _, _ = reference.advance(t2, dims, g, d[0], d[1], d[2], gamma, recon, 0, flags, steps, variant=variant, potential=pt,
                                    dropin=True)
    out["dropin"] = dict(equal=bool(np.array_equal(t2.cpu().numpy(), ref)), dt_equal=bool(dts_d == dts_ref))
oracle.set_gravity_potential(pot, bool(five))
orc = flat.copy()
dts_orc = oracle.advance(orc, *dims, g, d[0], d[1], d[2], gamma, recon, 0, flags, steps)
oracle.set_gravity_potential(None)
plain = flat.copy()
oracle.advance(plain, *dims, g, d[0], d[1], d[2], gamma, recon, 0, flags, steps)
out["gravity_effect_l1"] = float(rel_l1(real_view(orc, dims, g), real_view(plain, dims, g)).max())
out["oracle"] = dict(ulp=max_ulp(real_view(orc, dims, g), real_view(ref, dims, g)), dt_equal=bool(dts_orc == dts_ref))
for strict in (True, False):
    G = Grid3D(n, (1.0, 1.0, 1.0), recon=recon, solver=0, gamma=gamma, bc=flags, strict=strict)
    G.Set_Gravity_Potential(pot.reshape(dims[2], dims[1], dims[0]), five_point=bool(five))
    G.upload(flat)
    dts = G.advance(steps)
    got = G.download(); G.close()
    out["strict" if strict else "fast"] = dict(ulp=max_ulp(real_view(got, dims, g), real_view(ref, dims, g)),
                                               l1=float(rel_l1(real_view(got, dims, g), real_view(ref, dims, g)).max()),
                                               dt_equal=bool(dts == dts_ref))
print("RESULT " + json.dumps(out))
"""


@pytest.mark.parametrize("recon,five,dropin", [(2, 1, 1), (1, 0, 0)])
def test_potential_gradient_source_terms_match_the_gravity_build(recon, five, dropin):
    """-DGRAVITY -DGRAVITY_GPU (-DGRAVITY_5_POINTS_GRADIENT): the potential-gradient branch of Update_Conserved_Variables_3D
    (hydro_cuda.cu:317-379) fed with a caller-supplied potential: reference kernels == oracle == strict kernels, and the
    reference-signature driver linked against the shims gives the same bytes."""
    variant = "_grav5" if five else "_grav3"
    _need(recon, 0, variant)
    kw = dict(root=ROOT, recon=recon, n=(29, 21, 19), flags=[1, 1, 1, 1, 1, 1], gamma=5.0 / 3.0, steps=3, five=five, dropin=dropin)
    p = subprocess.run([sys.executable, "-c", POT_CHILD % kw], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    r = json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert "-DGRAVITY" in r["flags"].split() and ("-DGRAVITY_5_POINTS_GRADIENT" in r["flags"].split()) == bool(five), json.dumps(r)
    assert r["gravity_effect_l1"] > 1e-4, json.dumps(r)
    assert r["oracle"]["ulp"] == 0 and r["oracle"]["dt_equal"], json.dumps(r)
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], json.dumps(r)
    assert r["fast"]["l1"] <= 1e-11, json.dumps(r)
    if dropin:
        assert r["dropin"]["equal"] and r["dropin"]["dt_equal"], json.dumps(r)
