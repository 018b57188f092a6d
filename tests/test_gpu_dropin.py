"""Drop-in proof: ONE driver written against the reference's own host entry points (oracle/ref_build/ref_driver.cu
calls VL_Algorithm_3D_CUDA, Calc_dt_GPU, Apply_Temperature_Floor, SetGhostCells by their reference signatures) is
linked twice -- against the reference's kernel objects, and against cholla_b200/csrc/cholla_shims.cpp +
libcholla_b200.so.  Both builds must leave the same bytes in the caller's array and return the same dt history."""
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
from oracle import mhd, reference
import oracle
from cholla_b200 import initial_conditions as ic
from helpers import ghosted
recon, solver, n, flags, gamma, steps = %(recon)d, %(solver)d, %(n)r, %(flags)r, %(gamma)r, %(steps)d
simple = %(simple)r  # -DSIMPLE builds: the driver calls Simple_Algorithm_3D_CUDA
variant, custom_grav, lo = %(variant)r, %(custom_grav)d, %(lo)r  # -DDE / -DSCALAR / -DSTATIC_GRAV builds
is_mhd = solver == 2
g = {0: 2, 1: 3, 2: 4}[recon] + (1 if is_mhd else 0)
if is_mhd:
    full = ic.mhd_smooth(n, g, seed=5)
    dims = (n[0] + 2 * g, n[1] + 2 * g, n[2] + 2 * g)
    flat = np.ascontiguousarray(full.reshape(-1))
    mhd.set_boundaries(flat, *dims, g, flags)
else:
    fields = ic.with_passive(ic.spherical_overpressure(n), 1 if "_s1" in variant else 0, "_de" in variant, gamma)
    flat, dims = ghosted(fields, g, flags)
dx, dy, dz = [1.0 / v for v in n]
res = {}
for dropin in (False, True):
    t = torch.from_numpy(flat.copy()).cuda()
    dts, dti, _ = reference.advance(t, dims, g, dx, dy, dz, gamma, recon, solver, flags, steps, dropin=dropin, simple=simple,
                                    variant=variant, custom_grav=custom_grav, bounds=lo)
    res[dropin] = (t.cpu().numpy(), dts, dti)
a, b = res[False], res[True]
print("RESULT " + json.dumps(dict(equal=bool(np.array_equal(a[0], b[0])), dt_equal=bool(a[1] == b[1]),
      dti_equal=bool(a[2] == b[2]), changed=bool(not np.array_equal(a[0], flat)),
      flags=reference.lib(recon, solver, dropin=True, simple=simple, variant=variant).ref_build_flags().decode())))
"""


def run_child(**kw):
    kw["root"] = ROOT
    kw.setdefault("simple", False)
    kw.setdefault("variant", "")
    kw.setdefault("custom_grav", 0)
    kw.setdefault("lo", (0.0, 0.0, 0.0))
    p = subprocess.run([sys.executable, "-c", CHILD % kw], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return json.loads([l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])


@pytest.mark.parametrize("recon,solver,flags", [(2, 0, [3, 3, 2, 2, 1, 1]), (1, 1, [1] * 6), (2, 2, [1] * 6)])
def test_reference_driver_runs_unchanged_on_the_b200_library(recon, solver, flags):
    from oracle import reference
    if not (reference.available(recon, solver) and reference.available(recon, solver, dropin=True)):
        pytest.skip("oracle/_ref libraries for this combination were not built")
    r = run_child(recon=recon, solver=solver, n=(32, 24, 20), flags=flags, gamma=5.0 / 3.0, steps=4)
    assert r["changed"] and "cholla_shims" in r["flags"], r
    assert r["equal"] and r["dt_equal"] and r["dti_equal"], r


def test_simple_integrator_driver_runs_unchanged_on_the_b200_library():
    """The same proof for a -DSIMPLE -DPLMC -DHLLC build: the driver's Simple_Algorithm_3D_CUDA call lands in
    cholla_shims.cpp -> chl_simple_step."""
    from oracle import reference
    if not (reference.available(1, 0, simple=True) and reference.available(1, 0, dropin=True, simple=True)):
        pytest.skip("oracle/_ref libraries for this combination were not built")
    r = run_child(recon=1, solver=0, n=(32, 24, 20), flags=[3, 3, 2, 2, 1, 1], gamma=5.0 / 3.0, steps=4, simple=True)
    assert r["changed"] and "cholla_shims" in r["flags"] and "-DSIMPLE" in r["flags"], r
    assert r["equal"] and r["dt_equal"] and r["dti_equal"], r


@pytest.mark.parametrize("recon,variant,custom_grav,lo", [(1, "_de_sg", 1, (-0.5, -0.5, -0.5)), (2, "_de_s1", 0, (0.0, 0.0, 0.0)),
                                                          (1, "_de_cool", 0, (0.0, 0.0, 0.0))])
def test_dual_energy_scalar_and_static_gravity_builds_run_unchanged_on_the_b200_library(recon, variant, custom_grav, lo):
    """-DDE (-DSCALAR, -DSTATIC_GRAV, -DCOOLING_GPU) builds of the driver: the shims turn the macros into
    chl_config.dual_energy / n_scalars and chl_set_static_gravity (custom_grav, offsets and bounds arrive with every
    VL_Algorithm_3D_CUDA call); the driver's Cooling_Update call lands in chl_cooling_update."""
    from oracle import reference
    if not (reference.available(recon, 0, variant=variant) and reference.available(recon, 0, dropin=True, variant=variant)):
        pytest.skip("oracle/_ref libraries for this combination were not built")
    r = run_child(recon=recon, solver=0, n=(32, 24, 20), flags=[3, 3, 2, 2, 1, 1], gamma=5.0 / 3.0, steps=4, variant=variant,
                  custom_grav=custom_grav, lo=lo)
    assert r["changed"] and "cholla_shims" in r["flags"] and "-DDE" in r["flags"].split(), r
    assert r["equal"] and r["dt_equal"] and r["dti_equal"], r
