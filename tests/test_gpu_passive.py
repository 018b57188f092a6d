"""-DSCALAR / -DDE builds of the path: passive scalars and the dual-energy formalism through every
reconstruction / Riemann / update kernel (src/hydro/hydro_cuda.cu:754-1099, 1267-1307 and the `#ifdef DE / SCALAR` branches
of plmc / ppmc / pcm / hllc / roe / hll / VL_{1,2,3}D_cuda.cu).  Three arms on the same bytes: the reference's own kernels
built with those macros (oracle/_ref/*_de, *_s1, *_de_s1), the CPU oracle (orc_vl_step_x), and the B200-native kernels in
strict (must be bit-equal) and product arithmetic.  One subprocess per case: the reference keeps its scratch in globals."""
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
import cholla_b200 as cb
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for
from helpers import ghosted, real_view, rel_l1, max_ulp

recon, solver, n, flags, gamma, steps, icname = %(recon)d, %(solver)d, %(n)r, %(flags)r, %(gamma)r, %(steps)d, %(ic)r
ns, de, use_ref, dfloor = %(ns)d, %(de)d, %(use_ref)d, %(dfloor)r
variant = ("_de" if de else "") + ("_s%%d" %% ns if ns else "")
import inspect
gen = getattr(ic, icname)
base = gen(n, seed=4) if "seed" in inspect.signature(gen).parameters else gen(n)
fields = ic.with_passive(base, ns, bool(de), gamma)
if icname == "cold_flow" and de:  # an advected internal energy that has drifted from E - Ekin, as it does in a real run
    fields[-1] *= 1.0 + 0.05 * (np.random.default_rng(9).random(fields[-1].shape) - 0.5)
g = n_ghost_for(recon)
flat, dims = ghosted(fields, g, flags)
lengths = (1.0, 1.0, 1.0)
dx, dy, dz = [lengths[i] / n[i] for i in range(3)]
out = {}
# arm 1: CPU oracle
orc = flat.copy()
dts_orc = oracle.advance(orc, *dims, g, dx, dy, dz, gamma, recon, solver, flags, steps, n_scalars=ns, dual_energy=bool(de),
                         density_floor=dfloor)
want, dts_want = orc, dts_orc
# arm 2: the reference's kernels of the matching -DDE / -DSCALAR build
if use_ref:
    t = torch.from_numpy(flat.copy()).cuda()
    dts_ref, dti_ref, _ = reference.advance(t, dims, g, dx, dy, dz, gamma, recon, solver, flags, steps, variant=variant,
                                            density_floor=dfloor)
    ref = t.cpu().numpy()
    out["oracle_vs_ref"] = dict(ulp=max_ulp(real_view(orc, dims, g), real_view(ref, dims, g)), dt_equal=bool(dts_orc == dts_ref),
                                whole_array_equal=bool(np.array_equal(orc, ref)))
    out["flags"] = reference.lib(recon, solver, variant=variant).ref_build_flags().decode()
    want, dts_want = ref, dts_ref
# arm 3/4: B200-native kernels, strict and product arithmetic
for strict in (True, False):
    G = Grid3D(n, lengths, recon=recon, solver=solver, gamma=gamma, bc=flags, strict=strict, n_scalars=ns,
               dual_energy=bool(de), density_floor=dfloor)
    G.upload(flat)
    dts = G.advance(steps)
    got = G.download(); G.close()
    out["strict" if strict else "fast"] = dict(
        ulp=max_ulp(real_view(got, dims, g), real_view(want, dims, g)),
        l1=float(rel_l1(real_view(got, dims, g), real_view(want, dims, g)).max()),
        dt_equal=bool(dts == dts_want), whole_array_equal=bool(np.array_equal(got, want)))
# how many cells sit on each side of the dual-energy switch in the initial state (the test must exercise both)
if de:
    r = real_view(flat, dims, g)
    ut = r[4] - 0.5 * (r[1] ** 2 + r[2] ** 2 + r[3] ** 2) / r[0]
    out["frac_total_energy_cells"] = float((ut / r[4] > 0.001).mean())
print("RESULT " + json.dumps(out))
"""


def run_child(**kw):
    kw["root"] = ROOT
    kw.setdefault("dfloor", 0.0)
    p = subprocess.run([sys.executable, "-c", CHILD % kw], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def _have(recon, solver, ns, de):
    from oracle import reference
    return reference.available(recon, solver, variant=("_de" if de else "") + ("_s%d" % ns if ns else ""))


# (recon, solver, n_scalars, dual_energy): the -DDE / -DSCALAR reference builds of oracle/ref_build/Makefile
REF_BUILDS = [(2, 0, 0, 1), (2, 0, 1, 0), (2, 0, 1, 1), (1, 1, 1, 1), (1, 0, 0, 1), (2, 3, 1, 1), (3, 0, 1, 1),
              (0, 0, 1, 1)]  # recon 3 = PLMP, 0 = PCM


@pytest.mark.parametrize("recon,solver,ns,de", REF_BUILDS)
@pytest.mark.parametrize("icname", ["random_smooth", "cold_flow"])
def test_passive_3d_reference_oracle_b200(recon, solver, ns, de, icname):
    if not _have(recon, solver, ns, de):
        pytest.skip("oracle/_ref library for this combination was not built")
    r = run_child(recon=recon, solver=solver, n=(37, 19, 21), flags=[1, 1, 1, 1, 1, 1], gamma=5.0 / 3.0, steps=3,
                  ic=icname, ns=ns, de=de, use_ref=1)
    built = r["flags"].split()
    assert ("-DDE" in built) == bool(de) and ("-DSCALAR" in built) == bool(ns), json.dumps(r)
    assert r["oracle_vs_ref"]["ulp"] == 0 and r["oracle_vs_ref"]["dt_equal"], json.dumps(r)
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], json.dumps(r)
    if icname == "random_smooth":
        assert r["fast"]["l1"] <= 1e-12, json.dumps(r)
    if de and icname == "cold_flow":  # both branches of Get_Pressure_From_DE / Select_Internal_Energy were taken
        assert 0.02 < r["frac_total_energy_cells"] < 0.98, json.dumps(r)


@pytest.mark.parametrize("n,flags,icname,steps", [((45, 23, 1), [1, 1, 1, 1, 0, 0], "random_smooth", 4),
                                                  ((45, 23, 1), [3, 3, 2, 2, 0, 0], "cold_flow", 4),
                                                  ((100, 1, 1), [3, 3, 0, 0, 0, 0], "cold_flow", 12),
                                                  ((40, 32, 24), [3, 3, 2, 2, 1, 1], "spherical_overpressure", 4)])
def test_passive_de_scalar_other_shapes(n, flags, icname, steps):
    """1-D and 2-D grids (their half-step kernels take the plain pressure: VL_2D_cuda.cu:213, VL_1D_cuda.cu:183) and mixed
    boundaries, dual energy + one scalar, PPMC + HLLC."""
    if not _have(2, 0, 1, 1):
        pytest.skip("oracle/_ref library for this combination was not built")
    ic = icname
    r = run_child(recon=2, solver=0, n=n, flags=flags, gamma=5.0 / 3.0, steps=steps, ic=ic, ns=1, de=1, use_ref=1)
    assert r["oracle_vs_ref"]["ulp"] == 0 and r["oracle_vs_ref"]["dt_equal"], json.dumps(r)
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], json.dumps(r)


def test_passive_density_floor_scales_gas_energy_not_scalars():
    """hydro_cuda.cu:279-291 / VL_3D_cuda.cu:479-493: the density floor rescales the momenta, Energy and GasEnergy."""
    if not _have(2, 0, 1, 1):
        pytest.skip("oracle/_ref library for this combination was not built")
    r = run_child(recon=2, solver=0, n=(37, 19, 21), flags=[1, 1, 1, 1, 1, 1], gamma=5.0 / 3.0, steps=2,
                  ic="random_smooth", ns=1, de=1, use_ref=1, dfloor=0.8)
    assert r["oracle_vs_ref"]["ulp"] == 0 and r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], json.dumps(r)


@pytest.mark.parametrize("recon,solver,ns,de", [(0, 0, 1, 1), (1, 3, 1, 0), (2, 1, 0, 1), (0, 1, 1, 0), (1, 0, 1, 1),
                                                (4, 0, 1, 1), (3, 0, 1, 0), (4, 0, 0, 1)])  # 3 / 4 = PLMP / PPMP
def test_passive_combinations_without_a_reference_build_match_the_oracle(recon, solver, ns, de):
    """The rest of the built matrix against the CPU oracle alone (itself bit-equal to the reference on the builds above)."""
    r = run_child(recon=recon, solver=solver, n=(29, 17, 19), flags=[1, 1, 1, 1, 1, 1], gamma=1.4, steps=3,
                  ic="random_smooth", ns=ns, de=de, use_ref=0)
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], json.dumps(r)
    assert r["fast"]["l1"] <= 1e-12, json.dumps(r)


def test_scalar_floor_and_temperature_floor_with_dual_energy():
    import numpy as np
    import torch

    import oracle
    from cholla_b200 import _lib
    from cholla_b200 import initial_conditions as ic
    from cholla_b200.grid import Grid3D, n_ghost_for
    from helpers import ghosted
    import ctypes as C

    n, gamma, flags = (20, 18, 16), 5.0 / 3.0, [1, 1, 1, 1, 1, 1]
    fields = ic.with_passive(ic.random_smooth(n, seed=2), 1, True, gamma)
    g = n_ghost_for(2)
    flat, dims = ghosted(fields, g, flags)
    G = Grid3D(n, (1.0, 1.0, 1.0), recon=2, solver=0, gamma=gamma, bc=flags, n_scalars=1, dual_energy=True)
    G.upload(flat)
    G.Apply_Scalar_Floor(5, 0.6)
    _lib.check(_lib.lib().chl_apply_temperature_floor(G._ctx, G._ptr(), 1.2, G._stream()))
    got = G.download()
    G.close()
    want = flat.copy()
    oracle.scalar_floor(want, *dims, g, 5, 0.6)
    oracle.temperature_floor_x(want, *dims, g, 1.2, 7, True)
    assert np.array_equal(got, want)
    assert (want != flat).sum() > 100
