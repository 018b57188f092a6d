"""Validate the CPU oracle AND the B200-native kernels against the reference's OWN kernels (oracle/_ref: the
unmodified reference sources compiled for sm_100, see oracle/ref_build/Makefile) on the same inputs.

Each test runs in a fresh subprocess because the reference keeps its scratch arrays in process-wide globals
sized at the first call (src/integrators/VL_3D_cuda.cu:58-133): one process, one grid shape."""
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
integ = %(integ)r  # "vl" or "simple" (-DSIMPLE reference build, chl_simple_step, oracle.simple_step)
fields = getattr(ic, icname)(n) if icname != "random_smooth" else ic.random_smooth(n, seed=4)
g = n_ghost_for(recon)
flat, dims = ghosted(fields, g, flags)
lengths = (1.0, 1.0, 1.0)
dx, dy, dz = [lengths[i] / n[i] for i in range(3)]
# arm 1: the reference's kernels
t = torch.from_numpy(flat.copy()).cuda()
dts_ref, dti_ref, _ = reference.advance(t, dims, g, dx, dy, dz, gamma, recon, solver, flags, steps, simple=integ == "simple")
ref = t.cpu().numpy()
# arm 2: CPU oracle
orc = flat.copy()
dts_orc = oracle.advance(orc, *dims, g, dx, dy, dz, gamma, recon, solver, flags, steps, integrator=integ)
# arm 3/4: B200-native kernels, strict and product mode
out = {}
for strict in (True, False):
    G = Grid3D(n, lengths, recon=recon, solver=solver, gamma=gamma, bc=flags, strict=strict, integrator=integ)
    G.upload(flat)
    dts = G.advance(steps)
    got = G.download(); G.close()
    out["strict" if strict else "fast"] = dict(
        ulp=max_ulp(real_view(got, dims, g), real_view(ref, dims, g)),
        l1=float(rel_l1(real_view(got, dims, g), real_view(ref, dims, g)).max()),
        dt_equal=bool(dts == dts_ref), whole_array_equal=bool(np.array_equal(got, ref)))
out["oracle"] = dict(ulp=max_ulp(real_view(orc, dims, g), real_view(ref, dims, g)), dt_equal=bool(dts_orc == dts_ref),
                     l1=float(rel_l1(real_view(orc, dims, g), real_view(ref, dims, g)).max()),
                     whole_array_equal=bool(np.array_equal(orc, ref)))
out["flags"] = reference.lib(recon, solver, simple=integ == "simple").ref_build_flags().decode()
print("RESULT " + json.dumps(out))
"""


def run_child(**kw):
    kw["root"] = ROOT
    kw.setdefault("integ", "vl")
    p = subprocess.run([sys.executable, "-c", CHILD % kw], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[7:])


def _need(recon, solver, simple=False):
    from oracle import reference
    if not reference.available(recon, solver, simple=simple):
        pytest.skip("oracle/_ref library for this combination was not built")


@pytest.mark.parametrize("recon,solver", [(2, 0), (1, 0), (2, 1), (1, 1), (2, 3), (3, 0), (4, 0)])  # solver 3 = HLL (-DHLL); recon 3 = PLMP, 4 = PPMP
def test_reference_kernels_vs_oracle_vs_b200_3d(recon, solver):
    _need(recon, solver)
    r = run_child(recon=recon, solver=solver, n=(37, 19, 21), flags=[1, 1, 1, 1, 1, 1], gamma=5.0 / 3.0, steps=3,
                  ic="random_smooth")
    # the CPU restatement reproduces the reference's kernels bit for bit on the real cells (and the ghosts)
    assert r["oracle"]["ulp"] == 0 and r["oracle"]["dt_equal"], r
    # so does the strict arithmetic mode of the new kernels
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], r
    # and the product mode is inside the parity budget
    assert r["fast"]["l1"] <= 1e-12, r


def test_reference_kernels_blast_mixed_boundaries():
    _need(2, 0)
    r = run_child(recon=2, solver=0, n=(40, 32, 24), flags=[3, 3, 2, 2, 1, 1], gamma=1.66666667, steps=5,
                  ic="spherical_overpressure")
    assert r["oracle"]["whole_array_equal"] and r["strict"]["whole_array_equal"], r
    # no product-mode bar here: on discontinuous data the reference's PPM limiter turns a one-ulp perturbation into
    # an O(1e-3) change within a step (tests/test_oracle_sensitivity.py shows it with the oracle alone), so only
    # bit-reproducing arithmetic can match a reference snapshot of this problem


def test_reference_kernels_sod_1d_config1():
    _need(1, 0)
    r = run_child(recon=1, solver=0, n=(100, 1, 1), flags=[3, 3, 0, 0, 0, 0], gamma=1.4, steps=40, ic="riemann")
    assert r["oracle"]["ulp"] == 0 and r["strict"]["ulp"] == 0, r
    assert r["fast"]["l1"] <= 1e-12, r


def test_reference_kernels_2d():
    _need(2, 0)
    r = run_child(recon=2, solver=0, n=(45, 23, 1), flags=[1, 1, 1, 1, 0, 0], gamma=1.4, steps=4, ic="random_smooth")
    assert r["oracle"]["ulp"] == 0 and r["strict"]["ulp"] == 0, r
    assert r["fast"]["l1"] <= 1e-12, r


@pytest.mark.parametrize("n,flags,icname,steps", [((37, 19, 21), [1, 1, 1, 1, 1, 1], "random_smooth", 3),
                                                  ((40, 32, 24), [3, 3, 2, 2, 1, 1], "spherical_overpressure", 4),
                                                  ((45, 23, 1), [1, 1, 1, 1, 0, 0], "random_smooth", 4),
                                                  ((100, 1, 1), [3, 3, 0, 0, 0, 0], "riemann", 30)])
def test_simple_integrator_plmc_reference_vs_oracle_vs_b200(n, flags, icname, steps):
    """-DSIMPLE -DPLMC -DHLLC (builds/make.type.hydro with the cosmology integrator): Simple_Algorithm_{1,2,3}D_CUDA with
    PLMC's characteristic tracing, reference kernels == oracle == strict kernels, bit for bit."""
    _need(1, 0, simple=True)
    r = run_child(recon=1, solver=0, n=n, flags=flags, gamma=5.0 / 3.0 if icname != "riemann" else 1.4, steps=steps,
                  ic=icname, integ="simple")
    assert "-DSIMPLE" in r["flags"] and "-DVL" not in r["flags"]
    assert r["oracle"]["ulp"] == 0 and r["oracle"]["dt_equal"], r
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], r
    if icname == "random_smooth":
        assert r["fast"]["l1"] <= 1e-12, r


def test_simple_integrator_pcm_roe_reference():
    """-DSIMPLE -DPCM -DROE: first-order update.  The oracle and the strict kernels copy the conserved state like
    PCM_Reconstruction_3D and are bit-exact; the product arithmetic reconstructs from primitives (prim -> cons round-off)."""
    _need(0, 1, simple=True)
    r = run_child(recon=0, solver=1, n=(37, 19, 21), flags=[1, 1, 1, 1, 1, 1], gamma=5.0 / 3.0, steps=3,
                  ic="random_smooth", integ="simple")
    assert r["oracle"]["ulp"] == 0 and r["oracle"]["dt_equal"], r
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], r
    assert r["fast"]["l1"] <= 1e-12, r


@pytest.mark.parametrize("n,flags,icname,steps", [((37, 19, 21), [1, 1, 1, 1, 1, 1], "random_smooth", 3),
                                                  ((45, 23, 1), [1, 1, 1, 1, 0, 0], "random_smooth", 4),
                                                  ((100, 1, 1), [3, 3, 0, 0, 0, 0], "riemann", 30)])
def test_pcm_corrector_under_van_leer_reference_vs_oracle_vs_b200(n, flags, icname, steps):
    """-DVL -DPCM -DHLLC: the corrector's interface states are copies of the conserved half-step state
    (VL_3D_cuda.cu:226-229, pcm_cuda.cu); reference kernels == oracle == strict kernels, bit for bit, in 1-, 2- and 3-D."""
    _need(0, 0)
    r = run_child(recon=0, solver=0, n=n, flags=flags, gamma=5.0 / 3.0 if icname != "riemann" else 1.4, steps=steps,
                  ic=icname)
    assert "-DPCM" in r["flags"] and "-DVL" in r["flags"]
    assert r["oracle"]["ulp"] == 0 and r["oracle"]["dt_equal"], r
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], r
    if icname == "random_smooth":
        assert r["fast"]["l1"] <= 1e-12, r


@pytest.mark.parametrize("recon,n,flags,icname,steps,integ", [
    (2, (45, 23, 1), [1, 1, 1, 1, 0, 0], "random_smooth", 4, "vl"),
    (1, (100, 1, 1), [3, 3, 0, 0, 0, 0], "riemann", 30, "vl"),
    (1, (37, 19, 21), [1, 1, 1, 1, 1, 1], "random_smooth", 3, "simple"),
    (1, (40, 32, 24), [3, 3, 2, 2, 1, 1], "spherical_overpressure", 4, "simple")])
def test_exact_solver_reference_vs_b200(recon, n, flags, icname, steps, integ):
    """-DEXACT (exact_cuda.cu, Toro's solver with powf powers and a 1e-6 Newton tolerance): the strict kernels, built
    without FMA contraction and calling the same device powf, reproduce the reference's kernels bit for bit.  The CPU
    oracle and the product arithmetic can only follow to float round-off: powf is not correctly rounded on the device
    and has no CPU twin, and a one-ulp operand change can flip a float rounding.  The reference's 3-D Van Leer
    integrator does not compile with -DEXACT (VL_3D_cuda.cu:148-149), so its 3-D arm is the -DSIMPLE build."""
    _need(recon, 4, simple=integ == "simple")
    r = run_child(recon=recon, solver=4, n=n, flags=flags, gamma=1.4, steps=steps, ic=icname, integ=integ)
    assert "-DEXACT" in r["flags"]
    assert r["strict"]["ulp"] == 0 and r["strict"]["dt_equal"], r
    assert r["oracle"]["l1"] <= 1e-5 and r["fast"]["l1"] <= 1e-5, r
