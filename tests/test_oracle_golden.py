"""Pin the CPU oracle (oracle/vl_oracle.c) to the reference's own unit-test fiducials.

Tolerance is the reference's: testing_utilities::nearlyEqualDbl = <= 4 ULP or <= 1e-14 absolute
(src/utils/testing_utilities.cpp:63-75)."""
import json
import os

import numpy as np
import pytest

import oracle
from stdrandom import MT19937_64, UniformReal


def nearly_equal(a, b, ulps=4, abs_tol=1e-14):
    if a == b or abs(a - b) <= abs_tol:
        return True
    ia = np.float64(a).view(np.int64)
    ib = np.float64(b).view(np.int64)
    return abs(int(ia) - int(ib)) <= ulps


def load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as fh:
        return json.load(fh)


def test_prng_matches_libstdcpp():
    # first raw outputs of std::mt19937_64 seeded with 42 (well-known values)
    g = MT19937_64(42)
    assert g() == 13930160852258120406
    assert g() == 11788048577503494824


def test_ppm_single_variable_100_stencils(golden_dir):
    gd = load(golden_dir, "ppm_single_variable.json")
    g, dist = MT19937_64(42), UniformReal(-100, 100)
    data = [dist(g) for _ in range(500)]
    for i in range(100):
        L, R = oracle.ppm_single_variable(*data[5 * i:5 * i + 5])
        assert nearly_equal(L, gd["left"][i]), (i, L, gd["left"][i])
        assert nearly_equal(R, gd["right"][i]), (i, R, gd["right"][i])


def _check_sparse(arr, table, what, skip=()):
    table = {int(k): v for k, v in table.items() if int(k) not in skip}
    for i, v in enumerate(arr):
        want = table.get(i, 0.0)
        assert nearly_equal(float(v), want), (what, i, float(v), want)


def test_plmc_hydro_fiducials(golden_dir):
    gd = load(golden_dir, "plmc_hydro.json")
    g, dist = MT19937_64(42), UniformReal(0.1, 5)
    _dx, _dt = dist(g), dist(g)
    nx, ny, nz, nf = gd["nx"], gd["ny"], gd["nz"], gd["n_fields"]
    U = np.array([dist(g) for _ in range(nx * ny * nz * nf)])
    shapes = [(nx, ny, nz), (ny, nz, nx), (nz, nx, ny)]
    for d in range(3):
        QL, QR = oracle.reconstruct(oracle.PLMC, U, *shapes[d], gd["gamma"], d)
        _check_sparse(QL, gd["left"][d], "plmc L dir %d" % d)
        _check_sparse(QR, gd["right"][d], "plmc R dir %d" % d)


def test_ppmc_vl_hydro_fiducials(golden_dir):
    gd = load(golden_dir, "ppmc_vl.json")
    g, dist = MT19937_64(42), UniformReal(0.1, 5)
    n = gd["nx"]
    nf = gd["hydro"]["n_fields"]
    N = n ** 3
    U = np.array([dist(g) for _ in range(N * nf)])
    for d in range(3):
        QL, QR = oracle.reconstruct(oracle.PPMC, U, n, n, n, gd["gamma"], d)
        # the reference test only allocates (nf-1)*N interface words: the energy slot is never checked and
        # index 54 of the 'right' tables is an out-of-bounds artefact (see the fixture's note)
        _check_sparse(QL[: (nf - 1) * N], gd["hydro"]["left"][d], "ppmc L dir %d" % d)
        _check_sparse(QR[: (nf - 1) * N], gd["hydro"]["right"][d], "ppmc R dir %d" % d, skip=(54,))


def test_calc_dt_and_crossing_time(golden_dir):
    gd = load(golden_dir, "scalars.json")
    c = gd["calc_dt_3d"]
    U = np.array(c["conserved"], dtype=np.float64)
    dti = oracle.calc_dti(U, 1, 1, 1, c["n_ghost"], c["dx"], c["dy"], c["dz"], c["gamma"])
    assert nearly_equal(dti, c["dti"]), (dti, c["dti"])
    t = gd["inverse_crossing_time"]
    v = oracle.hydro_inverse_crossing_time(t["energy"], t["density"], 1.0 / t["density"], *t["velocity"],
                                           *t["cell_size"], t["gamma"])
    assert nearly_equal(v, t["hydro"]), (v, t["hydro"])


def test_hllc_uniform_state(golden_dir):
    h = load(golden_dir, "scalars.json")["hllc_uniform"]
    E = h["pressure"] / (h["gamma"] - 1)
    s = [h["density"], 0.0, 0.0, 0.0, E]
    f = oracle.riemann_single(oracle.HLLC, s, s, h["gamma"])
    for got, want in zip(f, h["flux"]):
        assert nearly_equal(float(got), want)
    f = oracle.riemann_single(oracle.ROE, s, s, h["gamma"])
    for got, want in zip(f, h["flux"]):
        assert nearly_equal(float(got), want)
