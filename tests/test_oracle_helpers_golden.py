"""Pin the building blocks of the CPU oracle (oracle/vl_oracle.c) to the reference's HELPER-level unit tests: the
known answers of src/reconstruction/reconstruction_tests.cu (Thread_Guard, Load_Data, Compute_Slope, Van_Leer_Slope,
Monotonize_Characteristic_Return_Primitive, Calc_Interface_Linear, Write_Data, Plm_Limit_Interfaces),
src/utils/hydro_utilities_tests.cpp (Get_Pressure_From_DE), src/utils/math_utilities_tests.cpp (dotProduct,
SquareMagnitude) and src/hydro/hydro_cuda_tests.cu (Scalar_Floor_Kernel), hydro build.  The numbers were taken from the
reference's test sources by tests/golden/extract_helper_fiducials.py (committed with its output, helpers.json).

Tolerance is the reference's: testing_utilities::nearlyEqualDbl = <= 4 ULP or <= 1e-14 absolute."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle
from test_oracle_golden import nearly_equal


@pytest.fixture(scope="module")
def gd(golden_dir):
    with open(os.path.join(golden_dir, "helpers.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def L():
    lib = oracle.lib()
    d, i, p, l = C.c_double, C.c_int, C.c_void_p, C.c_long
    lib.orc_test_load_data.argtypes = [p, l, l, i, i, i, d, p]
    lib.orc_test_write_data.argtypes = [p, p, l, l, i, i, i, d]
    lib.orc_test_thread_guard.argtypes = [i] * 7
    lib.orc_test_thread_guard.restype = i
    lib.orc_test_compute_slope.argtypes = [p, p, d, p]
    lib.orc_test_van_leer_slope.argtypes = [p, p, p]
    lib.orc_test_monotonize_characteristic.argtypes = [p, p, p, p, p, d, d, p]
    lib.orc_test_plm_limit_interfaces.argtypes = [p] * 5
    lib.orc_test_square_magnitude.argtypes = [d, d, d]
    lib.orc_test_square_magnitude.restype = d
    lib.orc_test_de_pressure.argtypes = [d, d, d, d]
    lib.orc_test_de_pressure.restype = d
    return lib


def _a(v):
    return np.ascontiguousarray(v, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _close(got, want, what):
    for k, (g, w) in enumerate(zip(np.atleast_1d(got), np.atleast_1d(want))):
        assert nearly_equal(float(g), float(w)), (what, k, float(g), float(w))


def test_thread_guard(gd, L):
    t = gd["thread_guard"]
    nx, ny, nz = t["nx"], t["ny"], t["nz"]
    inside = [x + nx * (y + ny * z) for x in range(nx) for y in range(ny) for z in range(nz)
              if not L.orc_test_thread_guard(t["order"], nx, ny, nz, x, y, z)]
    assert inside == t["inside_ids"]  # reconstruction_tests.cu:186-187: only cell 86 of the 6^3 grid is inside for order 3


def _iota_grid(t):
    n = t["n"]
    N = n ** 3
    U = np.arange(5 * N, dtype=np.float64)
    return U, N, t["xid"] + n * (t["yid"] + n * t["zid"])


def test_load_data(gd, L):
    t = gd["load_data"]
    U, N, cell = _iota_grid(t)
    U[4 * N:5 * N] *= t["energy_scale"]  # "Up the energy part of the grid to avoid negative pressure"
    out = np.zeros(5)
    L.orc_test_load_data(_p(U), cell, N, 1, 2, 3, t["gamma"], _p(out))
    _close(out, t["fiducial"], "Load_Data")


def test_compute_slope_and_van_leer_slope(gd, L):
    t = gd["compute_slope"]
    out = np.zeros(5)
    L.orc_test_compute_slope(_p(_a(t["left"])), _p(_a(t["right"])), t["coef"], _p(out))
    _close(out, [t["fiducial"]] * 5, "Compute_Slope")
    t = gd["van_leer_slope"]
    L.orc_test_van_leer_slope(_p(_a(t["left"])), _p(_a(t["right"])), _p(out))
    _close(out, t["fiducial"], "Van_Leer_Slope")


def test_monotonize_characteristic_return_primitive(gd, L):
    t = gd["monotonize_characteristic"]
    a = t["sound_speed"]
    out = np.zeros(5)
    L.orc_test_monotonize_characteristic(_p(_a(t["primitive"])), _p(_a(t["del_a_L"])), _p(_a(t["del_a_R"])),
                                         _p(_a(t["del_a_C"])), _p(_a(t["del_a_G"])), a, a * a, _p(out))
    _close(out, t["fiducial"], "Monotonize_Characteristic_Return_Primitive")


def test_calc_interface_linear(gd, L):
    """Calc_Interface_Linear(primitive, slopes, sign) = primitive + sign * 0.5 * slopes (reconstruction.h:645-667): the
    expression plmc_cell_t writes out per component, evaluated here the same way."""
    t = gd["calc_interface_linear"]
    got = [p + t["coef"] * 0.5 * s for p, s in zip(t["left"], t["right"])]
    _close(got, t["fiducial"], "Calc_Interface_Linear")


def test_write_data(gd, L):
    t = gd["write_data"]
    _, N, cell = _iota_grid(t)
    Q = np.zeros(5 * N)
    L.orc_test_write_data(_p(_a(t["interface"])), _p(Q), cell, N, 1, 2, 3, t["gamma"])
    want = {int(k): v for k, v in t["fiducial"].items()}
    for k, v in enumerate(Q):
        assert nearly_equal(float(v), want.get(k, 0.0)), ("Write_Data", k, float(v))


def test_plm_limit_interfaces(gd, L):
    t = gd["plm_limit_interfaces"]
    l, r = _a(t["interface_l_iph"]), _a(t["interface_r_imh"])
    L.orc_test_plm_limit_interfaces(_p(l), _p(r), _p(_a(t["cell_im1"])), _p(_a(t["cell_i"])), _p(_a(t["cell_ip1"])))
    _close(l, t["interface_l_iph_fiducial"], "Plm_Limit_Interfaces l_iph")
    _close(r, t["interface_r_imh_fiducial"], "Plm_Limit_Interfaces r_imh")


def test_get_pressure_from_de(gd, L):
    t = gd["get_pressure_from_de"]
    got = [L.orc_test_de_pressure(E, ut, ua, t["gamma"]) for E, ut, ua in zip(t["E"], t["U_total"], t["U_advected"])]
    _close(got, t["fiducial"], "Get_Pressure_From_DE")


def test_dot_product_and_square_magnitude_association(gd, L):
    """math_utils::dotProduct / SquareMagnitude sum a1 b1 + ((a2 b2) + (a3 b3)) (math_utilities.h:78-98); the oracle's
    square_mag is that association, and a dot product through it must hit the reference's fiducial as well."""
    t = gd["square_magnitude"]
    assert nearly_equal(L.orc_test_square_magnitude(*t["a"]), t["fiducial"])
    t = gd["dot_product"]
    a, b = t["a"], t["b"]
    assert nearly_equal(a[0] * b[0] + ((a[1] * b[1]) + (a[2] * b[2])), t["fiducial"])


def test_scalar_floor_kernel(gd):
    t = gd["scalar_floor"]
    for v in t["inputs"]:  # below, above and at the floor (hydro_cuda_tests.cu:172-195)
        U = np.zeros(t["n_fields"])
        U[t["field_num"]] = v
        oracle.scalar_floor(U, 1, 1, 1, 0, t["field_num"], t["scalar_floor"])
        assert U[t["field_num"]] == max(v, t["scalar_floor"]) and not U[:5].any()


# ---- the -DMHD branches of the same reference tests, against oracle/mhd_oracle.c ------------------------------------
@pytest.fixture(scope="module")
def M():
    from oracle import mhd
    lib = mhd.lib()
    d, i, p, l = C.c_double, C.c_int, C.c_void_p, C.c_long
    lib.morc_test_load_data.argtypes = [p, i, i, i, i, i, l, i, i, i, d, p]
    lib.morc_test_write_data.argtypes = [p, p, p, l, l, i, i, i, d]
    lib.morc_test_monotonize_characteristic.argtypes = [p, p, p, p, p, p, d, p]
    return lib


def test_mhd_load_data(gd, M):
    t = gd["mhd_load_data"]
    n = t["n"]
    N = n ** 3
    U = np.arange(8 * N, dtype=np.float64)
    U[4 * N:5 * N] *= t["energy_scale"]
    out = np.zeros(8)
    M.morc_test_load_data(_p(U), t["xid"], t["yid"], t["zid"], n, n, N, 1, 2, 3, t["gamma"], _p(out))
    _close(out, t["fiducial"], "Load_Data (MHD)")


def test_mhd_monotonize_characteristic_return_primitive(gd, M):
    t = gd["mhd_monotonize_characteristic"]
    out = np.zeros(8)
    M.morc_test_monotonize_characteristic(_p(_a(t["primitive"])), _p(_a(t["del_a_L"])), _p(_a(t["del_a_R"])),
                                          _p(_a(t["del_a_C"])), _p(_a(t["del_a_G"])), _p(_a(t["eigenvectors"])),
                                          t["sound_speed"] ** 2, _p(out))
    _close(out, t["fiducial"], "Monotonize_Characteristic_Return_Primitive (MHD)")


def test_mhd_write_data(gd, M):
    t = gd["mhd_write_data"]
    n = t["n"]
    N = n ** 3
    cell = t["xid"] + n * (t["yid"] + n * t["zid"])
    Q, U = np.zeros(8 * N), np.zeros(8 * N)  # the test's conserved array is all zeros: the face-normal field is 0
    M.morc_test_write_data(_p(_a(t["interface"])), _p(Q), _p(U), cell, N, 1, 2, 3, t["gamma"])
    want = {int(k): v for k, v in t["fiducial"].items()}
    for k, v in enumerate(Q[:7 * N]):  # the reference's interface array has 7 fields per side (no normal field)
        assert nearly_equal(float(v), want.get(k, 0.0)), ("Write_Data (MHD)", k, float(v))


@pytest.mark.parametrize("build", ["hydro", "mhd"])
def test_pressure_and_energy_from_primitives(gd, M, build):
    """hydro_utilities::Calc_Pressure_Primitive / Calc_Energy_Primitive (hydro_utilities.h:34-45, 106-118; known answers
    hydro_utilities_tests.cpp:61-77, 155-171, 191-206): the routines Load_Data and Write_Data call.  The MHD oracle's
    functions with a zero field are the hydro build's."""
    d_, p_ = C.c_double, C.c_void_p
    for fn in (M.morc_test_pressure_primitive, M.morc_test_energy_primitive):
        fn.argtypes, fn.restype = [d_] * 9, d_
    t = gd["pressure_energy_primitive"]
    q, g, fid = t["params"], t["gamma"], t[build]
    for i in range(3):
        b = (q["magnetic_x"][i], q["magnetic_y"][i], q["magnetic_z"][i]) if build == "mhd" else (0.0, 0.0, 0.0)
        v = (q["d"][i], q["vx"][i], q["vy"][i], q["vz"][i])
        assert nearly_equal(M.morc_test_pressure_primitive(q["E"][i], *v, g, *b), fid["pressure"][i]), ("pressure", i)
        assert nearly_equal(M.morc_test_energy_primitive(q["P"][i], *v, g, *b), fid["energy"][i]), ("energy", i)
        assert nearly_equal(M.morc_test_energy_primitive(-q["P"][i], *v, g, *b), fid["energy_negative_pressure"][i]), ("E(-P)", i)
        # a velocity 1e4 times larger makes the pressure negative: exactly TINY_NUMBER (hydro_utilities_tests.cpp:97-110)
        assert M.morc_test_pressure_primitive(q["E"][i], q["d"][i], 1e4 * q["vx"][i], q["vy"][i], q["vz"][i], g, *b) == 1e-20


def test_kinetic_energy_from_velocity(gd, L):
    """hydro_utilities::Calc_Kinetic_Energy_From_Velocity = 0.5 d SquareMagnitude(v) (hydro_utilities.h; the dual-energy
    passes call it, VL_3D_cuda.cu:418, hydro_cuda.cu:789): the oracle's 0.5 * d * square_mag(...) on the reference's inputs."""
    t, q = gd["kinetic_energy_from_velocity"], gd["pressure_energy_primitive"]["params"]
    c = t["coef"]
    for i in range(3):
        got = 0.5 * (c * q["d"][i]) * L.orc_test_square_magnitude(c * q["vx"][i], c * q["vy"][i], c * q["vz"][i])
        assert nearly_equal(got, t["fiducial"][i]), (i, got)
