"""ctypes front end of oracle/liborc.so (the C restatement in vl_oracle.c).

TEST INFRASTRUCTURE ONLY -- see vl_oracle.c's header.  Arrays are numpy float64, SoA
[field][z][y][x] flattened exactly like the reference's `dev_conserved`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PCM, PLMC, PPMC, PLMP, PPMP = 0, 1, 2, 3, 4
HLLC, ROE, HLL, EXACT = 0, 1, 3, 4
_lib = None


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "vl_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liborc.so"])
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        d, i, p, l = C.c_double, C.c_int, C.c_void_p, C.c_long
        _lib.orc_ppm_single_variable.argtypes = [d] * 5 + [p, p]
        _lib.orc_reconstruct.argtypes = [i, p, p, p, i, i, i, d, i]
        _lib.orc_riemann.argtypes = [i, p, p, p, i, i, i, d, i]
        _lib.orc_riemann_single.argtypes = [i, p, p, d, p]
        _lib.orc_vl_step.argtypes = [p, i, i, i, i, d, d, d, d, d, i, i, d]
        _lib.orc_simple_step.argtypes = [p, i, i, i, i, d, d, d, d, d, i, i, d]
        _lib.orc_plmc_trace_cell.argtypes = [p, p, p, d, d, p, p]
        _lib.orc_hydro_inverse_crossing_time.argtypes = [d] * 10
        _lib.orc_hydro_inverse_crossing_time.restype = d
        _lib.orc_calc_dti.argtypes = [p, i, i, i, i, d, d, d, d]
        _lib.orc_calc_dti.restype = d
        _lib.orc_temperature_floor.argtypes = [p, i, i, i, i, d]
        _lib.orc_vl_step_x.argtypes = [p, i, i, i, i, d, d, d, d, d, i, i, d, i, i]
        _lib.orc_vl_step_x.restype = i
        _lib.orc_temperature_floor_x.argtypes = [p, i, i, i, i, d, i, i]
        _lib.orc_scalar_floor.argtypes = [p, i, i, i, i, i, d]
        _lib.orc_cooling_update.argtypes = [p, i, i, i, i, i, i, d, d]
        _lib.orc_cie_cool.argtypes = [d, d]
        _lib.orc_cie_cool.restype = d
        _lib.orc_static_gravity_field.argtypes = [p, i, i, i, i, i, i, i, i, d, d, d, d, d, d]
        _lib.orc_set_gravity.argtypes = [p]
        _lib.orc_set_gravity_potential.argtypes = [p, i]
        _lib.orc_boundary_extents.argtypes = [i, i, i, i, i, p, p]
        _lib.orc_set_ghost_cells.argtypes = [p, i, i, i, i, i, p, i]
        _lib.orc_set_boundaries.argtypes = [p, i, i, i, i, i, p]
        _lib.orc_pack.argtypes = [p, p, i, i, i, l, l, i, i, i]
        _lib.orc_unpack.argtypes = [p, p, i, i, i, l, l, i, i, i]
    return _lib


def _ptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def ppm_single_variable(cm2, cm1, c0, cp1, cp2):
    L, R = C.c_double(), C.c_double()
    lib().orc_ppm_single_variable(cm2, cm1, c0, cp1, cp2, C.byref(L), C.byref(R))
    return L.value, R.value


def reconstruct(recon, U, nx, ny, nz, gamma, direction):
    """One launch of PLMC_cuda / PPMC_VL: returns (Q_L, Q_R), zero where the guard skips."""
    QL, QR = np.zeros_like(U), np.zeros_like(U)
    lib().orc_reconstruct(recon, _ptr(U), _ptr(QL), _ptr(QR), nx, ny, nz, gamma, direction)
    return QL, QR


def riemann(solver, QL, QR, nx, ny, nz, gamma, direction):
    F = np.zeros_like(QL)
    lib().orc_riemann(solver, _ptr(QL), _ptr(QR), _ptr(F), nx, ny, nz, gamma, direction)
    return F


def riemann_single(solver, left, right, gamma):
    l = np.ascontiguousarray(left, dtype=np.float64)
    r = np.ascontiguousarray(right, dtype=np.float64)
    f = np.zeros(5)
    lib().orc_riemann_single(solver, _ptr(l), _ptr(r), gamma, _ptr(f))
    return f


def vl_step(U, nx, ny, nz, n_ghost, dx, dy, dz, dt, gamma, recon, solver, density_floor=0.0):
    """In-place VL_Algorithm_{1,2,3}D_CUDA on U (flat SoA incl. ghosts)."""
    lib().orc_vl_step(_ptr(U), nx, ny, nz, n_ghost, dx, dy, dz, dt, gamma, recon, solver, density_floor)
    return U


def vl_step_x(U, nx, ny, nz, n_ghost, dx, dy, dz, dt, gamma, recon, solver, density_floor=0.0, n_scalars=0,
              dual_energy=False):
    """In-place VL step of a -DSCALAR / -DDE build: 5 + n_scalars + dual_energy fields (GasEnergy last)."""
    if lib().orc_vl_step_x(_ptr(U), nx, ny, nz, n_ghost, dx, dy, dz, dt, gamma, recon, solver, density_floor, n_scalars,
                           int(bool(dual_energy))) != 0:
        raise NotImplementedError("passive fields are restated for PCM / PLMC / PPMC with HLLC / Roe / HLL")
    return U


def temperature_floor_x(U, nx, ny, nz, n_ghost, U_floor, n_fields, dual_energy):
    lib().orc_temperature_floor_x(_ptr(U), nx, ny, nz, n_ghost, U_floor, n_fields, int(bool(dual_energy)))
    return U


def scalar_floor(U, nx, ny, nz, n_ghost, field_num, floor):
    lib().orc_scalar_floor(_ptr(U), nx, ny, nz, n_ghost, field_num, floor)
    return U


def cooling_update(U, nx, ny, nz, n_ghost, n_fields, dual_energy, dt, gamma):
    """Cooling_Update with the CIE curve (cooling_cuda.cu:20-196, 293-317), in place."""
    lib().orc_cooling_update(_ptr(U), nx, ny, nz, n_ghost, n_fields, int(bool(dual_energy)), dt, gamma)
    return U


def cie_cool(n, T):
    return lib().orc_cie_cool(n, T)


_gravity_keepalive = None


def static_gravity_field(nx, ny, nz, n_ghost, custom_grav, offsets, dx, dy, dz, bounds):
    """calc_g_{1,2,3}D (static_grav.h) for every cell of the local array: float64 [ndim * n_cells]."""
    ndim = 1 + (ny > 1) + (nz > 1)
    g = np.zeros(ndim * nx * ny * nz)
    lib().orc_static_gravity_field(_ptr(g), nx, ny, nz, n_ghost, custom_grav, offsets[0], offsets[1], offsets[2], dx, dy, dz,
                                   bounds[0], bounds[1], bounds[2])
    return g


def set_gravity(g):
    """The acceleration field every later vl_step / vl_step_x adds the -DSTATIC_GRAV source terms of (None: off)."""
    global _gravity_keepalive
    _gravity_keepalive = None if g is None else np.ascontiguousarray(g, dtype=np.float64)
    lib().orc_set_gravity(None if g is None else _ptr(_gravity_keepalive))


_potential_keepalive = None


def set_gravity_potential(pot, five_point=True):
    """The potential (ghost cells included) whose gradient every later 3-D vl_step / vl_step_x adds the -DGRAVITY source
    terms of (None: off)."""
    global _potential_keepalive
    _potential_keepalive = None if pot is None else np.ascontiguousarray(pot, dtype=np.float64).reshape(-1)
    lib().orc_set_gravity_potential(None if pot is None else _ptr(_potential_keepalive), int(bool(five_point)))


def simple_step(U, nx, ny, nz, n_ghost, dx, dy, dz, dt, gamma, recon, solver, density_floor=0.0):
    """In-place Simple_Algorithm_{1,2,3}D_CUDA on U (simple_3D_cuda.cu:25-180): PCM or PLMC with characteristic tracing."""
    if lib().orc_simple_step(_ptr(U), nx, ny, nz, n_ghost, dx, dy, dz, dt, gamma, recon, solver, density_floor) != 0:
        raise NotImplementedError("the SIMPLE integrator is restated for PCM and PLMC")
    return U


def plmc_trace_cell(wm, wi, wp, gamma, dtodx):
    """One PLMC cell of a -DSIMPLE build (plmc_cuda.cu:60-299): (L at i+1/2, R at i-1/2) primitives."""
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (wm, wi, wp)]
    L, R = np.zeros(5), np.zeros(5)
    lib().orc_plmc_trace_cell(_ptr(a[0]), _ptr(a[1]), _ptr(a[2]), gamma, dtodx, _ptr(L), _ptr(R))
    return L, R


def hydro_inverse_crossing_time(E, d, d_inv, vx, vy, vz, dx, dy, dz, gamma):
    return lib().orc_hydro_inverse_crossing_time(E, d, d_inv, vx, vy, vz, dx, dy, dz, gamma)


def calc_dti(U, nx, ny, nz, n_ghost, dx, dy, dz, gamma):
    return lib().orc_calc_dti(_ptr(U), nx, ny, nz, n_ghost, dx, dy, dz, gamma)


def temperature_floor(U, nx, ny, nz, n_ghost, U_floor):
    lib().orc_temperature_floor(_ptr(U), nx, ny, nz, n_ghost, U_floor)
    return U


def boundary_extents(direction, nx, ny, nz, n_ghost):
    lo, hi = (C.c_int * 3)(), (C.c_int * 3)()
    lib().orc_boundary_extents(direction, nx, ny, nz, n_ghost, lo, hi)
    return list(lo), list(hi)


def set_ghost_cells(U, nx, ny, nz, n_fields, n_ghost, flags, direction):
    fl = (C.c_int * 6)(*flags)
    lib().orc_set_ghost_cells(_ptr(U), nx, ny, nz, n_fields, n_ghost, fl, direction)
    return U


def set_boundaries(U, nx, ny, nz, n_fields, n_ghost, flags):
    fl = (C.c_int * 6)(*flags)
    lib().orc_set_boundaries(_ptr(U), nx, ny, nz, n_fields, n_ghost, fl)
    return U


def pack(U, nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize):
    buf = np.zeros(n_fields * isize * jsize * ksize)
    lib().orc_pack(_ptr(buf), _ptr(U), nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize)
    return buf


def unpack(buf, U, nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize):
    lib().orc_unpack(_ptr(buf), _ptr(U), nx, ny, n_fields, n_cells, idxoffset, isize, jsize, ksize)
    return U


def advance(U, nx, ny, nz, n_ghost, dx, dy, dz, gamma, recon, solver, flags, n_steps, cfl=0.3,
            density_floor=0.0, U_floor=0.0, integrator="vl", n_scalars=0, dual_energy=False, cooling=False):
    """The reference's main loop body (main.cpp:247-379) without I/O: dt from the CFL reduction, VL (or SIMPLE) step,
    temperature floor, ghost fill.  Returns the list of dt used."""
    dts = []
    step = {"vl": vl_step, "simple": simple_step}[integrator]
    n_fields = 5 + n_scalars + int(bool(dual_energy))
    if n_fields > 5:
        assert integrator == "vl"
        dti = calc_dti(U, nx, ny, nz, n_ghost, dx, dy, dz, gamma)
        for _ in range(n_steps):
            dt = cfl / dti
            vl_step_x(U, nx, ny, nz, n_ghost, dx, dy, dz, dt, gamma, recon, solver, density_floor, n_scalars, dual_energy)
            temperature_floor_x(U, nx, ny, nz, n_ghost, U_floor, n_fields, dual_energy)
            if cooling:
                cooling_update(U, nx, ny, nz, n_ghost, n_fields, dual_energy, dt, gamma)
            dti = calc_dti(U, nx, ny, nz, n_ghost, dx, dy, dz, gamma)
            set_boundaries(U, nx, ny, nz, n_fields, n_ghost, flags)
            dts.append(dt)
        return dts
    dti = calc_dti(U, nx, ny, nz, n_ghost, dx, dy, dz, gamma)
    for _ in range(n_steps):
        dt = cfl / dti
        step(U, nx, ny, nz, n_ghost, dx, dy, dz, dt, gamma, recon, solver, density_floor)
        temperature_floor(U, nx, ny, nz, n_ghost, U_floor)
        if cooling:
            cooling_update(U, nx, ny, nz, n_ghost, 5, False, dt, gamma)
        dti = calc_dti(U, nx, ny, nz, n_ghost, dx, dy, dz, gamma)
        set_boundaries(U, nx, ny, nz, 5, n_ghost, flags)
        dts.append(dt)
    return dts
