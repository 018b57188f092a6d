"""ctypes front end of the MHD part of oracle/liborc.so (mhd_oracle.c).  TEST INFRASTRUCTURE ONLY.

Arrays are numpy float64, SoA [field][z][y][x] flattened like the reference's `dev_conserved`; 8 fields
(density, momentum xyz, Energy, magnetic xyz with magnetic_d[id] on the +d face of cell id)."""
import ctypes as C

import numpy as np

from . import oracle as _o

PCM, PLMC, PPMC = 0, 1, 2
NF, NQ = 8, 7
_set = False


def lib():
    global _set
    L = _o.lib()
    if not _set:
        d, i, p = C.c_double, C.c_int, C.c_void_p
        for name in ("morc_fast_speed", "morc_slow_speed"):
            getattr(L, name).argtypes = [d] * 6
            getattr(L, name).restype = d
        L.morc_alfven_speed.argtypes = [d, d]
        L.morc_alfven_speed.restype = d
        L.morc_cell_centred_b.argtypes = [p, i, i, i, i, i, i, p]
        L.morc_reconstruct.argtypes = [i, p, p, p, i, i, i, d, i]
        L.morc_hlld_single.argtypes = [p, p, d, d, p]
        L.morc_hlld_single.restype = i
        L.morc_ct_electric_fields.argtypes = [p, p, p, p, p, i, i, i]
        L.morc_update_magnetic.argtypes = [p, p, p, i, i, i, d, d, d, d]
        L.morc_create.argtypes = [i, i, i]
        L.morc_create.restype = p
        L.morc_destroy.argtypes = [p]
        L.morc_vl_step.argtypes = [p, p, i, d, d, d, d, d, i, d]
        L.morc_mhd_inverse_crossing_time.argtypes = [d] * 13
        L.morc_mhd_inverse_crossing_time.restype = d
        L.morc_calc_dti.argtypes = [p, i, i, i, i, d, d, d, d]
        L.morc_calc_dti.restype = d
        L.morc_max_divergence.argtypes = [p, i, i, i, i, d, d, d]
        L.morc_max_divergence.restype = d
        L.morc_check_magnetic_divergence.argtypes = [p, i, i, i, d, d, d]
        L.morc_check_magnetic_divergence.restype = d
        L.morc_eigenvectors.argtypes = [p, d, p]
        L.morc_prim_to_char.argtypes = [p, p, p, d, p]
        L.morc_char_to_prim.argtypes = [p, p, p, d, p]
        for name in ("morc_flux", "morc_efield", "morc_half"):
            getattr(L, name).restype = p
        L.morc_flux.argtypes = [p, i]
        L.morc_efield.argtypes = [p]
        L.morc_half.argtypes = [p]
        _set = True
    return L


def _ptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _arr(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def fast_speed(d, p, bx, by, bz, gamma):
    return lib().morc_fast_speed(d, p, bx, by, bz, gamma)


def slow_speed(d, p, bx, by, bz, gamma):
    return lib().morc_slow_speed(d, p, bx, by, bz, gamma)


def alfven_speed(bx, d):
    return lib().morc_alfven_speed(bx, d)


def cell_centred_b(U, nx, ny, nz, x, y, z):
    out = np.zeros(3)
    lib().morc_cell_centred_b(_ptr(U), nx, ny, nz, x, y, z, _ptr(out))
    return out


def reconstruct(recon, U, nx, ny, nz, gamma, direction):
    """One launch of PLMC_cuda / PPMC_VL (MHD): (Q_L, Q_R), 7 interface fields, zero where the guard skips."""
    N = nx * ny * nz
    QL, QR = np.zeros(NQ * N), np.zeros(NQ * N)
    lib().morc_reconstruct(recon, _ptr(U), _ptr(QL), _ptr(QR), nx, ny, nz, gamma, direction)
    return QL, QR


def hlld_single(ql, qr, bx, gamma):
    """Sweep-frame states (d, m1, m2, m3, E, By, Bz) -> flux (same layout), or None when no branch fires."""
    ql, qr, f = _arr(ql), _arr(qr), np.zeros(7)
    ok = lib().morc_hlld_single(_ptr(ql), _ptr(qr), bx, gamma, _ptr(f))
    return f if ok else None


def ct_electric_fields(Fx, Fy, Fz, U, nx, ny, nz, fill=0.0):
    E = np.full(3 * nx * ny * nz, fill)
    lib().morc_ct_electric_fields(_ptr(Fx), _ptr(Fy), _ptr(Fz), _ptr(U), _ptr(E), nx, ny, nz)
    return E


def update_magnetic(src, dst, E, nx, ny, nz, dt, dx, dy, dz):
    lib().morc_update_magnetic(_ptr(src), _ptr(dst), _ptr(E), nx, ny, nz, dt, dx, dy, dz)
    return dst


def mhd_inverse_crossing_time(E, d, d_inv, vx, vy, vz, bx, by, bz, dx, dy, dz, gamma):
    return lib().morc_mhd_inverse_crossing_time(E, d, d_inv, vx, vy, vz, bx, by, bz, dx, dy, dz, gamma)


def calc_dti(U, nx, ny, nz, n_ghost, dx, dy, dz, gamma):
    return lib().morc_calc_dti(_ptr(U), nx, ny, nz, n_ghost, dx, dy, dz, gamma)


def max_divergence(U, nx, ny, nz, n_ghost, dx, dy, dz):
    return lib().morc_max_divergence(_ptr(U), nx, ny, nz, n_ghost, dx, dy, dz)


def check_magnetic_divergence(U, nx, ny, nz, dx, dy, dz):
    """mhd::checkMagneticDivergence's range: every cell with x, y, z > 1, ghost cells included."""
    return lib().morc_check_magnetic_divergence(_ptr(U), nx, ny, nz, dx, dy, dz)


def eigenvectors(prim, gamma):
    prim, out = _arr(prim), np.zeros(18)
    lib().morc_eigenvectors(_ptr(prim), gamma, _ptr(out))
    return out


def prim_to_char(prim, slope, eig, gamma):
    prim, slope, eig, out = _arr(prim), _arr(slope), _arr(eig), np.zeros(7)
    lib().morc_prim_to_char(_ptr(prim), _ptr(slope), _ptr(eig), gamma, _ptr(out))
    return out


def char_to_prim(prim, chr_, eig, gamma):
    prim, chr_, eig, out = _arr(prim), _arr(chr_), _arr(eig), np.zeros(8)
    lib().morc_char_to_prim(_ptr(prim), _ptr(chr_), _ptr(eig), gamma, _ptr(out))
    return out


def set_boundaries(U, nx, ny, nz, n_ghost, flags):
    """Ghost fill of an MHD build: all 8 fields copied, no transmissive diode (cuda_boundaries.cu:121-147)."""
    _o.lib().orc_set_mhd_build(1)
    try:
        _o.set_boundaries(U, nx, ny, nz, NF, n_ghost, flags)
    finally:
        _o.lib().orc_set_mhd_build(0)
    return U


class Stepper:
    """The MHD branch of VL_Algorithm_3D_CUDA with the reference's persistent, zero-initialised scratch."""

    def __init__(self, nx, ny, nz):
        self.dims = (nx, ny, nz)
        self.ctx = C.c_void_p(lib().morc_create(nx, ny, nz))

    def step(self, U, n_ghost, dx, dy, dz, dt, gamma, recon, density_floor=0.0):
        lib().morc_vl_step(self.ctx, _ptr(U), n_ghost, dx, dy, dz, dt, gamma, recon, density_floor)
        return U

    def close(self):
        if self.ctx:
            lib().morc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        self.close()


def advance(U, nx, ny, nz, n_ghost, dx, dy, dz, gamma, recon, flags, n_steps, cfl=0.3, density_floor=0.0):
    """main.cpp:247-379 for an MHD build without I/O.  Returns the dt history."""
    S = Stepper(nx, ny, nz)
    dts = []
    dti = calc_dti(U, nx, ny, nz, n_ghost, dx, dy, dz, gamma)
    for _ in range(n_steps):
        dt = cfl / dti
        S.step(U, n_ghost, dx, dy, dz, dt, gamma, recon, density_floor)
        dti = calc_dti(U, nx, ny, nz, n_ghost, dx, dy, dz, gamma)
        set_boundaries(U, nx, ny, nz, n_ghost, flags)
        dts.append(dt)
    S.close()
    return dts
