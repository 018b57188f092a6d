"""Closed-form initial conditions used by the tests and bench.py (host, numpy, float64).

Formulas follow the reference's generators (src/grid/initial_conditions.cpp: Constant :180, Sound_Wave :244,
Riemann :509, KH_res_ind :749, Spherical_Overpressure_3D :1203) with cell centres from the global-index rule
of Grid3D::Get_Position (grid3D.cpp:108-110), so the fields are independent of the rank count.  Each returns
conserved real cells as an array (5, nz, ny, nx).  They are inputs only: the same array is handed to the CUDA
path, the CPU oracle and the compiled reference, so parity never depends on how it was generated.
"""
import numpy as np


_BLOCK = None  # (offset_xyz, n_local_xyz): restrict the generators to one rank's block of the global grid


class block:
    """with block(offset, n_local): fields = sound_wave(n_global, ...)  -> only that rank's cells are built."""

    def __init__(self, offset, n_local):
        self.b = (tuple(offset), tuple(n_local))

    def __enter__(self):
        global _BLOCK
        self.prev, _BLOCK = _BLOCK, self.b
        return self

    def __exit__(self, *a):
        global _BLOCK
        _BLOCK = self.prev


def _centres(n, length, lo, axis):
    d = length / n
    i = np.arange(n) if _BLOCK is None else _BLOCK[0][axis] + np.arange(_BLOCK[1][axis])
    return lo + i * d + 0.5 * d  # Get_Position: bound + global_index * dx + dx/2 (grid3D.cpp:108-110)


def _shape(n):
    return (n[2], n[1], n[0]) if _BLOCK is None else (_BLOCK[1][2], _BLOCK[1][1], _BLOCK[1][0])


def _mesh(n, lengths, lo):
    x = _centres(n[0], lengths[0], lo[0], 0)
    y = _centres(n[1], lengths[1], lo[1], 1)
    z = _centres(n[2], lengths[2], lo[2], 2)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij", sparse=True)
    return X, Y, Z


def _pack(d, mx, my, mz, E):
    shape = np.broadcast_shapes(*(np.shape(a) for a in (d, mx, my, mz, E)))
    out = np.empty((5,) + shape, dtype=np.float64)
    for f, a in enumerate((d, mx, my, mz, E)):
        out[f] = a
    return out


def constant(n, rho=1.0, v=(0.0, 0.0, 0.0), P=1.0, gamma=5.0 / 3.0):
    d = np.full(_shape(n), rho)
    E = P / (gamma - 1.0) + 0.5 * rho * (v[0] ** 2 + v[1] ** 2 + v[2] ** 2)
    return _pack(d, d * v[0], d * v[1], d * v[2], np.full_like(d, E))


def sound_wave(n, lengths=(4.0, 4.0, 4.0), lo=(0.0, 0.0, 0.0), rho=1.0, v=(0.0, 0.0, 0.0), P=0.6, A=1e-4,
               gamma=1.666666666666667):
    """examples/3D/sound_wave.txt"""
    X, _, _ = _mesh(n, lengths, lo)
    s = A * np.sin(2.0 * np.pi * X) + 0.0 * (_mesh(n, lengths, lo)[1] + _mesh(n, lengths, lo)[2])
    d = rho + s
    E = P / (gamma - 1.0) + 0.5 * rho * (v[0] ** 2 + v[1] ** 2 + v[2] ** 2) + 1.5 * s
    return _pack(d, rho * v[0] + s, rho * v[1] + s, rho * v[2] + s, E)


def riemann(n, lengths=(1.0, 1.0, 1.0), lo=(0.0, 0.0, 0.0), left=(1.0, 0.0, 0.0, 0.0, 1.0),
            right=(0.1, 0.0, 0.0, 0.0, 0.1), diaph=0.5, gamma=1.4):
    """examples/1D/sod.txt and examples/3D/sod256.txt: (rho, vx, vy, vz, P) either side of x = diaph."""
    X, _, _ = _mesh(n, lengths, lo)
    X = X + 0.0 * (_mesh(n, lengths, lo)[1] + _mesh(n, lengths, lo)[2])
    L = X < diaph
    pick = lambda i: np.where(L, left[i], right[i])
    d, vx, vy, vz, P = (pick(i) for i in range(5))
    E = P / (gamma - 1.0) + 0.5 * d * (vx * vx + vy * vy + vz * vz)
    return _pack(d, d * vx, d * vy, d * vz, E)


def spherical_overpressure(n, lengths=(1.0, 1.0, 1.0), lo=(0.0, 0.0, 0.0), gamma=1.66666667):
    """Spherical_Overpressure_3D: rho=0.1, P=1 background; rho=1, P=11 inside r < 0.2 of the box centre."""
    X, Y, Z = _mesh(n, lengths, lo)
    c = [lo[i] + 0.5 * lengths[i] for i in range(3)]
    r = np.sqrt((X - c[0]) ** 2 + (Y - c[1]) ** 2 + (Z - c[2]) ** 2)
    inside = r < 0.2
    d = np.where(inside, 1.0, 0.1)
    P = np.where(inside, 11.0, 1.0)
    zero = np.zeros_like(d)
    return _pack(d, zero, zero, zero, P / (gamma - 1.0))


def kh_res_ind(n, lengths=(1.0, 1.0, 1.0), lo=(0.0, -0.5, -0.5), gamma=5.0 / 3.0):
    """KH_res_ind, 3-D branch (examples/3D/KH_res_ind_3D.txt): dense cylinder along x sheared against the
    ambient gas, resolution-independent ramp."""
    X, Y, Z = _mesh(n, lengths, lo)
    d1, d2, v1, v2, P, dy, A = 100.0, 1.0, 0.5, -0.5, 2.5, 0.05, 0.1
    r = np.sqrt(Z * Z + Y * Y)
    phi = np.arctan2(Z, Y)
    w = np.sqrt(-2.0 * dy * dy * np.log(0.5))
    e_m = np.exp(-0.5 * (r - 0.25 - w) ** 2 / (dy * dy))
    e_p = np.exp(-0.5 * (r - 0.25 + w) ** 2 / (dy * dy))
    inside = r < 0.25
    d = np.where(inside, d1 - (d1 - d2) * e_m, d2 + (d1 - d2) * e_p)
    mx = np.where(inside, v1 * d - d * e_m, v2 * d + d * e_p)
    pert = d * A * np.sin(4 * np.pi * X) * np.where(inside, e_p, 1.0 - e_p)
    my, mz = np.cos(phi) * pert, np.sin(phi) * pert
    E = P / (gamma - 1.0) + 0.5 * (mx * mx + my * my + mz * mz) / d
    return _pack(d, mx, my, mz, E)


def random_smooth(n, seed=0, gamma=5.0 / 3.0):
    """Seeded, strictly positive, smooth-ish random state that exercises every limiter branch."""
    rng = np.random.default_rng(seed)
    shape = _shape(n)
    d = 0.5 + rng.random(shape)
    vx, vy, vz = (rng.random(shape) - 0.5 for _ in range(3))
    P = 0.5 + rng.random(shape)
    E = P / (gamma - 1.0) + 0.5 * d * (vx * vx + vy * vy + vz * vz)
    return _pack(d, d * vx, d * vy, d * vz, E)


# ---- MHD (8 fields, face-centred magnetic field) ---------------------------------------------------------------
# These return the FULL local array, ghost cells included, as (8, nz + 2g, ny + 2g, nx + 2g): the face-centred field is
# the discrete curl of a vector potential sampled over the ghost zones too, so div B = 0 to round-off everywhere and
# periodic problems need no initial ghost fill.  magnetic_x[id] sits on the +x face of cell id (grid_enum.h:57-61).
def _mhd_mesh(n, g, lengths, lo):
    d = [lengths[i] / n[i] for i in range(3)]
    c = [lo[i] + (np.arange(n[i] + 2 * g) - g) * d[i] + 0.5 * d[i] for i in range(3)]  # Get_Position, ghosts included
    Z, Y, X = np.meshgrid(c[2], c[1], c[0], indexing="ij")
    return X, Y, Z, d


def _curl_faces(Ax, Ay, Az, d):
    """Init_Magnetic_Field_With_Vector_Potential (src/utils/mhd_utilities.cu:22-44): backward differences of a
    cell-sampled potential; index 0 of every axis is left at zero like the reference."""
    B = np.zeros((3,) + Ax.shape)
    s = (slice(1, None),) * 3
    zm, ym, xm = (slice(0, -1), s[1], s[2]), (s[0], slice(0, -1), s[2]), (s[0], s[1], slice(0, -1))
    B[0][s] = (Az[s] - Az[ym]) / d[1] - (Ay[s] - Ay[zm]) / d[2]
    B[1][s] = (Ax[s] - Ax[zm]) / d[2] - (Az[s] - Az[xm]) / d[0]
    B[2][s] = (Ay[s] - Ay[xm]) / d[0] - (Ax[s] - Ax[ym]) / d[1]
    return B


def _centred(B):
    """cellCenteredMagneticFields (mhd_utilities.h:205-236)."""
    c = B.copy()
    c[0][:, :, 1:] = 0.5 * (B[0][:, :, 1:] + B[0][:, :, :-1])
    c[1][:, 1:, :] = 0.5 * (B[1][:, 1:, :] + B[1][:, :-1, :])
    c[2][1:, :, :] = 0.5 * (B[2][1:, :, :] + B[2][:-1, :, :])
    return c


def _mhd_pack(d, vx, vy, vz, P, B, gamma):
    c = _centred(B)
    E = P / (gamma - 1.0) + 0.5 * d * (vx * vx + vy * vy + vz * vz) + 0.5 * (c[0] ** 2 + c[1] ** 2 + c[2] ** 2)
    out = np.empty((8,) + d.shape)
    for f, a in enumerate((d, d * vx, d * vy, d * vz, E, B[0], B[1], B[2])):
        out[f] = a
    return out


def orszag_tang(n, g, gamma=1.666666666666667):
    """Grid3D::Orszag_Tang_Vortex (initial_conditions.cpp:1926-1984, examples/3D/orszag_tang_vortex.txt)."""
    X, Y, Z, d = _mhd_mesh(n, g, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0))
    b0, d0, v0, p0 = 1.0 / np.sqrt(4.0 * np.pi), 25.0 / (36.0 * np.pi), 1.0, 5.0 / (12.0 * np.pi)
    Az = b0 / (4.0 * np.pi) * (np.cos(4.0 * np.pi * X) + 2.0 * np.cos(2.0 * np.pi * Y))
    B = _curl_faces(np.zeros_like(Az), np.zeros_like(Az), Az, d)
    rho = np.full_like(Az, d0)
    return _mhd_pack(rho, v0 * np.sin(2.0 * np.pi * Y), -v0 * np.sin(2.0 * np.pi * X), np.zeros_like(Az),
                     np.full_like(Az, p0), B, gamma)


def mhd_smooth(n, g, seed=0, gamma=5.0 / 3.0, amp=0.3, b_amp=0.5):
    """Seeded periodic state built from a few low Fourier modes: every wave family is excited in every direction,
    the solution stays smooth (well conditioned), and the field is the curl of a periodic potential."""
    rng = np.random.default_rng(seed)
    X, Y, Z, d = _mhd_mesh(n, g, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0))

    def modes(scale):
        f = np.zeros_like(X)
        for _ in range(4):
            k = rng.integers(-2, 3, size=3)
            ph = rng.random() * 2.0 * np.pi
            f = f + scale * (rng.random() - 0.5) * np.sin(2.0 * np.pi * (k[0] * X + k[1] * Y + k[2] * Z) + ph)
        return f
    rho = 1.0 + modes(amp)
    P = 1.0 + modes(amp)
    vx, vy, vz = modes(amp), modes(amp), modes(amp)
    A = [modes(b_amp / (2.0 * np.pi)) for _ in range(3)]
    B = _curl_faces(A[0], A[1], A[2], d)
    B[0] += 0.3
    B[1] += 0.2
    B[2] += 0.1
    return _mhd_pack(rho, vx, vy, vz, P, B, gamma)


def mhd_blast(n, g, gamma=5.0 / 3.0):
    """MHD_Spherical_Blast-like (initial_conditions.cpp:1865-1924): uniform oblique field, over-pressured sphere."""
    X, Y, Z, d = _mhd_mesh(n, g, (1.0, 1.0, 1.0), (-0.5, -0.5, -0.5))
    r = np.sqrt(X * X + Y * Y + Z * Z)
    P = np.where(r < 0.1, 10.0, 0.1)
    B = np.zeros((3,) + X.shape)
    B[0] += 1.0 / np.sqrt(2.0)
    B[1] += 1.0 / np.sqrt(2.0)
    z = np.zeros_like(X)
    return _mhd_pack(np.ones_like(X), z, z, z, P, B, gamma)
