"""Closed-form initial conditions used by the tests and bench.py (host, numpy, float64).

Formulas follow the reference's generators (src/grid/initial_conditions.cpp: Constant :180, Sound_Wave :244,
Linear_Wave :302, Square_Wave :437, Riemann :509, Shu_Osher :597, Blast_1D :635, KH :677, KH_res_ind :749,
Gresho :932, Implosion_2D :1033, Spherical_Overpressure_3D :1203) with cell centres from the global-index rule
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


def spherical_overdensity(n, lengths=(1.0, 1.0, 1.0), lo=(0.0, 0.0, 0.0), gamma=1.66666667):
    """Grid3D::Spherical_Overdensity_3D (initial_conditions.cpp:1253-1309): rho = 0.0005, P = 0.0005 background at rest;
    rho = 1 (same pressure) inside r < 0.2 of (0.5, 0.5, 0.5) -- the centre is a literal in the reference."""
    X, Y, Z = _mesh(n, lengths, lo)
    r = np.sqrt((X - 0.5) * (X - 0.5) + (Y - 0.5) * (Y - 0.5) + (Z - 0.5) * (Z - 0.5))
    d = np.where(r < 0.2, 1.0, 0.0005)
    zero = np.zeros_like(d)
    return _pack(d, zero, zero, zero, np.full_like(d, 0.0005) / (gamma - 1.0))


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


def _full(a, n, lengths, lo):
    """Broadcast a sparse-mesh expression to the block's full (nz, ny, nx) shape."""
    X, Y, Z = _mesh(n, lengths, lo)
    return a + 0.0 * (X + Y + Z)


def square_wave(n, lengths=(1.0, 1.0, 1.0), lo=(0.0, 0.0, 0.0), rho=1.0, v=(1.0, 0.0, 0.0), P=0.01, A=2.0,
                gamma=1.4):
    """Grid3D::Square_Wave (initial_conditions.cpp:437-502, examples/1D/square_wave.txt): density rho * A between a
    quarter and three quarters of the x extent, pressure equilibrium."""
    X, _, _ = _mesh(n, lengths, lo)
    X = _full(X, n, lengths, lo)
    d = np.where((X > 0.25 * lengths[0]) & (X < 0.75 * lengths[0]), rho * A, rho)
    E = P / (gamma - 1.0) + 0.5 * d * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2])
    return _pack(d, d * v[0], d * v[1], d * v[2], E)


def shu_osher(n, lengths=(2.0, 1.0, 1.0), lo=(-1.0, 0.0, 0.0), gamma=1.4):
    """Grid3D::Shu_Osher (initial_conditions.cpp:597-630, examples/1D/Shu_Osher.txt): Mach-3 shock into a sinusoidal
    density field.  1-D."""
    X, _, _ = _mesh(n, lengths, lo)
    X = _full(X, n, lengths, lo)
    shocked = X < -0.8
    d = np.where(shocked, 3.857143, 1.0 + 0.2 * np.sin(5.0 * np.pi * X))
    vx = np.where(shocked, 2.629369, 0.0)
    P = np.where(shocked, 10.33333, 1.0)
    z = np.zeros_like(d)
    return _pack(d, d * vx, z, z, P / (gamma - 1.0) + 0.5 * d * vx * vx)


def blast_1d(n, lengths=(1.0, 1.0, 1.0), lo=(0.0, 0.0, 0.0), gamma=1.4):
    """Grid3D::Blast_1D (initial_conditions.cpp:635-672, examples/1D/blast_1D.txt): two interacting blast waves (Stone
    2008, 8.1); the deck closes the tube with reflective boundaries."""
    X, _, _ = _mesh(n, lengths, lo)
    X = _full(X, n, lengths, lo)
    P = np.where(X < 0.1, 1000.0, np.where(X > 0.9, 100.0, 0.01))
    d = np.ones_like(P)
    z = np.zeros_like(P)
    return _pack(d, z, z, z, P / (gamma - 1.0))


def kh(n, lengths=(1.0, 1.0, 1.0), lo=(0.0, 0.0, 0.0), gamma=5.0 / 3.0):
    """Grid3D::KH (initial_conditions.cpp:677-744, examples/2D/KH_discontinuous_2D.txt): discontinuous shear layers at a
    quarter and three quarters of the y extent, d = 2 / 1, vx = +-0.5, vy = 0.1 sin(4 pi x), P = 2.5."""
    X, Y, _ = _mesh(n, lengths, lo)
    X, Y = _full(X, n, lengths, lo), _full(Y, n, lengths, lo)
    outer = (Y <= 1.0 * lengths[1] / 4.0) | (Y >= 3.0 * lengths[1] / 4.0)
    d = np.where(outer, 1.0, 2.0)
    mx = np.where(outer, -0.5, 0.5) * d
    my = d * 0.1 * np.sin(4 * np.pi * X)
    E = 2.5 / (gamma - 1.0) + 0.5 * (mx * mx + my * my) / d
    return _pack(d, mx, my, np.zeros_like(d), E)


def gresho(n, lengths=(1.0, 1.0, 1.0), lo=(-0.5, -0.5, 0.0), gamma=1.4):
    """Grid3D::Gresho (initial_conditions.cpp:932-1028, examples/2D/Gresho.txt): the stationary triangular vortex, point
    values at the cell centres."""
    X, Y, _ = _mesh(n, lengths, lo)
    X, Y = _full(X, n, lengths, lo), _full(Y, n, lengths, lo)
    r, phi = np.sqrt(X * X + Y * Y), np.arctan2(Y, X)
    inner, mid = r < 0.2, (r >= 0.2) & (r < 0.4)
    with np.errstate(divide="ignore", invalid="ignore"):
        p_mid = 9.0 - 4.0 * np.log(0.2) + 0.5 * 25.0 * r * r - 20.0 * r + 4.0 * np.log(r)
    vt = np.where(inner, 5.0 * r, np.where(mid, 2.0 - 5.0 * r, 0.0))
    P = np.where(inner, 5.0 + 0.5 * 25.0 * r * r, np.where(mid, p_mid, 3.0 + 4.0 * np.log(2.0)))
    vx, vy = -np.sin(phi) * vt, np.cos(phi) * vt
    d = np.ones_like(r)
    return _pack(d, d * vx, d * vy, np.zeros_like(d), P / (gamma - 1.0) + 0.5 * d * (vx * vx + vy * vy))


def implosion_2d(n, lengths=(0.3, 0.3, 1.0), lo=(0.0, 0.0, 0.0), gamma=1.4):
    """Grid3D::Implosion_2D (initial_conditions.cpp:1033-1071, examples/2D/Implosion_2D.txt; Liska & Wendroff 2003): low
    density and pressure below the diagonal x + y = 0.15."""
    X, Y, _ = _mesh(n, lengths, lo)
    X, Y = _full(X, n, lengths, lo), _full(Y, n, lengths, lo)
    low = Y < (0.1500001 - X)
    d = np.where(low, 0.125, 1.0)
    P = np.where(low, 0.14, 1.0)
    z = np.zeros_like(d)
    return _pack(d, z, z, z, P / (gamma - 1.0))


def linear_wave(n, lengths, lo, rho, vx, P, A, eigenvector, wave_length=1.0, pitch=0.0, yaw=0.0, gamma=5.0 / 3.0):
    """Grid3D::Linear_Wave, hydro build (initial_conditions.cpp:302-432): background (rho, vx, P) plus A sin(k x') times
    the right eigenvector (rho, momentum x / y / z, E) of the wave, x' rotated by pitch and yaw."""
    X, Y, Z = _mesh(n, lengths, lo)
    sy, cy, sp, cp = np.sin(yaw), np.cos(yaw), np.sin(pitch), np.cos(pitch)
    xr = cp * (X * cy + Y * sy) + Z * sp
    sine = _full(np.sin(xr * (2.0 * np.pi / wave_length)), n, lengths, lo)
    e_rho, e_mx, e_my, e_mz, e_E = eigenvector
    d = rho + A * e_rho * sine
    mx, my, mz = rho * vx + A * e_mx * sine, A * e_my * sine, A * e_mz * sine
    E = P / (gamma - 1.0) + 0.5 * rho * vx * vx + A * sine * e_E
    return _pack(d, mx * cp * cy - my * sy - mz * sp * cy, mx * cp * sy + my * cy - mz * sp * sy, mx * sp + mz * cp, E)


def random_smooth(n, seed=0, gamma=5.0 / 3.0):
    """Seeded, strictly positive, smooth-ish random state that exercises every limiter branch."""
    rng = np.random.default_rng(seed)
    shape = _shape(n)
    d = 0.5 + rng.random(shape)
    vx, vy, vz = (rng.random(shape) - 0.5 for _ in range(3))
    P = 0.5 + rng.random(shape)
    E = P / (gamma - 1.0) + 0.5 * d * (vx * vx + vy * vy + vz * vz)
    return _pack(d, d * vx, d * vy, d * vz, E)


def rayleigh_taylor(n, lengths=(0.33333333, 1.0, 1.0), lo=(0.0, 0.0, 0.0), gamma=1.4):
    """Grid3D::Rayleigh_Taylor (initial_conditions.cpp:881-931): a heavy slab (rho = 2) over a light one in hydrostatic
    balance with g = -0.1, a tapered cosine perturbation of v_y; 2-D, run with custom_grav = 2."""
    X, Y, _ = _mesh(n, lengths, lo)
    dl, du, g = 1.0, 2.0, -0.1
    vy = 0.01 * np.cos(6 * np.pi * X + np.pi) * np.exp(-(Y - 0.5 * lengths[1]) * (Y - 0.5 * lengths[1]) / 0.1)
    lower = Y <= 0.5 * lengths[1]
    d = np.where(lower, dl, du) + 0.0 * X
    P = np.where(lower, (1.0 / gamma - dl * g * 0.5) + dl * g * Y, (1.0 / gamma - du * g * 0.5) + du * g * Y) + 0.0 * X
    my = d * vy
    E = P / (gamma - 1.0) + 0.5 * (my * my) / d
    zero = np.zeros(np.broadcast(d, my).shape)
    return _pack(d + zero, zero, my + zero, zero, E + zero)


def basic_scalar_for(init, fields, P=None, lengths=None, lo=None, n=None):
    """The -DBASIC_SCALAR field the reference's generators set (initial_conditions.cpp:481-497 Square_Wave, 565-582
    Riemann, 719-732 KH): the density inside the marked region, zero elsewhere; generators that do not set it leave the
    calloc'ed zero."""
    d = fields[0]
    if init == "Riemann":
        X, _, _ = _mesh(n, lengths, lo)
        return np.where(X < P.diaph, 1.0 * P.rho_l, 0.0 * P.rho_r) + 0.0 * d
    if init == "Square_Wave":
        X, _, _ = _mesh(n, lengths, lo)
        return np.where((X > 0.25 * lengths[0]) & (X < 0.75 * lengths[0]), d * 1.0, d * 0.0)
    if init == "KH":
        _, Y, _ = _mesh(n, lengths, lo)
        return np.where((Y > 0.25 * lengths[1]) & (Y < 0.75 * lengths[1]), 1.0 * d, 0.0) + 0.0 * d
    return np.zeros_like(d)


def with_passive(fields, n_scalars=0, dual_energy=False, gamma=5.0 / 3.0, seed=1):
    """Append the passive fields of a -DSCALAR / -DDE build to a 5-field state: n_scalars scalars (conserved: density
    times a seeded mass fraction in (0.1, 0.9)), then GasEnergy = Energy - kinetic energy, which is what the reference's
    generators store (initial_conditions.cpp: `C.GasEnergy[id] = P / (gama - 1.0)`)."""
    f = np.asarray(fields)
    rng = np.random.default_rng(seed)
    extra = [f[0] * (0.1 + 0.8 * rng.random(f[0].shape)) for _ in range(int(n_scalars))]
    if dual_energy:
        extra.append(f[4] - 0.5 * (f[1] ** 2 + f[2] ** 2 + f[3] ** 2) / f[0])
    return np.concatenate([f, np.stack(extra)]) if extra else f.copy()


def cold_flow(n, seed=0, gamma=5.0 / 3.0, mach=40.0):
    """A fast, cold, perturbed stream: thermal energy is about a thousandth of the total, so that cells fall on both sides
    of the dual-energy switches (DE_ETA_1 = 0.001, DE_ETA_2 = 0.035; global.h:53-73)."""
    rng = np.random.default_rng(seed)
    shape = _shape(n)
    d = 0.8 + 0.4 * rng.random(shape)
    P = 0.6 * (0.5 + rng.random(shape))
    cs = np.sqrt(gamma * 0.6 / 1.0)
    v0 = mach * cs
    vx = v0 * (0.7 + 0.6 * rng.random(shape)) / np.sqrt(3.0)
    vy = v0 * (0.7 + 0.6 * rng.random(shape)) / np.sqrt(3.0) if len([m for m in n if m > 1]) > 1 else np.zeros(shape)
    vz = v0 * (0.7 + 0.6 * rng.random(shape)) / np.sqrt(3.0) if len([m for m in n if m > 1]) > 2 else np.zeros(shape)
    E = P / (gamma - 1.0) + 0.5 * d * (vx * vx + vy * vy + vz * vz)
    return _pack(d, d * vx, d * vy, d * vz, E)


# ---- MHD (8 fields, face-centred magnetic field) ---------------------------------------------------------------
# These return the FULL local array, ghost cells included, as (8, nz + 2g, ny + 2g, nx + 2g): the face-centred field is
# the discrete curl of a vector potential sampled over the ghost zones too, so div B = 0 to round-off everywhere and
# periodic problems need no initial ghost fill.  magnetic_x[id] sits on the +x face of cell id (grid_enum.h:57-61).
def _mhd_mesh(n, g, lengths, lo):
    """Cell centres of the local array, ghost zones included.  Inside `with block(offset, n_local)` only that rank's
    block (plus its ghost zones) of the global grid is built, from the same global-index rule, so the cells a rank
    owns do not depend on the rank count."""
    d = [lengths[i] / n[i] for i in range(3)]
    off, nl = ((0, 0, 0), n) if _BLOCK is None else _BLOCK
    # Get_Position: bound + global_index * dx + dx / 2 (grid3D.cpp:108-110), ghosts included
    c = [lo[i] + (off[i] + np.arange(nl[i] + 2 * g) - g) * d[i] + 0.5 * d[i] for i in range(3)]
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


def orszag_tang(n, g, gamma=1.666666666666667, lengths=(1.0, 1.0, 1.0), lo=(0.0, 0.0, 0.0)):
    """Grid3D::Orszag_Tang_Vortex (initial_conditions.cpp:1926-1984, examples/3D/orszag_tang_vortex.txt).  The
    reference's formulas take the cell positions of the deck's domain as they are (one period per unit length)."""
    X, Y, Z, d = _mhd_mesh(n, g, lengths, lo)
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


def mhd_blast(n, g, gamma=5.0 / 3.0, lengths=(1.0, 1.0, 1.0), lo=(-0.5, -0.5, -0.5), rho=1.0, v=(0.0, 0.0, 0.0), P=0.1,
              P_blast=10.0, radius=0.1, B=(0.70710678118654746, 0.70710678118654746, 0.0)):
    """Grid3D::MHD_Spherical_Blast (initial_conditions.cpp:1865-1924, examples/3D/mhd_blast.txt): uniform state and field
    from the deck (rho, vx.., P, Bx..), pressure P_blast inside `radius` of the origin.  Like the reference it insists
    on a domain centred on the origin and larger than the blast (its two asserts, :1869-1874)."""
    if any(lo[a] + 0.5 * lengths[a] != 0.0 for a in range(3)):
        raise ValueError("MHD_Spherical_Blast: the domain must be centred at zero (initial_conditions.cpp:1869)")
    if not np.sqrt(sum((0.5 * v_) ** 2 for v_ in lengths)) > radius:
        raise ValueError("MHD_Spherical_Blast: the domain must be larger than the blast radius (initial_conditions.cpp:1873)")
    X, Y, Z, d = _mhd_mesh(n, g, lengths, lo)
    r = np.sqrt(X * X + Y * Y + Z * Z)
    pres = np.where(r < radius, P_blast, P)
    Bf = np.zeros((3,) + X.shape)
    for a in range(3):
        Bf[a] += B[a]
    one = np.ones_like(X)
    return _mhd_pack(rho * one, v[0] * one, v[1] * one, v[2] * one, pres, Bf, gamma)


def mhd_riemann(n, g, lengths, lo, left, right, diaph, gamma):
    """Grid3D::Riemann, -DMHD build (initial_conditions.cpp:509-592; examples/3D/Brio_and_Wu.txt, Dai_and_Woodward.txt,
    Ryu_and_Jones_*.txt, Einfeldt_Strong_Rarefaction.txt): (rho, vx, vy, vz, P, Bx, By, Bz) either side of x = diaph.
    A face takes the side of the CENTRE of the cell it is stored with, and the energy uses the given field, not a
    face average -- both as in the reference."""
    X, Y, Z, d = _mhd_mesh(n, g, lengths, lo)
    L = X < diaph
    pick = lambda i: np.where(L, left[i], right[i])
    rho, vx, vy, vz, P, bx, by, bz = (pick(i) for i in range(8))
    E = np.maximum(P, 1e-20) / (gamma - 1.0) + 0.5 * rho * (vx * vx + ((vy * vy) + (vz * vz))) + 0.5 * (bx * bx + ((by * by) + (bz * bz)))
    out = np.empty((8,) + X.shape)
    for f, a in enumerate((rho, rho * vx, rho * vy, rho * vz, E, bx, by, bz)):
        out[f] = a
    return out


def mhd_linear_wave(n, g, lengths, lo, rho, vx, P, B0, A, eigenvector, wave_length=1.0, gamma=5.0 / 3.0, pitch=0.0,
                    yaw=0.0):
    """Grid3D::Linear_Wave, -DMHD build (initial_conditions.cpp:302-432): examples/3D/alfven_wave.txt,
    fast_magnetosonic.txt, slow_magnetosonic.txt, mhd_contact_wave.txt, and the reference's MHD linear-wave system
    tests, which also run the waves along y and z (yaw or pitch = pi / 2).  eigenvector = (rho, mx, my, mz, E, Bx, By,
    Bz).  Only the vector-potential variant of the rotated set-up is commented out in the reference (:312-358); the live
    code (:360-392) rotates the wave coordinate, the momenta and the field exactly as below, with the face fields
    evaluated at the cell-CENTRE coordinate.  The box is periodic, so the ghost zones are generated too."""
    X, Y, Z, d = _mhd_mesh(n, g, lengths, lo)
    sy, cy, sp, cp = np.sin(yaw), np.cos(yaw), np.sin(pitch), np.cos(pitch)
    xr = cp * (X * cy + Y * sy) + Z * sp
    sine = np.sin(xr * (2.0 * np.pi / wave_length)) + 0.0 * (X + Y + Z)
    e_rho, e_mx, e_my, e_mz, e_E, e_bx, e_by, e_bz = eigenvector
    out = np.empty((8,) + sine.shape)
    out[0] = rho + A * e_rho * sine
    mx, my, mz = rho * vx + A * e_mx * sine, A * e_my * sine, A * e_mz * sine
    out[1] = mx * cp * cy - my * sy - mz * sp * cy
    out[2] = mx * cp * sy + my * cy - mz * sp * sy
    out[3] = mx * sp + mz * cp
    out[4] = P / (gamma - 1.0) + 0.5 * rho * vx * vx + A * sine * e_E + 0.5 * (B0[0] * B0[0] + B0[1] * B0[1] + B0[2] * B0[2])
    bx, by, bz = B0[0] + A * e_bx * sine, B0[1] + A * e_by * sine, B0[2] + A * e_bz * sine
    out[5] = bx * cp * cy - by * sy - bz * sp * cy
    out[6] = bx * cp * sy + by * cy - bz * sp * sy
    out[7] = bx * sp + bz * cp
    return out


def mhd_constant(n, g, lengths, lo, rho=1.0, v=(0.0, 0.0, 0.0), P=1.0, B=(0.0, 0.0, 0.0), gamma=5.0 / 3.0):
    """Grid3D::Constant in an -DMHD build (initial_conditions.cpp:180-242): uniform state with a uniform field from the
    deck (Bx, By, Bz); used by the reference's MHD system tests (constant state must stay constant)."""
    X, Y, Z, d = _mhd_mesh(n, g, lengths, lo)
    one = np.ones_like(X + Y + Z)
    Bf = np.zeros((3,) + one.shape)
    for a in range(3):
        Bf[a] += B[a]
    return _mhd_pack(rho * one, v[0] * one, v[1] * one, v[2] * one, P * one, Bf, gamma)


def _mhd_pack_momentum(rho, mx, my, mz, P, B, gamma):
    """Calc_Energy_Conserved with the cell-centred field (hydro_utilities.h:120-135)."""
    c = _centred(B)
    E = np.maximum(P, 1e-20) / (gamma - 1.0) + 0.5 * (mx * mx + ((my * my) + (mz * mz))) / rho + 0.5 * (c[0] ** 2 + ((c[1] ** 2) + (c[2] ** 2)))
    out = np.empty((8,) + np.shape(B[0]))
    for f, a in enumerate((rho, mx, my, mz, E, B[0], B[1], B[2])):
        out[f] = a
    return out


def circularly_polarized_alfven_wave(n, g, lengths, lo, polarization=1.0, vx=0.0, pitch=0.0, yaw=0.0, gamma=5.0 / 3.0):
    """Grid3D::Circularly_Polarized_Alfven_Wave (initial_conditions.cpp:1680-1797,
    examples/3D/circularly_polarized_alfven_wave.txt): rho = 1, P = 0.1, B_x = 1, transverse amplitude 0.1, unit
    wavelength along the (pitch, yaw)-rotated x axis; vx = 0 travelling, 1 standing.  The field is the curl of a
    potential sampled on the cell edges, as in the reference."""
    X, Y, Z, d = _mhd_mesh(n, g, lengths, lo)
    X, Y, Z = X + 0.0 * (Y + Z), Y + 0.0 * (X + Z), Z + 0.0 * (X + Y)
    sy, cy, sp, cp = np.sin(yaw), np.cos(yaw), np.sin(pitch), np.cos(pitch)
    amp, k, bx0, rho, P = 0.1, 2.0 * np.pi, 1.0, 1.0, 0.1

    def potential(x, y, z):
        x_rot = x * cp * cy + y * cp * sy + z * sp
        y_rot = -x * sy + y * cy
        return polarization * (amp / k) * np.sin(k * x_rot), (amp / k) * np.cos(k * x_rot) + bx0 * y_rot

    ay, az = potential(X, Y + d[1] / 2.0, Z + d[2] / 2.0)
    Ax = -ay * sy - az * sp * cy
    ay, az = potential(X + d[0] / 2.0, Y, Z + d[2] / 2.0)
    Ay = ay * cy - az * sp * sy
    ay, az = potential(X + d[0] / 2.0, Y + d[1] / 2.0, Z)
    Az = az * cp
    B = _curl_faces(Ax, Ay, Az, d)
    x_rot = X * cp * cy + Y * cp * sy + Z * sp
    mx = rho * vx + 0.0 * X
    my = -polarization * rho * amp * np.sin(k * x_rot)
    mz = -rho * amp * np.cos(k * x_rot)
    return _mhd_pack_momentum(np.full_like(X, rho), mx * cp * cy - my * sy - mz * sp * cy,
                              mx * cp * sy + my * cy - mz * sp * sy, mx * sp + mz * cp, P, B, gamma)


def advecting_field_loop(n, g, lengths, lo, rho=1.0, v=(1.0, 1.0, 2.0), P=1.0, A=1e-3, radius=0.3, gamma=5.0 / 3.0):
    """Grid3D::Advecting_Field_Loop (initial_conditions.cpp:1799-1863, examples/3D/advecting_field_loop.txt): a weak field
    loop, A_y = A_z = A (R - r) inside r < R, advected diagonally through a periodic box centred on the origin."""
    X, Y, Z, d = _mhd_mesh(n, g, lengths, lo)
    X, Y, Z = X + 0.0 * (Y + Z), Y + 0.0 * (X + Z), Z + 0.0 * (X + Y)
    r = np.sqrt((X + d[0] / 2.0) ** 2 + Y ** 2 + (Z + d[2] / 2.0) ** 2)
    Ay = np.where(r < radius, A * (radius - r), 0.0)
    r = np.sqrt((X + d[0] / 2.0) ** 2 + (Y + d[1] / 2.0) ** 2 + Z ** 2)
    Az = np.where(r < radius, A * (radius - r), 0.0)
    B = _curl_faces(np.zeros_like(Ay), Ay, Az, d)
    one = np.ones_like(X)
    return _mhd_pack_momentum(rho * one, rho * v[0] * one, rho * v[1] * one, rho * v[2] * one, P, B, gamma)
