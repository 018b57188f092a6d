"""Host-side initial-condition generators added for the reference's shipped decks (formulas from
src/grid/initial_conditions.cpp).  No GPU: the generators only build the arrays the path starts from."""
import glob
import os

import numpy as np
import pytest

from cholla_b200 import initial_conditions as ic
from cholla_b200 import parameters as pm
from cholla_b200 import run

G = 1.4


def prim(f, gamma):
    d = f[0]
    v = f[1:4] / d
    return d, v, (f[4] - 0.5 * d * (v ** 2).sum(axis=0)) * (gamma - 1.0)


def test_square_wave_shu_osher_blast():
    f = ic.square_wave((64, 1, 1), rho=1.0, v=(1.0, 0.0, 0.0), P=0.01, A=2.0, gamma=G)
    d, v, p = prim(f, G)
    assert np.allclose(p, 0.01) and np.all(v[0] == 1.0)
    assert np.all(d[0, 0, 16:48] == 2.0) and np.all(d[0, 0, :16] == 1.0) and np.all(d[0, 0, 48:] == 1.0)
    f = ic.shu_osher((200, 1, 1), gamma=G)
    d, v, p = prim(f, G)
    x = -1.0 + (np.arange(200) + 0.5) * 0.01
    assert np.all(d[0, 0, x < -0.8] == 3.857143) and np.allclose(v[0][0, 0, x < -0.8], 2.629369)
    assert np.allclose(p[0, 0, x < -0.8], 10.33333) and np.allclose(d[0, 0, x >= -0.8], 1.0 + 0.2 * np.sin(5 * np.pi * x[x >= -0.8]))
    f = ic.blast_1d((100, 1, 1), gamma=G)
    d, v, p = prim(f, G)
    assert np.allclose(p[0, 0, :10], 1000.0) and np.allclose(p[0, 0, 90:], 100.0) and np.allclose(p[0, 0, 10:90], 0.01)


def test_kh_gresho_implosion():
    f = ic.kh((32, 32, 1))
    assert np.all(f[0][0, :8] == 1.0) and np.all(f[0][0, 8:24] == 2.0) and np.all(f[0][0, 24:] == 1.0)
    assert np.allclose(f[1][0, 12] / f[0][0, 12], 0.5) and np.allclose(f[1][0, 2] / f[0][0, 2], -0.5)
    f = ic.gresho((64, 64, 1), gamma=G)
    d, v, p = prim(f, G)
    x = -0.5 + (np.arange(64) + 0.5) / 64
    X, Y = np.meshgrid(x, x)
    r = np.sqrt(X * X + Y * Y)
    assert np.allclose((v[0][0] * X + v[1][0] * Y), 0.0, atol=1e-15)  # purely azimuthal
    speed = np.sqrt(v[0][0] ** 2 + v[1][0] ** 2)
    assert np.allclose(speed[r < 0.2], 5.0 * r[r < 0.2]) and np.allclose(speed[r >= 0.4], 0.0)
    assert np.allclose(p[0][r >= 0.4], 3.0 + 4.0 * np.log(2.0))
    # the pressure profile balances the centrifugal force: dP/dr = v^2 / r, checked across the two kinks
    for r0 in (0.2, 0.4):
        lo_, hi_ = (9.0 - 4.0 * np.log(0.2) + 12.5 * r0 * r0 - 20.0 * r0 + 4.0 * np.log(r0)), None
        assert np.isclose(lo_, 5.0 + 12.5 * r0 * r0 if r0 == 0.2 else 3.0 + 4.0 * np.log(2.0))
    f = ic.implosion_2d((30, 30, 1), gamma=G)
    d, v, p = prim(f, G)
    assert d[0, 0, 0] == 0.125 and np.isclose(p[0, 0, 0], 0.14) and d[0, 29, 29] == 1.0 and np.isclose(p[0, 29, 29], 1.0)
    assert (d == 0.125).sum() == sum(1 for j in range(30) for i in range(30) if (j + 0.5) * 0.01 < 0.1500001 - (i + 0.5) * 0.01)


def test_hydro_linear_wave_is_the_sound_wave_eigenvector():
    g = 5.0 / 3.0
    a = np.sqrt(g * 0.6 / 1.0)
    f = ic.linear_wave((64, 4, 4), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), 1.0, 0.0, 0.6, 1e-6, (1.0, a, 0.0, 0.0, a * a / (g - 1.0)), gamma=g)
    x = (np.arange(64) + 0.5) / 64
    assert np.allclose(f[0][0, 0], 1.0 + 1e-6 * np.sin(2 * np.pi * x)) and np.allclose(f[1][2, 1], 1e-6 * a * np.sin(2 * np.pi * x))
    assert np.all(f[2] == 0) and np.all(f[3] == 0)
    # rotated by 90 degrees about z: the wave runs along y and the momentum perturbation is in y
    r = ic.linear_wave((4, 64, 4), (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), 1.0, 0.0, 0.6, 1e-6, (1.0, a, 0.0, 0.0, 0.0), yaw=np.pi / 2, gamma=g)
    assert np.allclose(r[2][0, :, 0], f[1][0, 0], atol=1e-20) and np.allclose(r[1], 0.0, atol=1e-20)


def test_mhd_riemann_and_linear_wave():
    from oracle import mhd
    g, n = 5, (16, 6, 6)
    brio_l, brio_r = (1.0, 0, 0, 0, 1.0, 0.75, 1.0, 0.0), (0.125, 0, 0, 0, 0.1, 0.75, -1.0, 0.0)
    f = ic.mhd_riemann(n, g, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), brio_l, brio_r, 0.5, 2.0)
    dims = (26, 16, 16)
    assert f.shape == (8, 16, 16, 26)
    assert np.all(f[5] == 0.75) and np.all(f[6][:, :, :g + 8] == 1.0) and np.all(f[6][:, :, g + 8:] == -1.0)
    assert np.allclose(f[4][:, :, :g + 8], 1.0 / 1.0 + 0.5 * (0.75 ** 2 + 1.0))  # gamma = 2: P / (gamma - 1) + B^2 / 2
    d = [1.0 / v for v in n]
    assert mhd.max_divergence(np.ascontiguousarray(f.reshape(-1)), *dims, g, *d) == 0.0
    ev = (0.4472135954999579, 0.8944271909999159, -0.4472135954999579, 0.0, 2.0124611797498106, 0.0, 0.8944271909999159, 0.0)
    w = ic.mhd_linear_wave(n, g, (1.0, 1.0, 1.0), (0.0, 0.0, 0.0), 1.0, 0.0, 0.6, (1.0, 1.5, 0.0), 1e-6, ev)
    assert np.all(w[5] == 1.0) and np.allclose(w[6].mean(), 1.5, atol=1e-9) and np.ptp(w[6]) > 1e-6
    assert mhd.max_divergence(np.ascontiguousarray(w.reshape(-1)), *dims, g, *d) <= 1e-12
    # the MHD oracle runs the shock tube (transmissive in x, periodic across) without producing a NaN
    flat = np.ascontiguousarray(f.reshape(-1))
    flags = [3, 3, 1, 1, 1, 1]
    mhd.set_boundaries(flat, *dims, g, flags)
    mhd.advance(flat, *dims, g, *d, 2.0, 2, flags, 3)
    assert np.isfinite(flat).all()


DECK = """nx=%d
ny=%d
nz=%d
tout=0.1
outstep=0.1
gamma=1.4
init=%s
xmin=0.0
ymin=0.0
zmin=0.0
xlen=1.0
ylen=1.0
zlen=1.0
xl_bcnd=1
xu_bcnd=1
yl_bcnd=1
yu_bcnd=1
zl_bcnd=1
zu_bcnd=1
outdir=./
rho=1.0
vx=1.0
P=0.01
A=2.0
"""


@pytest.mark.parametrize("name,n", [("Square_Wave", (32, 1, 1)), ("Shu_Osher", (32, 1, 1)), ("Blast_1D", (32, 1, 1)),
                                    ("KH", (16, 16, 1)), ("Gresho", (16, 16, 1)), ("Implosion_2D", (16, 16, 1)),
                                    ("Linear_Wave", (16, 8, 8))])
def test_deck_dispatch(name, n):
    P = pm.parse_text(DECK % (n + (name,)))
    f = run.initial_fields(P)
    assert f.shape == (5, n[2], n[1], n[0]) and np.isfinite(f).all() and (f[0] > 0).all()


def test_shipped_decks_build_their_initial_state():
    """Every hydro / MHD deck under the reference's examples/ whose generator is on the path parses and produces a
    finite state (reduced to a few cells per axis).  Skipped where the reference tree is absent (the GPU box)."""
    root = "/root/reference/examples"
    if not os.path.isdir(root):
        pytest.skip("reference tree not present")
    hydro = {"Constant", "Sound_Wave", "Riemann", "Square_Wave", "Shu_Osher", "Blast_1D", "KH", "KH_res_ind", "Gresho",
             "Implosion_2D", "Spherical_Overpressure_3D", "Spherical_Overdensity_3D"}
    mhd_only = {"Orszag_Tang_Vortex", "MHD_Spherical_Blast", "Linear_Wave", "Circularly_Polarized_Alfven_Wave",
                "Advecting_Field_Loop"}
    built = 0
    for deck in sorted(glob.glob(root + "/[123]D/*.txt")):
        P = pm.parse_params(deck, [])
        small = ["nx=%d" % min(P.nx, 16)] + (["ny=%d" % min(P.ny, 8)] if P.ny > 1 else []) + (["nz=%d" % min(P.nz, 8)] if P.nz > 1 else [])
        P = pm.parse_params(deck, small)
        is_mhd = P.init in mhd_only or (P.init == "Riemann" and (P.Bx_l or P.By_l or P.Bz_l or P.By_r))
        if P.init not in hydro | mhd_only:
            continue
        f = run.initial_fields(P, bool(is_mhd), 5 if is_mhd else 0)
        assert np.isfinite(f).all() and (f[0] > 0).all(), deck
        built += 1
    assert built >= 27


def test_alfven_wave_and_field_loop():
    from oracle import mhd
    g, n = 5, (16, 8, 8)
    lengths, lo = (1.0, 0.5, 0.5), (0.0, 0.0, 0.0)
    d = [lengths[i] / n[i] for i in range(3)]
    dims = tuple(v + 2 * g for v in n)
    w = ic.circularly_polarized_alfven_wave(n, g, lengths, lo, polarization=1.0, vx=0.0)
    flat = np.ascontiguousarray(w.reshape(-1))
    assert mhd.max_divergence(flat, *dims, g, *d) <= 1e-12
    real = w[:, g:-g, g:-g, g:-g]
    assert np.allclose(real[5], 1.0, atol=1e-12)  # B_x = 1 everywhere
    x = (np.arange(16) + 0.5) / 16
    # B_y = -dA_z/dx at the y face: 0.1 sin(2 pi x) up to the finite difference; momentum is anti-parallel to the field
    assert np.allclose(real[6][3, 3], 0.1 * np.sin(2 * np.pi * x), atol=0.1 * (2 * np.pi / 16) ** 2)
    assert np.allclose(real[2][3, 3], -0.1 * np.sin(2 * np.pi * x)) and np.allclose(real[3][3, 3], -0.1 * np.cos(2 * np.pi * x))
    assert np.allclose(real[6] ** 2 + 0 * real[7], real[6] ** 2)
    # periodic: the generated ghost zones equal the periodic images (index 0 of every axis stays zero, like the
    # reference's Init_Magnetic_Field_With_Vector_Potential)
    assert np.allclose(w[5:, 1:, 1:, 1:g], w[5:, 1:, 1:, n[0] + 1:n[0] + g], atol=1e-13)
    loop = ic.advecting_field_loop((16, 16, 16), g, (1.0, 1.0, 1.0), (-0.5, -0.5, -0.5), A=1e-3, radius=0.3)
    flat = np.ascontiguousarray(loop.reshape(-1))
    assert mhd.max_divergence(flat, 26, 26, 26, g, 1 / 16, 1 / 16, 1 / 16) <= 1e-12
    b2 = (loop[5:] ** 2).sum(axis=0)
    assert b2.max() > 0 and b2[g, g, g] == 0.0 and np.isclose(np.sqrt(b2.max()), np.sqrt(2) * 1e-3, rtol=0.5)
    assert np.all(loop[1] == 1.0) and np.all(loop[3] == 2.0)


@pytest.mark.parametrize("gen", ["orszag_tang", "mhd_riemann", "cpaw", "loop"])
def test_mhd_generators_are_rank_count_independent(gen):
    """Inside ic.block(offset, n_local) an MHD generator builds one rank's array (ghost zones included) from the global
    index rule of Get_Position: the cells and faces a rank owns equal the single-domain ones, so a run does not depend
    on the decomposition."""
    from cholla_b200 import _lib
    g, n = 5, (16, 12, 8)
    lengths, lo = (1.0, 1.0, 1.0), (-0.5, -0.5, -0.5)
    make = {"orszag_tang": lambda: ic.orszag_tang(n, g),
            "mhd_riemann": lambda: ic.mhd_riemann(n, g, lengths, lo, (1.0, 0, 0, 0, 1.0, 0.75, 1.0, 0.0),
                                                  (0.125, 0, 0, 0, 0.1, 0.75, -1.0, 0.0), 0.0, 2.0),
            "cpaw": lambda: ic.circularly_polarized_alfven_wave(n, g, lengths, lo),
            "loop": lambda: ic.advecting_field_loop(n, g, lengths, lo, radius=0.3)}[gen]
    whole = make()
    np_grid = _lib.tile_decomposition(4, *n)
    for rank in range(4):
        coords = (rank % np_grid[0], (rank // np_grid[0]) % np_grid[1], rank // (np_grid[0] * np_grid[1]))
        ext = [_lib.local_extent(n[a], np_grid[a], coords[a]) for a in range(3)]
        nl, off = [e[0] for e in ext], [e[1] for e in ext]
        with ic.block(off, nl):
            part = make()
        assert part.shape == (8, nl[2] + 2 * g, nl[1] + 2 * g, nl[0] + 2 * g)
        # what the rank owns -- its real cells and the faces around them -- and all but the two outermost ghost layers
        # (the curl helper leaves index 0 at zero and the energy of index 1 averages with it; the exchange refills both)
        want = whole[:, off[2] + 2:off[2] + nl[2] + 2 * g, off[1] + 2:off[1] + nl[1] + 2 * g, off[0] + 2:off[0] + nl[0] + 2 * g]
        assert np.array_equal(part[:, 2:, 2:, 2:], want), rank


MHD_BLAST_DECK = """
nx=20
ny=30
nz=20
tout=0.2
outstep=0.2
init=MHD_Spherical_Blast
xmin=-0.5
ymin=-0.75
zmin=-0.5
xlen=1.0
ylen=1.5
zlen=1.0
xl_bcnd=1
xu_bcnd=1
yl_bcnd=1
yu_bcnd=1
zl_bcnd=1
zu_bcnd=1
outdir=./
rho=2.0
vx=0.0
vy=0.25
vz=0.0
P=0.1
P_blast=10.0
radius=0.2
Bx=0.70710678118654746
By=0.70710678118654746
Bz=0.0
gamma=1.666666666666667
"""


def test_mhd_blast_takes_the_decks_domain_and_parameters():
    """examples/3D/mhd_blast.txt is not cubic (ny = 300, ylen = 1.5): the blast must be a sphere of the deck's radius in
    the deck's coordinates, with the deck's rho, v, P, P_blast and B (initial_conditions.cpp:1865-1924)."""
    P = pm.parse_text(MHD_BLAST_DECK)
    g = 5
    f = run.initial_fields(P, True, g)
    assert f.shape == (8, 20 + 2 * g, 30 + 2 * g, 20 + 2 * g)
    real = f[:, g:-g, g:-g, g:-g]
    assert np.all(real[0] == 2.0) and np.all(real[2] == 2.0 * 0.25) and np.all(real[5] == P.Bx)
    d = 0.05  # the same cell size on every axis: dy = 1.5 / 30
    x = -0.5 + (np.arange(20) + 0.5) * d
    y = -0.75 + (np.arange(30) + 0.5) * d
    Z, Y, X = np.meshgrid(x, y, x, indexing="ij")
    inside = np.sqrt(X * X + Y * Y + Z * Z) < 0.2
    e_mag, e_kin = 0.5 * (P.Bx ** 2 + P.By ** 2), 0.5 * 2.0 * 0.25 ** 2
    pres = (real[4] - e_mag - e_kin) * (P.gamma - 1.0)
    assert np.allclose(pres[inside], 10.0) and np.allclose(pres[~inside], 0.1)
    assert inside.sum() > 0 and inside.sum() == (np.abs(pres - 10.0) < 1e-9).sum()
    # the blast is round in index space too (equal cell sizes): as many cells across in y as in x
    ii = np.argwhere(inside)
    assert np.ptp(ii[:, 1]) == np.ptp(ii[:, 2]) == np.ptp(ii[:, 0])
    # the reference's asserts: a domain that is not centred on the origin is refused
    with pytest.raises(ValueError):
        run.initial_fields(pm.parse_text(MHD_BLAST_DECK.replace("ymin=-0.75", "ymin=-0.5")), True, g)


def test_mhd_linear_wave_rotations_and_mhd_constant():
    """The reference's MHD linear-wave system tests run every wave along x, y and z (yaw or pitch = pi / 2,
    mhd_system_tests.cpp:129-505): the rotated generator must be the x-aligned one with its axes permuted, and
    `Constant` must build a uniform 8-field state with the deck's field."""
    n, g = (8, 8, 8), 5
    L, lo = (1.0, 1.0, 1.0), (0.0, 0.0, 0.0)
    ev = (0.3, -0.2, 0.1, 0.05, 0.4, 0.0, 0.25, -0.15)
    B0 = (1.0, 1.5, 0.5)
    base = ic.mhd_linear_wave(n, g, L, lo, 1.0, 0.2, 0.6, B0, 1e-3, ev, 1.0, 5.0 / 3.0)
    ywave = ic.mhd_linear_wave(n, g, L, lo, 1.0, 0.2, 0.6, B0, 1e-3, ev, 1.0, 5.0 / 3.0, yaw=0.5 * np.pi)
    # yaw = pi/2: x' = y, and (vx', vy') -> (-vy', vx'): fields as functions of y equal the base as functions of x
    a = base[:, 0, 0, :]     # vary along x
    b = ywave[:, 0, :, 0]    # vary along y
    assert np.allclose(b[0], a[0], atol=1e-15) and np.allclose(b[4], a[4], atol=1e-15)
    assert np.allclose(b[2], a[1], atol=1e-15) and np.allclose(b[1], -a[2], atol=1e-15) and np.allclose(b[3], a[3], atol=1e-15)
    assert np.allclose(b[6], a[5], atol=1e-15) and np.allclose(b[5], -a[6], atol=1e-15) and np.allclose(b[7], a[7], atol=1e-15)
    zwave = ic.mhd_linear_wave(n, g, L, lo, 1.0, 0.2, 0.6, B0, 1e-3, ev, 1.0, 5.0 / 3.0, pitch=0.5 * np.pi)
    c = zwave[:, :, 0, 0]    # vary along z
    assert np.allclose(c[0], a[0], atol=1e-15) and np.allclose(c[3], a[1], atol=1e-15) and np.allclose(c[1], -a[3], atol=1e-15)
    P = pm.parse_text(MHD_BLAST_DECK.replace("init=MHD_Spherical_Blast", "init=Constant"))
    f = run.initial_fields(P, True, g)
    assert f.shape[0] == 8 and np.all(f[5] == P.Bx) and np.all(f[0] == 2.0) and np.all(f[2] == 0.5)
    with pytest.raises(NotImplementedError):
        run.initial_fields(pm.parse_text(MHD_BLAST_DECK.replace("init=MHD_Spherical_Blast", "init=KH")), True, g)


def test_spherical_overdensity():
    """Grid3D::Spherical_Overdensity_3D (initial_conditions.cpp:1253-1309; examples/3D/Spherical_Collapse.txt): a unit-density
    sphere of radius 0.2 about (0.5, 0.5, 0.5) in a rho = 0.0005 background, uniform pressure 0.0005, at rest."""
    n, gamma = (20, 20, 20), 5.0 / 3.0
    f = ic.spherical_overdensity(n, gamma=gamma)
    assert f.shape == (5, 20, 20, 20) and not f[1:4].any()
    assert np.array_equal(f[4], np.full((20, 20, 20), 0.0005 / (gamma - 1.0)))
    x = (np.arange(20) + 0.5) / 20
    Z, Y, X = np.meshgrid(x, x, x, indexing="ij")
    inside = np.sqrt((X - 0.5) ** 2 + (Y - 0.5) ** 2 + (Z - 0.5) ** 2) < 0.2
    assert np.array_equal(f[0] == 1.0, inside) and set(np.unique(f[0])) == {0.0005, 1.0}
    assert abs(inside.mean() - 4.0 / 3.0 * np.pi * 0.2 ** 3) < 3e-3
    P = pm.parse_text(DECK % ((16, 16, 16) + ("Spherical_Overdensity_3D",)))
    assert run.initial_fields(P).shape == (5, 16, 16, 16)
