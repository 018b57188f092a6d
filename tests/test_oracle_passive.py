"""CPU checks of the oracle's -DSCALAR / -DDE restatement (orc_vl_step_x): properties that hold whatever the reference
build, so that the restatement is constrained here as well as by the side-by-side GPU runs of tests/test_gpu_passive.py."""
import numpy as np
import pytest

import oracle
from cholla_b200 import initial_conditions as ic
from helpers import ghosted, real_view


def _state(n, ns, de, gamma, recon, icname="random_smooth"):
    g = {0: 2, 1: 3, 2: 4, 3: 3, 4: 4}[recon]
    fields = ic.with_passive(getattr(ic, icname)(n, seed=3), ns, de, gamma)
    flat, dims = ghosted(fields, g, [1, 1, 1, 1, 1, 1] if n[2] > 1 else ([1, 1, 1, 1, 0, 0] if n[1] > 1 else [1, 1, 0, 0, 0, 0]))
    return flat, dims, g


@pytest.mark.parametrize("recon,solver", [(r, s) for r in (0, 1, 2) for s in (0, 1, 3)] + [(3, 0), (4, 0)])
def test_scalars_do_not_feed_back_and_follow_the_mass(recon, solver):
    n, gamma = (14, 12, 10), 1.4
    flat, dims, g = _state(n, 1, False, gamma, recon)
    N = dims[0] * dims[1] * dims[2]
    plain = flat[:5 * N].copy()
    flat[5 * N:] = 0.7 * flat[:N]  # a constant mass fraction
    oracle.vl_step(plain, *dims, g, 0.1, 0.1, 0.1, 0.004, gamma, recon, solver)
    oracle.vl_step_x(flat, *dims, g, 0.1, 0.1, 0.1, 0.004, gamma, recon, solver, 0.0, 1, False)
    assert np.array_equal(flat[:5 * N], plain)                     # the five hydro fields are untouched by the scalar
    r = real_view(flat, dims, g)
    assert np.abs(r[5] - 0.7 * r[0]).max() < 1e-14                # and a constant mass fraction stays constant


@pytest.mark.parametrize("n", [(14, 12, 10), (20, 14, 1), (40, 1, 1)])
def test_dual_energy_warm_gas_keeps_total_energy_branch(n):
    """Thermal energy far above DE_ETA_1 of E: every pressure comes from E - Ekin, so the five hydro fields equal the
    plain build's up to the last Sync_Energies (E = Ekin + (E - Ekin), one rounding), and GasEnergy ends as E - Ekin."""
    gamma = 5.0 / 3.0
    flat, dims, g = _state(n, 0, True, gamma, 2)
    N = dims[0] * dims[1] * dims[2]
    plain = flat[:5 * N].copy()
    d = [1.0 / m for m in n]
    oracle.vl_step(plain, *dims, g, d[0], d[1], d[2], 0.002, gamma, 2, 0)
    oracle.vl_step_x(flat, *dims, g, d[0], d[1], d[2], 0.002, gamma, 2, 0, 0.0, 0, True)
    r, p = real_view(flat, dims, g), real_view(plain, dims, g)
    assert np.array_equal(r[:4], p[:4])
    assert np.abs(r[4] - p[4]).max() <= 4e-16 * np.abs(p[4]).max()
    ekin = 0.5 * (r[1] ** 2 + r[2] ** 2 + r[3] ** 2) / r[0]
    assert np.abs(r[5] - (p[4] - ekin)).max() <= 1e-14 * np.abs(p[4]).max()


def test_dual_energy_cold_flow_uses_the_advected_energy():
    n, gamma = (16, 12, 10), 5.0 / 3.0
    flat, dims, g = _state(n, 0, True, gamma, 2, "cold_flow")
    N = dims[0] * dims[1] * dims[2]
    flat[5 * N:] *= 1.03  # the advected energy has drifted from E - Ekin
    before = real_view(flat.copy(), dims, g)
    oracle.advance(flat, *dims, g, 1 / 16, 1 / 12, 1 / 10, gamma, 2, 0, [1] * 6, 2, n_scalars=0, dual_energy=True)
    r = real_view(flat, dims, g)
    ut = r[4] - 0.5 * (r[1] ** 2 + r[2] ** 2 + r[3] ** 2) / r[0]
    assert np.isfinite(flat).all() and (r[5] > 0).all()
    frac = (before[4] - 0.5 * (before[1] ** 2 + before[2] ** 2 + before[3] ** 2) / before[0]) / before[4]
    assert 0.02 < (frac > 0.001).mean() < 0.98          # the input straddles DE_ETA_1
    assert np.abs(r[5] - ut).max() <= 1e-9 * r[4].max()  # Sync_Energies: E = Ekin + GasEnergy afterwards


def test_floors():
    n, gamma = (10, 9, 8), 1.4
    flat, dims, g = _state(n, 1, True, gamma, 2)
    N = dims[0] * dims[1] * dims[2]
    a = flat.copy()
    oracle.scalar_floor(a, *dims, g, 5, 0.9)
    assert real_view(a, dims, g)[5].min() == 0.9 and np.array_equal(a[:5 * N], flat[:5 * N])
    b = flat.copy()
    oracle.temperature_floor_x(b, *dims, g, 3.0, 7, True)
    r = real_view(b, dims, g)
    assert (r[6] / r[0]).min() >= 3.0 * (1 - 1e-15)


def test_static_gravity_oracle_uniform_field_accelerates_a_uniform_gas():
    """Rayleigh-Taylor field (custom_grav = 2, g = (0, -1)) on a uniform gas at rest, periodic: after a step every real cell
    has momentum_y = -rho dt; the work term averages the velocities BEFORE the momentum source is added (hydro_cuda.cu:
    146-152 reads v^{n+1} first), so the energy only moves from the second step on: 0.25 dt g (2 rho) (v_y + v_y) = 2 dt^2."""
    nx, ny, g, gamma, dt = 16, 12, 3, 1.4, 0.01
    fields = ic.constant((nx, ny, 1), rho=2.0, P=1.0, gamma=gamma)
    flat, dims = ghosted(fields, g, [1, 1, 1, 1, 0, 0])
    field = oracle.static_gravity_field(*dims, g, 2, (0, 0, 0), 0.1, 0.1, 1.0, (0.0, 0.0, 0.0))
    assert set(np.unique(field)) == {-1.0, 0.0}
    oracle.set_gravity(field)
    try:
        oracle.vl_step(flat, *dims, g, 0.1, 0.1, 1.0, dt, gamma, 1, 0)
        r = real_view(flat, dims, g).copy()
        assert np.all(r[0] == 2.0) and np.all(r[1] == 0.0)
        assert np.allclose(r[2], -2.0 * dt, rtol=1e-14) and np.all(r[4] == fields[4].flat[0])
        oracle.set_boundaries(flat, *dims, 5, g, [1, 1, 1, 1, 0, 0])
        oracle.vl_step(flat, *dims, g, 0.1, 0.1, 1.0, dt, gamma, 1, 0)
    finally:
        oracle.set_gravity(None)
    r = real_view(flat, dims, g)
    assert np.allclose(r[2], -4.0 * dt, rtol=1e-13)
    assert np.allclose(r[4] - fields[4].flat[0], 2.0 * dt * dt, rtol=1e-9)


def test_static_gravity_fields_are_odd_and_point_inwards():
    n, g = (12, 12, 12), 3
    dims = tuple(m + 2 * g for m in n)
    f = oracle.static_gravity_field(*dims, g, 1, (0, 0, 0), 1.0, 1.0, 1.0, (-6.0, -6.0, -6.0)).reshape(3, *dims[::-1])
    assert np.isfinite(f).all()
    assert np.allclose(f[0], -f[0][:, :, ::-1], rtol=1e-12) and np.allclose(f[2], -f[2][::-1], rtol=1e-12)
    assert (f[0][:, :, dims[0] // 2 + 1:] < 0).all() and (f[2][dims[2] // 2 + 1:] < 0).all()


def test_cie_cooling_oracle_curve_and_energy_loss():
    """CIE_cool (cooling_cuda.cu:293-317): no cooling below 10^4 K, continuous at the 10^5.9 K and 10^7.4 K joints of the fit
    to a few per cent, n^2 scaling; and Cooling_Update lowers E by exactly n k dT / ((gamma - 1) ENERGY_UNIT)."""
    assert oracle.cie_cool(1.0, 9.0e3) == 0.0
    assert oracle.cie_cool(2.0, 1e6) == pytest.approx(4.0 * oracle.cie_cool(1.0, 1e6), rel=1e-15)
    for Tj in (10 ** 5.9, 10 ** 7.4):
        assert oracle.cie_cool(1.0, Tj * 0.999999) == pytest.approx(oracle.cie_cool(1.0, Tj * 1.000001), rel=0.15)
    n, g, gamma = (6, 5, 4), 3, 5.0 / 3.0
    MP, KB = 1.672622e-24, 1.380658e-16
    DU = 1.98847e33 / 3.08567758e21 ** 3
    PU = DU * (3.08567758e21 / 3.15569e10) ** 2
    d = 1.0 * 0.6 * MP / DU
    P = 1.0 * KB * 3.0e5 / PU
    flat, dims = ghosted(ic.constant(n, rho=d, P=P, gamma=gamma), g, [1] * 6)
    before = flat.copy()
    oracle.cooling_update(flat, *dims, g, 5, False, 0.05, gamma)
    r0, r1 = real_view(before, dims, g), real_view(flat, dims, g)
    assert np.array_equal(r0[:4], r1[:4]) and (r1[4] < r0[4]).all()
    T1 = r1[4] * (gamma - 1.0) * PU / (1.0 * KB)
    assert np.all(T1 < 3.0e5) and np.all(T1 > 1.0e4)
    assert np.array_equal(flat.reshape(5, -1)[:, 0], before.reshape(5, -1)[:, 0])  # ghost cells untouched


def test_gravity_potential_oracle_linear_potential_is_a_uniform_field():
    """phi = a x + b y + c z: both gradient forms give g = -(a, b, c) -- except that the five-point y and z components divide
    by 12 dx like the reference's (hydro_cuda.cu:345-365), so with dy = 2 dx the y acceleration doubles."""
    n, g, gamma, dt = (12, 10, 8), 4, 1.4, 0.003
    dx, dy, dz = 0.1, 0.2, 0.1
    flat, dims = ghosted(ic.constant(n, rho=1.5, P=1.0, gamma=gamma), g, [1] * 6)
    c = [np.arange(dims[i]) * s for i, s in enumerate((dx, dy, dz))]
    Z, Y, X = np.meshgrid(c[2], c[1], c[0], indexing="ij")
    pot = 0.3 * X + 0.5 * Y - 0.2 * Z
    for five, gy in ((False, -0.5), (True, -0.5 * dy / dx)):
        U = flat.copy()
        oracle.set_gravity_potential(pot, five)
        try:
            oracle.vl_step(U, *dims, g, dx, dy, dz, dt, gamma, 2, 0)
        finally:
            oracle.set_gravity_potential(None)
        r = real_view(U, dims, g)
        assert np.allclose(r[1], 1.5 * dt * -0.3, rtol=1e-11) and np.allclose(r[3], 1.5 * dt * 0.2, rtol=1e-11)
        assert np.allclose(r[2], 1.5 * dt * gy, rtol=1e-11)
