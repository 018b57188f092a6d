"""The exact Riemann solver's oracle restatement (oracle/vl_oracle.c: exact_flux; reference
src/riemann_solvers/exact_cuda.cu, -DEXACT) against the published star-region values of Toro (1999), table 4.2 / 4.3.
The reference ships no test for this solver and evaluates its powers with the device's powf, which no CPU libm
reproduces bit for bit: on the CPU this restatement is pinned to float accuracy only ("parity unpinned", see its
header); the bit-level check of the CUDA path is the reference's own -DEXACT build in tests/test_gpu_reference.py."""
import numpy as np
import pytest

import oracle

G = 1.4


def cons(d, u, p, v=0.0, w=0.0):
    return [d, d * u, d * v, d * w, p / (G - 1.0) + 0.5 * d * (u * u + v * v + w * w)]


def test_sod_flux_is_the_star_left_state():
    # Toro test 1: p* = 0.30313, u* = 0.92745, rho*L = 0.42632; x/t = 0 lies between the fan's tail and the contact
    f = oracle.riemann_single(oracle.EXACT, cons(1.0, 0.0, 1.0), cons(0.125, 0.0, 0.1), G)
    d, u, p = 0.42632, 0.92745, 0.30313
    E = p / (G - 1.0) + 0.5 * d * u * u
    assert np.allclose(f, [d * u, d * u * u + p, 0.0, 0.0, (E + p) * u], rtol=2e-5, atol=1e-12)


def test_two_rarefactions_and_strong_shock():
    # Toro test 2 (123 problem): symmetric, u* = 0, p* = 0.00189 -> no mass flux, momentum flux = p*
    f = oracle.riemann_single(oracle.EXACT, cons(1.0, -2.0, 0.4), cons(1.0, 2.0, 0.4), G)
    assert abs(f[0]) < 1e-6 and abs(f[1] - 0.00189) < 2e-5 and abs(f[4]) < 1e-6
    # Toro test 3: p* = 460.894, u* = 19.5975, rho*L = 0.57506
    f = oracle.riemann_single(oracle.EXACT, cons(1.0, 0.0, 1000.0), cons(1.0, 0.0, 0.01), G)
    assert np.isclose(f[0], 0.57506 * 19.5975, rtol=1e-4) and np.isclose(f[1], 0.57506 * 19.5975 ** 2 + 460.894, rtol=1e-4)


def test_supersonic_and_mirror_cases():
    # both states moving right faster than sound: the flux is the left state's own flux, transverse momenta upwinded
    L, R = cons(1.0, 5.0, 1.0, 0.3, -0.2), cons(0.8, 5.5, 0.9, -0.1, 0.4)
    f = oracle.riemann_single(oracle.EXACT, L, R, G)
    assert np.allclose(f, [5.0, 26.0, 1.5, -1.0, (L[4] + 1.0) * 5.0], rtol=1e-14)
    # mirror symmetry: reflecting the problem flips the sign of the mass and energy fluxes
    a = oracle.riemann_single(oracle.EXACT, cons(1.0, 0.3, 1.0, 0.1, 0.2), cons(0.5, -0.2, 0.4, -0.3, 0.1), G)
    b = oracle.riemann_single(oracle.EXACT, cons(0.5, 0.2, 0.4, -0.3, 0.1), cons(1.0, -0.3, 1.0, 0.1, 0.2), G)
    assert np.allclose([a[0], a[1], a[4]], [-b[0], b[1], -b[4]], rtol=1e-6)


@pytest.mark.parametrize("solver", [oracle.HLLC, oracle.ROE])
def test_agrees_with_the_approximate_solvers_on_weak_jumps(solver):
    rng = np.random.default_rng(2)
    for _ in range(50):
        d, u, p = 1.0 + 0.2 * rng.random(), rng.random() - 0.5, 1.0 + 0.2 * rng.random()
        L = cons(d, u, p, 0.1, -0.1)
        R = cons(d * (1 + 1e-3 * rng.standard_normal()), u + 1e-3 * rng.standard_normal(), p * (1 + 1e-3 * rng.standard_normal()), 0.1, -0.1)
        a, b = oracle.riemann_single(oracle.EXACT, L, R, G), oracle.riemann_single(solver, L, R, G)
        assert np.allclose(a, b, rtol=0, atol=2e-5)  # O(jump^2) + the solver's float-precision powers


def test_vl_step_with_the_exact_solver_conserves():
    from cholla_b200 import initial_conditions as ic
    from helpers import ghosted, real_view
    n, g, flags = (16, 12, 10), 4, [1] * 6
    flat, dims = ghosted(ic.random_smooth(n, seed=9), g, flags)
    before = real_view(flat, dims, g).sum(axis=(1, 2, 3))
    oracle.advance(flat, *dims, g, 1.0 / 16, 1.0 / 12, 1.0 / 10, 5.0 / 3.0, oracle.PPMC, oracle.EXACT, flags, 2)
    assert np.allclose(before, real_view(flat, dims, g).sum(axis=(1, 2, 3)), rtol=1e-13, atol=1e-12)
