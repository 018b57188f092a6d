"""Why the strict (bit-reproducing) arithmetic mode exists.

The reference's PPMC limiter is discontinuous in its inputs: on a problem with a jump (the blast wave of
BASELINE configs 3/5) a +-1 ulp perturbation of the state grows to O(1e-3) absolute within ONE step.  The
north-star parity bar (relative L1 <= 1e-12 against the reference's snapshot) is therefore only reachable on
such problems by reproducing the reference's floating-point results exactly -- which is what `strict=1` does
and what tests/test_gpu_reference.py checks against the reference's own compiled kernels.  This test pins the
sensitivity itself, using nothing but the CPU oracle."""
import numpy as np

import oracle
from cholla_b200 import initial_conditions as ic
from helpers import ghosted, real_view


def test_one_ulp_noise_is_amplified_on_discontinuous_data():
    n, flags, gamma, g = (24, 20, 16), [3, 3, 2, 2, 1, 1], 1.66666667, 4
    flat, dims = ghosted(ic.spherical_overpressure(n), g, flags)
    dx, dy, dz = [1.0 / v for v in n]
    a = flat.copy()
    oracle.advance(a, *dims, g, dx, dy, dz, gamma, oracle.PPMC, oracle.HLLC, flags, 1)
    b = a * (1.0 + np.random.default_rng(0).integers(-1, 2, size=a.shape) * 1.1e-16)
    oracle.advance(a, *dims, g, dx, dy, dz, gamma, oracle.PPMC, oracle.HLLC, flags, 1)
    oracle.advance(b, *dims, g, dx, dy, dz, gamma, oracle.PPMC, oracle.HLLC, flags, 1)
    d = np.abs(real_view(a, dims, g) - real_view(b, dims, g))
    assert d.max() > 1e-6  # thirteen orders of magnitude above the perturbation


def test_smooth_data_is_well_conditioned():
    n, flags, gamma, g = (24, 20, 16), [1] * 6, 5.0 / 3.0, 4
    flat, dims = ghosted(ic.sound_wave(n, A=1e-2), g, flags)
    dx, dy, dz = [4.0 / v for v in n]
    a = flat.copy()
    b = a * (1.0 + np.random.default_rng(0).integers(-1, 2, size=a.shape) * 1.1e-16)
    oracle.advance(a, *dims, g, dx, dy, dz, gamma, oracle.PPMC, oracle.HLLC, flags, 2)
    oracle.advance(b, *dims, g, dx, dy, dz, gamma, oracle.PPMC, oracle.HLLC, flags, 2)
    d = np.abs(real_view(a, dims, g) - real_view(b, dims, g))
    assert d.max() < 1e-13
