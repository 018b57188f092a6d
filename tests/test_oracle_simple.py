"""The SIMPLE integrator's oracle restatement (oracle/vl_oracle.c: orc_simple_step, plmc_cell_t; reference
src/integrators/simple_{1,2,3}D_cuda.cu and the #ifndef VL block of src/reconstruction/plmc_cuda.cu:130-285) without
a GPU.  The reference ships no fiducial for this build, so the bit-level pin is tests/test_gpu_reference.py, which runs
the reference's own -DSIMPLE kernels next to this code on the GPU; here are the properties that need no reference."""
import numpy as np
import pytest

import oracle
from cholla_b200 import initial_conditions as ic
from helpers import ghosted, real_view


def _prims(U, N, id_, gamma):
    d = U[id_]
    vx, vy, vz = U[N + id_] / d, U[2 * N + id_] / d, U[3 * N + id_] / d
    p = (U[4 * N + id_] - 0.5 * d * (vx * vx + ((vy * vy) + (vz * vz)))) * (gamma - 1.0)
    return np.array([d, vx, vy, vz, p])


def test_tracing_with_zero_time_step_is_the_van_leer_interface():
    """dt = 0 switches the characteristic tracing off: the SIMPLE build's PLMC cell is then the VL build's."""
    n, g, gamma = (24, 1, 1), 3, 1.4
    flat, dims = ghosted(ic.random_smooth(n, seed=3), g, [1, 1, 0, 0, 0, 0])
    N = dims[0]
    QL, QR = oracle.reconstruct(oracle.PLMC, flat, *dims, gamma, 0)
    for i in range(3, 20):
        L, R = oracle.plmc_trace_cell(_prims(flat, N, i - 1, gamma), _prims(flat, N, i, gamma), _prims(flat, N, i + 1, gamma),
                                      gamma, 0.0)
        assert L[0] == QL[i] and R[0] == QR[i - 1]  # density of the two interface states
        assert np.isclose(L[0] * L[1], QL[N + i], rtol=1e-15) and np.isclose(R[0] * R[1], QR[N + i - 1], rtol=1e-15)


def test_tracing_moves_the_interface_values_upwind():
    """A right-moving supersonic flow: the state on the low side of face i+1/2 is evolved (all three waves leave
    through it), the state on the high side of face i-1/2 is not."""
    wm, wi, wp = [1.0, 3.0, 0.1, 0.0, 1.0], [1.1, 3.1, 0.2, 0.1, 1.1], [1.2, 3.2, 0.3, 0.2, 1.2]
    L0, R0 = oracle.plmc_trace_cell(wm, wi, wp, 1.4, 0.0)
    L1, R1 = oracle.plmc_trace_cell(wm, wi, wp, 1.4, 0.2)
    assert np.array_equal(R0, R1) and not np.array_equal(L0, L1)
    assert L1[0] < L0[0]  # density rises with x and the flow comes from the left: the traced value is lower


@pytest.mark.parametrize("n,flags", [((20, 12, 10), [1] * 6), ((32, 14, 1), [1, 1, 1, 1, 0, 0]), ((64, 1, 1), [1, 1, 0, 0, 0, 0])])
@pytest.mark.parametrize("recon", [oracle.PCM, oracle.PLMC])
def test_simple_step_conserves_and_keeps_a_uniform_state(n, flags, recon):
    g, gamma = 3, 5.0 / 3.0
    d = [1.0 / v for v in n]
    flat, dims = ghosted(ic.random_smooth(n, seed=5), g, flags)
    before = real_view(flat, dims, g).sum(axis=(1, 2, 3))
    oracle.advance(flat, *dims, g, *d, gamma, recon, oracle.HLLC, flags, 3, integrator="simple")
    after = real_view(flat, dims, g).sum(axis=(1, 2, 3))
    assert np.allclose(before, after, rtol=1e-13, atol=1e-12)  # periodic box: mass, momentum, energy
    uni, dims = ghosted(ic.constant(n, 1.3, (0.2, -0.1, 0.3), 0.7, gamma), g, flags)
    ref = uni.copy()
    oracle.advance(uni, *dims, g, *d, gamma, recon, oracle.ROE, flags, 2, integrator="simple")
    assert np.allclose(uni, ref, rtol=1e-14, atol=0)


def test_simple_and_van_leer_agree_to_truncation_error_on_a_sound_wave():
    n, g, gamma, flags = (64, 1, 1), 3, 5.0 / 3.0, [1, 1, 0, 0, 0, 0]
    a, dims = ghosted(ic.sound_wave(n, (1.0, 1.0, 1.0)), g, flags)
    b = a.copy()
    oracle.advance(a, *dims, g, 1.0 / 64, 1.0, 1.0, gamma, oracle.PLMC, oracle.HLLC, flags, 20, integrator="simple")
    oracle.advance(b, *dims, g, 1.0 / 64, 1.0, 1.0, gamma, oracle.PLMC, oracle.HLLC, flags, 20, integrator="vl")
    amp = np.abs(real_view(b, dims, g)[0] - 1.0).max()
    assert 0 < np.abs(real_view(a, dims, g)[0] - real_view(b, dims, g)[0]).max() < 0.05 * amp


def test_simple_step_is_restated_for_pcm_and_plmc_only():
    flat, dims = ghosted(ic.random_smooth((16, 1, 1), seed=1), 4, [1, 1, 0, 0, 0, 0])
    with pytest.raises(NotImplementedError):
        oracle.simple_step(flat, *dims, 4, 1.0 / 16, 1.0, 1.0, 1e-3, 1.4, oracle.PPMC, oracle.HLLC)


@pytest.mark.parametrize("solver", [oracle.HLLC, oracle.ROE])
def test_pcm_two_d_van_leer_step_is_the_first_order_step(solver):
    """VL_Algorithm_2D_CUDA has no -DPCM branch in its second reconstruction (VL_2D_cuda.cu:95-120; the 1-D and 3-D files
    have one), so a -DPCM 2-D build solves the corrector's Riemann problems on the interface arrays of step 1 -- copies of
    U^n -- and its step is the SIMPLE integrator's first-order step, bit for bit.  (The GPU pin against the reference's own
    -DPCM build is tests/test_gpu_reference.py::test_pcm_corrector_under_van_leer_reference_vs_oracle_vs_b200.)  In 3-D the
    corrector really uses the half-step state and the two integrators differ."""
    gamma, dt = 5.0 / 3.0, 2.0e-3
    for n, flags, same in (((33, 21, 1), [1, 1, 1, 1, 0, 0], True), ((17, 13, 11), [1, 1, 1, 1, 1, 1], False)):
        g = 2
        flat, dims = ghosted(ic.random_smooth(n, seed=11), g, flags)
        dx, dy, dz = 1.0 / n[0], 1.0 / n[1], 1.0 / n[2]
        a, b = flat.copy(), flat.copy()
        oracle.vl_step(a, *dims, g, dx, dy, dz, dt, gamma, oracle.PCM, solver)
        oracle.simple_step(b, *dims, g, dx, dy, dz, dt, gamma, oracle.PCM, solver)
        ra, rb = real_view(a, dims, g), real_view(b, dims, g)
        assert np.array_equal(ra, rb) == same
        assert not np.array_equal(ra, real_view(flat, dims, g))
