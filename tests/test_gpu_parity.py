"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * strict mode (reference-order arithmetic, no FMA): identical bits to the oracle on every real cell;
  * product (fast) mode: relative L1 per field <= 1e-12 in fp64;
  * fp32: relative L1 <= 1e-5 against the fp64 oracle (the reference has no fp32 build: parity unpinned).
"""
import ctypes as C

import numpy as np
import pytest

import cholla_b200 as cb
import oracle
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for
from helpers import ghosted, max_ulp, real_view, rel_l1

pytestmark = pytest.mark.gpu

TOL_F64 = 1e-12
TOL_F32 = 1e-5
PERIODIC = [1, 1, 1, 1, 1, 1]


def make_grid(n, recon, solver, flags, gamma, strict, precision=2, lengths=(1.0, 1.0, 1.0), fused=True, **kw):
    return Grid3D(n, lengths, recon=recon, solver=solver, gamma=gamma, bc=flags, precision=precision,
                  strict=strict, fused=fused, **kw)


def run_both(fields, recon, solver, flags, gamma, n_steps, strict, precision=2, lengths=(1.0, 1.0, 1.0),
             fused=True, U_floor=0.0, integrator="vl"):
    nf, nz, ny, nx = fields.shape
    g = n_ghost_for(recon)
    flat, dims = ghosted(fields, g, flags)
    G = make_grid((nx, ny, nz), recon, solver, flags, gamma, strict, precision, lengths, fused,
                  temperature_floor_U=U_floor, integrator=integrator)
    assert (G.H.nx, G.H.ny, G.H.nz) == dims
    G.upload(flat)
    dts_gpu = G.advance(n_steps)
    got = G.download()
    G.close()
    ref = flat.copy()
    dts_cpu = oracle.advance(ref, *dims, g, G.H.dx, G.H.dy, G.H.dz, gamma, recon, solver, flags, n_steps,
                             U_floor=U_floor, integrator=integrator)
    return got, ref, dims, g, dts_gpu, dts_cpu


CASES_3D = [(cb.PPMC, cb.HLLC), (cb.PLMC, cb.HLLC), (cb.PPMC, cb.ROE), (cb.PLMC, cb.ROE), (cb.PCM, cb.HLLC),
            (cb.PPMC, cb.HLL), (cb.PLMC, cb.HLL), (cb.PLMP, cb.HLLC), (cb.PLMP, cb.ROE), (cb.PPMP, cb.HLLC)]


@pytest.mark.parametrize("recon,solver", CASES_3D)
def test_strict_mode_is_bit_exact_3d(recon, solver):
    # ragged sizes: not multiples of the 32x8 CTA column, several z-chunks
    fields = ic.random_smooth((37, 19, 21), seed=1)
    got, ref, dims, g, dts_gpu, dts_cpu = run_both(fields, recon, solver, PERIODIC, 5.0 / 3.0, 3, strict=True)
    assert dts_gpu == dts_cpu  # CFL reduction is bit-exact too
    a, b = real_view(got, dims, g), real_view(ref, dims, g)
    # every combination reproduces the reference's bits -- a PCM corrector too: strict contexts keep its half-step state
    # conserved, the values PCM_Reconstruction_3D copies (pcm_cuda.cu:256-411)
    assert max_ulp(a, b) == 0
    assert np.array_equal(got, ref)  # ghosts too: same boundary routine semantics


@pytest.mark.parametrize("recon,solver", CASES_3D)
def test_fast_mode_within_tolerance_3d(recon, solver):
    fields = ic.random_smooth((37, 19, 21), seed=2)
    got, ref, dims, g, dts_gpu, dts_cpu = run_both(fields, recon, solver, PERIODIC, 5.0 / 3.0, 3, strict=False)
    err = rel_l1(real_view(got, dims, g), real_view(ref, dims, g))
    assert err.max() <= TOL_F64, err
    assert np.allclose(dts_gpu, dts_cpu, rtol=1e-13, atol=0)


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("recon,solver", [(cb.PPMC, cb.HLLC), (cb.PLMC, cb.ROE)])
def test_2d(recon, solver, strict):
    fields = ic.random_smooth((45, 23, 1), seed=3)
    got, ref, dims, g, dg, dc = run_both(fields, recon, solver, [1, 1, 1, 1, 0, 0], 1.4, 4, strict=strict)
    a, b = real_view(got, dims, g), real_view(ref, dims, g)
    if strict:
        assert max_ulp(a, b) == 0 and dg == dc
    else:
        assert rel_l1(a, b).max() <= TOL_F64


@pytest.mark.parametrize("strict", [True, False])
def test_config1_sod_1d_plmc_hllc(strict):
    """BASELINE config 1: examples/1D/sod.txt (nx=100, gamma=1.4, transmissive), PLMC + HLLC, to t = 0.2."""
    n = (100, 1, 1)
    fields = ic.riemann(n, gamma=1.4)
    flags = [3, 3, 0, 0, 0, 0]
    g = n_ghost_for(cb.PLMC)
    flat, dims = ghosted(fields, g, flags)
    G = make_grid(n, cb.PLMC, cb.HLLC, flags, 1.4, strict)
    G.upload(flat)
    ref = flat.copy()
    t_gpu = t_cpu = 0.0
    steps = 0
    dti_g = G.Calc_Inverse_Timestep()
    dti_c = oracle.calc_dti(ref, *dims, g, G.H.dx, G.H.dy, G.H.dz, 1.4)
    while t_cpu < 0.2 and steps < 1000:
        dt_c = min(0.3 / dti_c, 0.2 - t_cpu)  # main.cpp:258-261 clips the last step to the output time
        G.set_dt(dti_g)
        G.H.dt = min(G.H.dt, 0.2 - t_gpu)
        dti_g = G.Update_Hydro_Grid()
        G.Set_Boundary_Conditions_Grid()
        oracle.vl_step(ref, *dims, g, G.H.dx, 1.0, 1.0, dt_c, 1.4, cb.PLMC, cb.HLLC)
        dti_c = oracle.calc_dti(ref, *dims, g, G.H.dx, G.H.dy, G.H.dz, 1.4)
        oracle.set_boundaries(ref, *dims, 5, g, flags)
        t_gpu += G.H.dt
        t_cpu += dt_c
        steps += 1
    got = G.download()
    G.close()
    assert steps > 20 and abs(t_gpu - t_cpu) < 1e-15
    a, b = real_view(got, dims, g), real_view(ref, dims, g)
    if strict:
        assert np.array_equal(got, ref)
    else:
        err = rel_l1(a[[0, 1, 4]], b[[0, 1, 4]])  # my, mz are identically zero
        assert err.max() <= TOL_F64, err


@pytest.mark.parametrize("strict", [True, False])
def test_sound_wave_64_ppmc_hllc(strict):
    """Scaled-down BASELINE config 2 (examples/3D/sound_wave.txt at 64^3): 10 steps against the oracle."""
    n = (64, 64, 64)
    fields = ic.sound_wave(n)
    got, ref, dims, g, dg, dc = run_both(fields, cb.PPMC, cb.HLLC, PERIODIC, 1.666666666666667, 10, strict=strict,
                                         lengths=(4.0, 4.0, 4.0))
    a, b = real_view(got, dims, g), real_view(ref, dims, g)
    if strict:
        assert np.array_equal(got, ref) and dg == dc
        return
    err = rel_l1(a, b)
    assert err[[0, 4]].max() <= TOL_F64, err
    # the wave's momentum is a 1e-4 perturbation riding on an O(1) pressure: one ulp of P moves it by 1e-12 of
    # its own size, so its error is measured against the natural momentum scale rho0 * c0 instead
    scale = 1.0 * np.sqrt(1.666666666666667 * 0.6 / 1.0)
    assert (np.abs(a[1:4] - b[1:4]).mean(axis=(1, 2, 3)) / scale).max() <= TOL_F64


def test_blast_48_transmissive_reflective_mix_strict():
    n = (48, 40, 36)
    fields = ic.spherical_overpressure(n)
    flags = [3, 3, 2, 2, 1, 1]
    got, ref, dims, g, dg, dc = run_both(fields, cb.PPMC, cb.HLLC, flags, 1.66666667, 6, strict=True)
    assert np.array_equal(got, ref)


def test_fused_epilogue_equals_separate_passes():
    """chl_vl_step_fused (temperature floor + Calc_dt folded into the corrector) == the three separate calls."""
    fields = ic.spherical_overpressure((40, 24, 24))
    # a floor that actually bites: U_floor above the ambient specific internal energy
    kw = dict(recon=cb.PPMC, solver=cb.HLLC, flags=PERIODIC, gamma=1.66666667, n_steps=3, strict=True, U_floor=20.0)
    a, ref, dims, g, da, dc = run_both(fields, fused=True, **kw)
    b, _, _, _, db, _ = run_both(fields, fused=False, **kw)
    assert np.array_equal(a, b) and da == db
    assert np.array_equal(a, ref) and da == dc


def test_fp32_against_fp64_oracle():
    if not cb.lib().chl_has_precision(1):
        pytest.skip("fp32 kernels not built")
    fields = ic.sound_wave((48, 32, 32), A=1e-2)
    got, ref, dims, g, dg, dc = run_both(fields, cb.PPMC, cb.HLLC, PERIODIC, 5.0 / 3.0, 5, strict=False,
                                         precision=1, lengths=(4.0, 4.0, 4.0))
    err = rel_l1(real_view(got, dims, g), real_view(ref, dims, g))
    assert err.max() <= TOL_F32, err


# ---- the small entry points ----
@pytest.mark.parametrize("dims", [(30, 1, 1), (30, 22, 1), (30, 22, 18)])
@pytest.mark.parametrize("strict", [True, False])
def test_calc_dti(dims, strict):
    n = dims
    fields = ic.random_smooth(n, seed=5)
    ndim = 1 + (n[1] > 1) + (n[2] > 1)
    flags = [1, 1] * ndim + [0, 0] * (3 - ndim)
    flat, gd = ghosted(fields, 3, flags)
    G = make_grid(n, cb.PLMC, cb.HLLC, flags, 1.4, strict, lengths=(1.0, 2.0, 3.0))
    G.upload(flat)
    got = G.Calc_Inverse_Timestep()
    want = oracle.calc_dti(flat, *gd, 3, G.H.dx, G.H.dy, G.H.dz, 1.4)
    G.close()
    if strict:
        assert got == want
    else:
        assert abs(got - want) <= 4e-16 * want


def test_calc_dt_reference_fiducial(golden_dir):
    """The reference's own Calc_dt_3D known answer (hydro_cuda_tests.cu:28-80) on the GPU path."""
    import json
    import os
    c = json.load(open(os.path.join(golden_dir, "scalars.json")))["calc_dt_3d"]
    n = (9, 9, 9)
    fields = ic.constant(n, rho=1.0, P=(5.0 / 3.0 - 1.0) * 1.0, gamma=5.0 / 3.0)  # E = 1 like the fiducial cell
    flat, gd = ghosted(fields, 2, PERIODIC)
    G = make_grid(n, cb.PCM, cb.HLLC, PERIODIC, 5.0 / 3.0, True, lengths=(9.0, 9.0, 9.0))
    G.upload(flat)
    got = G.Calc_Inverse_Timestep()
    G.close()
    assert abs(got - c["dti"]) <= 4 * np.spacing(c["dti"])


@pytest.mark.parametrize("flags", [[1, 1, 1, 1, 1, 1], [2, 2, 2, 2, 2, 2], [3, 3, 3, 3, 3, 3], [2, 3, 1, 1, 3, 2],
                                   [4, 1, 5, 2, 3, 3]])
def test_ghost_cells_all_boundary_kinds(flags):
    n = (14, 12, 10)
    g = 4
    fields = ic.random_smooth(n, seed=7)
    fields[1:4] *= 3.0  # make the diode bite on some cells
    NX, NY, NZ = n[0] + 2 * g, n[1] + 2 * g, n[2] + 2 * g
    a = np.random.default_rng(11).random((5, NZ, NY, NX)) + 0.5  # ghosts start as junk, identical in both arms
    a[:, g:-g, g:-g, g:-g] = fields
    flat = np.ascontiguousarray(a.reshape(-1))
    G = make_grid(n, cb.PPMC, cb.HLLC, [1] * 6, 1.4, True)
    G.flags = list(flags)
    G.upload(flat)
    for d in range(6):
        G.Set_Boundaries(d)
    got = G.download()
    G.close()
    want = flat.copy()
    oracle.set_boundaries(want, NX, NY, NZ, 5, g, flags)
    assert np.array_equal(got, want)


def test_pack_unpack_roundtrip_and_layout():
    import torch
    from cholla_b200.grid import halo_slab
    n = (12, 10, 8)
    g = 3
    NX, NY, NZ = n[0] + 2 * g, n[1] + 2 * g, n[2] + 2 * g
    N = NX * NY * NZ
    flat = np.random.default_rng(3).random(5 * N)
    G = make_grid(n, cb.PLMC, cb.HLLC, [1] * 6, 1.4, True)
    G.upload(flat)
    L = cb.lib()
    for face in range(6):
        s = halo_slab(NX, NY, NZ, g, face, recv=False)
        r = halo_slab(NX, NY, NZ, g, face ^ 1, recv=True)
        buf = torch.zeros(5 * s.ncells, dtype=torch.float64, device=G.device)
        assert L.chl_pack_buffer(C.c_void_p(buf.data_ptr()), G._ptr(), NX, NY, 5, N, s.idxoffset, s.isize, s.jsize,
                                 s.ksize, 2, G._stream()) == 0
        want = oracle.pack(flat, NX, NY, 5, N, s.idxoffset, s.isize, s.jsize, s.ksize)
        assert np.array_equal(buf.cpu().numpy(), want)
        # unpacking my own low slab into my own high ghosts == periodic wrap along that axis
        assert L.chl_unpack_buffer(C.c_void_p(buf.data_ptr()), G._ptr(), NX, NY, 5, N, r.idxoffset, r.isize,
                                   r.jsize, r.ksize, 2, G._stream()) == 0
        oracle.unpack(want, flat, NX, NY, 5, N, r.idxoffset, r.isize, r.jsize, r.ksize)
        assert np.array_equal(G.download(), flat)
    G.close()


def test_temperature_floor():
    n = (16, 12, 10)
    g = 4
    fields = ic.random_smooth(n, seed=9)
    flat, gd = ghosted(fields, g, PERIODIC)
    G = make_grid(n, cb.PPMC, cb.HLLC, PERIODIC, 5.0 / 3.0, True)
    G.upload(flat)
    assert cb.lib().chl_apply_temperature_floor(G._ctx, G._ptr(), 1.5, G._stream()) == 0
    got = G.download()
    G.close()
    want = oracle.temperature_floor(flat.copy(), *gd, g, 1.5)
    assert np.array_equal(got, want) and not np.array_equal(want, flat)


# ---- full-size, size-independent properties (BASELINE config 2 size: 256^3 PPMC + HLLC) ----
def test_256_cubed_conservation_and_uniform_state():
    import torch
    n = (256, 256, 256)
    G = make_grid(n, cb.PPMC, cb.HLLC, PERIODIC, 1.666666666666667, False, lengths=(4.0, 4.0, 4.0))
    # (1) a uniform moving state is a fixed point of the scheme
    G.Set_Initial_Conditions(ic.constant(n, rho=1.3, v=(0.3, -0.2, 0.1), P=0.7, gamma=1.666666666666667))
    G.Set_Boundary_Conditions_Grid()
    before = G.real_cells()
    G.advance(2)
    after = G.real_cells()
    assert np.abs(after - before).max() <= 1e-13
    # (2) periodic box: mass, momentum and energy are conserved to round-off
    f = ic.sound_wave(n, A=1e-2)
    G.Set_Initial_Conditions(f)
    G.Set_Boundary_Conditions_Grid()
    tot0 = f.reshape(5, -1).sum(axis=1)
    G.advance(3)
    tot1 = G.real_cells().reshape(5, -1).sum(axis=1)
    scale = np.abs(f).reshape(5, -1).sum(axis=1)
    assert (np.abs(tot1 - tot0) / scale).max() <= 1e-13
    # (3) the IC varies along x only and the scheme is dimensionally unsplit: the solution stays x-only
    r = G.real_cells()
    assert np.abs(r - r[:, :1, :1, :]).max() <= 1e-12
    G.close()
    torch.cuda.empty_cache()


@pytest.mark.parametrize("n,mhd", [((37, 19, 21), False), ((33, 40, 1), False), ((50, 1, 1), False), ((20, 35, 18), True)])
def test_snapshot_pack_kernel_matches_host_layout(n, mhd):
    """chl_pack_snapshot_field (CopyReal3D_GPU_Kernel's counterpart, io_gpu.cu:45-95) against the NumPy statement of the
    same layout in cholla_b200/snapshot.py: every field, face-centred magnetic fields included, fp64 and -> fp32."""
    from cholla_b200 import snapshot
    recon = cb.PPMC
    G = Grid3D(n, (1.0, 1.0, 1.0), recon=recon, solver=cb.HLLD if mhd else cb.HLLC, gamma=1.4, bc=[1] * 6, strict=True,
               device="cuda:0", mhd=mhd)
    H = G.H
    host = np.random.default_rng(11).standard_normal(H.n_fields * H.n_cells)
    G.upload(host)
    want = snapshot.grid_datasets(H, host, mhd=mhd)
    for f, nm in enumerate(snapshot.HYDRO_FIELDS):
        got = G.snapshot_field(f)
        assert got.shape == want[nm].shape and np.array_equal(got, want[nm].astype(np.float64)), nm
        assert np.array_equal(G.snapshot_field(f, out_precision=1), want[nm].astype(np.float32)), nm
    if mhd:
        for d, nm in enumerate(snapshot.MAGNETIC_FIELDS):
            got = G.snapshot_field(5 + d, mhd_direction=d)
            assert got.shape == want[nm].shape and np.array_equal(got, want[nm].astype(np.float64)), nm
    G.close()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("solver", [cb.HLLC, cb.ROE, cb.HLL])
@pytest.mark.parametrize("shape", [(37, 19, 21), (45, 23, 1), (90, 1, 1)])
def test_simple_integrator_plmc_strict_is_bit_exact(shape, solver, fused):
    """chl_simple_step(_fused) (Simple_Algorithm_{1,2,3}D_CUDA, -DSIMPLE): PLMC with characteristic tracing against the
    oracle, every solver, 1-/2-/3-D, with the temperature floor and the CFL reduction in the epilogue or as passes."""
    fields = ic.random_smooth(shape, seed=7)
    flags = [1, 1] + ([1, 1] if shape[1] > 1 else [0, 0]) + ([1, 1] if shape[2] > 1 else [0, 0])
    got, ref, dims, g, dts_gpu, dts_cpu = run_both(fields, cb.PLMC, solver, flags, 1.4, 4, strict=True, fused=fused,
                                                   U_floor=1e-3, integrator="simple")
    assert dts_gpu == dts_cpu
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("recon,solver", [(cb.PLMC, cb.HLLC), (cb.PLMC, cb.ROE), (cb.PCM, cb.HLLC)])
def test_simple_integrator_fast_mode_within_tolerance(recon, solver):
    fields = ic.random_smooth((37, 19, 21), seed=8)
    got, ref, dims, g, dts_gpu, dts_cpu = run_both(fields, recon, solver, PERIODIC, 5.0 / 3.0, 3, strict=False,
                                                   integrator="simple")
    assert rel_l1(real_view(got, dims, g), real_view(ref, dims, g)).max() <= TOL_F64
    assert np.allclose(dts_gpu, dts_cpu, rtol=1e-13, atol=0)


def test_simple_integrator_rejects_what_it_does_not_build():
    """PPMC under -DSIMPLE is PPMC_CTU, which is not on the path: the call fails loudly instead of running PPMC_VL."""
    G = make_grid((24, 24, 24), cb.PPMC, cb.HLLC, PERIODIC, 1.4, True, integrator="simple")
    G.upload(ghosted(ic.random_smooth((24, 24, 24), seed=1), n_ghost_for(cb.PPMC), PERIODIC)[0])
    G.H.dt = 1e-3
    with pytest.raises(cb.ChollaB200Error):
        G.Execute_Hydro_Integrator()
    G.close()


@pytest.mark.parametrize("recon", [cb.PCM, cb.PLMC, cb.PPMC])
@pytest.mark.parametrize("strict", [True, False])
def test_exact_solver_against_oracle(recon, strict):
    """chl_config.solver = CHL_SOLVER_EXACT against the CPU restatement: float round-off apart (the device's powf has no
    CPU twin -- the bit-level check is tests/test_gpu_reference.py::test_exact_solver_reference_vs_b200)."""
    fields = ic.random_smooth((37, 19, 21), seed=3)
    got, ref, dims, g, dts_gpu, dts_cpu = run_both(fields, recon, cb.EXACT, PERIODIC, 1.4, 3, strict=strict)
    assert rel_l1(real_view(got, dims, g), real_view(ref, dims, g)).max() <= 1e-5
    assert np.allclose(dts_gpu, dts_cpu, rtol=1e-5, atol=0)


@pytest.mark.parametrize("recon,solver", [(cb.PPMC, cb.HLLC), (cb.PLMC, cb.ROE), (cb.PPMC, cb.HLL)])
def test_strict_ieee_redo_path_extreme_operands(recon, solver):
    """The strict sweeps evaluate their quotients branch-free and redo a block with IEEE divisions when a quotient
    leaves the fast-path window (csrc/exact_div.cuh, kDivFlag).  Ordinary data never gets there; this state does, in
    most cells: velocities of 1e-300 (numerators below 2^-969, denormal quotients) on top of a smooth state, plus a
    slab at rest (exact zeros, which must NOT need the redo).  Same bits as the oracle either way."""
    rng = np.random.default_rng(5)
    fields = ic.random_smooth((24, 20, 18), seed=3)
    d = fields[0].copy()
    v = rng.standard_normal((3,) + d.shape) * 1e-300
    v[:, :, :6, :] = 0.0
    p = fields[4] - 0.5 * (fields[1] ** 2 + fields[2] ** 2 + fields[3] ** 2) / d
    for a in range(3):
        fields[1 + a] = d * v[a]
    fields[4] = p  # the kinetic energy underflows: E = e_int
    got, ref, dims, g, dts_gpu, dts_cpu = run_both(fields, recon, solver, PERIODIC, 5.0 / 3.0, 2, strict=True)
    assert dts_gpu == dts_cpu
    assert np.array_equal(got, ref)


def test_bad_cell_counter():
    """chl_bad_cells: a corrector result with density <= 0 is counted (the reference prints and averages neighbours
    in place, a race -- hydro_cuda.cu:292-297); healthy data counts zero."""
    fields = ic.random_smooth((20, 16, 12), seed=4)
    G = make_grid((20, 16, 12), cb.PLMC, cb.HLLC, PERIODIC, 5.0 / 3.0, True)
    G.Set_Initial_Conditions(fields)
    G.Set_Boundary_Conditions_Grid()
    G.advance(2)
    assert G.bad_cells() == 0
    # an absurd time step drives densities negative
    G.H.dt = 50.0 * G.H.dt
    G.Update_Hydro_Grid()
    assert G.bad_cells() > 0
    assert G.bad_cells() == 0  # reading resets the counter
    G.close()


@pytest.mark.parametrize("solver", [cb.HLLC, cb.ROE, cb.HLL])
@pytest.mark.parametrize("strict", [True, False])
def test_fused_predictor_equals_the_three_sweeps(solver, strict):
    """The single-launch predictor (csrc/vl_predictor.cuh: 32 x 8 columns marching along z, every cell converted once;
    an option, chl_set_tuning) against the default three direction sweeps: strict arithmetic bit for bit, on a ragged grid with several tiles
    and z chunks, mixed boundaries; product arithmetic within the parity budget."""
    n = (70, 21, 19)
    fields = ic.random_smooth(n, seed=9)
    # product arithmetic is compared on smooth data only: a boundary kink lets the PPM limiter amplify a rounding difference
    flags = [3, 3, 2, 2, 1, 1] if strict else PERIODIC
    out = []
    for fused, chunk in ((False, 0), (True, 0), (True, 5)):
        G = make_grid(n, cb.PPMC, solver, flags, 5.0 / 3.0, strict)
        G.set_tuning(dir_chunk=chunk, fused_predictor=fused)
        G.Set_Initial_Conditions(fields)
        G.Set_Boundary_Conditions_Grid()
        dts = G.advance(3)
        out.append((G.download(), dts))
        G.close()
    if strict:
        assert np.array_equal(out[0][0], out[1][0]) and out[0][1] == out[1][1]
        assert np.array_equal(out[0][0], out[2][0])
    else:
        g = n_ghost_for(cb.PPMC)
        dims = (n[0] + 2 * g, n[1] + 2 * g, n[2] + 2 * g)
        assert rel_l1(real_view(out[1][0], dims, g), real_view(out[0][0], dims, g)).max() <= TOL_F64
        assert np.array_equal(out[1][0], out[2][0])  # the z chunking never changes a bit
