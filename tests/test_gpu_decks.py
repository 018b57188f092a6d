"""Deck-driven runs of the generators added for the reference's shipped examples (tests/test_initial_conditions.py builds
them on the CPU): each deck goes through cholla_b200.run on the GPU for a few steps and must leave the bytes (hydro) or
the 1e-12-close state (MHD: rsqrt / rhypot have no correctly rounded CPU twin) of the CPU oracle advanced by
the reference's loop, and a readable HDF5 snapshot."""
import numpy as np
import pytest

import cholla_b200 as cb
from cholla_b200 import h5lite, parameters as pm, run
from cholla_b200.grid import n_ghost_for
from helpers import ghosted, real_view

pytestmark = pytest.mark.gpu

BASE = """tout=10.0
outstep=10.0
xmin=%(xmin)s
ymin=%(ymin)s
zmin=%(zmin)s
xlen=%(xlen)s
ylen=%(ylen)s
zlen=%(zlen)s
outdir=./
"""


def deck(tmp_path, init, n, bc, gamma, lo=(0.0, 0.0, 0.0), ln=(1.0, 1.0, 1.0), **extra):
    text = "nx=%d\nny=%d\nnz=%d\ngamma=%r\ninit=%s\n" % (n + (gamma, init))
    text += "xl_bcnd=%d\nxu_bcnd=%d\nyl_bcnd=%d\nyu_bcnd=%d\nzl_bcnd=%d\nzu_bcnd=%d\n" % tuple(bc)
    text += BASE % dict(xmin=lo[0], ymin=lo[1], zmin=lo[2], xlen=ln[0], ylen=ln[1], zlen=ln[2])
    text += "".join("%s=%r\n" % kv for kv in extra.items())
    p = tmp_path / (init + ".txt")
    p.write_text(text)
    return pm.parse_params(str(p), ["outdir=%s/" % tmp_path])


@pytest.mark.parametrize("init,n,bc,gamma,lo,ln,recon,solver", [
    ("Shu_Osher", (200, 1, 1), [3, 3, 0, 0, 0, 0], 1.4, (-1.0, 0.0, 0.0), (2.0, 1.0, 1.0), "ppmc", "hllc"),
    ("Blast_1D", (128, 1, 1), [2, 2, 0, 0, 0, 0], 1.4, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), "plmc", "hllc"),
    ("Gresho", (48, 48, 1), [1, 1, 1, 1, 0, 0], 1.4, (-0.5, -0.5, 0.0), (1.0, 1.0, 1.0), "ppmc", "hllc"),
    ("Implosion_2D", (40, 40, 1), [2, 2, 2, 2, 0, 0], 1.4, (0.0, 0.0, 0.0), (0.3, 0.3, 1.0), "plmc", "hllc"),
    ("KH", (32, 32, 1), [1, 1, 1, 1, 0, 0], 5.0 / 3.0, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), "ppmc", "roe")])
def test_hydro_deck_matches_the_oracle_loop(tmp_path, init, n, bc, gamma, lo, ln, recon, solver):
    import oracle
    P = deck(tmp_path, init, n, bc, gamma, lo, ln)
    steps = 8
    G, files = run.run(P, recon=recon, solver=solver, strict=True, max_steps=steps, log=lambda *_: None)
    got, H = G.real_cells(), G.H
    G.close()
    assert H.n_step == steps
    r, s = run.RECON[recon], run.SOLVER[solver]
    g = n_ghost_for(r)
    flat, dims = ghosted(run.initial_fields(P), g, P.boundary_flags())
    oracle.advance(flat, *dims, g, H.dx, H.dy, H.dz, gamma, r, s, P.boundary_flags(), steps)
    assert np.array_equal(got, real_view(flat, dims, g))
    with h5lite.File(files[0]) as f:  # the initial snapshot, in the reference's layout
        assert f["density"].shape == tuple(v for v in n if v > 1) and f.attrs["n_step"][0] == 0


@pytest.mark.parametrize("init,bc,extra", [
    ("Riemann", [3, 3, 1, 1, 1, 1], dict(rho_l=1.0, P_l=1.0, Bx_l=0.75, By_l=1.0, rho_r=0.125, P_r=0.1, Bx_r=0.75, By_r=-1.0,
                                         diaph=0.5)),
    ("Circularly_Polarized_Alfven_Wave", [1, 1, 1, 1, 1, 1], dict(polarization=1.0, vx=0.0)),
    ("Advecting_Field_Loop", [1, 1, 1, 1, 1, 1], dict(rho=1.0, vx=1.0, vy=1.0, vz=2.0, P=1.0, A=0.001, radius=0.3))])
def test_mhd_deck_matches_the_oracle_loop(tmp_path, init, bc, extra):
    from oracle import mhd
    n = (32, 16, 16)
    lo, ln = ((-0.5, -0.5, -0.5), (1.0, 1.0, 1.0)) if init == "Advecting_Field_Loop" else ((0.0, 0.0, 0.0), (1.0, 0.5, 0.5))
    gamma = 2.0 if init == "Riemann" else 5.0 / 3.0
    P = deck(tmp_path, init, n, bc, gamma, lo, ln, **extra)
    steps = 5
    G, files = run.run(P, recon="plmc", solver="hlld", strict=True, max_steps=steps, log=lambda *_: None)
    got, H = G.download(), G.H
    G.close()
    assert H.n_step == steps and H.n_fields == 8
    g = n_ghost_for(cb.PLMC, mhd=True)
    dims = tuple(v + 2 * g for v in n)
    flat = np.ascontiguousarray(run.initial_fields(P, True, g).reshape(-1))
    mhd.set_boundaries(flat, *dims, g, bc)
    mhd.advance(flat, *dims, g, H.dx, H.dy, H.dz, gamma, cb.PLMC, bc, steps)
    real = lambda a: a.reshape(8, dims[2], dims[1], dims[0])[:, g:-g, g:-g, g:-g]
    a, b = real(got), real(flat)
    # smooth wave: the 1e-12 parity bar.  The shock tube and the loop (kinks in the field) sit next to limiter decisions
    # that a last-bit difference can flip (tests/test_oracle_sensitivity.py), and the MHD kernels are bit-exact against
    # the reference's kernels, not against the CPU's libm: a looser bar there
    tol = {"Riemann": 1e-8, "Advecting_Field_Loop": 1e-10}.get(init, 1e-12)
    for q in ([0], [1, 2, 3], [4], [5, 6, 7]):
        err = sum(np.abs(a[f] - b[f]).sum() for f in q) / max(sum(np.abs(b[f]).sum() for f in q), 1e-300)
        assert err <= tol, (q, err)
    assert mhd.max_divergence(got, *dims, g, H.dx, H.dy, H.dz) <= 1e-11
    with h5lite.File(files[0]) as f:
        assert f["magnetic_x"].shape == (n[0] + 1, n[1], n[2])


@pytest.mark.parametrize("mhd", [False, True])
def test_float32_output_files(tmp_path, mhd):
    """n_out_float32 / out_float32_* (Output_Float32, io.cpp:266-375): `<n>.float32.h5.<rank>` beside the full snapshot,
    transposed and narrowed on the device; the MHD datasets in the reference's (n+1)^3 format."""
    import os
    from cholla_b200 import snapshot
    init = "Orszag_Tang_Vortex" if mhd else "Sound_Wave"
    extra = dict(n_out_float32=1, out_float32_density=1, out_float32_Energy=1, rho=1.0, P=0.6, A=0.01)
    if mhd:
        extra.update(out_float32_magnetic_x=1, out_float32_magnetic_z=1)
    P = deck(tmp_path, init, (24, 16, 12), [1] * 6, 5.0 / 3.0, **extra)
    G, files = run.run(P, recon="plmc", solver="hlld" if mhd else "hllc", strict=True, max_steps=2, log=lambda *_: None)
    G.close()
    path = os.path.join(os.path.dirname(files[0]), "0.float32.h5.0")
    with h5lite.File(files[0]) as full, h5lite.File(path) as f32:
        want = ["density", "Energy"] + (["magnetic_x", "magnetic_z"] if mhd else [])
        assert sorted(f32.keys()) == sorted(want) and list(f32.attrs) == list(full.attrs)
        for nm in ("density", "Energy"):
            assert f32[nm].dtype == np.dtype(">f4") and np.array_equal(f32[nm].read(), full[nm].read().astype(np.float32))
        if mhd:
            bx, bz = f32["magnetic_x"].read(), f32["magnetic_z"].read()
            assert bx.shape == (25, 17, 13) and bz.shape == (25, 17, 13)
            # the real faces are a sub-block of the float32 box: x faces start at the same x index, one further in y and z
            assert np.array_equal(bx[:, 1:, 1:], full["magnetic_x"].read().astype(np.float32))
            assert np.array_equal(bz[1:, 1:, :], full["magnetic_z"].read().astype(np.float32))


@pytest.mark.parametrize("mhd", [False, True])
def test_slice_output_files(tmp_path, mhd):
    """n_slice (Output_Slices / Write_Slices_HDF5, io.cpp:518-552, 1835-2250): `<n>_slice.h5.<rank>` holds the three global
    mid-planes of the state the full snapshot holds."""
    import os
    init = "Orszag_Tang_Vortex" if mhd else "Sound_Wave"
    P = deck(tmp_path, init, (24, 16, 12), [1] * 6, 5.0 / 3.0, n_slice=1, n_projection=1, rho=1.0, P=0.6, A=0.01)
    G, files = run.run(P, recon="plmc", solver="hlld" if mhd else "hllc", strict=True, max_steps=2, log=lambda *_: None)
    G.close()
    # n_projection (Output_Projected_Data, io.cpp:380-430): column densities and density-weighted temperatures, reduced on
    # the device
    with h5lite.File(files[0]) as full, h5lite.File(os.path.join(os.path.dirname(files[0]), "0_proj.h5.0")) as pr:
        d = full["density"].read()
        assert sorted(pr.keys()) == ["T_xy", "T_xz", "d_xy", "d_xz"] and pr["d_xy"].shape == (24, 16) and pr["T_xz"].shape == (24, 12)
        assert np.allclose(pr["d_xy"].read(), d.sum(axis=2) / 12.0, rtol=1e-13) and np.allclose(pr["d_xz"].read(), d.sum(axis=1) / 16.0, rtol=1e-13)
        assert (pr["T_xy"].read() > 0).all()
    with h5lite.File(files[0]) as full, h5lite.File(os.path.join(os.path.dirname(files[0]), "0_slice.h5.0")) as s:
        assert list(s.attrs) == list(full.attrs)
        for nm, ds in (("d", "density"), ("mx", "momentum_x"), ("E", "Energy")):
            a = full[ds].read()
            assert np.array_equal(s[nm + "_xy"].read(), a[:, :, 6]) and np.array_equal(s[nm + "_xz"].read(), a[:, 8, :])
            assert np.array_equal(s[nm + "_yz"].read(), a[12, :, :])
        if mhd:
            bx = full["magnetic_x"].read()  # faces 0 .. 24: cell i lies between faces i and i + 1
            assert np.array_equal(s["magnetic_x_xy"].read(), 0.5 * (bx[1:, :, 6] + bx[:-1, :, 6]))


WAVES = {  # hydro_system_tests.cpp:205-258: eigenvector (rho, mx, my, mz, E), vx, fiducial number of time steps, L1 factor
    "SoundWaveRightMoving": ((1.0, 1.0, 1.0, 1.0, 1.5), 0.0, 214, 2.0),
    "SoundWaveLeftMoving": ((1.0, -1.0, 1.0, 1.0, 1.5), 0.0, 214, 2.0),
    "HydroContactWave": ((1.0, 1.0, 0.0, 0.0, 0.5), 1.0, 427, 1.0),
}


@pytest.mark.parametrize("recon,allowed", [("plmc", 1e-7), ("ppmc", 2.7e-8)])
@pytest.mark.parametrize("wave", sorted(WAVES))
def test_reference_linear_wave_system_tests(tmp_path, wave, recon, allowed):
    """tHYDROtMHDSYSTEMLinearWavesParameterizedMpi (hydro_system_tests.cpp:142-258) with the reference's pass criteria,
    deck-driven on the GPU: one period on 64 x 32 x 32, the fiducial number of time steps, L1 error against the initial
    snapshot inside the Gardiner & Stone bound of the reconstruction (tests/test_system_tests.py runs the same on the
    CPU oracle)."""
    from cholla_b200 import system_test
    ev, vx, steps, factor = WAVES[wave]
    P = deck(tmp_path, "Linear_Wave", (64, 32, 32), [1] * 6, 5.0 / 3.0, ln=(1.0, 0.5, 0.5), rho=1.0, vx=vx, vy=0.0, vz=0.0,
             P=1.0 / (5.0 / 3.0), A=1e-6, rEigenVec_rho=ev[0], rEigenVec_MomentumX=ev[1], rEigenVec_MomentumY=ev[2],
             rEigenVec_MomentumZ=ev[3], rEigenVec_E=ev[4])
    P = pm.parse_params(str(tmp_path / "Linear_Wave.txt"), ["outdir=%s" % tmp_path, "tout=1.0", "outstep=1.0"])
    G, files = run.run(P, recon=recon, solver="hllc", strict=True, log=lambda *_: None)
    G.close()
    assert len(files) == 2
    fails, info = system_test.l1_error_test(str(tmp_path), factor * allowed, allowed, expected_n_steps=steps)
    assert not fails, (fails, info)


def test_reference_sound_wave_3d_system_test(tmp_path):
    """tHYDROtMHDSYSTEMSoundWave3D (hydro_system_tests.cpp:96-128): 64^3 to t = 0.05, analytic sine to 1e-7."""
    from cholla_b200 import system_test
    deck(tmp_path, "Linear_Wave", (64, 64, 64), [1] * 6, 1.666666666666667, rho=1.0, vx=0.0, vy=0.0, vz=0.0, P=0.6, A=1e-5,
         rEigenVec_rho=1.0, rEigenVec_MomentumX=1.0, rEigenVec_MomentumY=1.0, rEigenVec_MomentumZ=1.0, rEigenVec_E=1.5)
    P = pm.parse_params(str(tmp_path / "Linear_Wave.txt"), ["outdir=%s" % tmp_path, "tout=0.05", "outstep=0.05"])
    G, files = run.run(P, recon="plmc", solver="hllc", strict=True, log=lambda *_: None)
    G.close()
    kx = 2 * np.pi / 64.0
    phase = kx * 0.5 - 1.0 * 0.05 * 2 * np.pi
    assert system_test.analytic_sine(str(tmp_path), "density", 1.0, 1e-5, kx, 0.0, 0.0, phase, 1e-7)
    assert system_test.analytic_sine(str(tmp_path), "momentum_x", 0.0, 1e-5, kx, 0.0, 0.0, phase, 1e-7)


def test_rayleigh_taylor_deck_with_static_gravity_matches_the_oracle_loop(tmp_path):
    """examples/2D/Rayleigh_Taylor.txt at a smaller size: init=Rayleigh_Taylor, custom_grav=2 under -DSTATIC_GRAV (PLMC + HLLC
    as builds/make.type.static_grav selects).  The field is exact (g = (0, -1)), so the oracle loop must give the bytes."""
    import oracle
    n, gamma, bc = (40, 80, 1), 1.4, [1, 1, 2, 2, 0, 0]
    P = deck(tmp_path, "Rayleigh_Taylor", n, bc, gamma, (0.0, 0.0, 0.0), (0.33333333, 1.0, 1.0), custom_grav=2)
    steps = 10
    G, files = run.run(P, recon="plmc", solver="hllc", strict=True, max_steps=steps, log=lambda *_: None, static_gravity=True)
    got, H = G.real_cells(), G.H
    G.close()
    g = n_ghost_for(cb.PLMC)
    flat, dims = ghosted(run.initial_fields(P), g, bc)
    start = flat.copy()
    oracle.set_gravity(oracle.static_gravity_field(*dims, g, 2, (0, 0, 0), H.dx, H.dy, H.dz, (0.0, 0.0, 0.0)))
    try:
        oracle.advance(flat, *dims, g, H.dx, H.dy, H.dz, gamma, cb.PLMC, cb.HLLC, bc, steps)
    finally:
        oracle.set_gravity(None)
    assert np.array_equal(got, real_view(flat, dims, g))
    assert np.abs(real_view(flat, dims, g)[2] - real_view(start, dims, g)[2]).max() > 1e-3  # the slab is falling


def test_sod_deck_with_dual_energy_and_scalar_snapshot_and_restart(tmp_path):
    """A -DDE -DSCALAR -DBASIC_SCALAR run of the Riemann generator: oracle bytes, the GasEnergy and scalar0 datasets of
    Write_Grid_HDF5 (io.cpp:1413-1424), and a restart from them (Read_Grid_HDF5, io.cpp:2538-2544)."""
    import oracle
    n, gamma, bc = (64, 1, 1), 1.4, [3, 3, 0, 0, 0, 0]
    extra = dict(rho_l=1.0, vx_l=0.0, vy_l=0.0, vz_l=0.0, P_l=1.0, rho_r=0.1, vx_r=0.0, vy_r=0.0, vz_r=0.0, P_r=0.1, diaph=0.5)
    P = deck(tmp_path, "Riemann", n, bc, gamma, **extra)
    P.outstep = 0.02
    G, files = run.run(P, recon="ppmc", solver="hllc", strict=True, max_steps=12, log=lambda *_: None, dual_energy=True,
                       scalar=True)
    got, H, t_end = G.real_cells(), G.H, G.H.t
    G.close()
    assert H.n_fields == 7 and len(files) >= 2
    with h5lite.File(files[0]) as f:
        assert f["GasEnergy"].shape == (64,) and f["scalar0"].shape == (64,)
        assert np.allclose(f["scalar0"].read()[:32], 1.0) and np.all(f["scalar0"].read()[32:] == 0.0)
        assert np.allclose(f["GasEnergy"].read()[:32], 1.0 / (gamma - 1.0))
    # the scalar marks the left state: it stays within [0, rho] (PPM undershoots by denormal amounts ahead of the contact)
    assert got[5].min() >= -1e-100 and np.all(got[5] <= got[0] * (1 + 1e-12))
    # restart from the last full snapshot and continue to the same time: same bytes as the uninterrupted run
    with h5lite.File(files[-1]) as f:
        n_last = int(f.attrs["n_step"][0])
    assert 0 < n_last < 12
    import os
    P2 = deck(tmp_path, "Read_Grid", n, bc, gamma, **extra)
    P2.outstep, P2.nfile = 0.02, len(files) - 1
    P2.indir = os.path.dirname(files[-1]) + "/"
    G2, _ = run.run(P2, recon="ppmc", solver="hllc", strict=True, max_steps=12, log=lambda *_: None, dual_energy=True,
                    scalar=True, outdir="none")
    again = G2.real_cells()
    assert G2.H.n_step == 12
    G2.close()
    assert np.array_equal(again, got)
