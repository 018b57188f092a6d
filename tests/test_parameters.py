"""Parameter-file reader (cholla_b200/parameters.py) against the rules of the reference's Parse_Params / Parse_Param
(src/global/global.cpp:127-470), and the deck-driven main loop (cholla_b200/run.py, src/main.cpp:63-379).
The deck text below is written for this test in the format of the reference's examples/1D/sod.txt (config C1)."""
import os

import numpy as np
import pytest

from cholla_b200 import parameters as pm

SOD_DECK = """#
# Parameter File for 1D Sod Shock tube
#

################################################
# number of grid cells in the x dimension
nx=100
ny=1
nz=1
# final output time
tout=0.2
outstep=0.2
init=Riemann
xmin=0.0
ymin=0.0
zmin=0.0
xlen=1.0
ylen=1.0
zlen=1.0
xl_bcnd=3
xu_bcnd=3
yl_bcnd=0
yu_bcnd=0
zl_bcnd=0
zu_bcnd=0
outdir=./
; a semicolon comment
rho_l=1.0
vx_l=0.0
vy_l=0.0
vz_l=0.0
P_l=1.0
rho_r=0.1
vx_r=0.0
vy_r=0.0
vz_r=0.0
P_r=0.1
diaph=0.5
gamma=1.4
"""


def test_sod_deck_values():
    P = pm.parse_text(SOD_DECK)
    assert (P.nx, P.ny, P.nz) == (100, 1, 1) and P.n_global() == [100, 1, 1]
    assert P.tout == 0.2 and P.outstep == 0.2 and P.init == "Riemann" and P.gamma == 1.4
    assert P.boundary_flags() == [3, 3, 0, 0, 0, 0]
    assert (P.rho_l, P.P_l, P.rho_r, P.P_r, P.diaph) == (1.0, 1.0, 0.1, 0.1, 0.5)
    assert P.domain() == ((0.0, 0.0, 0.0), (1.0, 1.0, 1.0)) and P.outdir == "./" and not P.unknown


def test_line_rules_and_c_number_parsing():
    text = ("nx=  64 \n"          # the value is trimmed (Trim, global.cpp:82-101)
            "nfile = 3\n"         # ... the name is not: 'nfile ' matches nothing
            " ny=8\n"             # the NAME is used verbatim (global.cpp:153-157): ' ny' is not 'ny'
            "#nz=3\n"             # comment: first character
            " #nz=5\n"            # NOT a comment for the reference (first character is a blank) -> unknown name ' #nz'
            "tout=1e-1xyz\n"      # atof reads the numeric prefix
            "gamma=abc\n"         # no digits: atof -> 0
            "nz=7.9\n"            # atoi stops at the point
            "init=Sound_Wave\r\n"  # Trim removes the carriage return too
            "=5\n"                # strtok: a single token -> skipped
            "justaword\n"
            "xlen=2=3\n")         # third token ignored
    warned = []
    P = pm.parse_text(text, warn=warned.append)
    assert P.nx == 64 and P.ny == 0 and P.nz == 7
    assert P.tout == 0.1 and P.gamma == 0.0 and P.init == "Sound_Wave" and P.xlen == 2.0
    assert set(P.unknown) == {"nfile ", " ny", " #nz"} and len(warned) == 3 and P.nfile == 0
    assert pm.atoi("  -12abc") == -12 and pm.atoi("x1") == 0 and pm.atof(".5e1q") == 5.0 and pm.atof("+3.") == 3.0


def test_command_line_overrides_and_optional_names():
    warned = []
    P = pm.parse_text(SOD_DECK + "n_proc_x=2\ntile_length=3\nbogus=1\nn_slice=4\n", argv=["nx=256", "tout=0.05", "noequals"],
                      warn=warned.append)
    assert P.nx == 256 and P.tout == 0.05 and P.n_proc_x == 2
    assert list(P.unknown) == ["bogus"] and len(warned) == 1  # optionalParams and other-physics names do not warn
    with pytest.raises(ValueError):
        pm.parse_text("output_always=2\n")
    with pytest.raises(OSError):
        pm.parse_params("/nonexistent/deck.txt")


def test_temperature_floor_units():
    from cholla_b200 import run
    # grid3D.cpp:497: U_floor = T * KB / (gamma - 1) / MP / SP_ENERGY_UNIT, with SP_ENERGY_UNIT = (kpc / kyr)^2 in cgs
    T, gamma = 1.0e4, 5.0 / 3.0
    want = T * 1.380658e-16 / (gamma - 1.0) / 1.672622e-24 / (3.08567758e21 / 3.15569e10) ** 2
    assert run.u_floor(T, gamma) == want and run.u_floor(0.0, gamma) == 0.0


def test_initial_condition_dispatch():
    from cholla_b200 import initial_conditions as ic, run
    P = pm.parse_text(SOD_DECK)
    f = run.initial_fields(P)
    assert f.shape == (5, 1, 1, 100) and np.array_equal(f, ic.riemann((100, 1, 1), gamma=1.4))
    P.init = "Disk_3D"
    with pytest.raises(NotImplementedError):
        run.initial_fields(P)


@pytest.mark.gpu
def test_deck_driven_run_matches_oracle_loop(tmp_path):
    """run.py on the Sod deck (PLMC + HLLC, config C1): same bytes as the oracle advanced by the reference's loop
    (dt = 0.3 / max_dti, clipped to the output time), snapshot 0 and 1 written with the reference's dataset names."""
    import oracle
    import cholla_b200 as cb
    from cholla_b200 import run
    from cholla_b200.grid import n_ghost_for
    from helpers import ghosted, real_view
    deck = tmp_path / "sod.txt"
    deck.write_text(SOD_DECK)
    P = pm.parse_params(str(deck), ["outdir=%s/" % tmp_path])
    G, files = run.run(P, recon="plmc", solver="hllc", strict=True, log=lambda *_: None)
    got = G.real_cells()
    n_step, t = G.H.n_step, G.H.t
    G.close()
    assert t == 0.2 and n_step > 20 and len(files) == 2
    g = n_ghost_for(cb.PLMC)
    flat, dims = ghosted(run.initial_fields(P), g, P.boundary_flags())
    tc, steps = 0.0, 0
    dti = oracle.calc_dti(flat, *dims, g, 0.01, 1.0, 1.0, 1.4)
    while tc < 0.2:
        dt = 0.3 / dti
        if tc + dt > 0.2:
            dt = 0.2 - tc
        oracle.vl_step(flat, *dims, g, 0.01, 1.0, 1.0, dt, 1.4, cb.PLMC, cb.HLLC)
        dti = oracle.calc_dti(flat, *dims, g, 0.01, 1.0, 1.0, 1.4)
        oracle.set_boundaries(flat, *dims, 5, g, P.boundary_flags())
        tc += dt
        steps += 1
    assert steps == n_step and np.array_equal(got, real_view(flat, dims, g))
    from cholla_b200 import h5lite
    with h5lite.File(files[1]) as z:  # the reference's HDF5 layout: 1-D datasets of nx_real big-endian doubles
        assert z["density"].shape == (100,) and z["density"].dtype == np.dtype(">f8")
        assert float(z.attrs["t"][0]) == 0.2 and int(z.attrs["n_step"][0]) == n_step and z.attrs["dims"].tolist() == [100, 1, 1]
        assert np.array_equal(z["density"].read(), got[0, 0, 0, :]) and np.array_equal(z["Energy"].read(), got[4, 0, 0, :])
    assert os.path.basename(files[1]) == "1.h5.0" and os.path.basename(os.path.dirname(files[1])) == "1"


@pytest.mark.gpu
def test_restart_from_snapshot_continues_bit_for_bit(tmp_path):
    """init=Read_Grid (io.cpp:2255-2310, main.cpp:89-128): a run restarted from its own snapshot 1 (t = 0.1) reaches
    t = 0.2 with the bytes, step count and snapshot of the uninterrupted run."""
    from cholla_b200 import h5lite, run
    deck = tmp_path / "sod.txt"
    deck.write_text(SOD_DECK)
    P = pm.parse_params(str(deck), ["outdir=%s/a/" % tmp_path, "outstep=0.1"])
    G, files = run.run(P, recon="plmc", solver="hllc", strict=True, log=lambda *_: None)
    full, n_full = G.real_cells(), G.H.n_step
    G.close()
    assert len(files) == 3 and os.path.normpath(files[1]).endswith("/1/1.h5.0")
    R = pm.parse_params(str(deck), ["outdir=%s/b/" % tmp_path, "outstep=0.1", "init=Read_Grid", "nfile=1",
                                    "indir=%s/a/1/" % tmp_path])
    G2, files2 = run.run(R, recon="plmc", solver="hllc", strict=True, log=lambda *_: None)
    again, n_again, t_again = G2.real_cells(), G2.H.n_step, G2.H.t
    G2.close()
    assert t_again == 0.2 and n_again == n_full and np.array_equal(again, full)
    assert len(files2) == 1 and os.path.normpath(files2[0]).endswith("/2/2.h5.0")  # a restart does not rewrite its input (main.cpp:198)
    with h5lite.File(files[2]) as a, h5lite.File(files2[0]) as b:
        for nm in a.keys():
            assert np.array_equal(a[nm].read(), b[nm].read())
        assert a.attrs["n_step"][0] == b.attrs["n_step"][0] and a.attrs["t"][0] == b.attrs["t"][0]


OT_DECK = """# Orszag-Tang vortex, in the format of the reference's examples/3D/orszag_tang_vortex.txt (config C4), small grid
nx=24
ny=24
nz=24
tout=0.5
outstep=0.5
init=Orszag_Tang_Vortex
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
gamma=1.666666666666667
"""


@pytest.mark.gpu
def test_deck_driven_mhd_run(tmp_path):
    """run.py on an Orszag-Tang deck (PPMC + HLLD + CT): the same bytes as the MHD oracle advanced over the same number
    of steps, div B at round-off, magnetic datasets in the snapshot."""
    from oracle import mhd
    import cholla_b200 as cb
    from cholla_b200 import initial_conditions as ic, run
    from cholla_b200.grid import n_ghost_for
    deck = tmp_path / "ot.txt"
    deck.write_text(OT_DECK)
    P = pm.parse_params(str(deck), ["outdir=%s/" % tmp_path])
    steps = 6
    G, files = run.run(P, recon="ppmc", solver="hlld", strict=True, max_steps=steps, log=lambda *_: None)
    got = G.download()
    H = G.H
    assert H.n_step == steps and H.n_fields == 8 and len(files) == 1  # only the initial snapshot: tout not reached
    G.close()
    g = n_ghost_for(cb.PPMC, mhd=True)
    dims = (24 + 2 * g,) * 3
    flat = np.ascontiguousarray(ic.orszag_tang((24, 24, 24), g, P.gamma).reshape(-1))
    mhd.set_boundaries(flat, *dims, g, [1] * 6)
    d = 1.0 / 24
    mhd.advance(flat, *dims, g, d, d, d, P.gamma, cb.PPMC, [1] * 6, steps)
    real = lambda a: a.reshape(8, dims[2], dims[1], dims[0])[:, g:-g, g:-g, g:-g]
    a, b = real(got), real(flat)
    groups = [[0], [1, 2, 3], [4], [5, 6, 7]]
    err = max(sum(np.abs(a[f] - b[f]).sum() for f in q) / sum(np.abs(b[f]).sum() for f in q) for q in groups)
    assert err <= 1e-12, err  # rsqrt / rhypot have no correctly rounded libm twin: tolerance, not bits
    assert mhd.max_divergence(got, *dims, g, d, d, d) <= 1e-11
    from cholla_b200 import h5lite
    with h5lite.File(files[0]) as z:  # face-centred fields carry one extra face along their own direction
        assert z["magnetic_x"].shape == (25, 24, 24) and z["magnetic_z"].shape == (24, 24, 25)
        assert int(z.attrs["n_fields"][0]) == 8 and "magnetic_field_unit" in z.attrs and "-DMHD" in z.attrs["Macro Flags"]
