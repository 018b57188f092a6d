"""`python -m cholla_b200.run deck.txt [name=value ...] [--recon ppmc --solver hllc ...]` -- the body of the
reference's `main()` (src/main.cpp:63-379) around the B200 library: read the deck, build the grid, set the initial
conditions, fill the ghost cells, then

    dti = G.Calc_Inverse_Timestep()
    while t < tout:   set_dt(dti) -> clip dt to the next output time -> Update_Hydro_Grid -> Update_Time
                      -> Set_Boundary_Conditions_Grid -> snapshot when t == outtime

What a deck cannot say -- the reference chooses reconstruction, Riemann solver, MHD and precision at compile time
through `builds/make.type.*` -- comes from the options, which map one-to-one to `chl_config`.  Snapshots carry the
reference's dataset and attribute names (`density, momentum_x, momentum_y, momentum_z, Energy[, magnetic_*]`; `gamma, t,
dt, n_step, dims, ...`, src/io/io.cpp:605-737,1350-1497) in HDF5 files of the reference's layout (`[i][j][k]` axis order,
big-endian types, one file per rank), written by cholla_b200/snapshot.py; `init=Read_Grid` restarts from one.
"""
import argparse
import os
import sys

import numpy as np

from . import _lib
from . import initial_conditions as ic
from .parameters import Parameters, parse_params

# global.h:26-47
MP, KB = 1.672622e-24, 1.380658e-16
TIME_UNIT, LENGTH_UNIT = 3.15569e10, 3.08567758e21
SP_ENERGY_UNIT = (LENGTH_UNIT / TIME_UNIT) ** 2

RECON = {"pcm": _lib.PCM, "plmc": _lib.PLMC, "ppmc": _lib.PPMC, "plmp": _lib.PLMP, "ppmp": _lib.PPMP}
SOLVER = {"hllc": _lib.HLLC, "roe": _lib.ROE, "hll": _lib.HLL, "hlld": _lib.HLLD, "exact": _lib.EXACT}


def u_floor(temperature_floor, gamma):
    """Grid3D::Update_Hydro_Grid, grid3D.cpp:493-503 (non-cosmological branch)."""
    return temperature_floor * KB / (gamma - 1.0) / MP / SP_ENERGY_UNIT


def initial_fields(P: Parameters, mhd=False, n_ghost=0):
    """Grid3D::Set_Initial_Conditions (src/grid/initial_conditions.cpp:30-100): dispatch on `init`.  Hydro generators
    return the real cells (5, nz, ny, nx) (honouring an enclosing `ic.block`); MHD ones the full local array."""
    n = P.n_global()
    lo, ln = P.domain()
    ln = tuple(v if v != 0.0 else 1.0 for v in ln)
    name = P.init
    if mhd and name == "Riemann":  # -DMHD build of Grid3D::Riemann: face fields either side of the diaphragm
        return ic.mhd_riemann(n, n_ghost, ln, lo, (P.rho_l, P.vx_l, P.vy_l, P.vz_l, P.P_l, P.Bx_l, P.By_l, P.Bz_l),
                              (P.rho_r, P.vx_r, P.vy_r, P.vz_r, P.P_r, P.Bx_r, P.By_r, P.Bz_r), P.diaph, P.gamma)
    if mhd and name == "Linear_Wave":
        return ic.mhd_linear_wave(n, n_ghost, ln, lo, P.rho, P.vx, P.P, (P.Bx, P.By, P.Bz), P.A,
                                  (P.rEigenVec_rho, P.rEigenVec_MomentumX, P.rEigenVec_MomentumY, P.rEigenVec_MomentumZ,
                                   P.rEigenVec_E, P.rEigenVec_Bx, P.rEigenVec_By, P.rEigenVec_Bz), P.wave_length, P.gamma,
                                  P.pitch, P.yaw)
    if mhd and name == "Constant":
        return ic.mhd_constant(n, n_ghost, ln, lo, P.rho, (P.vx, P.vy, P.vz), P.P, (P.Bx, P.By, P.Bz), P.gamma)
    if mhd and name == "Circularly_Polarized_Alfven_Wave":
        return ic.circularly_polarized_alfven_wave(n, n_ghost, ln, lo, P.polarization, P.vx, P.pitch, P.yaw, P.gamma)
    if mhd and name == "Advecting_Field_Loop":
        return ic.advecting_field_loop(n, n_ghost, ln, lo, P.rho, (P.vx, P.vy, P.vz), P.P, P.A, P.radius, P.gamma)
    if mhd and name not in ("Orszag_Tang_Vortex", "MHD_Spherical_Blast"):
        # the remaining hydro generators return real cells of a 5-field state: say so instead of failing in the upload
        raise NotImplementedError("init=%s has no MHD (8-field, face-centred B) generator here; the MHD ones are Constant, "
                                  "Riemann, Linear_Wave, Orszag_Tang_Vortex, MHD_Spherical_Blast, "
                                  "Circularly_Polarized_Alfven_Wave, Advecting_Field_Loop" % name)
    if name == "Constant":
        return ic.constant(n, P.rho, (P.vx, P.vy, P.vz), P.P, P.gamma)
    if name == "Sound_Wave":
        return ic.sound_wave(n, ln, lo, P.rho, (P.vx, P.vy, P.vz), P.P, P.A, P.gamma)
    if name == "Riemann":
        return ic.riemann(n, ln, lo, (P.rho_l, P.vx_l, P.vy_l, P.vz_l, P.P_l), (P.rho_r, P.vx_r, P.vy_r, P.vz_r, P.P_r),
                          P.diaph, P.gamma)
    if name == "Spherical_Overpressure_3D":
        return ic.spherical_overpressure(n, ln, lo, P.gamma)
    if name == "Spherical_Overdensity_3D":
        return ic.spherical_overdensity(n, ln, lo, P.gamma)
    if name == "KH_res_ind":
        return ic.kh_res_ind(n, ln, lo, P.gamma)
    if name == "Square_Wave":
        return ic.square_wave(n, ln, lo, P.rho, (P.vx, P.vy, P.vz), P.P, P.A, P.gamma)
    if name == "Shu_Osher":
        return ic.shu_osher(n, ln, lo, P.gamma)
    if name == "Blast_1D":
        return ic.blast_1d(n, ln, lo, P.gamma)
    if name == "KH":
        return ic.kh(n, ln, lo, P.gamma)
    if name == "Gresho":
        return ic.gresho(n, ln, lo, P.gamma)
    if name == "Implosion_2D":
        return ic.implosion_2d(n, ln, lo, P.gamma)
    if name == "Rayleigh_Taylor":
        return ic.rayleigh_taylor(n, ln, lo, P.gamma)
    if name == "Linear_Wave" and not mhd:
        return ic.linear_wave(n, ln, lo, P.rho, P.vx, P.P, P.A, (P.rEigenVec_rho, P.rEigenVec_MomentumX,
                              P.rEigenVec_MomentumY, P.rEigenVec_MomentumZ, P.rEigenVec_E), P.wave_length, P.pitch, P.yaw,
                              P.gamma)
    if name == "Orszag_Tang_Vortex" and mhd:
        return ic.orszag_tang(n, n_ghost, P.gamma, ln, lo)
    if name == "MHD_Spherical_Blast" and mhd:
        return ic.mhd_blast(n, n_ghost, P.gamma, ln, lo, P.rho, (P.vx, P.vy, P.vz), P.P, P.P_blast, P.radius,
                            (P.Bx, P.By, P.Bz))
    raise NotImplementedError("init=%s is not one of the generators of the hot-path scope (SURVEY 8d): Constant, Sound_Wave, "
                              "Riemann, Square_Wave, Shu_Osher, Blast_1D, KH, KH_res_ind, Gresho, Implosion_2D, Rayleigh_Taylor, Linear_Wave, "
                              "Spherical_Overpressure_3D, Spherical_Overdensity_3D, Orszag_Tang_Vortex, MHD_Spherical_Blast, Circularly_Polarized_Alfven_Wave, "
                              "Advecting_Field_Loop" % name)


def snapshot(G, P: Parameters, nfile, outdir):
    """Write_Data -> Output_Data -> Write_Header_HDF5 + Write_Grid_HDF5 (io.cpp:92-260, 605-737, 1350-1497): one HDF5
    file per rank, `<outdir>/<nfile>/<nfile>.h5.<rank>` -- see cholla_b200/snapshot.py."""
    from . import snapshot as snap
    path = None
    if nfile % max(P.n_hydro, 1) == 0:  # io.cpp:126-130
        path = snap.write_data(G, P, nfile, outdir)
    f32 = snap.write_float32(G, P, nfile, outdir)  # io.cpp:133-137: n_out_float32 and the out_float32_* selection
    snap.write_slices(G, P, nfile, outdir)  # io.cpp:152-156: n_slice
    snap.write_projection(G, P, nfile, outdir)  # io.cpp:140-144: n_projection
    return path if path is not None else f32


def read_grid(G, P: Parameters):
    """Grid3D::Read_Grid (io.cpp:2255-2310): `init=Read_Grid` restarts from `<indir><nfile>.h5.<rank>`; gamma, t and
    n_step come from the file."""
    from . import snapshot as snap
    path = "%s%d.h5.%d" % (P.indir, P.nfile, G.rank)
    host, gamma, t, n_step = snap.read_grid(path, G.H, G.mhd, G.n_scalars, G.dual_energy)
    if gamma != G.gamma:
        raise ValueError("%s was written with gamma = %r, the deck says %r" % (path, gamma, G.gamma))
    G.upload(host)
    G.H.t, G.H.n_step = t, n_step


def run(P: Parameters, recon="ppmc", solver="hllc", strict=True, precision=2, mhd=False, outdir=None, device=None,
        dist_group=None, rank=0, nproc=1, max_steps=None, log=print, integrator="vl", dual_energy=False, scalar=False,
        static_gravity=False, cooling=False):
    """main.cpp:63-379.  Returns (Grid3D, list of snapshot paths).  dual_energy / scalar / static_gravity / cooling stand
    for the build macros -DDE, -DSCALAR -DBASIC_SCALAR, -DSTATIC_GRAV (the field is the deck's custom_grav) and
    -DCOOLING_GPU (CIE curve) of the reference's make.type files."""
    from .grid import Grid3D, n_ghost_for
    mhd = mhd or solver == "hlld"
    lo, ln = P.domain()
    ln = tuple(v if v != 0.0 else 1.0 for v in ln)
    G = Grid3D(P.n_global(), ln, bounds_min=lo, recon=RECON[recon], solver=SOLVER[solver], gamma=P.gamma,
               bc=P.boundary_flags(), precision=precision, strict=strict, density_floor=P.density_floor,
               temperature_floor_U=u_floor(P.temperature_floor, P.gamma), device=device, dist_group=dist_group,
               rank=rank, nproc=nproc, mhd=mhd, integrator=integrator, n_scalars=1 if scalar else 0,
               dual_energy=dual_energy, cooling="cie" if cooling else None)
    if static_gravity:
        G.Set_Static_Gravity(P.custom_grav)
    is_restart = P.init == "Read_Grid"  # main.cpp:89-95
    if is_restart:
        read_grid(G, P)
    elif mhd:
        with ic.block(G.decomp.offset, G.decomp.n_local):
            G.upload(initial_fields(P, True, n_ghost_for(RECON[recon], True)))
    else:
        with ic.block(G.decomp.offset, G.decomp.n_local):
            f = initial_fields(P)
            extra = []
            if scalar:
                extra.append(ic.basic_scalar_for(P.init, f, P, ln, lo, P.n_global()))
            if dual_energy:  # the generators store P / (gamma - 1): E minus the kinetic energy, to round-off
                extra.append(f[4] - 0.5 * (f[1] ** 2 + f[2] ** 2 + f[3] ** 2) / f[0])
            G.Set_Initial_Conditions_Local(np.concatenate([f, np.stack(extra)]) if extra else f)
    G.Set_Boundary_Conditions_Grid()
    outdir = outdir if outdir is not None else (P.outdir or "./")
    files, nfile, outtime = [], P.nfile, 0.0
    if is_restart:
        outtime += G.H.t  # main.cpp:124-128

    def output(n):
        path = snapshot(G, P, n, outdir)
        if path is not None:  # n_hydro may skip the full snapshot of this output number (io.cpp:126-130)
            files.append(path)

    if outdir != "none" and not is_restart:
        output(nfile)  # main.cpp:205-215: the initial state is output 0
    nfile += 1
    outtime += P.outstep
    dti = G.Calc_Inverse_Timestep()
    while G.H.t < P.tout and (max_steps is None or G.H.n_step < max_steps):
        G.set_dt(dti)
        nxt = min(outtime, P.tout)  # main.cpp:257-261
        if G.H.t + G.H.dt > nxt:
            G.H.dt = nxt - G.H.t
        dti = G.Update_Hydro_Grid()
        G.Update_Time()
        G.Set_Boundary_Conditions_Grid()
        if G.mhd:  # main.cpp:375-378: mhd::checkMagneticDivergence after every step, limit 1e-9
            divb = G.max_magnetic_divergence()
            if not divb <= 1.0e-9:
                raise FloatingPointError("The magnetic divergence has exceeded the maximum allowed value. Divergence = "
                                         "%7.4e, the maximum allowed divergence = %7.4e" % (divb, 1.0e-9))
        bad = G.bad_cells()
        if bad:  # the reference prints and averages neighbours in place (hydro_cuda.cu:292-297); here the host decides
            raise FloatingPointError("%d cell(s) came out of the corrector with density <= 0 at step %d" % (bad, G.H.n_step))
        if P.output_always or G.H.t == outtime:  # main.cpp:327 (Output_Now every step), 341-350
            if outdir != "none":
                output(nfile)
            nfile += 1
            if G.H.t == outtime:
                outtime += P.outstep
    log("n_step: %d  sim time: %.10f  last dt: %.6e" % (G.H.n_step, G.H.t, G.H.dt))
    return G, files


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("deck")
    ap.add_argument("overrides", nargs="*", help="name=value pairs that override the deck (main.cpp argv handling)")
    ap.add_argument("--recon", default="ppmc", choices=sorted(RECON))
    ap.add_argument("--solver", default="hllc", choices=sorted(SOLVER))
    ap.add_argument("--mhd", action="store_true")
    ap.add_argument("--integrator", default="vl", choices=["vl", "simple"], help="-DVL / -DSIMPLE (make.type.hydro:8-12)")
    ap.add_argument("--fast", action="store_true", help="product arithmetic instead of the reference's bits")
    ap.add_argument("--dual-energy", action="store_true", help="-DDE")
    ap.add_argument("--scalar", action="store_true", help="-DSCALAR -DBASIC_SCALAR")
    ap.add_argument("--static-gravity", action="store_true", help="-DSTATIC_GRAV with the deck's custom_grav")
    ap.add_argument("--cooling", action="store_true", help="-DCOOLING_GPU (CIE curve)")
    ap.add_argument("--precision", type=int, default=2, choices=[1, 2])
    ap.add_argument("--outdir", default=None, help="overrides the deck's outdir; 'none' writes nothing")
    ap.add_argument("--max-steps", type=int, default=None)
    a = ap.parse_args(argv)
    P = parse_params(a.deck, a.overrides)
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", 1), ("RANK", 0), ("LOCAL_RANK", 0)))
    group = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        group = dist.group.WORLD
    G, files = run(P, a.recon, a.solver, strict=not a.fast and a.precision == 2, precision=a.precision, mhd=a.mhd,
                   outdir=a.outdir, device="cuda:%d" % local, dist_group=group, rank=rank, nproc=world,
                   max_steps=a.max_steps, log=print if rank == 0 else (lambda *_: None), integrator=a.integrator,
                   dual_energy=a.dual_energy, scalar=a.scalar, static_gravity=a.static_gravity, cooling=a.cooling)
    G.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
