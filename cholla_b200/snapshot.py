"""The reference's HDF5 snapshot layout: writer, restart reader and the per-rank concatenation.

    Write_Data / Output_Data       src/io/io.cpp:92-260      -> write_data()
    FnameTemplate                  src/io/io.cpp:2779-2828   -> effective_output_dir_path(), format_fname()
    Grid3D::Write_Header_HDF5      src/io/io.cpp:605-737     -> header_attributes()
    Grid3D::Write_Grid_HDF5        src/io/io.cpp:1350-1497   -> grid_datasets()
    CopyReal3D_GPU_Kernel          src/io/io_gpu.cu:45-95    -> the [i][j][k] transposition and the MHD face shift
    Grid3D::Read_Grid / _HDF5      src/io/io.cpp:2255-2310, 2508-2720 -> read_grid()
    python_scripts/concat_3d_data.py, concat_internals.py    -> concat_3d_dataset()

Files are real HDF5 (cholla_b200/h5lite.py writes the format directly; this image has no libhdf5): one file per rank,
`<outdir>/<nfile>/<nfile>.h5.<rank>`, datasets `H5T_IEEE_F64BE` (or F32BE for a PRECISION=1 build) of shape
(nx_real, ny_real, nz_real) -- x slowest, "indexing that plays well with Python" -- and face-centred magnetic fields
with one extra face along their own direction.  Attributes carry the reference's names, types (F64BE / I32BE /
variable-length strings), shapes ((1,) and (3,)) and creation order.  Everything here works on host arrays, so it is
exercised without a GPU.
"""
import os
from collections import OrderedDict

import numpy as np

from . import h5lite

# src/global/global.h:40-48
TIME_UNIT = 3.15569e10
LENGTH_UNIT = 3.08567758e21
MASS_UNIT = 1.98847e33
DENSITY_UNIT = MASS_UNIT / (LENGTH_UNIT * LENGTH_UNIT * LENGTH_UNIT)
VELOCITY_UNIT = LENGTH_UNIT / TIME_UNIT
ENERGY_UNIT = DENSITY_UNIT * VELOCITY_UNIT * VELOCITY_UNIT
MAGNETIC_FIELD_UNIT = (MASS_UNIT / LENGTH_UNIT) ** 0.5 / TIME_UNIT

HYDRO_FIELDS = ["density", "momentum_x", "momentum_y", "momentum_z", "Energy"]
MAGNETIC_FIELDS = ["magnetic_x", "magnetic_y", "magnetic_z"]
_RECON_FLAG = {0: "-DPCM", 1: "-DPLMC", 2: "-DPPMC", 3: "-DPLMP", 4: "-DPPMP"}
_SOLVER_FLAG = {0: "-DHLLC", 1: "-DROE", 2: "-DHLLD", 3: "-DHLL", 4: "-DEXACT"}


def effective_output_dir_path(outdir, nfile, legacy_flat_outdir=False):
    """FnameTemplate::effective_output_dir_path, io.cpp:2779-2795 (always ends in '/')."""
    if not outdir:
        return "./"
    if not legacy_flat_outdir:
        return outdir + "/" + str(nfile) + "/"
    return os.path.dirname(outdir) + "/"


def format_fname(outdir, nfile, proc_id, pre_extension_suffix="", legacy_flat_outdir=False):
    """FnameTemplate::format_fname, io.cpp:2807-2828, for an HDF5 build."""
    prefix = outdir if legacy_flat_outdir else effective_output_dir_path(outdir, nfile) + "/"
    return prefix + str(nfile) + pre_extension_suffix + ".h5" + "." + str(proc_id)


def macro_flags(recon, solver, precision=2, mhd=False, integrator="vl", G=None):
    """The MACRO_FLAGS string of the equivalent reference build (Makefile:173-175, builds/make.type.*).  G: the live grid,
    for the macros of a dual-energy / scalar / static-gravity / cooling context."""
    flags = ["-DCUDA", "-DMPI_CHOLLA", "-DPRECISION=%d" % precision, _RECON_FLAG[recon], _SOLVER_FLAG[solver], "-DSIMPLE" if integrator == "simple" else "-DVL",
             "-DDENSITY_FLOOR", "-DTEMPERATURE_FLOOR", "-DHDF5", "-DOUTPUT"]
    if mhd:
        flags.append("-DMHD")
    if G is not None:
        if getattr(G, "dual_energy", False):
            flags.append("-DDE")
        if getattr(G, "n_scalars", 0):
            flags += ["-DSCALAR", "-DBASIC_SCALAR"]
        if getattr(G, "static_gravity", False):
            flags.append("-DSTATIC_GRAV")
        if getattr(G, "cooling", None):
            flags.append("-DCOOLING_GPU")
    return " ".join(flags)


def _f64(*v):
    return np.array(v, dtype=">f8")


def _i32(*v):
    return np.array(v, dtype=">i4")


def header_attributes(H, gamma, n_global, n_local, offset, nprocs, bounds, domain, mhd=False, git_hash="",
                      macro=""):
    """Grid3D::Write_Header_HDF5 of an MPI_CHOLLA build, io.cpp:605-737, in creation order."""
    a = OrderedDict()
    a["gamma"] = _f64(gamma)
    a["Git Commit Hash"] = h5lite.VarString(git_hash)
    a["Macro Flags"] = h5lite.VarString(macro)
    a["cholla"] = h5lite.VarString("")
    a["t"] = _f64(H.t)
    a["dt"] = _f64(H.dt)
    a["n_step"] = _i32(H.n_step)
    a["n_fields"] = _i32(H.n_fields)
    a["time_unit"] = _f64(TIME_UNIT)
    a["length_unit"] = _f64(LENGTH_UNIT)
    a["mass_unit"] = _f64(MASS_UNIT)
    a["velocity_unit"] = _f64(VELOCITY_UNIT)
    a["density_unit"] = _f64(DENSITY_UNIT)
    a["energy_unit"] = _f64(ENERGY_UNIT)
    if mhd:
        a["magnetic_field_unit"] = _f64(MAGNETIC_FIELD_UNIT)
    a["dims"] = _i32(*n_global)
    a["dims_local"] = _i32(*n_local)
    a["offset"] = _i32(*offset)
    a["nprocs"] = _i32(*nprocs)
    a["bounds"] = _f64(*bounds)
    a["domain"] = _f64(*domain)
    a["dx"] = _f64(H.dx, H.dy, H.dz)
    return a


def passive_datasets(H, mhd, n_scalars=0, dual_energy=False):
    """(dataset name, field index) of the -DDE / -DSCALAR fields in the order Write_Grid_HDF5 emits them (io.cpp:1413-1424):
    GasEnergy (the last field) right after Energy, then scalar0 (BASIC_SCALAR, field 5)."""
    out = []
    if not mhd and dual_energy:
        out.append(("GasEnergy", H.n_fields - 1))
    if not mhd and n_scalars:
        out.append(("scalar0", 5))
    return out


def _ghosts(H):
    g = H.n_ghost
    return g, (g if H.ny > 1 else 0), (g if H.nz > 1 else 0)


def grid_datasets(H, C_host, mhd=False, precision=2, n_scalars=0, dual_energy=False):
    """Grid3D::Write_Grid_HDF5 with Output_Complete_Data, io.cpp:1350-1497.  C_host: (n_fields, nz, ny, nx) host array,
    ghost cells included.  1-D grids give (nx_real,), 2-D (nx_real, ny_real), 3-D (nx_real, ny_real, nz_real)."""
    C_host = np.asarray(C_host).reshape(H.n_fields, H.nz, H.ny, H.nx)
    gx, gy, gz = _ghosts(H)
    dt = ">f8" if precision == 2 else ">f4"
    ndim = 1 + (H.ny > 1) + (H.nz > 1)
    out = OrderedDict()
    for f, nm in enumerate(HYDRO_FIELDS):
        real = C_host[f, gz:gz + H.nz_real, gy:gy + H.ny_real, gx:gx + H.nx_real]
        t = real.transpose(2, 1, 0)  # [i][j][k]: dest = k + j*nz_real + i*nz_real*ny_real
        out[nm] = np.ascontiguousarray(t.reshape(t.shape[:ndim])).astype(dt)
    for nm, f in passive_datasets(H, mhd, n_scalars, dual_energy):
        t = C_host[f, gz:gz + H.nz_real, gy:gy + H.ny_real, gx:gx + H.nx_real].transpose(2, 1, 0)
        out[nm] = np.ascontiguousarray(t.reshape(t.shape[:ndim])).astype(dt)
    if mhd and ndim == 3:  # io.cpp:1478-1487; the face on the low side of the first real cell is part of the output
        for d, nm in enumerate(MAGNETIC_FIELDS):
            sx, sy, sz = (1 if d == 0 else 0), (1 if d == 1 else 0), (1 if d == 2 else 0)
            real = C_host[5 + d, gz - sz:gz + H.nz_real, gy - sy:gy + H.ny_real, gx - sx:gx + H.nx_real]
            out[nm] = np.ascontiguousarray(real.transpose(2, 1, 0)).astype(dt)
    return out


def write_snapshot(path, H, C_host, gamma, n_global, n_local, offset, nprocs, bounds, domain, mhd=False, precision=2,
                   git_hash="", macro=""):
    """Output_Data, io.cpp:198-260: header then grid, one file."""
    attrs = header_attributes(H, gamma, n_global, n_local, offset, nprocs, bounds, domain, mhd, git_hash, macro)
    h5lite.write(path, attrs, grid_datasets(H, C_host, mhd, precision))
    return path


def write_data(G, P, nfile, outdir):
    """Write_Data -> Output_Data for one rank of a live Grid3D (device -> host -> file)."""
    H = G.H
    os.makedirs(effective_output_dir_path(outdir, nfile, getattr(P, "legacy_flat_outdir", False)),
                exist_ok=True)  # Ensure_Dir_Exists, io.cpp:2830-
    path = format_fname(outdir, nfile, G.rank, "", getattr(P, "legacy_flat_outdir", False))
    attrs = header_attributes(H, P.gamma, G.n_global, G.decomp.n_local, G.decomp.offset, G.decomp.np,
                              (P.xmin, P.ymin, P.zmin), (P.xlen, P.ylen, P.zlen), G.mhd,
                              macro=macro_flags(G.recon, G.solver, G.precision, G.mhd, getattr(G, "integrator", "vl"), G))
    dt = ">f8" if G.precision == 2 else ">f4"
    data = OrderedDict()
    for f, nm in enumerate(HYDRO_FIELDS):  # device-side transposition, one field at a time like the reference
        data[nm] = G.snapshot_field(f).astype(dt)
    for nm, f in passive_datasets(H, G.mhd, getattr(G, "n_scalars", 0), getattr(G, "dual_energy", False)):
        data[nm] = G.snapshot_field(f).astype(dt)
    if G.mhd and H.nz > 1:
        for d, nm in enumerate(MAGNETIC_FIELDS):
            data[nm] = G.snapshot_field(5 + d, mhd_direction=d).astype(dt)
    h5lite.write(path, attrs, data)
    return path


FLOAT32_SELECT = [("density", 0, "out_float32_density"), ("momentum_x", 1, "out_float32_momentum_x"),
                  ("momentum_y", 2, "out_float32_momentum_y"), ("momentum_z", 3, "out_float32_momentum_z"),
                  ("Energy", 4, "out_float32_Energy")]
FLOAT32_SELECT_MHD = [("magnetic_x", 5, "out_float32_magnetic_x"), ("magnetic_y", 6, "out_float32_magnetic_y"),
                      ("magnetic_z", 7, "out_float32_magnetic_z")]


def float32_datasets(H, C_host, P, mhd=False):
    """Output_Float32 (io.cpp:266-375): the fields the deck selects with out_float32_<name> > 0, as H5T_IEEE_F32BE.  3-D
    only.  In an MHD build the magnetic datasets have the reference's "different output format": (n+1)^3 cells starting
    one below the real cells on EVERY axis (io.cpp:347-362 passes n_ghost - 1 and no direction)."""
    C_host = np.asarray(C_host).reshape(H.n_fields, H.nz, H.ny, H.nx)
    g = H.n_ghost
    out = OrderedDict()
    for nm, f, key in FLOAT32_SELECT:
        if getattr(P, key, 0) > 0:
            out[nm] = np.ascontiguousarray(C_host[f, g:g + H.nz_real, g:g + H.ny_real, g:g + H.nx_real].transpose(2, 1, 0)).astype(">f4")
    if mhd:
        for nm, f, key in FLOAT32_SELECT_MHD:
            if getattr(P, key, 0) > 0:
                out[nm] = np.ascontiguousarray(C_host[f, g - 1:g + H.nz_real, g - 1:g + H.ny_real, g - 1:g + H.nx_real].transpose(2, 1, 0)).astype(">f4")
    return out


def write_float32(G, P, nfile, outdir):
    """Write_Data's Output_Float32 leg for one rank of a live Grid3D: `<outdir>/<n>/<n>.float32.h5.<rank>`; returns the
    path, or None where the reference returns without writing (1-D / 2-D grids, n_out_float32 unset or not dividing
    nfile)."""
    H = G.H
    if H.ny_real == 1 or H.nz_real == 1 or not P.n_out_float32 or nfile % P.n_out_float32 != 0:
        return None
    flat = getattr(P, "legacy_flat_outdir", False)
    os.makedirs(effective_output_dir_path(outdir, nfile, flat), exist_ok=True)
    path = format_fname(outdir, nfile, G.rank, ".float32", flat)
    attrs = header_attributes(H, P.gamma, G.n_global, G.decomp.n_local, G.decomp.offset, G.decomp.np,
                              (P.xmin, P.ymin, P.zmin), (P.xlen, P.ylen, P.zlen), G.mhd,
                              macro=macro_flags(G.recon, G.solver, G.precision, G.mhd, getattr(G, "integrator", "vl"), G))
    data = OrderedDict()
    for nm, f, key in FLOAT32_SELECT:
        if getattr(P, key, 0) > 0:
            data[nm] = G.snapshot_field(f, out_precision=1).astype(">f4")
    if G.mhd:
        for nm, f, key in FLOAT32_SELECT_MHD:
            if getattr(P, key, 0) > 0:
                data[nm] = G.snapshot_field(f, out_precision=1, extra=1).astype(">f4")
    h5lite.write(path, attrs, data)
    return path


def slice_datasets(H, get, n_global, offset, mhd=False, precision=2):
    """Grid3D::Write_Slices_HDF5, 3-D MPI_CHOLLA branch (io.cpp:1835-2250): the xy, xz and yz planes through the middle
    of the GLOBAL grid (index n_global / 2), real cells only, as d_/mx_/my_/mz_/E_<plane> (+ cell-centred
    magnetic_{x,y,z}_<plane> in an MHD build).  A rank the plane does not cross writes zeros.  `get(f, ks, js, is_)`
    returns field f over the given local index slices as a numpy array, so the planes can come straight from device
    memory."""
    g = H.n_ghost
    dt = ">f8" if precision == 2 else ">f4"
    n_local = (H.nx_real, H.ny_real, H.nz_real)
    real = [slice(g, g + H.nx_real), slice(g, g + H.ny_real), slice(g, g + H.nz_real)]
    names = [("d", 0), ("mx", 1), ("my", 2), ("mz", 3), ("E", 4)]
    out = OrderedDict()
    for plane, axis in (("xy", 2), ("xz", 1), ("yz", 0)):
        mid = n_global[axis] // 2
        inside = offset[axis] <= mid < offset[axis] + n_local[axis]
        loc = mid - offset[axis] + g
        keep = [a for a in (0, 1, 2) if a != axis]
        shape = tuple(n_local[a] for a in keep)

        def plane_of(f, shift=(0, 0, 0)):
            sl = [slice(real[a].start - shift[a], real[a].stop - shift[a]) for a in range(3)]
            sl[axis] = slice(loc - shift[axis], loc - shift[axis] + 1)
            a = np.asarray(get(f, sl[2], sl[1], sl[0]))  # (nz, ny, nx) with one axis of length 1
            return np.squeeze(a, axis=2 - axis).T  # -> [first kept axis][second kept axis], x before y before z

        for nm, f in names:
            out["%s_%s" % (nm, plane)] = (plane_of(f) if inside else np.zeros(shape)).astype(dt)
        if mhd:
            for d, nm in enumerate(MAGNETIC_FIELDS):
                sh = tuple(1 if a == d else 0 for a in range(3))
                v = 0.5 * (plane_of(5 + d) + plane_of(5 + d, sh)) if inside else np.zeros(shape)
                out["%s_%s" % (nm, plane)] = v.astype(dt)
    return out


def write_slices(G, P, nfile, outdir):
    """Output_Slices (io.cpp:518-552): `<outdir>/<n>/<n>_slice.h5.<rank>`, header + the three mid-planes.  Only the cells
    of the planes leave the device.  3-D grids; returns None when n_slice is unset or does not divide nfile."""
    H = G.H
    if not getattr(P, "n_slice", 0) or nfile % P.n_slice != 0 or H.nz_real == 1 or H.ny_real == 1:
        return None
    flat = getattr(P, "legacy_flat_outdir", False)
    os.makedirs(effective_output_dir_path(outdir, nfile, flat), exist_ok=True)
    path = format_fname(outdir, nfile, G.rank, "_slice", flat)
    C4 = G.C_device.view(H.n_fields, H.nz, H.ny, H.nx)
    get = lambda f, ks, js, is_: C4[f, ks, js, is_].to("cpu", G.torch.float64).numpy()
    attrs = header_attributes(H, P.gamma, G.n_global, G.decomp.n_local, G.decomp.offset, G.decomp.np,
                              (P.xmin, P.ymin, P.zmin), (P.xlen, P.ylen, P.zlen), G.mhd,
                              macro=macro_flags(G.recon, G.solver, G.precision, G.mhd, getattr(G, "integrator", "vl"), G))
    h5lite.write(path, attrs, slice_datasets(H, get, G.n_global, G.decomp.offset, G.mhd, G.precision))
    return path


MP, KB = 1.672622e-24, 1.380658e-16  # global.h:26-27
PRESSURE_UNIT = DENSITY_UNIT * VELOCITY_UNIT * VELOCITY_UNIT


def projection_datasets(H, C4, gamma, mhd=False, precision=2):
    """Grid3D::Write_Projection_HDF5, 3-D branch without DE / DUST (io.cpp:1502-1660): column density along z and along y,
    d_xy = sum_k d dz, d_xz = sum_j d dy, and the density-weighted temperature T_xy = sum_k T d dz, T_xz = sum_j T d dy
    with T = P PRESSURE_UNIT / (n KB), n = d DENSITY_UNIT / (mu MP), mu = 0.6 (hydro_utilities.h:60-88).  C4 is the
    (n_fields, nz, ny, nx) state as a numpy array or a torch tensor: only arithmetic, slicing and sums are used, so the
    reduction runs where the data lives and four small planes come back.  (The sums associate differently from the
    reference's sequential loops: equal to round-off.)"""
    g, mu, tiny = H.n_ghost, 0.6, 1e-20
    r = (slice(g, g + H.nz_real), slice(g, g + H.ny_real), slice(g, g + H.nx_real))
    d, mx, my, mz, E = (C4[f][r] for f in range(5))
    e = E - 0.5 * (mx * mx + my * my + mz * mz) / d
    if mhd:  # cellCenteredMagneticFields: average of the two faces of the cell
        lo = [(slice(g, g + H.nz_real), slice(g, g + H.ny_real), slice(g - 1, g - 1 + H.nx_real)),
              (slice(g, g + H.nz_real), slice(g - 1, g - 1 + H.ny_real), slice(g, g + H.nx_real)),
              (slice(g - 1, g - 1 + H.nz_real), slice(g, g + H.ny_real), slice(g, g + H.nx_real))]
        for a in range(3):
            b = 0.5 * (C4[5 + a][r] + C4[5 + a][lo[a]])
            e = e - 0.5 * b * b
    P = e * (gamma - 1.0)
    P = P * (P > tiny) + tiny * (P <= tiny)  # fmax(P, TINY_NUMBER), hydro_utilities.h:48-58
    n = d * DENSITY_UNIT / (mu * MP)
    Td = (P * PRESSURE_UNIT / (n * KB)) * d
    dt = ">f8" if precision == 2 else ">f4"
    host = lambda a: np.asarray(a.cpu() if hasattr(a, "cpu") else a, dtype=np.float64).T.astype(dt)  # -> [x][y] / [x][z]
    out = OrderedDict()
    out["d_xy"] = host(d.sum(0) * H.dz)
    out["d_xz"] = host(d.sum(1) * H.dy)
    out["T_xy"] = host(Td.sum(0) * H.dz)
    out["T_xz"] = host(Td.sum(1) * H.dy)
    return out


def write_projection(G, P, nfile, outdir):
    """Output_Projected_Data (io.cpp:380-430): `<outdir>/<n>/<n>_proj.h5.<rank>`, header + the four projections, reduced on
    the device.  3-D grids; None when n_projection is unset or does not divide nfile."""
    H = G.H
    if not getattr(P, "n_projection", 0) or nfile % P.n_projection != 0 or H.nz_real == 1 or H.ny_real == 1:
        return None
    flat = getattr(P, "legacy_flat_outdir", False)
    os.makedirs(effective_output_dir_path(outdir, nfile, flat), exist_ok=True)
    path = format_fname(outdir, nfile, G.rank, "_proj", flat)
    attrs = header_attributes(H, P.gamma, G.n_global, G.decomp.n_local, G.decomp.offset, G.decomp.np,
                              (P.xmin, P.ymin, P.zmin), (P.xlen, P.ylen, P.zlen), G.mhd,
                              macro=macro_flags(G.recon, G.solver, G.precision, G.mhd, getattr(G, "integrator", "vl"), G))
    C4 = G.C_device.view(H.n_fields, H.nz, H.ny, H.nx).to(G.torch.float64)
    h5lite.write(path, attrs, projection_datasets(H, C4, P.gamma, G.mhd, G.precision))
    return path


def read_grid(path, H, mhd=False, n_scalars=0, dual_energy=False):
    """Grid3D::Read_Grid_HDF5, io.cpp:2508-2720: returns (C_host (n_fields, nz, ny, nx) with the real cells -- and, for
    MHD, the real faces -- filled and the ghost cells zero, gamma, t, n_step).  The caller sets the boundaries next, as
    main.cpp does."""
    gx, gy, gz = _ghosts(H)
    ndim = 1 + (H.ny > 1) + (H.nz > 1)
    out = np.zeros((H.n_fields, H.nz, H.ny, H.nx), dtype=np.float64)
    with h5lite.File(path) as f:
        gamma = float(f.attrs["gamma"][0])
        t = float(f.attrs["t"][0])
        n_step = int(f.attrs["n_step"][0])
        want = (H.nx_real, H.ny_real, H.nz_real)[:ndim]
        for nm, i in [(nm, i) for i, nm in enumerate(HYDRO_FIELDS)] + passive_datasets(H, mhd, n_scalars, dual_energy):
            a = f[nm].read()
            if a.shape != want:
                raise ValueError("%s: dataset %s has shape %s, the grid expects %s" % (path, nm, a.shape, want))
            a = a.reshape(a.shape + (1,) * (3 - ndim)).astype(np.float64)
            out[i, gz:gz + H.nz_real, gy:gy + H.ny_real, gx:gx + H.nx_real] = a.transpose(2, 1, 0)
        if mhd and ndim == 3:
            for d, nm in enumerate(MAGNETIC_FIELDS):
                sx, sy, sz = (1 if d == 0 else 0), (1 if d == 1 else 0), (1 if d == 2 else 0)
                a = f[nm].read().astype(np.float64)
                if a.shape != (H.nx_real + sx, H.ny_real + sy, H.nz_real + sz):
                    raise ValueError("%s: dataset %s has shape %s" % (path, nm, a.shape))
                out[5 + d, gz - sz:gz + H.nz_real, gy - sy:gy + H.ny_real, gx - sx:gx + H.nx_real] = a.transpose(2, 1, 0)
    return out, gamma, t, n_step


def concat_3d_dataset(source_directory, output_directory, num_processes, output_number, skip_fields=(),
                      destination_dtype=None):
    """python_scripts/concat_3d_data.py: merge `<src>/<n>/<n>.h5.<rank>` of all ranks into `<out>/<n>.h5`.  The header
    is rank 0's without `dims_local` and `offset` (concat_internals.copy_header); magnetic datasets overlap by the
    shared face."""
    os.makedirs(output_directory, exist_ok=True)
    dest_path = os.path.join(output_directory, "%d.h5" % output_number)
    data, attrs = OrderedDict(), None
    for rank in range(num_processes):
        src = os.path.join(source_directory, str(output_number), "%d.h5.%d" % (output_number, rank))
        with h5lite.File(src) as f:
            if attrs is None:
                attrs = OrderedDict((k, v) for k, v in f.attrs.items() if k not in ("dims_local", "offset"))
                dims = tuple(int(v) for v in f.attrs["dims"])
            nloc = [int(v) for v in f.attrs["dims_local"]]
            off = [int(v) for v in f.attrs["offset"]]
            for nm in f.keys():
                if nm in skip_fields:
                    continue
                ds = f[nm]
                a = ds.read()
                extra = [1 if nm == "magnetic_" + "xyz"[d] else 0 for d in range(3)]
                if nm not in data:
                    shape = tuple(dims[d] + extra[d] for d in range(a.ndim))
                    data[nm] = np.zeros(shape, dtype=destination_dtype or ds.dtype)
                sl = tuple(slice(off[d], off[d] + nloc[d] + extra[d]) for d in range(a.ndim))
                data[nm][sl] = a
    out_attrs = OrderedDict()
    for k, v in attrs.items():
        if isinstance(v, str):
            out_attrs[k] = h5lite.VarString(v)
        else:
            out_attrs[k] = np.asarray(v).astype(">f8" if np.asarray(v).dtype.kind == "f" else ">i4")
    h5lite.write(dest_path, out_attrs, data)
    return dest_path


def concat_2d_dataset(source_directory, output_directory, num_processes, output_number, dataset_kind, skip_fields=(),
                      destination_dtype=None):
    """python_scripts/concat_2d_data.py for `slice` and `proj` outputs: `<src>/<n>/<n>_<kind>.h5.<rank>` of all ranks into
    `<out>/<n>_<kind>.h5`.  A rank's plane is ADDED at its offset -- projections of ranks stacked along the line of sight
    sum, slices of ranks the plane misses are zeros (and, like the script, are skipped)."""
    if dataset_kind not in ("slice", "proj"):
        raise ValueError("dataset_kind can only be 'slice' or 'proj' (rotated projections are not written)")
    os.makedirs(output_directory, exist_ok=True)
    dest_path = os.path.join(output_directory, "%d_%s.h5" % (output_number, dataset_kind))
    data, attrs = OrderedDict(), None
    axes = {"xy": (0, 1, 2), "yz": (1, 2, 0), "xz": (0, 2, 1)}  # the two kept axes, then the one through the plane
    for rank in range(num_processes):
        src = os.path.join(source_directory, str(output_number), "%d_%s.h5.%d" % (output_number, dataset_kind, rank))
        with h5lite.File(src) as f:
            if attrs is None:
                attrs = OrderedDict((k, v) for k, v in f.attrs.items() if k not in ("dims_local", "offset"))
            dims = [int(v) for v in f.attrs["dims"]]
            nloc = [int(v) for v in f.attrs["dims_local"]]
            off = [int(v) for v in f.attrs["offset"]]
            for nm in f.keys():
                if nm in skip_fields:
                    continue
                a0, a1, ax = axes[nm[-2:]]
                ds = f[nm]
                if nm not in data:
                    data[nm] = np.zeros((dims[a0], dims[a1]), dtype=destination_dtype or ds.dtype)
                if dataset_kind == "slice" and not (off[ax] <= dims[ax] // 2 <= off[ax] + nloc[ax]):
                    continue
                data[nm][off[a0]:off[a0] + nloc[a0], off[a1]:off[a1] + nloc[a1]] += ds.read().astype(data[nm].dtype)
    out_attrs = OrderedDict()
    for k, v in attrs.items():
        if isinstance(v, str):
            out_attrs[k] = h5lite.VarString(v)
        else:
            out_attrs[k] = np.asarray(v).astype(">f8" if np.asarray(v).dtype.kind == "f" else ">i4")
    h5lite.write(dest_path, out_attrs, data)
    return dest_path


def main(argv=None):
    """`python -m cholla_b200.snapshot -s <source dir> -o <output dir> -n <ranks> -c <outputs> [-d slice|proj]`: the
    command lines of python_scripts/concat_3d_data.py and concat_2d_data.py (concat_internals.common_cli: -c takes single
    numbers and inclusive ranges, e.g. `-c 0 4-6`)."""
    import argparse
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("-s", "--source-directory", required=True)
    ap.add_argument("-o", "--output-directory", required=True)
    ap.add_argument("-n", "--num-processes", type=int, required=True)
    ap.add_argument("-c", "--concat-outputs", nargs="+", required=True)
    ap.add_argument("-d", "--dataset-kind", choices=["slice", "proj"], default=None, help="2-D outputs; default: the 3-D snapshot")
    ap.add_argument("--skip-fields", nargs="*", default=[])
    a = ap.parse_args(argv)
    outputs = []
    for tok in a.concat_outputs:
        lo, _, hi = tok.partition("-")
        outputs.extend(range(int(lo), int(hi or lo) + 1))
    for n in outputs:
        if a.dataset_kind:
            print(concat_2d_dataset(a.source_directory, a.output_directory, a.num_processes, n, a.dataset_kind, a.skip_fields))
        else:
            print(concat_3d_dataset(a.source_directory, a.output_directory, a.num_processes, n, a.skip_fields))
    return 0


if __name__ == "__main__":
    import sys
    sys.exit(main())

