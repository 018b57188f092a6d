"""HDF5 snapshot layout without a GPU: the format writer / reader (cholla_b200/h5lite.py) and the
reference's snapshot structure on top of it (cholla_b200/snapshot.py; src/io/io.cpp:605-737, 1350-1497, 2508-2720).

The reader is pinned on a file that libhdf5 itself wrote: tests/golden/libhdf5_written_testdouble.mat is SciPy's
`scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat` (a MATLAB 7.3 file = HDF5 behind a 512-byte user block, BSD
licence), whose one dataset is 0 : pi/4 : 2 pi.  The writer is then checked against that reader, against the
byte-level encodings libhdf5 used in that file, and structurally against the format specification."""
import os
import struct
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import pytest

from cholla_b200 import h5lite, snapshot

HERE = os.path.dirname(os.path.abspath(__file__))
LIBHDF5_FILE = os.path.join(HERE, "golden", "libhdf5_written_testdouble.mat")


def test_reader_on_a_libhdf5_written_file():
    with h5lite.File(LIBHDF5_FILE) as f:
        assert f._base == 512 and f.group_leaf_k == 4 and f.group_internal_k == 16
        assert f.keys() == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.dtype("<f8") and d.attrs["MATLAB_class"] == "double"
        assert np.array_equal(d.read().ravel(), np.arange(9) * (np.pi / 4))
        with pytest.raises(KeyError):
            f["nope"]


def test_message_encodings_equal_libhdf5s():
    """Datatype / dataspace / attribute messages as libhdf5 encoded them in the pinned file, byte for byte."""
    with h5lite.File(LIBHDF5_FILE) as f:
        msgs = {t: bytes(b) for t, _fl, b in f._messages(f.root._links["testdouble"])}
    assert msgs[0x0003][:20] == h5lite._dtype_msg("<f8")
    assert msgs[0x0001] == h5lite._space_msg((9, 1))
    # big-endian variants differ in the byte-order bit only
    le, be = h5lite._dtype_msg("<f8"), h5lite._dtype_msg(">f8")
    assert be[1] == le[1] | 1 and be[:1] + be[2:] == le[:1] + le[2:]
    assert h5lite._dtype_msg(">i4") == bytes([0x10, 0x09, 0, 0, 4, 0, 0, 0, 0, 0, 32, 0])
    assert h5lite._dtype_msg(">f4")[8:] == struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)


def _example():
    rng = np.random.default_rng(3)
    attrs = OrderedDict()
    attrs["gamma"] = np.array([5.0 / 3.0], dtype=">f8")
    attrs["Git Commit Hash"] = h5lite.VarString("0123456789abcdef0123456789abcdef01234567")
    attrs["Macro Flags"] = h5lite.VarString("-DPRECISION=2 -DPPMC -DHLLC -DVL")
    attrs["cholla"] = h5lite.VarString("")
    attrs["n_step"] = np.array([17], dtype=">i4")
    attrs["dims"] = np.array([6, 5, 4], dtype=">i4")
    attrs["dx"] = np.array([0.5, 0.25, 0.125], dtype=">f8")
    data = OrderedDict()
    data["density"] = rng.random((6, 5, 4)).astype(">f8")
    data["momentum_x"] = rng.standard_normal((6, 5, 4)).astype(">f8")
    data["Energy"] = rng.random((6, 5, 4)).astype(">f4")
    data["counts"] = rng.integers(-5, 5, (7,)).astype("<i4")
    data["empty"] = np.zeros((0,), dtype=">f8")
    return attrs, data


def test_write_read_round_trip(tmp_path):
    attrs, data = _example()
    p = str(tmp_path / "a.h5")
    eof = h5lite.write(p, attrs, data)
    assert os.path.getsize(p) == eof
    with h5lite.File(p) as f:
        assert list(f.attrs) == list(attrs)  # creation order
        assert f.attrs["Git Commit Hash"] == str(attrs["Git Commit Hash"]) and f.attrs["cholla"] == ""
        assert f.attrs["Macro Flags"] == str(attrs["Macro Flags"])
        assert f.attrs["gamma"].shape == (1,) and f.attrs["gamma"][0] == 5.0 / 3.0
        assert f.attrs["n_step"].dtype.kind == "i" and f.attrs["n_step"][0] == 17
        assert f.attrs["dims"].tolist() == [6, 5, 4] and f.attrs["dx"].tolist() == [0.5, 0.25, 0.125]
        assert f.keys() == sorted(data)  # symbol table order
        for k, v in data.items():
            d = f[k]
            assert d.shape == v.shape and d.dtype == v.dtype
            assert np.array_equal(d.read(), v)
        assert f["density"][2, 3, 1] == data["density"][2, 3, 1]
        assert "density" in f and "Density" not in f


def test_written_file_structure(tmp_path):
    """The structures libhdf5 walks when it opens the file: superblock v0, root symbol-table entry, B-tree / heap /
    symbol node signatures and sizes, name-sorted entries, everything inside the end-of-file address, 8-byte aligned
    raw data stored big-endian."""
    attrs, data = _example()
    p = str(tmp_path / "a.h5")
    h5lite.write(p, attrs, data)
    b = open(p, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])
    leaf_k, int_k, flags = struct.unpack_from("<HHI", b, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert (base, free, eof, drv) == (0, h5lite.UNDEF, len(b), h5lite.UNDEF) and flags == 0
    name_off, root, cache, _r, btree, heap = struct.unpack_from("<QQIIQQ", b, 56)
    assert name_off == 0 and root == 96 and cache == 1
    ver, _res, nmsgs, refs, hsize = struct.unpack_from("<BBHII", b, root)
    assert ver == 1 and refs == 1 and nmsgs == 1 + len(attrs) and hsize % 8 == 0 and root + 16 + hsize == btree
    # every message is 8-byte aligned and the first is the symbol-table message pointing at the same B-tree and heap
    t, sz, fl = struct.unpack_from("<HHB", b, root + 16)
    assert (t, sz) == (0x0011, 16) and struct.unpack_from("<QQ", b, root + 24) == (btree, heap)
    assert b[btree:btree + 4] == b"TREE" and b[btree + 4] == 0 and b[btree + 5] == 0
    used, left, right = struct.unpack_from("<HQQ", b, btree + 6)
    assert used == 1 and left == right == h5lite.UNDEF
    key0, child, key1 = struct.unpack_from("<QQQ", b, btree + 24)
    assert heap == btree + 24 + 2 * int_k * 8 + (2 * int_k + 1) * 8  # full-size node
    assert b[heap:heap + 4] == b"HEAP"
    dsize, free_head, daddr = struct.unpack_from("<QQQ", b, heap + 8)
    assert free_head == 1 and daddr == heap + 32 and dsize % 8 == 0
    seg = b[daddr:daddr + dsize]
    assert seg[:8] == bytes(8) and key0 == 0
    assert b[child:child + 4] == b"SNOD" and child == daddr + dsize
    nsym, = struct.unpack_from("<H", b, child + 6)
    names = []
    for j in range(nsym):
        noff, oaddr, ctype = struct.unpack_from("<QQI", b, child + 8 + 40 * j)
        assert noff % 8 == 0 and ctype == 0 and oaddr < eof and b[oaddr] == 1
        names.append(seg[noff:seg.index(b"\0", noff)])
    assert names == sorted(names) and names == sorted(k.encode() for k in data)
    assert seg[key1:seg.index(b"\0", key1)] == names[-1]  # right key = greatest name of the node
    gcol = child + 8 + 2 * leaf_k * 40
    assert b[gcol:gcol + 4] == b"GCOL" and struct.unpack_from("<Q", b, gcol + 8)[0] == 4096
    idx, ref, osz = struct.unpack_from("<HH4xQ", b, gcol + 16)
    assert (idx, ref, osz) == (1, 1, 40) and b[gcol + 32:gcol + 72] == str(attrs["Git Commit Hash"]).encode()
    # raw data: contiguous, aligned, big-endian on disk
    raw = data["density"].tobytes()
    at = b.index(raw)
    assert at % 8 == 0 and struct.unpack_from(">d", b, at)[0] == float(data["density"][0, 0, 0])


def test_reader_rejects_what_it_does_not_understand(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not an hdf5 file" * 40)
    with pytest.raises(h5lite.H5Error):
        h5lite.File(str(p))
    attrs, data = _example()
    q = str(tmp_path / "a.h5")
    h5lite.write(q, attrs, data)
    b = open(q, "rb").read()
    (tmp_path / "t.h5").write_bytes(b[:-64])  # truncated: end-of-file address beyond the file
    with pytest.raises(h5lite.H5Error):
        h5lite.File(str(tmp_path / "t.h5"))
    with pytest.raises(h5lite.H5Error):
        h5lite.write(q, {}, {"s": np.array(["a"])})


def _header(n_real, g, n_fields):
    nx, ny, nz = n_real
    H = SimpleNamespace(nx_real=nx, ny_real=ny, nz_real=nz, n_ghost=g, n_fields=n_fields, nx=nx + 2 * g,
                        ny=ny + 2 * g if ny > 1 else 1, nz=nz + 2 * g if nz > 1 else 1, dx=1.0 / nx, dy=1.0 / max(ny, 1),
                        dz=1.0 / max(nz, 1), t=0.125, dt=0.001, n_step=42)
    return H


@pytest.mark.parametrize("n_real,mhd", [((7, 5, 4), False), ((6, 5, 4), True), ((9, 6, 1), False), ((11, 1, 1), False)])
def test_snapshot_layout_and_restart(tmp_path, n_real, mhd):
    """Write_Header_HDF5 + Write_Grid_HDF5 layout, and Read_Grid_HDF5 putting it back."""
    g, nf = (5 if mhd else 4), (8 if mhd else 5)
    H = _header(n_real, g, nf)
    C = np.random.default_rng(1).standard_normal((nf, H.nz, H.ny, H.nx))
    p = str(tmp_path / "0.h5.0")
    snapshot.write_snapshot(p, H, C, 1.4, n_real, n_real, (0, 0, 0), (1, 1, 1), (0.0, -0.5, 0.25), (1.0, 2.0, 3.0),
                            mhd=mhd, macro=snapshot.macro_flags(2, 2 if mhd else 0, 2, mhd))
    ndim = 1 + (n_real[1] > 1) + (n_real[2] > 1)
    gy, gz = (g if n_real[1] > 1 else 0), (g if n_real[2] > 1 else 0)
    with h5lite.File(p) as f:
        want = ["gamma", "Git Commit Hash", "Macro Flags", "cholla", "t", "dt", "n_step", "n_fields", "time_unit",
                "length_unit", "mass_unit", "velocity_unit", "density_unit", "energy_unit"]
        want += ["magnetic_field_unit"] if mhd else []
        want += ["dims", "dims_local", "offset", "nprocs", "bounds", "domain", "dx"]
        assert list(f.attrs) == want  # io.cpp:605-737, creation order
        assert f.attrs["gamma"][0] == 1.4 and f.attrs["t"][0] == 0.125 and f.attrs["n_step"][0] == 42
        assert f.attrs["n_fields"][0] == nf and f.attrs["dims"].tolist() == list(n_real)
        assert f.attrs["bounds"].tolist() == [0.0, -0.5, 0.25] and f.attrs["domain"].tolist() == [1.0, 2.0, 3.0]
        assert f.attrs["velocity_unit"][0] == 3.08567758e21 / 3.15569e10 and ("-DPPMC" in f.attrs["Macro Flags"])
        names = ["density", "momentum_x", "momentum_y", "momentum_z", "Energy"]
        assert sorted(f.keys()) == sorted(names + (["magnetic_x", "magnetic_y", "magnetic_z"] if mhd else []))
        for fi, nm in enumerate(names):
            d = f[nm]
            assert d.dtype == np.dtype(">f8") and d.shape == tuple(n_real[:ndim])
            a = d.read().reshape(n_real)
            for (i, j, k) in [(0, 0, 0), (n_real[0] - 1, n_real[1] - 1, n_real[2] - 1), (2, n_real[1] // 2, n_real[2] // 3)]:
                assert a[i, j, k] == C[fi, k + gz, j + gy, i + g]  # io_gpu.cu:62-66
        if mhd:
            bx, by, bz = f["magnetic_x"].read(), f["magnetic_y"].read(), f["magnetic_z"].read()
            assert bx.shape == (7, 5, 4) and by.shape == (6, 6, 4) and bz.shape == (6, 5, 5)
            assert bx[0, 1, 2] == C[5, 2 + g, 1 + g, g - 1] and bx[6, 4, 3] == C[5, 3 + g, 4 + g, g + 5]
            assert by[1, 0, 2] == C[6, 2 + g, g - 1, 1 + g] and bz[1, 2, 0] == C[7, g - 1, 2 + g, 1 + g]
    back, gamma, t, n_step = snapshot.read_grid(p, H, mhd)
    assert (gamma, t, n_step) == (1.4, 0.125, 42)
    real = (slice(None), slice(gz, gz + n_real[2]), slice(gy, gy + n_real[1]), slice(g, g + n_real[0]))
    assert np.array_equal(back[real][:5], C[real][:5])
    if mhd:
        assert np.array_equal(back[5, g:-g, g:-g, g - 1:-g], C[5, g:-g, g:-g, g - 1:-g])
        assert np.array_equal(back[7, g - 1:-g, g:-g, g:-g], C[7, g - 1:-g, g:-g, g:-g])
    mask = np.ones_like(back, dtype=bool)
    mask[real] = False
    assert not back[:5][mask[:5]].any()  # ghost cells are left for Set_Boundary_Conditions
    H2 = _header((n_real[0] + 1,) + tuple(n_real[1:]), g, nf)
    with pytest.raises(ValueError):
        snapshot.read_grid(p, H2, mhd)


def test_file_names():
    """FnameTemplate, io.cpp:2779-2828."""
    assert snapshot.format_fname("out", 3, 2) == "out/3//3.h5.2"
    assert snapshot.effective_output_dir_path("", 3) == "./"
    assert snapshot.format_fname("run/prefix_", 3, 0, legacy_flat_outdir=True) == "run/prefix_3.h5.0"
    assert snapshot.effective_output_dir_path("run/prefix_", 3, True) == "run/"
    assert snapshot.format_fname("out", 7, 1, ".float32") == "out/7//7.float32.h5.1"


def test_concat_of_rank_files(tmp_path):
    """python_scripts/concat_3d_data.py on a 2 x 2 x 1 decomposition (hydro + face-centred magnetic fields)."""
    from cholla_b200 import _lib
    n_glob, g, nf = (8, 6, 4), 5, 8
    rng = np.random.default_rng(5)
    cells = rng.standard_normal((5, n_glob[2], n_glob[1], n_glob[0]))
    faces = [rng.standard_normal((n_glob[2] + (d == 2), n_glob[1] + (d == 1), n_glob[0] + (d == 0))) for d in range(3)]
    np_grid = _lib.tile_decomposition(4, *n_glob)
    assert list(np_grid) == [2, 2, 1]
    src = tmp_path / "raw"
    for rank in range(4):
        cx, cy = rank % 2, rank // 2
        ext = [_lib.local_extent(n_glob[a], np_grid[a], c) for a, c in zip(range(3), (cx, cy, 0))]
        nloc, off = [e[0] for e in ext], [e[1] for e in ext]
        H = _header(nloc, g, nf)
        C = np.zeros((nf, H.nz, H.ny, H.nx))
        C[:5, g:-g, g:-g, g:-g] = cells[:, off[2]:off[2] + nloc[2], off[1]:off[1] + nloc[1], off[0]:off[0] + nloc[0]]
        for d in range(3):
            s = [(d == 0), (d == 1), (d == 2)]
            C[5 + d, g - s[2]:g + nloc[2], g - s[1]:g + nloc[1], g - s[0]:g + nloc[0]] = \
                faces[d][off[2]:off[2] + nloc[2] + s[2], off[1]:off[1] + nloc[1] + s[1], off[0]:off[0] + nloc[0] + s[0]]
        os.makedirs(src / "2", exist_ok=True)
        snapshot.write_snapshot(str(src / "2" / ("2.h5.%d" % rank)), H, C, 5.0 / 3.0, n_glob, nloc, off, np_grid,
                                (0, 0, 0), (1, 1, 1), mhd=True)
    out = snapshot.concat_3d_dataset(str(src), str(tmp_path / "cat"), 4, 2)
    assert os.path.basename(out) == "2.h5"
    with h5lite.File(out) as f:
        assert "dims_local" not in f.attrs and "offset" not in f.attrs and f.attrs["nprocs"].tolist() == [2, 2, 1]
        for i, nm in enumerate(snapshot.HYDRO_FIELDS):
            assert np.array_equal(f[nm].read(), cells[i].transpose(2, 1, 0))
        for d, nm in enumerate(snapshot.MAGNETIC_FIELDS):
            assert np.array_equal(f[nm].read(), faces[d].transpose(2, 1, 0))


def test_float32_output_selection_and_mhd_format():
    """Output_Float32 (io.cpp:266-375): only the selected fields, big-endian floats; the MHD datasets cover (n+1)^3 cells
    from one below the real cells on every axis."""
    n_real, g = (6, 5, 4), 5
    H = _header(n_real, g, 8)
    C = np.random.default_rng(2).standard_normal((8, H.nz, H.ny, H.nx))
    P = SimpleNamespace(out_float32_density=1, out_float32_momentum_x=0, out_float32_momentum_y=0, out_float32_momentum_z=0,
                        out_float32_Energy=1, out_float32_magnetic_x=0, out_float32_magnetic_y=1, out_float32_magnetic_z=0)
    d = snapshot.float32_datasets(H, C, P, mhd=True)
    assert list(d) == ["density", "Energy", "magnetic_y"] and all(v.dtype == np.dtype(">f4") for v in d.values())
    assert d["density"].shape == n_real and d["magnetic_y"].shape == (7, 6, 5)
    assert d["Energy"][2, 3, 1] == np.float32(C[4, 1 + g, 3 + g, 2 + g])
    assert d["magnetic_y"][0, 0, 0] == np.float32(C[6, g - 1, g - 1, g - 1]) and d["magnetic_y"][6, 5, 4] == np.float32(C[6, g + 3, g + 4, g + 5])
    assert snapshot.format_fname("out", 4, 0, ".float32") == "out/4//4.float32.h5.0"


@pytest.mark.parametrize("mhd", [False, True])
def test_slice_datasets(mhd):
    """Write_Slices_HDF5 (io.cpp:1835-2250): global mid-planes, [x][y] / [x][z] / [y][z] order, zeros on ranks the plane
    misses, cell-centred magnetic fields."""
    g, nf = (5 if mhd else 4), (8 if mhd else 5)
    n_glob, n_loc, off = (12, 10, 8), (6, 10, 8), (6, 0, 0)  # the upper-x rank of a 2 x 1 x 1 decomposition
    H = _header(n_loc, g, nf)
    C = np.random.default_rng(4).standard_normal((nf, H.nz, H.ny, H.nx))
    get = lambda f, ks, js, is_: C[f, ks, js, is_]
    d = snapshot.slice_datasets(H, get, n_glob, off, mhd)
    want = ["%s_%s" % (a, p) for p in ("xy", "xz", "yz") for a in (["d", "mx", "my", "mz", "E"] + (["magnetic_x", "magnetic_y", "magnetic_z"] if mhd else []))]
    assert list(d) == want
    assert d["d_xy"].shape == (6, 10) and d["E_xz"].shape == (6, 8) and d["mx_yz"].shape == (10, 8)
    assert d["d_xy"][2, 3] == C[0, 4 + g, 3 + g, 2 + g]        # z = 8 / 2 = 4
    assert d["E_xz"][1, 7] == C[4, 7 + g, 5 + g, 1 + g]        # y = 10 / 2 = 5
    assert d["mx_yz"][9, 0] == C[1, g, 9 + g, 0 + g]           # x = 12 / 2 = 6 -> local 0 on this rank
    lower = snapshot.slice_datasets(H, get, n_glob, (0, 0, 0), mhd)  # the lower-x rank: the yz plane is not its own
    assert not lower["d_yz"].any() and lower["d_xy"].any()
    if mhd:
        assert d["magnetic_x_xy"][2, 3] == 0.5 * (C[5, 4 + g, 3 + g, 2 + g] + C[5, 4 + g, 3 + g, 1 + g])
        assert d["magnetic_z_yz"][4, 2] == 0.5 * (C[7, 2 + g, 4 + g, g] + C[7, 1 + g, 4 + g, g])


@pytest.mark.parametrize("mhd", [False, True])
def test_projection_datasets_numpy_and_torch(mhd):
    """Write_Projection_HDF5 (io.cpp:1502-1660) restated with array arithmetic: checked against the reference's loops
    written out in Python, and the torch path (what runs on the device) against the numpy path."""
    import torch
    g, nf, n_real, gamma = (5 if mhd else 4), (8 if mhd else 5), (5, 4, 3), 5.0 / 3.0
    H = _header(n_real, g, nf)
    rng = np.random.default_rng(6)
    C = rng.standard_normal((nf, H.nz, H.ny, H.nx)) * 0.1
    C[0] = 1.0 + rng.random(C[0].shape)
    C[4] = 3.0 + rng.random(C[0].shape)
    d = snapshot.projection_datasets(H, C, gamma, mhd)
    assert list(d) == ["d_xy", "d_xz", "T_xy", "T_xz"] and d["d_xy"].shape == (5, 4) and d["T_xz"].shape == (5, 3)
    mu = 0.6
    for (i, j) in [(0, 0), (4, 3), (2, 1)]:
        dxy = Txy = 0.0
        for k in range(n_real[2]):
            z, y, x = k + g, j + g, i + g
            rho = C[0, z, y, x]
            e = C[4, z, y, x] - 0.5 * (C[1, z, y, x] ** 2 + C[2, z, y, x] ** 2 + C[3, z, y, x] ** 2) / rho
            if mhd:
                bx = 0.5 * (C[5, z, y, x] + C[5, z, y, x - 1]); by = 0.5 * (C[6, z, y, x] + C[6, z, y - 1, x])
                bz = 0.5 * (C[7, z, y, x] + C[7, z - 1, y, x])
                e -= 0.5 * (bx * bx + by * by + bz * bz)
            P = max(e * (gamma - 1.0), 1e-20)
            n = rho * snapshot.DENSITY_UNIT / (mu * snapshot.MP)
            T = P * snapshot.PRESSURE_UNIT / (n * snapshot.KB)
            dxy += rho * H.dz
            Txy += T * rho * H.dz
        assert np.isclose(d["d_xy"][i, j], dxy, rtol=1e-14) and np.isclose(d["T_xy"][i, j], Txy, rtol=1e-13)
    t = snapshot.projection_datasets(H, torch.from_numpy(C), gamma, mhd)
    for k in d:
        assert np.allclose(t[k].astype(np.float64), d[k].astype(np.float64), rtol=1e-14)


def test_concat_2d_slices_and_projections(tmp_path):
    """python_scripts/concat_2d_data.py: slice and projection files of a 2 x 2 x 1 decomposition back to the planes of the
    single-domain state (projections of ranks along the line of sight add up)."""
    from cholla_b200 import _lib
    n_glob, g, nf, gamma = (8, 6, 4), 4, 5, 5.0 / 3.0
    rng = np.random.default_rng(8)
    cells = rng.random((5, n_glob[2], n_glob[1], n_glob[0])) + 0.5
    cells[4] += 3.0
    np_grid = _lib.tile_decomposition(4, *n_glob)
    Hg = _header(n_glob, g, nf)
    Cg = np.zeros((nf, Hg.nz, Hg.ny, Hg.nx))
    Cg[:, g:-g, g:-g, g:-g] = cells
    whole_slices = snapshot.slice_datasets(Hg, lambda f, ks, js, is_: Cg[f, ks, js, is_], n_glob, (0, 0, 0))
    whole_proj = snapshot.projection_datasets(Hg, Cg, gamma)
    for rank in range(4):
        cx, cy = rank % 2, rank // 2
        ext = [_lib.local_extent(n_glob[a], np_grid[a], c) for a, c in zip(range(3), (cx, cy, 0))]
        nloc, off = [e[0] for e in ext], [e[1] for e in ext]
        H = _header(nloc, g, nf)
        H.dx, H.dy, H.dz = Hg.dx, Hg.dy, Hg.dz
        C = np.zeros((nf, H.nz, H.ny, H.nx))
        C[:, g:-g, g:-g, g:-g] = cells[:, off[2]:off[2] + nloc[2], off[1]:off[1] + nloc[1], off[0]:off[0] + nloc[0]]
        attrs = snapshot.header_attributes(H, gamma, n_glob, nloc, off, np_grid, (0, 0, 0), (1, 1, 1))
        os.makedirs(tmp_path / "raw" / "3", exist_ok=True)
        h5lite.write(str(tmp_path / "raw" / "3" / ("3_slice.h5.%d" % rank)), attrs,
                     snapshot.slice_datasets(H, lambda f, ks, js, is_, C=C: C[f, ks, js, is_], n_glob, off))
        h5lite.write(str(tmp_path / "raw" / "3" / ("3_proj.h5.%d" % rank)), attrs, snapshot.projection_datasets(H, C, gamma))
    assert snapshot.main(["-s", str(tmp_path / "raw"), "-o", str(tmp_path / "cli"), "-n", "4", "-c", "3-3", "-d", "proj"]) == 0
    assert os.path.exists(tmp_path / "cli" / "3_proj.h5")
    s = snapshot.concat_2d_dataset(str(tmp_path / "raw"), str(tmp_path / "cat"), 4, 3, "slice")
    p = snapshot.concat_2d_dataset(str(tmp_path / "raw"), str(tmp_path / "cat"), 4, 3, "proj")
    assert os.path.basename(s) == "3_slice.h5" and os.path.basename(p) == "3_proj.h5"
    with h5lite.File(s) as f:
        assert "offset" not in f.attrs
        for nm, want in whole_slices.items():
            assert np.array_equal(f[nm].read(), want), nm
    with h5lite.File(p) as f:
        for nm, want in whole_proj.items():
            assert np.allclose(f[nm].read(), want, rtol=1e-13), nm
