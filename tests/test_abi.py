"""CPU-side checks of the drop-in boundary: the library loads, exports every declared symbol, refuses to
compute without a GPU, and its host-only helpers mirror the reference's host logic."""
import ctypes as C

import pytest

import cholla_b200 as cb
import oracle
from cholla_b200 import _lib
from cholla_b200.grid import Decomposition, halo_slab


def test_library_exports_every_declared_symbol():
    L = cb.lib()
    names = cb.abi_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "libcholla_b200.so does not export %s" % n
    assert L.chl_abi_version() == 2
    assert L.chl_has_precision(2) == 1


def test_no_silent_cpu_fallback():
    L = cb.lib()
    if L.chl_device_count() > 0:
        pytest.skip("a GPU is present; the refusal path is for CPU-only boxes")
    cfg = cb.Config(40, 40, 40, 4, 5, cb.PPMC, cb.HLLC, 2, 0, 0, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0)
    ctx = C.c_void_p()
    rc = L.chl_create(C.byref(cfg), C.byref(ctx))
    assert rc == -3 and not ctx.value
    assert b"no CUDA device" in L.chl_last_error()


def test_config_validation():
    L = cb.lib()
    ctx = C.c_void_p()
    bad = cb.Config(40, 40, 40, 2, 5, cb.PPMC, cb.HLLC, 2, 0, 0, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0)  # too few ghosts
    assert L.chl_create(C.byref(bad), C.byref(ctx)) == -1
    bad = cb.Config(40, 40, 40, 4, 8, cb.PPMC, cb.HLLC, 2, 0, 0, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0)  # 8 fields = MHD
    assert L.chl_create(C.byref(bad), C.byref(ctx)) == -1  # a hydro grid has 5 + n_scalars + dual_energy fields
    # -DSCALAR / -DDE contexts: the field count must match, only the built combinations exist, MHD carries none
    bad = cb.Config(40, 40, 40, 4, 6, cb.PPMC, cb.HLLC, 2, 1, 0, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0, 1, 1)
    assert L.chl_create(C.byref(bad), C.byref(ctx)) == -1
    bad = cb.Config(40, 40, 40, 4, 7, cb.PPMC, cb.HLLC, 2, 1, 0, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0, 2, 0)  # two scalars
    assert L.chl_create(C.byref(bad), C.byref(ctx)) == -4
    bad = cb.Config(40, 40, 40, 4, 6, cb.PPMP, cb.ROE, 2, 1, 0, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0, 0, 1)  # PPMP: HLLC only
    assert L.chl_create(C.byref(bad), C.byref(ctx)) == -4
    bad = cb.Config(40, 40, 40, 5, 8, cb.PPMC, cb.HLLD, 2, 1, 1, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0, 0, 1)  # MHD + DE
    assert L.chl_create(C.byref(bad), C.byref(ctx)) == -4
    # the reference's own MHD constraints (src/utils/error_handling.cpp:83-118): HLLD only, 3-D only, one more ghost
    bad = cb.Config(40, 40, 40, 5, 8, cb.PPMC, cb.HLLC, 2, 0, 1, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0)
    assert L.chl_create(C.byref(bad), C.byref(ctx)) == -1
    bad = cb.Config(40, 40, 1, 5, 8, cb.PPMC, cb.HLLD, 2, 0, 1, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0)
    assert L.chl_create(C.byref(bad), C.byref(ctx)) == -1
    bad = cb.Config(40, 40, 40, 4, 8, cb.PPMC, cb.HLLD, 2, 0, 1, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0)
    assert L.chl_create(C.byref(bad), C.byref(ctx)) == -1
    bad = cb.Config(40, 1, 40, 4, 5, cb.PPMC, cb.HLLC, 2, 0, 0, 5.0 / 3.0, 0.1, 0.1, 0.1, 0.0, 0.0)
    assert L.chl_create(C.byref(bad), C.byref(ctx)) == -1


@pytest.mark.parametrize("dims", [(20, 1, 1), (20, 18, 1), (20, 18, 16)])
def test_boundary_extents_match_oracle(dims):
    nx, ny, nz = dims
    ndim = 1 + (ny > 1) + (nz > 1)
    for d in range(2 * ndim):
        assert tuple(cb.boundary_extents(nx, ny, nz, 3, d)) == tuple(oracle.boundary_extents(d, nx, ny, nz, 3))


def test_tile_decomposition_reference_cases():
    # SURVEY 2.3 / mpi_routines.cpp:852-930: 8 -> 2x2x2, 4 -> 2x2x1, 2 -> 2x1x1, sorted np_x >= np_y >= np_z
    assert cb.tile_decomposition(1, 64, 64, 64) == [1, 1, 1]
    assert cb.tile_decomposition(2, 64, 64, 64) == [2, 1, 1]
    assert cb.tile_decomposition(4, 64, 64, 64) == [2, 2, 1]
    assert cb.tile_decomposition(8, 64, 64, 64) == [2, 2, 2]
    assert cb.tile_decomposition(6, 64, 64, 64) == [3, 2, 1]
    assert cb.tile_decomposition(12, 64, 64, 64) == [3, 2, 2]
    assert cb.tile_decomposition(4, 64, 1, 1) == [4, 1, 1]
    assert cb.tile_decomposition(6, 64, 64, 1) == [3, 2, 1]
    assert cb.tile_decomposition(2, 32, 64, 128) == [1, 1, 2]  # nz_global > nx_global swaps x and z


def test_local_extent_uneven_split():
    # mpi_routines.cpp:373-415: the first n % np ranks get one extra cell
    got = [cb.local_extent(10, 4, c) for c in range(4)]
    assert got == [(3, 0), (3, 3), (2, 6), (2, 8)]
    assert [cb.local_extent(8, 2, c) for c in range(2)] == [(4, 0), (4, 4)]


def test_decomposition_flags_and_neighbours():
    # periodic global faces become exchange faces when the axis is split; unsplit axes keep their flag
    d = Decomposition.build(4, 3, (64, 64, 64), [1, 1, 3, 3, 2, 2])
    assert d.np == [2, 2, 1] and d.coords == [1, 1, 0]
    assert d.flags == [5, 5, 5, 3, 2, 2]
    assert d.dest == [2, 2, 1, 1, 3, 3]
    assert d.n_local == [32, 32, 64] and d.offset == [32, 32, 0]
    d0 = Decomposition.build(4, 0, (64, 64, 64), [1, 1, 3, 3, 2, 2])
    assert d0.flags == [5, 5, 3, 5, 2, 2]


def test_halo_slab_geometry_matches_reference_offsets():
    nx, ny, nz, g = 20, 18, 16, 4
    # Load_Hydro_DeviceBuffer_X0 (mpi_boundaries.cpp:113-131)
    s = halo_slab(nx, ny, nz, g, 0, recv=False)
    assert (s.idxoffset, s.isize, s.jsize, s.ksize) == (g + g * nx + g * nx * ny, g, ny - 2 * g, nz - 2 * g)
    # Load_Hydro_DeviceBuffer_Y1
    s = halo_slab(nx, ny, nz, g, 3, recv=False)
    assert (s.idxoffset, s.isize, s.jsize, s.ksize) == ((ny - 2 * g) * nx + g * nx * ny, nx, g, nz - 2 * g)
    # Unload_Hydro_DeviceBuffer_Z1
    s = halo_slab(nx, ny, nz, g, 5, recv=True)
    assert (s.idxoffset, s.isize, s.jsize, s.ksize) == ((nz - g) * nx * ny, nx, ny, g)
    # Unload_Hydro_DeviceBuffer_X1, 2-D
    s = halo_slab(nx, ny, 1, g, 1, recv=True)
    assert (s.idxoffset, s.isize, s.jsize, s.ksize) == (nx - g + g * nx, g, ny - 2 * g, 1)


def test_reference_builds_resolve_every_symbol():
    """oracle/_ref/*.so (the reference's own kernels behind ref_driver.cu, and the drop-in builds) must load with
    immediate binding: an unresolved symbol (a host function of the reference the driver forgot to supply) would
    otherwise only show on the GPU box."""
    import ctypes
    import glob
    import os
    import pytest
    libs = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "*.so")))
    if not libs:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    for f in libs:
        L = ctypes.CDLL(f, mode=os.RTLD_NOW)
        assert hasattr(L, "ref_advance") and hasattr(L, "ref_build_flags"), f
