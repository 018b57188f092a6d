"""ctypes binding of libcholla_b200.so -- exactly the symbols include/cholla_b200.h declares."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
# CHOLLA_B200_LIB selects a tuning variant built with `make VARIANT=...` (development only)
LIB_PATH = os.environ.get("CHOLLA_B200_LIB") or os.path.join(_HERE, "libcholla_b200.so")
HEADER_PATH = os.path.join(_ROOT, "include", "cholla_b200.h")

CHL_OK = 0
PCM, PLMC, PPMC, PLMP, PPMP = 0, 1, 2, 3, 4
HLLC, ROE, HLLD, HLL, EXACT = 0, 1, 2, 3, 4
BC_PERIODIC, BC_REFLECTIVE, BC_TRANSMISSIVE, BC_CUSTOM, BC_MPI = 1, 2, 3, 4, 5


class ChollaB200Error(RuntimeError):
    pass


class Config(C.Structure):
    """struct chl_config (include/cholla_b200.h)."""
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("n_ghost", C.c_int), ("n_fields", C.c_int),
                ("recon", C.c_int), ("solver", C.c_int), ("precision", C.c_int), ("strict", C.c_int), ("mhd", C.c_int),
                ("gamma", C.c_double), ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
                ("density_floor", C.c_double), ("temperature_floor_U", C.c_double),
                ("n_scalars", C.c_int), ("dual_energy", C.c_int)]


_lib = None


def abi_symbols():
    """Every function name declared in include/cholla_b200.h."""
    with open(HEADER_PATH) as fh:
        text = fh.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(chl_[a-z0-9_]+)\s*\(", text)))


def lib():
    """Load the CUDA library.  Fails loudly if it has not been built -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ChollaB200Error("%s is missing: run `make -C cholla_b200/csrc` (or __graft_entry__.build())"
                              % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    i, d, p, l = C.c_int, C.c_double, C.c_void_p, C.c_long
    ip = C.POINTER(C.c_int)
    L.chl_abi_version.restype = i
    L.chl_last_error.restype = C.c_char_p
    L.chl_device_count.restype = i
    L.chl_has_precision.argtypes = [i]
    L.chl_create.argtypes = [C.POINTER(Config), C.POINTER(p)]
    L.chl_destroy.argtypes = [p]
    L.chl_scratch_bytes.argtypes = [p]
    L.chl_scratch_bytes.restype = C.c_size_t
    L.chl_vl_step.argtypes = [p, p, d, p]
    L.chl_vl_step_fused.argtypes = [p, p, d, p]
    L.chl_simple_step.argtypes = [p, p, d, p]
    L.chl_simple_step_fused.argtypes = [p, p, d, p]
    L.chl_dti_result.argtypes = [p, C.POINTER(d), p]
    L.chl_dti_device_ptr.argtypes = [p, C.POINTER(p)]
    L.chl_dti_decode.argtypes = [C.c_ulonglong]
    L.chl_dti_decode.restype = d
    L.chl_calc_dti.argtypes = [p, p, C.POINTER(d), p]
    L.chl_set_profiling.argtypes = [p, i]
    L.chl_stage_times.argtypes = [p, C.POINTER(d), ip]
    L.chl_apply_temperature_floor.argtypes = [p, p, d, p]
    L.chl_apply_scalar_floor.argtypes = [p, p, i, d, p]
    L.chl_cooling_update.argtypes = [p, p, d, p]
    L.chl_set_static_gravity.argtypes = [p, i, i, i, i, d, d, d, p]
    L.chl_set_gravity_field.argtypes = [p, p, p]
    L.chl_gravity_field_ptr.argtypes = [p, C.POINTER(p)]
    L.chl_copy_gravity_field.argtypes = [p, p, p]
    L.chl_set_gravity_potential.argtypes = [p, p, i]
    L.chl_set_ghost_cells.argtypes = [p, p, ip, i, p]
    L.chl_set_boundaries.argtypes = [p, p, ip, p]
    L.chl_set_ghost_cells_ext.argtypes = [p, i, i, i, i, i, i, ip, i, i, i, i, i, i, i, i, i, p]
    L.chl_pack_buffer.argtypes = [p, p, i, i, i, l, l, i, i, i, i, p]
    L.chl_pack_snapshot_field.argtypes = [p, p, i, i, i, i, i, i, i, i, i, i, p]
    L.chl_unpack_buffer.argtypes = [p, p, i, i, i, l, l, i, i, i, i, p]
    L.chl_boundary_extents.argtypes = [i, i, i, i, i, ip, ip]
    L.chl_tile_decomposition.argtypes = [i, i, i, i, ip]
    L.chl_local_extent.argtypes = [i, i, i, ip, ip]
    L.chl_debug_exact_div.argtypes = [p, p, p, p, l, p]
    L.chl_debug_exact_sqrt.argtypes = [p, p, p, l, p]
    L.chl_debug_exact_div_flagged.argtypes = [p, p, p, p, i, l, p]
    L.chl_bad_cells.argtypes = [p, C.POINTER(C.c_ulonglong), p]
    L.chl_set_tuning.argtypes = [p, i, i]
    L.chl_mhd_max_divergence.argtypes = [p, p, C.POINTER(d), p]
    L.chl_halo_blob_bytes.restype = C.c_size_t
    L.chl_halo_create.argtypes = [p, ip, i, C.POINTER(p)]
    L.chl_halo_destroy.argtypes = [p]
    L.chl_halo_export.argtypes = [p, p]
    L.chl_halo_connect.argtypes = [p, i, p]
    L.chl_halo_exchange.argtypes = [p, p, p]
    L.chl_halo_error.argtypes = [p, p]
    L.chl_vl_step_exchange.argtypes = [p, p, p, d, p]
    L.chl_dti_copy_key.argtypes = [p, p, p]
    _lib = L
    return L


def check(code, what="cholla_b200"):
    if code != CHL_OK:
        raise ChollaB200Error("%s failed (%d): %s" % (what, code, lib().chl_last_error().decode()))


def boundary_extents(nx, ny, nz, n_ghost, direction):
    lo, hi = (C.c_int * 3)(), (C.c_int * 3)()
    check(lib().chl_boundary_extents(nx, ny, nz, n_ghost, direction, lo, hi), "chl_boundary_extents")
    return list(lo), list(hi)


def tile_decomposition(nproc, nx_global, ny_global, nz_global):
    np_ = (C.c_int * 3)()
    check(lib().chl_tile_decomposition(nproc, nx_global, ny_global, nz_global, np_), "chl_tile_decomposition")
    return list(np_)


def local_extent(n_global, nproc_axis, coord):
    n, off = C.c_int(), C.c_int()
    check(lib().chl_local_extent(n_global, nproc_axis, coord, C.byref(n), C.byref(off)), "chl_local_extent")
    return n.value, off.value
