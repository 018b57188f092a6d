"""ctypes front end of oracle/_ref/libcholla_ref_<recon>_<solver>.so -- the reference's own, unmodified hydro
kernels compiled for sm_100 by oracle/ref_build/Makefile.  TEST/BENCH INFRASTRUCTURE ONLY.  Needs a GPU to run
(the reference has no CPU hydro path); the .so files are prebuilt in the build container and travel to the GPU
box, where /root/reference does not exist."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")
_NAMES = {0: "pcm", 1: "plmc", 2: "ppmc", 3: "plmp", 4: "ppmp"}
_SOLV = {0: "hllc", 1: "roe", 2: "hlld", 3: "hll", 4: "exact"}
_libs = {}


def lib_path(recon, solver, nosync=False, dropin=False, simple=False, variant=""):
    """dropin=True: the same driver linked against cholla_b200/csrc/cholla_shims.cpp + libcholla_b200.so instead of
    the reference's kernel objects (oracle/ref_build/Makefile, target `dropin`)."""
    sfx = "_simple" if simple else ""  # a -DSIMPLE build (Simple_Algorithm_*_CUDA) instead of -DVL
    sfx += variant  # "_de", "_s1", "_de_s1": -DDE / -DSCALAR -DBASIC_SCALAR builds (oracle/ref_build/Makefile: PASSIVE)
    if dropin:
        return os.path.join(REF_DIR, "libcholla_dropin_%s_%s%s.so" % (_NAMES[recon], _SOLV[solver], sfx))
    return os.path.join(REF_DIR, "libcholla_ref_%s_%s%s%s.so" % (_NAMES[recon], _SOLV[solver], sfx, "_nosync" if nosync else ""))


def available(recon, solver, nosync=False, dropin=False, simple=False, variant=""):
    return os.path.exists(lib_path(recon, solver, nosync, dropin, simple, variant))


def lib(recon, solver, nosync=False, dropin=False, simple=False, variant=""):
    key = (recon, solver, nosync, dropin, simple, variant)
    if key not in _libs:
        L = C.CDLL(lib_path(recon, solver, nosync, dropin, simple, variant))  # RTLD_LOCAL: several builds of the same symbols coexist
        i, d, p = C.c_int, C.c_double, C.c_void_p
        L.ref_build_flags.restype = C.c_char_p
        L.ref_set_gamma.argtypes = [d]
        L.ref_vl_step.argtypes = [p, i, i, i, i, d, d, d, d, i, d]
        L.ref_calc_dti.argtypes = [p, i, i, i, i, i, d, d, d]
        L.ref_calc_dti.restype = d
        L.ref_temperature_floor.argtypes = [p, i, i, i, i, i, d]
        L.ref_set_boundaries.argtypes = [p, i, i, i, i, i, p]
        L.ref_advance.argtypes = [p, i, i, i, i, d, d, d, p, i, d, d, d, p, p]
        L.ref_advance.restype = d
        L.ref_n_fields.restype = i
        L.ref_set_gravity.argtypes = [i, d, d, d]
        L.ref_set_potential.argtypes = [p]
        _libs[key] = L
    return _libs[key]


def advance(dev_tensor, dims, n_ghost, dx, dy, dz, gamma, recon, solver, flags, n_steps, cfl=0.3, density_floor=0.0,
            U_floor=0.0, nosync=False, timed=False, dropin=False, simple=False, variant="", custom_grav=0,
            bounds=(0.0, 0.0, 0.0), potential=None):
    """Run the reference's step loop on a torch CUDA float64 tensor (flat SoA, ghosts included), in place.
    Returns (dts, last_dti, ms or None).  One process may only use ONE grid shape per library: the reference
    keeps its scratch in globals sized at the first call (VL_3D_cuda.cu:58-133)."""
    L = lib(recon, solver, nosync, dropin, simple, variant)
    assert dev_tensor.numel() == L.ref_n_fields() * dims[0] * dims[1] * dims[2], "field count of this reference build"
    L.ref_set_gamma(gamma)
    L.ref_set_gravity(custom_grav, bounds[0], bounds[1], bounds[2])  # read by -DSTATIC_GRAV builds (variant "..._sg") only
    # -DGRAVITY builds ("..._grav5", "..._grav3"): a CUDA float64 tensor of n_cells, handed over as d_grav_potential
    L.ref_set_potential(C.c_void_p(potential.data_ptr()) if potential is not None else None)
    fl = (C.c_int * 6)(*flags)
    dts = (C.c_double * max(n_steps, 1))()
    ms = C.c_float(0.0)
    dti = L.ref_advance(C.c_void_p(dev_tensor.data_ptr()), dims[0], dims[1], dims[2], n_ghost, dx, dy, dz, fl, n_steps,
                        cfl, density_floor, U_floor, dts, C.byref(ms) if timed else None)
    return list(dts)[:n_steps], dti, (ms.value if timed else None)
