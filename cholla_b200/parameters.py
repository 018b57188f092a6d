"""Parameter-file reader: the reference's `Parse_Params` / `Parse_Param` (src/global/global.cpp:127-470) for the
keys the Van Leer hot path and its initial conditions read, so that the decks shipped under the reference's
`examples/` drive this library unmodified.

Same rules as the reference:
  * a line whose FIRST character is a newline, '#' or ';' is skipped (global.cpp:146-149);
  * `name=value`, split at the first two '='-separated tokens (strtok, :153-165); the value is trimmed of
    surrounding white space (`Trim`, :88-101), the name is taken verbatim;
  * integers go through `atoi`, reals through `atof` (leading numeric prefix, 0 when there is none);
  * command-line `name=value` arguments override the file (:171-187);
  * an unknown name is reported once and ignored unless it is one of the reference's `optionalParams` (:104-108).
Host-only, no GPU.  Compile-time choices of the reference's `make.type.*` files (reconstruction, Riemann solver,
MHD, precision) are not in a deck; `run.py` takes them as options, mirroring `chl_config`.
"""
import re
import sys
from dataclasses import dataclass, field, fields
from typing import Dict, List

_INT = re.compile(r"\s*[+-]?\d+")
_FLOAT = re.compile(r"\s*[+-]?(?:(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?|inf(?:inity)?|nan)", re.I)


def atoi(s):
    """C atoi: optional white space, sign, digits; anything else ends the number; no digits -> 0."""
    m = _INT.match(s)
    return int(m.group(0)) if m else 0


def atof(s):
    """C atof (strtod): the longest numeric prefix; none -> 0.0."""
    m = _FLOAT.match(s)
    return float(m.group(0)) if m else 0.0


@dataclass
class Parameters:
    """The subset of `struct Parameters` (src/global/global.h:183-330) this path reads; defaults as in the struct."""
    nx: int = 0
    ny: int = 0
    nz: int = 0
    tout: float = 0.0
    outstep: float = 0.0
    n_steps_output: int = 0
    gamma: float = 0.0
    init: str = ""
    nfile: int = 0
    n_hydro: int = 1
    n_out_float32: int = 0
    n_slice: int = 0
    n_projection: int = 0
    out_float32_density: int = 0
    out_float32_momentum_x: int = 0
    out_float32_momentum_y: int = 0
    out_float32_momentum_z: int = 0
    out_float32_Energy: int = 0
    out_float32_magnetic_x: int = 0
    out_float32_magnetic_y: int = 0
    out_float32_magnetic_z: int = 0
    output_always: int = 0
    xmin: float = 0.0
    ymin: float = 0.0
    zmin: float = 0.0
    xlen: float = 0.0
    ylen: float = 0.0
    zlen: float = 0.0
    xl_bcnd: int = 0
    xu_bcnd: int = 0
    yl_bcnd: int = 0
    yu_bcnd: int = 0
    zl_bcnd: int = 0
    zu_bcnd: int = 0
    custom_bcnd: str = ""
    legacy_flat_outdir: int = 0
    outdir: str = ""
    indir: str = ""
    rho: float = 0.0
    vx: float = 0.0
    vy: float = 0.0
    vz: float = 0.0
    P: float = 0.0
    A: float = 0.0
    Bx: float = 0.0
    By: float = 0.0
    Bz: float = 0.0
    rho_l: float = 0.0
    vx_l: float = 0.0
    vy_l: float = 0.0
    vz_l: float = 0.0
    P_l: float = 0.0
    Bx_l: float = 0.0
    By_l: float = 0.0
    Bz_l: float = 0.0
    rho_r: float = 0.0
    vx_r: float = 0.0
    vy_r: float = 0.0
    vz_r: float = 0.0
    P_r: float = 0.0
    Bx_r: float = 0.0
    By_r: float = 0.0
    Bz_r: float = 0.0
    diaph: float = 0.0
    radius: float = 0.0
    P_blast: float = 0.0
    wave_length: float = 1.0
    pitch: float = 0.0
    yaw: float = 0.0
    polarization: float = 0.0
    rEigenVec_rho: float = 0.0
    rEigenVec_MomentumX: float = 0.0
    rEigenVec_MomentumY: float = 0.0
    rEigenVec_MomentumZ: float = 0.0
    rEigenVec_E: float = 0.0
    rEigenVec_Bx: float = 0.0
    rEigenVec_By: float = 0.0
    rEigenVec_Bz: float = 0.0
    prng_seed: int = 0
    temperature_floor: float = 0.0
    density_floor: float = 0.0
    scalar_floor: float = 0.0
    custom_grav: int = 0  # global.h:210: selects the analytic field of a -DSTATIC_GRAV build (static_grav.h)
    n_proc_x: int = 0
    n_proc_y: int = 0
    n_proc_z: int = 0
    unknown: Dict[str, str] = field(default_factory=dict)  # names Parse_Param would warn about

    def boundary_flags(self):
        """flags[6] in SetGhostCells order (-x, +x, -y, +y, -z, +z)."""
        return [self.xl_bcnd, self.xu_bcnd, self.yl_bcnd, self.yu_bcnd, self.zl_bcnd, self.zu_bcnd]

    def n_global(self):
        return [self.nx, max(self.ny, 1), max(self.nz, 1)]

    def domain(self):
        return (self.xmin, self.ymin, self.zmin), (self.xlen, self.ylen, self.zlen)


# names the reference accepts that have no meaning for this path (other physics, output selection): parsed, not stored
_IGNORED = {"n_particle", "n_rotated_projection", "out_float32_GasEnergy",
            "snr_filename", "scale_outputs_file", "Omega_b", "bc_potential_type", "UVB_rates_file"}
# optionalParams, global.cpp:104-108
_OPTIONAL = {"flag_delta", "ddelta_dt", "n_delta", "Lz", "Lx", "phi", "theta", "delta", "nzr", "nxr", "H0", "Omega_M",
             "Omega_L", "Init_redshift", "End_redshift", "tile_length", "n_proc_x", "n_proc_y", "n_proc_z"}
_TYPES = {f.name: f.type for f in fields(Parameters) if f.name != "unknown"}


def parse_param(name, value, parms, warn=None):
    """Parse_Param (global.cpp:192-): one name/value pair into `parms`."""
    t = _TYPES.get(name)
    if t is int:
        v = atoi(value)
        if name == "output_always" and v not in (0, 1):
            raise ValueError("output_always must be 1 or 0.")  # CHOLLA_ASSERT, global.cpp:263
        setattr(parms, name, v)
    elif t is float:
        setattr(parms, name, atof(value))
    elif t is str:
        setattr(parms, name, value)
    elif name in _IGNORED or name in _OPTIONAL:
        pass
    else:
        parms.unknown[name] = value
        if warn is not None:
            warn("WARNING: %s/%s: Unknown parameter/value pair!" % (name, value))


def _split(line):
    """strtok(buff, "=") twice: tokens are maximal runs of non-'=' characters; fewer than two -> None."""
    tok = [t for t in line.split("=") if t != ""]
    return (tok[0], tok[1]) if len(tok) >= 2 else None


def parse_params(param_file, argv=(), warn=lambda m: print(m, file=sys.stderr)):
    """Parse_Params (global.cpp:127-188): the deck, then the command-line overrides."""
    parms = Parameters()
    with open(param_file) as fh:  # a missing file raises, like the reference's exit(1)
        for buff in fh:
            if buff[:1] in ("\n", "#", ";") or buff == "":
                continue
            nv = _split(buff)
            if nv is None:
                continue
            name, value = nv
            parse_param(name, value.strip(), parms, warn)  # Trim(value); the name is used as read
    for arg in argv:
        nv = _split(arg)
        if nv is None:
            continue
        parse_param(nv[0], nv[1], parms, warn)  # overrides are not trimmed (global.cpp:171-187)
    return parms


def parse_text(text, argv=(), warn=None):
    """The same from a string (tests)."""
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as fh:
        fh.write(text)
        path = fh.name
    try:
        return parse_params(path, argv, warn)
    finally:
        import os
        os.unlink(path)
