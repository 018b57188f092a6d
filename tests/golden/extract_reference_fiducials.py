#!/usr/bin/env python3
"""Extract the hard-coded known-answer values ("fiducials") of the reference's own unit
tests into small JSON fixtures.

Run in the build container only (it reads /root/reference, which does not exist on the
GPU box); the JSON it writes is committed under tests/golden/ and is what the tests read.

Sources (reference @ 8478c9e):
  src/reconstruction/plmc_cuda_tests.cu:26-183      PLMC_cuda hydro, 5x4x4 grid
  src/reconstruction/ppmc_cuda_tests.cu:121-259     PPMC_VL hydro (non-MHD branch), 6^3 grid
  src/reconstruction/reconstruction_tests.cu:501-572 PPM_Single_Variable, 100 stencils
  src/hydro/hydro_cuda_tests.cu:28-143              Calc_dt_3D / inverse crossing times
  src/riemann_solvers/hllc_cuda_tests.cu:152-178    HLLC uniform state

Only numbers are taken (index -> value tables and scalar constants); no code.
"""
import json
import os
import re
import sys

REF = os.environ.get("CHOLLA_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))

NUM = r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?"


def read(path):
    with open(os.path.join(REF, path)) as fh:
        return fh.read()


def parse_map_lists(block):
    """Parse `{{{i, v}, {i, v}}, {{...}}, {{...}}}` into a list of {str(i): v} dicts."""
    out = []
    depth = 0
    cur = None
    for m in re.finditer(r"\{|\}|(\d+)\s*,\s*(" + NUM + r")", block):
        tok = m.group(0)
        if tok == "{":
            depth += 1
            if depth == 2:
                cur = {}
        elif tok == "}":
            if depth == 2:
                out.append(cur)
                cur = None
            depth -= 1
        else:
            cur[m.group(1)] = float(m.group(2))
    return out


def between(text, start_pat, end_pat, pos=0):
    s = re.search(start_pat, text[pos:])
    if not s:
        raise RuntimeError("pattern not found: " + start_pat)
    a = pos + s.end()
    e = re.search(end_pat, text[a:])
    return text[a : a + e.start()], a + e.end()


def plmc():
    t = read("src/reconstruction/plmc_cuda_tests.cu")
    a = t.index("TEST(tHYDROPlmcReconstructor")
    b = t.index("TEST(", a + 10)
    t = t[a:b]
    left, p = between(t, r"fiducial_interface_left\s*=", r";")
    right, _ = between(t, r"fiducial_interface_right\s*=", r"\s;", p)
    return {
        "source": "src/reconstruction/plmc_cuda_tests.cu:26-183",
        "prng": "std::mt19937_64(42); std::uniform_real_distribution<double>(0.1, 5); draws: dx, dt, then grid",
        "nx": 5, "ny": 4, "nz": 4, "n_fields": 5, "gamma": 5.0 / 3.0,
        "left": parse_map_lists(left), "right": parse_map_lists(right),
    }


def ppmc_vl():
    t = read("src/reconstruction/ppmc_cuda_tests.cu")
    a = t.index("TEST(tALLPpmcVLReconstructor")
    t = t[a:]
    # hydro branch is after '#else   // not MHD'
    h = t.index("#else", t.index("fiducial_interface_left"))
    t_h = t[h : t.index("#endif", h)]
    left, p = between(t_h, r"fiducial_interface_left\s*=", r";")
    right, _ = between(t_h, r"fiducial_interface_right\s*=", r";", p)
    m = t[t.index("#ifdef MHD", t.index("// Fiducial Data")) : h]
    mleft, p = between(m, r"fiducial_interface_left\s*=", r";")
    mright, _ = between(m, r"fiducial_interface_right\s*=", r";", p)
    return {
        "source": "src/reconstruction/ppmc_cuda_tests.cu:121-259",
        "prng": "std::mt19937_64(42); std::uniform_real_distribution<double>(0.1, 5); draws: grid only",
        "note": "hydro test allocates (n_fields-1)*N interface words, so the energy slot is out of bounds "
                "and the index-54 entries of the hydro 'right' tables are an allocator artefact; ignore them",
        "nx": 6, "ny": 6, "nz": 6, "gamma": 5.0 / 3.0,
        "hydro": {"n_fields": 5, "left": parse_map_lists(left), "right": parse_map_lists(right)},
        "mhd": {"n_fields": 8, "left": parse_map_lists(mleft), "right": parse_map_lists(mright)},
    }


def ppm_single_variable():
    t = read("src/reconstruction/reconstruction_tests.cu")
    a = t.index("TEST(tALLReconstructionPPMSingleVariable")
    t = t[a : t.index("TEST(", a + 10)]
    left, p = between(t, r"fiducial_left_interface\s*\{", r"\}")
    right, _ = between(t, r"fiducial_right_interface\s*\{", r"\}", p)
    return {
        "source": "src/reconstruction/reconstruction_tests.cu:501-572",
        "prng": "std::mt19937_64(42); std::uniform_real_distribution<double>(-100, 100); 500 draws, 5 per stencil",
        "left": [float(x) for x in re.findall(NUM, left)],
        "right": [float(x) for x in re.findall(NUM, right)],
    }


def scalars():
    t = read("src/hydro/hydro_cuda_tests.cu")
    dt3d = float(re.search(r"fiducialDt\s*=\s*(" + NUM + ")", t).group(1))
    icts = [float(x) for x in re.findall(r"fiducialInverseCrossingTime\s*=\s*(" + NUM + ")", t)]
    return {
        "source": "src/hydro/hydro_cuda_tests.cu:28-143; src/riemann_solvers/hllc_cuda_tests.cu:152-178",
        "calc_dt_3d": {"conserved": [1.0, 0.0, 0.0, 0.0, 1.0], "dx": 1.0, "dy": 1.0, "dz": 1.0,
                       "gamma": 5.0 / 3.0, "n_ghost": 0, "dti": dt3d},
        "inverse_crossing_time": {
            "energy": 7.6976906577e2, "density": 1.6756968986,
            "velocity": [7.0829278656, 5.9283073464, 8.8417748226],
            "magnetic": [9.2400807786, 8.0382409757, 3.3284839263],
            "cell_size": [8.1019429453e2, 7.1254780684e2, 7.5676716066e2],
            "gamma": 5.0 / 3.0, "hydro": icts[0], "mhd": icts[1]},
        "hllc_uniform": {"density": 1.0, "pressure": 1.0, "velocity": [0.0, 0.0, 0.0], "gamma": 1.4,
                         "flux": [0.0, 1.0, 0.0, 0.0, 0.0]},
    }


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found at %s (this script only runs in the build container)" % REF)
    fixtures = {
        "plmc_hydro.json": plmc(),
        "ppmc_vl.json": ppmc_vl(),
        "ppm_single_variable.json": ppm_single_variable(),
        "scalars.json": scalars(),
    }
    for name, data in fixtures.items():
        with open(os.path.join(OUT, name), "w") as fh:
            json.dump(data, fh, indent=1)
        print("wrote", name)


if __name__ == "__main__":
    main()
