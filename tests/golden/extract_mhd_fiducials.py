#!/usr/bin/env python3
"""Extract the known-answer values of the reference's own MHD unit tests into tests/golden/mhd_*.json.

Run in the build container only (reads /root/reference); the JSON it writes is committed and is what
tests/test_oracle_mhd_golden.py reads.  Only numbers are taken -- initial states, index -> value tables, scalar
constants -- no code.

Sources (reference @ 8478c9e):
  src/riemann_solvers/hlld_cuda_tests.cu:288-1876    Calculate_HLLD_Fluxes_CUDA: Brio & Wu, Dai & Woodward,
                                                     Ryu & Jones 4d, Einfeldt rarefaction, constant / degenerate states
  src/reconstruction/plmc_cuda_tests.cu:184-293      PLMC_cuda, MHD, 4^3 grid
  src/reconstruction/ppmc_cuda_tests.cu:121-259      PPMC_VL, MHD branch, 6^3 grid
  src/reconstruction/reconstruction_tests.cu:51-175  Primitive_To_Characteristic, Characteristic_To_Primitive,
                                                     Compute_Eigenvectors (MHD)
  src/mhd/ct_electric_fields_tests.cu:128-181        Calculate_CT_Electric_Fields on iota grids
  src/mhd/magnetic_update_tests.cu:116-123           Update_Magnetic_Field_3D on iota grids
  src/hydro/hydro_cuda_tests.cu:116-143              mhdInverseCrossingTime
  src/utils/mhd_utilities_tests.cu                   fast / slow magnetosonic and Alfven speeds, cell-centred fields
"""
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from extract_reference_fiducials import NUM, OUT, between, parse_map_lists, read  # noqa: E402


def _eval(expr, consts):
    expr = expr.strip()
    if not re.fullmatch(r"[-+*/ ().\w]+", expr):
        raise ValueError(expr)
    return float(eval(expr, {"__builtins__": {}}, dict(consts)))  # arithmetic on literals / named constants only


def hlld():
    t = read("src/riemann_solvers/hlld_cuda_tests.cu")
    starts = [m.start() for m in re.finditer(r"^TEST_F\(tMHDCalculateHLLDFluxesCUDA, (\w+)\)", t, flags=re.M)]
    end = t.index("TEST(tMHDHlldInternalApproximateLRWaveSpeeds")
    cases, skipped = [], []
    for a, b in zip(starts, starts[1:] + [end]):
        body = t[a:b]
        name = re.match(r"TEST_F\(tMHDCalculateHLLDFluxesCUDA, (\w+)\)", body).group(1)
        body_nc = re.sub(r"//[^\n]*", "", body)
        consts = {}
        for m in re.finditer(r"(?:Real|double)\s+const\s+(\w+)\s*=\s*([^;{]+);", body_nc):
            try:
                consts[m.group(1)] = _eval(m.group(2), consts)
            except Exception:
                pass
        if "gamma" not in consts:
            skipped.append(name + ": no gamma")
            continue
        states = {}
        for m in re.finditer(r"(\w+)\s*=\s*Primitive_2_Conserved\(\s*\{([^}]*)\}\s*,\s*gamma", body_nc):
            try:
                prim = [_eval(x, consts) for x in m.group(2).split(",")]
            except Exception:
                continue
            if len(prim) == 8:
                states[m.group(1)] = prim
        for m in re.finditer(r"std::vector<Real>\s+(?:const\s+)?(\w+)\s*\{([^}]*)\}", body_nc):
            pass  # raw conserved vectors (negative-pressure cases) are not transcribed
        # (fiducialFlux, Compute_Fluxes(l, r, gamma, direction)) pairs in source order
        n = 0
        for m in re.finditer(r"fiducialFlux\s*\{([^}]*)\}.*?Compute_Fluxes\(\s*(\w+)\s*,\s*(\w+)\s*,\s*gamma\s*,\s*direction\s*\)",
                             body_nc, flags=re.S):
            l, r = m.group(2), m.group(3)
            if l not in states or r not in states:
                continue
            try:
                flux = [_eval(x, consts) for x in m.group(1).split(",")]
            except Exception:
                continue
            if len(flux) != 8:
                continue
            cases.append({"test": name, "left": l, "right": r, "gamma": consts["gamma"], "prim_left": states[l],
                          "prim_right": states[r], "flux": flux})
            n += 1
        if n == 0:
            skipped.append(name + ": no transcribable case")
    return {"source": "src/riemann_solvers/hlld_cuda_tests.cu:288-1876",
            "layout": "prim = (rho, vx, vy, vz, P, Bx, By, Bz); flux = (rho, mx, my, mz, E, Bx=0, By, Bz), direction 0",
            "tolerance": "fixedEpsilon 2.7e-15 or 7 ulp (hlld_cuda_tests.cu:196-197)", "skipped": skipped, "cases": cases}


def plmc_mhd():
    t = read("src/reconstruction/plmc_cuda_tests.cu")
    a = t.index("TEST(tMHDPlmcReconstructor")
    t = t[a:]
    left, p = between(t, r"fiducial_interface_left\s*=", r";")
    right, _ = between(t, r"fiducial_interface_right\s*=", r";", p)
    return {"source": "src/reconstruction/plmc_cuda_tests.cu:184-293",
            "prng": "std::mt19937_64(42); uniform_real_distribution<double>(0.1, 5); draws: dx, dt, then grid",
            "nx": 4, "ny": 4, "nz": 4, "n_fields": 8, "gamma": 5.0 / 3.0,
            "left": parse_map_lists(left), "right": parse_map_lists(right)}


def ppmc_mhd():
    t = read("src/reconstruction/ppmc_cuda_tests.cu")
    a = t.index("TEST(tALLPpmcVLReconstructor")
    t = t[a:]
    a = t.index("// Fiducial Data")
    b = t.index("#else", a)
    t = t[a:b]
    left, p = between(t, r"fiducial_interface_left\s*=", r";")
    right, _ = between(t, r"fiducial_interface_right\s*=", r";", p)
    return {"source": "src/reconstruction/ppmc_cuda_tests.cu:121-259 (MHD branch)",
            "prng": "std::mt19937_64(42); uniform_real_distribution<double>(0.1, 5); draws: grid only",
            "nx": 6, "ny": 6, "nz": 6, "n_fields": 8, "gamma": 5.0 / 3.0,
            "left": parse_map_lists(left), "right": parse_map_lists(right)}


def _braced_numbers(text, pat):
    m = re.search(pat + r"\s*\{([^}]*)\}", text, flags=re.S)
    return [float(x) for x in re.findall(NUM, m.group(1))]


def scalars():
    out = {}
    t = read("src/reconstruction/reconstruction_tests.cu")
    a = t.index("TEST(tMHDReconstructionPrimitive2Characteristic")
    b = t.index("TEST(tMHDReconstructionCharacteristic2Primitive")
    c = t.index("TEST(tMHDReconstructionComputeEigenvectors")
    d = t.index("#endif", c)
    out["prim_to_char"] = {"source": "src/reconstruction/reconstruction_tests.cu:51-80", "gamma": 5.0 / 3.0,
                           "primitive": list(range(1, 9)), "slope": list(range(9, 17)), "eigen": list(range(17, 35)),
                           "fiducial": _braced_numbers(t[a:b], r"fiducial_results")}
    out["char_to_prim"] = {"source": "src/reconstruction/reconstruction_tests.cu:82-111", "gamma": 5.0 / 3.0,
                           "primitive": list(range(1, 9)), "characteristic": list(range(17, 24)),
                           "eigen": list(range(17, 35)), "fiducial": _braced_numbers(t[b:c], r"fiducial_results"),
                           "fiducial_order": "density, vx, vy, vz, pressure, magnetic_x (unset), magnetic_y, magnetic_z"}
    out["eigenvectors"] = {"source": "src/reconstruction/reconstruction_tests.cu:113-175", "gamma": 5.0 / 3.0,
                           "primitive": list(range(1, 9)), "fiducial": _braced_numbers(t[c:d], r"fiducial_results")}

    t = read("src/mhd/ct_electric_fields_tests.cu")
    ct = {}
    for name in ("PositiveVelocity", "NegativeVelocity", "ZeroVelocity"):
        a = t.index("TEST_F(tMHDCalculateCTElectricFields, %sExpectCorrectOutput)" % name)
        b = t.index("Run_Test();", a)
        ct[name] = {m.group(1): float(m.group(2)) for m in
                    re.finditer(r"fiducialData\.at\((\d+)\)\s*=\s*(" + NUM + r")", t[a:b])}
    out["ct_electric_fields"] = {"source": "src/mhd/ct_electric_fields_tests.cu:25-181", "n": 2,
                                 "setup": "fluxX, fluxY, fluxZ (7 fields each), grid (8 fields): one running iota; "
                                          "Negative: first n_cells entries of each flux negated; Zero: zeroed; "
                                          "untouched outputs keep -999", "cases": ct}

    t = read("src/mhd/magnetic_update_tests.cu")
    a = t.index("TEST_F(tMHDUpdateMagneticField3D")
    out["magnetic_update"] = {"source": "src/mhd/magnetic_update_tests.cu:25-123", "n": 3, "dt": 3.2, "dx": 2.5,
                              "setup": "sourceGrid iota from 0 (8 fields), ctElectricFields iota continuing; "
                                       "destination prefilled with -999",
                              "fiducial": {m.group(1): float(m.group(2)) for m in
                                           re.finditer(r"fiducialData\.at\((\d+)\)\s*=\s*(" + NUM + r")", t[a:])}}

    t = read("src/hydro/hydro_cuda_tests.cu")
    a = t.index("TEST(tMHDMhdInverseCrossingTime")
    b = t.index("// Function to test", a)
    vals = {m.group(1): _eval(m.group(2), {}) for m in re.finditer(r"double const (\w+)\s*=\s*([^;]+);", t[a:b])}
    out["mhd_inverse_crossing_time"] = {"source": "src/hydro/hydro_cuda_tests.cu:116-143", **vals}

    t = read("src/utils/mhd_utilities_tests.cu")
    a = t.index("struct TestParams")
    b = t.index("std::vector<std::string> names", a)
    params = {m.group(1): [float(x) for x in re.findall(NUM, m.group(2))] for m in
              re.finditer(r"std::vector<double>\s+(\w+)\{([^}]*)\}", t[a:b])}

    def fid(test, var):
        a = t.index(test)
        return _braced_numbers(t[a:], var)
    out["mhd_utilities"] = {
        "source": "src/utils/mhd_utilities_tests.cu", "gamma": 5.0 / 3.0, "params": params,
        "fast_speed": {"coef": [1.0, 1.0, 1.0e-25],
                       "fiducial": fid("TEST(tMHDFastMagnetosonicSpeed, CorrectInput", "fiducialFastMagnetosonicSpeed")},
        "fast_speed_negative_density": {"coef": [1.0, 1.0, 1.0e-25],
                                        "fiducial": fid("TEST(tMHDFastMagnetosonicSpeed, NegativeDensity",
                                                        "fiducialFastMagnetosonicSpeed")},
        "slow_speed": {"coef": 1e-95, "first_index": 2,
                       "fiducial": fid("TEST(tMHDSlowMagnetosonicSpeed, CorrectInput", "fiducialSlowMagnetosonicSpeed")},
        "slow_speed_negative_density": {"coef": 1e-95, "first_index": 2,
                                        "fiducial": fid("TEST(tMHDSlowMagnetosonicSpeed, NegativeDensity",
                                                        "fiducialSlowMagnetosonicSpeed")},
        "alfven_speed": {"fiducial": fid("TEST(tMHDAlfvenSpeed, CorrectInput", "fiducialAlfvenSpeed")},
        "alfven_speed_negative_density": {"fiducial": fid("TEST(tMHDAlfvenSpeed, NegativeDensity", "fiducialAlfvenSpeed")},
        "cell_centred": {"nx": 3, "ny": 3, "n_cells": 125, "xid": 1, "yid": 1, "zid": 1, "fiducial": [637.5, 761.5, 883.5],
                         "setup": "8-field grid of 125 cells per field filled with iota; indices use nx = ny = 3"},
    }
    t = read("src/mhd/magnetic_divergence_tests.cu")
    m = re.search(r"fiducialDivergence\s*=\s*(" + NUM + r")\s*/\s*(" + NUM + r")", t)
    out["magnetic_divergence"] = {
        "source": "src/mhd/magnetic_divergence_tests.cu:27-70",
        "grid_size": int(re.search(r"gridSize\s*=\s*(\d+)", t).group(1)),
        "n_ghost": int(re.search(r"n_ghost\s*=\s*(\d+)", t).group(1)),
        "dx": float(re.search(r"G\.H\.dx\s*=\s*(" + NUM + r")", t).group(1)), "n_fields": 8,
        "prng": "std::mt19937(1); std::uniform_real_distribution<double>(1, 5); every value of the 8-field grid, in "
                "memory order, divided by 1E15",
        "fiducial_numerator": float(m.group(1)), "fiducial_denominator": float(m.group(2)),
    }
    return out


def main():
    for name, fn in (("mhd_hlld", hlld), ("mhd_plmc", plmc_mhd), ("mhd_ppmc_vl", ppmc_mhd), ("mhd_scalars", scalars)):
        data = fn()
        with open(os.path.join(OUT, name + ".json"), "w") as fh:
            json.dump(data, fh, indent=1)
        if name == "mhd_hlld":
            print(name, len(data["cases"]), "cases; skipped:", data["skipped"])
        else:
            print(name, "ok")


if __name__ == "__main__":
    main()
