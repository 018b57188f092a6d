#!/usr/bin/env python3
"""Extract the inputs and known answers of the reference's HELPER-level unit tests (hydro build: neither MHD nor DE
defined) into tests/golden/helpers.json.

Run in the build container only (it reads /root/reference); the JSON it writes is committed and is what
tests/test_oracle_helpers_golden.py reads.  Only numbers are taken -- brace-initialised structs, vectors and scalar
constants of the test bodies -- no code.

Sources (reference @ 8478c9e):
  src/reconstruction/reconstruction_tests.cu:177-203   Thread_Guard
  src/reconstruction/reconstruction_tests.cu:205-249   Load_Data
  src/reconstruction/reconstruction_tests.cu:251-334   Compute_Slope, Van_Leer_Slope
  src/reconstruction/reconstruction_tests.cu:336-392   Monotonize_Characteristic_Return_Primitive
  src/reconstruction/reconstruction_tests.cu:425-458   Calc_Interface_Linear
  src/reconstruction/reconstruction_tests.cu:574-611   Write_Data
  src/reconstruction/reconstruction_tests.cu:613-690   Plm_Limit_Interfaces
  src/utils/hydro_utilities_tests.cpp:36-59, 225-236   Get_Pressure_From_DE
  src/utils/math_utilities_tests.cpp:45-76             dotProduct, SquareMagnitude
  src/hydro/hydro_cuda_tests.cu:147-196                Scalar_Floor_Kernel
"""
import json
import os
import re

REF = os.environ.get("CHOLLA_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
NUM = r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?"


def read(path):
    with open(os.path.join(REF, path)) as fh:
        return fh.read()


def preprocess(text, defined=()):
    """Keep the branches a build without MHD / DE compiles (#ifdef / #ifndef / #else / #endif only)."""
    out, stack = [], []
    for line in text.splitlines():
        t = line.strip()
        m = re.match(r"#\s*(ifdef|ifndef)\s+(\w+)", t)
        if m:
            on = (m.group(2) in defined) == (m.group(1) == "ifdef")
            stack.append(on)
            continue
        if re.match(r"#\s*else", t):
            stack[-1] = not stack[-1]
            continue
        if re.match(r"#\s*endif", t):
            stack.pop()
            continue
        if all(stack):
            out.append(line)
    return "\n".join(out)


def test_body(text, name, defined=()):
    a = text.index(name)
    b = re.search(r"\nTEST|\n__global__|\Z", text[a + 10:])
    return preprocess(text[a:a + 10 + b.start()], defined)


def braces(body, var):
    m = re.search(r"\b" + re.escape(var) + r"\s*(?:=\s*)?\{([^{}]*)\}", body)
    if not m:
        raise RuntimeError("no initialiser for " + var)
    return [float(x) for x in re.findall(NUM, m.group(1).replace("5. / 3.", ""))]


def scalar(body, var):
    m = re.search(r"\b" + re.escape(var) + r"\s*=\s*(" + NUM + r")\s*[;,]", body)
    return float(m.group(1))


def members(body, var):
    """`var.density = x; var.velocity_x = ...` assignments in struct order."""
    vals = []
    for f in ("density", "velocity_x", "velocity_y", "velocity_z", "pressure"):
        m = re.search(r"\b" + re.escape(var) + r"\." + f + r"\s*=\s*(" + NUM + r")\s*;", body)
        vals.append(float(m.group(1)))
    return vals


def main():
    rec = read("src/reconstruction/reconstruction_tests.cu")
    out = {"source": "reference unit tests, hydro build (no MHD, no DE); see the header of extract_helper_fiducials.py"}

    b = test_body(rec, "TEST(tALLReconstructionThreadGuard")
    out["thread_guard"] = {"order": int(scalar(b, "order")), "nx": int(scalar(b, "nx")), "ny": int(scalar(b, "ny")),
                           "nz": int(scalar(b, "nz")),
                           "inside_ids": [int(x) for x in re.findall(r"fiducial_vals\.at\((\d+)\)\s*=\s*0", b)]}

    b = test_body(rec, "TEST(tALLReconstructionLoadData")
    out["load_data"] = {"n": 3, "energy_scale": 5.0e2, "gamma": 5.0 / 3.0, "xid": 1, "yid": 1, "zid": 1,
                        "fiducial": braces(b, "fiducial_data")}

    b = test_body(rec, "TEST(tALLReconstructionComputeSlope")
    out["compute_slope"] = {"left": braces(b, "left"), "right": braces(b, "right"), "coef": scalar(b, "coef"),
                            "fiducial": scalar(b, "fiducial_data")}

    b = test_body(rec, "TEST(tALLReconstructionVanLeerSlope")
    out["van_leer_slope"] = {"left": braces(b, "left"), "right": braces(b, "right"), "fiducial": braces(b, "fiducial_data")}

    b = test_body(rec, "TEST(tALLReconstructionMonotonizeCharacteristicReturnPrimitive")
    out["monotonize_characteristic"] = {
        "primitive": braces(b, "primitive"), "del_a_L": braces(b, "del_a_L"), "del_a_R": braces(b, "del_a_R"),
        "del_a_C": braces(b, "del_a_C"), "del_a_G": braces(b, "del_a_G"), "sound_speed": scalar(b, "sound_speed"),
        "fiducial": braces(b, "fiducial_data")}

    b = test_body(rec, "TEST(tALLReconstructionCalcInterfaceLinear")
    out["calc_interface_linear"] = {"left": braces(b, "left"), "right": braces(b, "right"), "coef": scalar(b, "coef"),
                                    "fiducial": braces(b, "fiducial_data")}

    b = test_body(rec, "TEST(tALLReconstructionWriteData")
    pairs = re.search(r"fiducial_interface\s*=\s*\{(.*?)\};", b, re.S).group(1)
    out["write_data"] = {"interface": braces(b, "interface"), "n": 3, "gamma": 5.0 / 3.0, "xid": 1, "yid": 1, "zid": 1,
                         "fiducial": {m.group(1): float(m.group(2))
                                      for m in re.finditer(r"\{(\d+)\s*,\s*(" + NUM + r")\}", pairs)}}

    b = test_body(rec, "TEST(tHYDROReconstructionPlmLimitInterfaces")
    out["plm_limit_interfaces"] = {k: members(b, k) for k in
                                   ("interface_l_iph", "interface_r_imh", "cell_im1", "cell_i", "cell_ip1",
                                    "interface_l_iph_fiducial", "interface_r_imh_fiducial")}

    # the -DMHD branches of the same tests (8-field states; the eigenvector struct of the monotonisation test is 17 .. 34)
    M = ("MHD",)
    b = test_body(rec, "TEST(tALLReconstructionLoadData", M)
    out["mhd_load_data"] = {"n": 3, "energy_scale": 5.0e2, "gamma": 5.0 / 3.0, "xid": 1, "yid": 1, "zid": 1,
                            "fiducial": braces(b, "fiducial_data")}
    b = test_body(rec, "TEST(tALLReconstructionMonotonizeCharacteristicReturnPrimitive", M)
    out["mhd_monotonize_characteristic"] = {
        "primitive": braces(b, "primitive"), "del_a_L": braces(b, "del_a_L"), "del_a_R": braces(b, "del_a_R"),
        "del_a_C": braces(b, "del_a_C"), "del_a_G": braces(b, "del_a_G"), "sound_speed": scalar(b, "sound_speed"),
        "eigenvectors": braces(b, "eigenvectors"), "fiducial": braces(b, "fiducial_data")}
    b = test_body(rec, "TEST(tALLReconstructionWriteData", M)
    pairs = re.search(r"fiducial_interface\s*=\s*\{(.*?)\};", b, re.S).group(1)
    out["mhd_write_data"] = {"interface": braces(b, "interface"), "n": 3, "gamma": 5.0 / 3.0, "xid": 1, "yid": 1, "zid": 1,
                             "fiducial": {m.group(1): float(m.group(2))
                                          for m in re.finditer(r"\{(\d+)\s*,\s*(" + NUM + r")\}", pairs)}}

    hu = preprocess(read("src/utils/hydro_utilities_tests.cpp"))
    b = test_body(hu, "TEST(tHYDROHydroUtilsGetPressureFromDE")
    out["get_pressure_from_de"] = {"E": braces(hu, "E"), "U_total": braces(hu, "U_total"),
                                   "U_advected": braces(hu, "U_advected"), "gamma": 5.0 / 3.0,
                                   "fiducial": braces(b, "fiducial_Ps")}

    hu_raw = read("src/utils/hydro_utilities_tests.cpp")
    params = {k: braces(hu, k) for k in ("d", "vx", "vy", "vz", "E", "P", "magnetic_x", "magnetic_y", "magnetic_z")}
    out["pressure_energy_primitive"] = {"gamma": 5.0 / 3.0, "params": params}
    for build, defined in (("hydro", ()), ("mhd", ("MHD",))):
        tp = test_body(hu_raw, "TEST(tHYDROtMHDHydroUtilsCalcPressurePrimitive, CorrectInputExpectCorrectOutput", defined)
        te = test_body(hu_raw, "TEST(tHYDROtMHDHydroUtilsCalcEnergyPrimitive, CorrectInputExpectCorrectOutput", defined)
        tn = test_body(hu_raw, "TEST(tHYDROtMHDHydroUtilsCalcEnergyPrimitive, NegativePressureExpectAutomaticFix", defined)
        out["pressure_energy_primitive"][build] = {"pressure": braces(tp, "fiducial_pressure"),
                                                   "energy": braces(te, "fiducial_energy"),
                                                   "energy_negative_pressure": braces(tn, "fiducial_energy")}

    b = test_body(hu_raw, "TEST(tHYDROtMHDCalcKineticEnergyFromVelocity")
    out["kinetic_energy_from_velocity"] = {"coef": scalar(b, "coef"), "fiducial": braces(b, "fiducialEnergies")}

    mu = read("src/utils/math_utilities_tests.cpp")
    b = test_body(mu, "TEST(tALLDotProduct")
    ab = re.search(r"a\{([^{}]*)\},\s*b\{([^{}]*)\}", b)
    out["dot_product"] = {"a": [float(x) for x in re.findall(NUM, ab.group(1))],
                          "b": [float(x) for x in re.findall(NUM, ab.group(2))],
                          "fiducial": scalar(b, "fiducialDotProduct")}
    b = test_body(mu, "TEST(tALLSquareMagnitude")
    out["square_magnitude"] = {"a": braces(b, "a"), "fiducial": scalar(b, "fiducial_square_magnitude")}

    b = test_body(read("src/hydro/hydro_cuda_tests.cu"), "TEST(tHYDROScalarFloor")
    out["scalar_floor"] = {"n_fields": int(scalar(b, "n_fields")), "field_num": int(scalar(b, "field_num")),
                           "scalar_floor": scalar(b, "scalar_floor"),
                           "inputs": [float(x) for x in re.findall(r"host_conserved\.at\(field_num\)\s*=\s*(" + NUM + r")", b)]}

    with open(os.path.join(OUT, "helpers.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote helpers.json:", ", ".join(k for k in out if k != "source"))


if __name__ == "__main__":
    main()
