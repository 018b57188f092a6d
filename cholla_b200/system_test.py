"""The comparison rules of the reference's system-test harness (src/system_tests/system_tester.cpp,
src/utils/testing_utilities.cpp), on the snapshots this package writes (or the reference wrote):

    loadTestFieldData       system_tester.cpp:430-507   -> load_field(): per-rank files reassembled with offset / dims_local
    runTest                 system_tester.cpp:34-211    -> compare_to_fiducial(): every fiducial dataset equal to 4 ulps or
                                                           1e-14 absolute (nearlyEqualDbl), or the L2 norm of the per-dataset
                                                           mean absolute errors and the max error below given bounds
    runL1ErrorTest          system_tester.cpp:215-331   -> l1_error_test(): output 0 against output 1 (a wave after one
                                                           period), same norms, plus the expected number of time steps
    analyticConstant / analyticSine   testing_utilities.cpp:94-123

The reference's fiducial snapshots (cholla-tests-data) are Git-LFS objects that are not in this image; the tests that
need none of them -- the linear-wave, sound-wave and constant-state system tests, whose pass criteria are numbers in
hydro_system_tests.cpp -- run in tests/test_system_tests.py."""
import os

import numpy as np

from . import h5lite
from .snapshot import format_fname

I64_MAX = np.iinfo(np.int64).max


def ulps_distance(a, b):
    """testing_utilities::ulpsDistanceDbl, element-wise."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    ia, ib = a.view(np.int64), b.view(np.int64)
    with np.errstate(over="ignore"):
        d = np.abs(ia - ib)
    bad = np.isnan(a) | np.isnan(b) | np.isinf(a) | np.isinf(b) | ((ia < 0) != (ib < 0))
    return np.where(a == b, 0, np.where(bad, I64_MAX, d))


def nearly_equal(a, b, fixed_epsilon=1e-14, ulps_epsilon=4):
    """testing_utilities::nearlyEqualDbl: within ulps_epsilon ULPs (numbers far from zero) or fixed_epsilon absolutely."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return (ulps_distance(a, b) <= ulps_epsilon) | (np.abs(a - b) <= fixed_epsilon)


def load_field(outdir, nfile, name, n_ranks=1):
    """SystemTestRunner::loadTestFieldData: the dataset of every rank placed at its `offset` in an array of the global
    `dims`; face-centred magnetic fields are one longer along their own axis and neighbouring ranks share a face."""
    out = None
    for rank in range(n_ranks):
        with h5lite.File(format_fname(outdir, nfile, rank)) as f:
            dims = [int(v) for v in f.attrs["dims"]]
            nl = [int(v) for v in f.attrs["dims_local"]]
            off = [int(v) for v in f.attrs["offset"]]
            a = f[name].read().astype(np.float64)
        a = a.reshape(a.shape + (1,) * (3 - a.ndim))
        extra = [1 if name == "magnetic_" + "xyz"[d] else 0 for d in range(3)]
        if out is None:
            out = np.zeros([dims[d] + extra[d] for d in range(3)])
        out[off[0]:off[0] + nl[0] + extra[0], off[1]:off[1] + nl[1] + extra[1], off[2]:off[2] + nl[2] + extra[2]] = a
    return out


def dataset_names(outdir, nfile):
    with h5lite.File(format_fname(outdir, nfile, 0)) as f:
        return f.keys()


def n_steps(outdir, nfile):
    with h5lite.File(format_fname(outdir, nfile, 0)) as f:
        return int(f.attrs["n_step"][0])


def _norms(pairs):
    """Per-dataset mean absolute error, the L2 norm of those, and the largest single error (system_tester.cpp:153-207)."""
    l1 = {nm: float(np.abs(a - b).mean()) for nm, a, b in pairs}
    max_err = max(float(np.abs(a - b).max()) for _, a, b in pairs)
    return l1, float(np.sqrt(sum(v * v for v in l1.values()))), max_err


def compare_to_fiducial(outdir, fiducial, n_ranks=1, l2_norm_only=False, max_l1_error=0.0, max_error=0.0,
                        fixed_epsilon=1e-14, compare_n_steps=True, nfile=1):
    """SystemTestRunner::runTest.  `fiducial`: path of a single-file snapshot (the layout of cholla-tests-data) or a dict
    name -> array (setFiducialData).  Returns a list of failure messages (empty = pass)."""
    fails = []
    if isinstance(fiducial, str):
        with h5lite.File(fiducial) as f:
            fid = {nm: f[nm].read().astype(np.float64) for nm in f.keys()}
            fid_steps = int(f.attrs["n_step"][0])
        if compare_n_steps and n_steps(outdir, nfile) != fid_steps:
            fails.append("n_step: test %d, fiducial %d" % (n_steps(outdir, nfile), fid_steps))
    else:
        fid = {k: np.asarray(v, dtype=np.float64) for k, v in fiducial.items()}
    have = dataset_names(outdir, nfile)
    pairs = []
    for nm, want in fid.items():
        if nm not in have:
            fails.append("the test data does not contain the dataset '%s'" % nm)
            continue
        got = load_field(outdir, nfile, nm, n_ranks)
        want = want.reshape(want.shape + (1,) * (3 - want.ndim))
        if got.size != want.size:
            fails.append("the fiducial and test '%s' datasets are not the same length" % nm)
            continue
        got = got.reshape(want.shape)
        if l2_norm_only:
            pairs.append((nm, want, got))
        else:
            ok = nearly_equal(want, got, fixed_epsilon)
            if not ok.all():
                i = tuple(int(v[0]) for v in np.nonzero(~ok))
                fails.append("difference in %s at %s: fiducial %r, test %r (%d cells differ)"
                             % (nm, list(i), float(want[i]), float(got[i]), int((~ok).sum())))
    if l2_norm_only and pairs:
        _, l2, mx = _norms(pairs)
        if not l2 < max_l1_error:
            fails.append("the norm of the L1 error vector is %.3e, allowed %.3e" % (l2, max_l1_error))
        if not mx < max_error:
            fails.append("the maximum error is %.3e, allowed %.3e" % (mx, max_error))
    return fails


def l1_error_test(outdir, max_l1_error, max_error, n_ranks=1, expected_n_steps=None, skip=("GasEnergy",)):
    """SystemTestRunner::runL1ErrorTest: output 1 against output 0 of the same run.  Returns (failures, details)."""
    fails = []
    if expected_n_steps is not None and n_steps(outdir, 1) != expected_n_steps:
        fails.append("n_step: test %d, expected %d" % (n_steps(outdir, 1), expected_n_steps))
    pairs = [(nm, load_field(outdir, 0, nm, n_ranks), load_field(outdir, 1, nm, n_ranks))
             for nm in dataset_names(outdir, 0) if nm not in skip]
    l1, l2, mx = _norms(pairs)
    for nm, v in l1.items():
        if not v < max_l1_error:
            fails.append("the L1 error of %s is %.3e, allowed %.3e" % (nm, v, max_l1_error))
    if not l2 < max_l1_error:
        fails.append("the norm of the L1 error vector is %.3e, allowed %.3e" % (l2, max_l1_error))
    if not mx < max_error:
        fails.append("the maximum error is %.3e, allowed %.3e" % (mx, max_error))
    return fails, dict(l1=l1, l2=l2, max_error=mx, n_step=n_steps(outdir, 1))


def analytic_constant(outdir, name, value, n_ranks=1, fixed_epsilon=5.0e-12, nfile=1):
    """testing_utilities::analyticConstant (Check_Results: nearlyEqualDbl with a 5e-12 absolute allowance)."""
    return bool(nearly_equal(load_field(outdir, nfile, name, n_ranks), value, fixed_epsilon).all())


def analytic_sine(outdir, name, constant, amplitude, kx, ky, kz, phase, tolerance, n_ranks=1, nfile=1):
    """testing_utilities::analyticSine: constant + amplitude sin(kx i + ky j + kz k + phase) in cell-index space."""
    got = load_field(outdir, nfile, name, n_ranks)
    i, j, k = np.meshgrid(*(np.arange(n) for n in got.shape), indexing="ij")
    want = constant + amplitude * np.sin(kx * i + ky * j + kz * k + phase)
    return bool(nearly_equal(want, got, tolerance).all())
