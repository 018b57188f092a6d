"""Strict mode's shared-reciprocal division (cholla_b200/csrc/exact_div.cuh) must return the bits of the
compiler's own IEEE division for EVERY operand pair -- it replaces the reference's `a / b` (built -fmad=false,
Makefile:148) inside the bit-exact kernels.  Random operands over the whole exponent range plus the edge cases the
fast-path test is about: zero / tiny numerators, denormal and infinite quotients, huge and denormal denominators,
infinities and NaNs, signed zeros."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(a, b):
    import torch
    import cholla_b200 as cb
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    qs, qp = torch.empty_like(ta), torch.empty_like(ta)
    rc = cb.lib().chl_debug_exact_div(C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(qs.data_ptr()),
                                      C.c_void_p(qp.data_ptr()), a.size, None)
    assert rc == 0
    torch.cuda.synchronize()
    return qs.cpu().numpy(), qp.cpu().numpy()


def _same_bits(x, y):
    xi, yi = x.view(np.uint64), y.view(np.uint64)
    both_nan = np.isnan(x) & np.isnan(y)  # NaN payloads are not part of the contract
    return bool(np.all((xi == yi) | both_nan))


def test_random_operands_whole_range():
    rng = np.random.default_rng(7)
    n = 1 << 22
    # uniformly random bit patterns: every exponent, both signs, denormals, infinities, NaNs
    a = rng.integers(0, 1 << 64, n, dtype=np.uint64).view(np.float64).copy()
    b = rng.integers(0, 1 << 64, n, dtype=np.uint64).view(np.float64).copy()
    qs, qp = _run(a, b)
    assert _same_bits(qs, qp)
    with np.errstate(all="ignore"):
        assert _same_bits(qp, a / b)  # and both are the IEEE quotient


def test_physical_range_and_edge_cases():
    rng = np.random.default_rng(11)
    n = 1 << 21
    a = (rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 12, n))
    b = (rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 12, n))
    a[::7] = 0.0
    a[1::7] = -0.0
    edge = np.array([0.0, -0.0, 1.0, -1.0, 6.0, 3.0, 1e-300, -1e-300, 5e-324, 2.2250738585072014e-308, 1e300, -1e300,
                     1.7976931348623157e308, np.inf, -np.inf, np.nan, 2.0 ** -969, 2.0 ** -970, 2.0 ** 1017, 2.0 ** 1016,
                     1.0 + 2.0 ** -52, 2.0 - 2.0 ** -52, 1.0 / 3.0])
    A, B = np.meshgrid(edge, edge)
    a = np.concatenate([a, A.ravel()])
    b = np.concatenate([b, B.ravel()])
    qs, qp = _run(a, b)
    assert _same_bits(qs, qp)
    with np.errstate(all="ignore"):
        assert _same_bits(qp, a / b)


def test_constant_denominators_of_the_path():
    """x / 6 (PPM_Single_Variable) and x / (gamma - 1) with zero and round-off-sized numerators."""
    rng = np.random.default_rng(3)
    n = 1 << 20
    a = rng.standard_normal(n) * 10.0 ** rng.uniform(-30, 5, n)
    a[::5] = 0.0
    a[1::5] = -0.0
    for den in (6.0, 1.66666667 - 1.0, 5.0 / 3.0 - 1.0, 1.4 - 1.0):
        qs, qp = _run(a, np.full(n, den))
        assert _same_bits(qs, qp)


def _run_sqrt(x):
    import torch
    import cholla_b200 as cb
    tx = torch.from_numpy(x).cuda()
    ro, rp = torch.empty_like(tx), torch.empty_like(tx)
    rc = cb.lib().chl_debug_exact_sqrt(C.c_void_p(tx.data_ptr()), C.c_void_p(ro.data_ptr()), C.c_void_p(rp.data_ptr()),
                                       x.size, None)
    assert rc == 0
    torch.cuda.synchronize()
    return ro.cpu().numpy(), rp.cpu().numpy()


def test_sqrt_whole_range_and_edges():
    """sqrt_r (strict mode) must return the bits of sqrt() for every operand: random bit patterns (every exponent,
    negatives, denormals, infinities, NaNs), the physical range, and the edges of the fast-path window."""
    rng = np.random.default_rng(5)
    n = 1 << 22
    x = rng.integers(0, 1 << 64, n, dtype=np.uint64).view(np.float64).copy()
    phys = np.abs(rng.standard_normal(n)) * 10.0 ** rng.uniform(-25, 25, n)
    edge = np.array([0.0, -0.0, 1.0, 4.0, 2.0, 1e-20, 1e-300, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308,
                     np.inf, -np.inf, np.nan, -1.0, 2.0 ** -969, 2.0 ** -970, 2.0 ** 1000, 2.0 ** 1016, 1.0 + 2.0 ** -52,
                     2.0 - 2.0 ** -52, 4.0 - 2.0 ** -50])
    x = np.concatenate([x, phys, edge])
    ro, rp = _run_sqrt(x)
    assert _same_bits(ro, rp)
    with np.errstate(all="ignore"):
        assert _same_bits(rp, np.sqrt(x))


def _run_flagged(a, b, zero_ok=True):
    import torch
    import cholla_b200 as cb
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    q = torch.empty_like(ta)
    fl = torch.empty(a.size, dtype=torch.int32, device="cuda")
    rc = cb.lib().chl_debug_exact_div_flagged(C.c_void_p(ta.data_ptr()), C.c_void_p(tb.data_ptr()), C.c_void_p(q.data_ptr()),
                                              C.c_void_p(fl.data_ptr()), int(zero_ok), a.size, None)
    assert rc == 0
    torch.cuda.synchronize()
    return q.cpu().numpy(), fl.cpu().numpy()


def test_branch_free_policy_flag_clear_means_exact():
    """kDivFlag (the hydro sweeps): whenever the sticky flag stays clear the raw quotient carries the bits of a / b and
    the raw square root those of sqrt(a); the flag never comes up on the operands the path actually sees (normal
    magnitudes and exact zeros of either sign), so the IEEE redo stays out of the hot loop."""
    rng = np.random.default_rng(23)
    n = 1 << 22
    a = rng.integers(0, 1 << 64, n, dtype=np.uint64).view(np.float64).copy()
    b = rng.integers(0, 1 << 64, n, dtype=np.uint64).view(np.float64).copy()
    edge = np.array([0.0, -0.0, 1.0, -1.0, 6.0, 3.0, 1e-300, -1e-300, 5e-324, 2.2250738585072014e-308, 1e300, -1e300,
                     1.7976931348623157e308, np.inf, -np.inf, np.nan, 2.0 ** -969, 2.0 ** -970, 2.0 ** 1017, 2.0 ** 1016,
                     2.0 ** -1016, 2.0 ** -1017, 2.0 ** -1022, 1.0 + 2.0 ** -52, 2.0 - 2.0 ** -52, 1.0 / 3.0])
    A, B = np.meshgrid(edge, edge)
    a = np.concatenate([a, A.ravel()])
    b = np.concatenate([b, B.ravel()])
    for zero_ok in (True, False):
        q, fl = _run_flagged(a, b, zero_ok)
        with np.errstate(all="ignore"):
            want = a / b
        clear = (fl & 1) == 0
        assert clear.sum() > 0.15 * a.size  # random bit patterns: a sizeable share stays in the fast-path window
        assert _same_bits(q[clear], want[clear])
        assert not np.any(fl & 4)          # clear sqrt flag => the bits of sqrt(a)
    # the physical range, exact zeros of both signs included: no flag at all
    m = 1 << 21
    a = rng.standard_normal(m) * 10.0 ** rng.uniform(-30, 30, m)
    b = rng.standard_normal(m) * 10.0 ** rng.uniform(-30, 30, m)
    a[::5] = 0.0
    a[1::5] = -0.0
    q, fl = _run_flagged(a, b, True)
    assert not np.any(fl & 1)
    with np.errstate(all="ignore"):
        assert _same_bits(q, a / b)
    # the nonzero hint: a zero numerator only raises the flag (the block is redone), it never passes silently
    q, fl = _run_flagged(a, b, False)
    z = a == 0.0
    assert np.all(fl[z] & 1) and not np.any(fl[~z] & 1)
    pos = np.abs(a) + 1e-250               # inside the square root's fast-path window (2^-969 ... 2^1016)
    _, fl = _run_flagged(pos, b, True)
    assert not np.any(fl & 2)              # positive normal arguments never flag the square root


def test_one_test_window_of_the_branch_free_policy():
    """kDivFlag tests the denominator once (2^-400 <= |b| <= 2^400, on its reciprocal) and then only the quotient
    (2^-560 < |q| < 2^1017): the numerator |a| = |q b| is then inside the stock fast-path window (>= 2^-969, finite) without
    a test of its own.  Operands straddling every edge of that window: flag clear => the bits of a / b, and the window is
    not wider than claimed (denominators outside it, quotients below it and non-finite operands always raise the flag)."""
    rng = np.random.default_rng(29)
    eb = np.array([-1022, -1000, -402, -401, -400, -399, -1, 0, 1, 399, 400, 401, 402, 1000, 1016, 1023], dtype=np.float64)
    eq = np.array([-1074, -1023, -1022, -970, -969, -562, -561, -560, -559, -558, 0, 559, 560, 600, 1015, 1017], dtype=np.float64)
    mb = 1.0 + rng.random((64, 1, 1))
    mq = 1.0 + rng.random((64, 1, 1))
    with np.errstate(all="ignore"):
        b = (mb * 2.0 ** eb[None, :, None]) * np.ones((1, 1, eq.size))
        a = (mq * 2.0 ** eq[None, None, :]) * b          # a = q b (rounded): q lands next to the edge exponent
    a, b = a.ravel().copy(), b.ravel().copy()
    sign = rng.integers(0, 2, a.size) * 2.0 - 1.0
    a *= sign
    for zero_ok in (True, False):
        q, fl = _run_flagged(a, b, zero_ok)
        with np.errstate(all="ignore"):
            want = a / b
        clear = (fl & 1) == 0
        assert _same_bits(q[clear], want[clear])
        assert clear.sum() > 0.15 * a.size
        outside = (np.abs(b) < 2.0 ** -401) | (np.abs(b) > 2.0 ** 401) | ~np.isfinite(a) | ~np.isfinite(want) | \
                  ((np.abs(want) < 2.0 ** -561) & (a != 0.0))
        assert np.all(fl[outside] & 1)
    # division by the constant 6 (PPM_Single_Variable) over the whole range of numerators
    a = rng.integers(0, 1 << 64, 1 << 22, dtype=np.uint64).view(np.float64).copy()
    a = np.concatenate([a, rng.standard_normal(1 << 21) * 10.0 ** rng.uniform(-160, 300, 1 << 21),
                        np.array([0.0, -0.0, 6.0, -6.0, 1.0, 3.0, 2.0 ** -559, 2.0 ** -557, 1.7976931348623157e308])])
    with np.errstate(all="ignore"):
        want = a / 6.0
    for mode in (True,):
        q, fl = _run_flagged(a, np.full_like(a, 6.0), mode)
        clear = (fl & 1) == 0
        assert clear.sum() > 0.7 * a.size
        assert _same_bits(q[clear], want[clear]) and np.all(fl[~np.isfinite(a)] & 1)
