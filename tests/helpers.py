"""Shared helpers for the parity tests: build a ghosted local array and run both arms on it."""
import numpy as np

import oracle


def ghosted(fields, g, flags):
    """fields (5, nz, ny, nx) real cells -> flat SoA array with ghost layers filled by the ORACLE's boundary
    routine (so both arms start from the very same bytes)."""
    nf, nz, ny, nx = fields.shape
    gz = g if nz > 1 else 0
    gy = g if ny > 1 else 0
    NX, NY, NZ = nx + 2 * g, ny + 2 * gy, nz + 2 * gz
    a = np.zeros((nf, NZ, NY, NX))
    a[:, gz:gz + nz, gy:gy + ny, g:g + nx] = fields
    flat = np.ascontiguousarray(a.reshape(-1))
    oracle.set_boundaries(flat, NX, NY, NZ, nf, g, flags)
    return flat, (NX, NY, NZ)


def real_view(flat, dims, g):
    NX, NY, NZ = dims
    a = flat.reshape(-1, NZ, NY, NX)
    gz = g if NZ > 1 else 0
    gy = g if NY > 1 else 0
    return a[:, gz:NZ - gz, gy:NY - gy, g:NX - g]


def rel_l1(a, b):
    """relative L1 error per field over real cells: sum|a-b| / sum|b| (SURVEY 8d parity gate)."""
    out = []
    for f in range(a.shape[0]):
        den = np.abs(b[f]).sum()
        num = np.abs(a[f] - b[f]).sum()
        out.append(num / den if den > 0 else num)
    return np.array(out)


def max_ulp(a, b):
    ia = np.ascontiguousarray(a, dtype=np.float64).view(np.int64)
    ib = np.ascontiguousarray(b, dtype=np.float64).view(np.int64)
    return int(np.abs(ia - ib).max())
