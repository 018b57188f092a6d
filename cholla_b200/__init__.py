"""cholla_b200 -- B200-native (sm_100a) implementation of Cholla's Van Leer hydro hot path.

The product is the C-ABI shared library `libcholla_b200.so` (include/cholla_b200.h) built from
cholla_b200/csrc.  This package is the thin host-side mirror of the reference's `Grid3D` call sites that
tests, bench.py and multi-GPU drivers use; PyTorch only supplies device memory, streams and
torch.distributed.  There is no CPU fallback: importing works anywhere, computing needs a CUDA device.
"""
from ._lib import (CHL_OK, ChollaB200Error, Config, EXACT, HLL, HLLC, HLLD, PCM, PLMC, PLMP, PPMC, PPMP, ROE, BC_PERIODIC, BC_REFLECTIVE,
                   BC_TRANSMISSIVE, BC_CUSTOM, BC_MPI, abi_symbols, boundary_extents, lib, local_extent,
                   tile_decomposition)
from .grid import Grid3D, n_ghost_for

__all__ = ["CHL_OK", "ChollaB200Error", "Config", "HLLC", "ROE", "HLLD", "HLL", "EXACT", "PCM", "PLMC", "PLMP", "PPMC", "PPMP", "BC_PERIODIC",
           "BC_REFLECTIVE", "BC_TRANSMISSIVE", "BC_CUSTOM", "BC_MPI", "abi_symbols", "boundary_extents", "lib",
           "local_extent", "tile_decomposition", "Grid3D", "n_ghost_for"]
