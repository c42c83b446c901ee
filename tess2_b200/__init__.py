"""tess2_b200 -- B200-native density-estimation stage of tess2 (dense() and its per-tet /
per-cell geometry) behind a C ABI.  See DESIGN.md and include/tess_b200.h.

Only what the hot path needs lives here:
  csrc/      CUDA kernels (sm_100a) + the C ABI  -> libtess_b200.so
  lib.py     ctypes binding of the C ABI
  dense.py   host-side mirror of the reference's dense()/WriteGrid/volume interface
  multi.py   one-process-per-GPU driver (torch.distributed for the plumbing, NCCL inside the library)
  harness/   host-side input staging: particle generators, block decomposition, SciPy-Qhull
"""
from .dense import (DENSE_TESS, DENSE_CIC, DENSE_DTFE, Context, DenseResult, dense, WriteGrid, fill_vert_to_tet,  # noqa: F401
                    fill_circumcenters, volume, complete, default_context, check_blocks)
from .lib import TessB200Error, LIB_PATH  # noqa: F401
