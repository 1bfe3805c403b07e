"""spmv_test_b200 — B200-native sparse SGEMV (Y = x·A) behind the host interfaces of
PACTHEMAN123/spMV-test.

The product is `lib/libspmv_b200.so` (C-ABI in include/spmv_b200.h, hand-written sm_100a
kernels in csrc/).  This package is the thin Python host side over that C-ABI: plan objects
for tests and bench.py, and the column-slab partitioner used for the multi-GPU config.  There is
no CPU compute path: every call that produces y needs the CUDA library and a GPU.
"""
from ._cabi import LIB_PATH, SpmvError, lib, VARIANTS, LAYOUTS  # noqa: F401
from .plan import Plan, compact_x, pack_dump, ref_pack  # noqa: F401
from .partition import column_bounds, ShardedSgemv  # noqa: F401
from .group import Group, local_group, group_run_host  # noqa: F401
