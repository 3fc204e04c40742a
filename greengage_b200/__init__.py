"""greengage_b200 — B200-native executor for Greengage's scan / hash-join / aggregate / sort / motion hot path.

The product is libggb200.so (hand-written sm_100a CUDA behind the C-ABI of include/ggb200.h) plus
libgghost.so (host C: executor-node surface, synthetic loader).  This Python package is plumbing for
tests and benchmarks: ctypes bindings, plan builders, torch.distributed glue for Motion.
"""
from . import capi  # noqa: F401
