"""fpocket_b200 — B200-native (sm_100a) implementation of fpocket's alpha-sphere pocket-detection hot path.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C ABI of include/fpocket_cuda.h),
`capi.py` (ctypes mirror of that ABI), `hostprep.py` (the flattening of parsed atoms the host shim does)
and `api.py` (Python mirror of the reference's search_pocket / mdpocket grid interface).
"""
__version__ = "0.1.0"
