"""genima_b200 — B200-native (sm_100a) implementation of Genima's per-step inference hot path.

Python here is plumbing (device memory, streams, weight binding, the reference-facing call signatures); all arithmetic
runs in hand-written CUDA kernels reached through the C ABI in include/genima_b200.h (genima_b200/csrc/*.cu).
"""
__version__ = "0.1.0"
