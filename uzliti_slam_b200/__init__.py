"""uzliti_slam_b200 — B200-native feature-edge estimation (Hamming kNN-2 + ratio + RANSAC rigid transform).

The product is the C-ABI shared library built from csrc/ (include/uzliti_edge.h); the host-side mirror of
the reference's C++ interface lives in adapter/.  This Python package is only the thin ctypes binding the
tests and the bench harness use — it adds no compute and has NO CPU fallback: if the CUDA library is
missing or no B200 is visible, calls raise.
"""
from .binding import (EdgeEstimator, UzError, Params, EdgeResult, Features, lib_path, load_library,  # noqa: F401
                      build_library, EXPORTED_SYMBOLS, GroupEstimator)
