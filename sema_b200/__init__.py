"""sema_b200 — B200-native exact vector search for Sema's retrieval hot path.

The product is the CUDA library behind ``include/sema_b200.h``; this package is the
Python binding plus the host-side mirror of the reference's storage interface.
Importing the package does not load the library; the first use does, and fails
loudly when it has not been built (there is no CPU fallback).
"""
from ._lib import METRIC_COSINE, METRIC_L2, SemaError  # noqa: F401
from .index import GpuIndex, ShardGroup  # noqa: F401

__all__ = ["GpuIndex", "ShardGroup", "SemaError", "METRIC_COSINE", "METRIC_L2"]
