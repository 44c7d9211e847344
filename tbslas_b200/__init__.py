"""tbslas_b200: B200-native semi-Lagrangian advection hot path of arashb/tbslas.

The product is ``libtbslas_b200.so`` (hand-written CUDA for sm_100a behind the C ABI in
``include/tbslas_b200.h``).  This package holds the build script, the ctypes binding and a
host-side mirror of the reference's functor surface (``api``), plus ``flat_tree`` -- the
leaf-list data format and synthetic-input builders used by tests and the benchmark.
"""
from .capi import FREESPACE, PERIODIC, TbslasError  # noqa: F401
from .flat_tree import FlatTree  # noqa: F401

__all__ = ["FlatTree", "FREESPACE", "PERIODIC", "TbslasError"]
