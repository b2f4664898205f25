"""cnosdb_b200 — B200-native (sm_100a) implementation of CnosDB's tskv scan hot path:
TSM page decode -> time-range + series-selection filter -> time-bucketed aggregate.

  cabi     ctypes mirror of include/tskv_gpu.h (the drop-in C ABI)
  engine   host-side mirror of the reference's reader interface (QueryOption, ScanResult, ...)
  datagen  seeded synthetic TSM pages + host-side page writer
  build    builds the in-tree shared libraries
"""
from . import cabi  # noqa: F401

__all__ = ["cabi", "engine", "datagen", "build"]
