"""rtpp_b200 — B200 (sm_100a) implementation of the data-parallel hot path shared by the four
planners of TZY-H/RTPP.  The product is the C-ABI library librtpp_gpu.so (include/rtpp_gpu.h) and
the C++ planner classes above it (rtpp_b200/planners); this Python package is only a thin ctypes
mirror of that ABI for tests and bench.py.  There is no CPU fallback: importing works anywhere,
but every call needs the built library and a CUDA device.
"""
from .capi import CdrtEngine, CdrtPlanner, Ctx, Engine, FmtEngine, RtppError, lib, lib_path, build  # noqa: F401
