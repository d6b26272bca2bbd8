"""vvflow_b200 — B200-native (sm_100a CUDA) implementation of libvvhd's per-step particle hot path.

Only what the path needs lives here: csrc/ (CUDA kernels + the C ABI of include/vvgpu.h),
capi.py (ctypes binding), vvhd.py (host-side mirror of the reference's class interface) and
host/ (the C++ adapter classes for the vvflow binary). See DESIGN.md.
"""
from . import capi  # noqa: F401
from .vvhd import (MConvectiveFast, MDiffusiveFast, MEpsilonFast, MFlowmove, Space, TBody,  # noqa: F401
                   TSortedTree, XPressure, XVorticity)
