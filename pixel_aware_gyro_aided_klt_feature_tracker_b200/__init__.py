"""pagk: B200-native pixel-aware gyro-aided KLT hot path (CUDA, sm_100a) behind a C-ABI.

``capi``    ctypes mirror of include/pagk.h and the loader of csrc/libpagk_cuda.so
``tracker`` Python mirror of the reference's GyroAidedTracker / PatchMatch interface
``synth``   seeded synthetic frame pairs + gyro of the BASELINE.json shapes
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
