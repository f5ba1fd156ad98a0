"""eic-opticks_b200 : B200-native drop-in for the eic-opticks simulate path.

Gensteps (torch, Cerenkov, scintillation) or input photons in, detector hits out, behind the
reference's own conventions: CSGFoundry geometry arrays, bnd/optical/icdf tables, curand Philox
streams per absolute photon index, sphoton/sseq/record layouts.  The compute lives in
csrc/libphox.so (hand-written sm_100a CUDA behind the C ABI of include/phox.h); this package is the
host-side mirror of the reference interface plus the array builders.
"""
from . import lib                       # noqa: F401  ctypes binding (no CPU fallback)
from . import foundry, tables, gensteps, geometries   # noqa: F401
from .simulator import Simulator, Event  # noqa: F401
from .lib import (PhoxError, MODE_MINIMAL, MODE_HITPHOTON, MODE_HITPHOTONSEQ, MODE_DEBUGLITE, MODE_DEBUGHEAVY,  # noqa: F401
                  RNG_PRODUCTION, RNG_DEBUG_TAG, ACCEL_BVH, ACCEL_BRUTE, ACCEL_BVH_NOHOME,
                  KERNEL_AUTO, KERNEL_PERSISTENT, KERNEL_WAVEFRONT)

__all__ = ["Simulator", "Event", "PhoxError", "lib", "foundry", "tables", "gensteps", "geometries"]
