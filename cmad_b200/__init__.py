"""cmad_b200 - B200-native (sm_100a) constitutive-update hot path for CMAD.

Host side in Python (as the reference is), hot path in hand-written CUDA behind
the C-ABI of ``include/cmad_b200.h``.  There is no CPU fallback: compute entry
points raise if ``libcmad_b200.so`` has not been built.
"""
from . import _lib
from .material import NewtonSettings, active_param_ids, material_from_values
from .parameters import Parameters

__all__ = ["Parameters", "NewtonSettings", "material_from_values", "active_param_ids", "_lib"]
