"""positionbaseddynamics_b200 -- B200-native PBD/XPBD constraint-projection engine.

Hot path only: the inner solver loop of the reference's TimeStepController::step
as hand-written sm_100a kernels behind a C ABI (include/pbd_b200.h), plus a host-side mirror of the
reference's SimulationModel / TimeStepController interface.  No CPU fallback exists.
"""
from . import _capi  # noqa: F401
from ._capi import Engine, PbdError  # noqa: F401
