"""aircraftoptimalcontrol_b200 -- B200-native batched trajectory optimiser behind the Python entry points of
MohamedAtwan/AirCraftOptimalControl (regularized-Newton optimal control of a 2-D longitudinal aircraft model).

    from aircraftoptimalcontrol_b200.aircraft_simplified import Dynamics, Cost      # drop-ins
    from aircraftoptimalcontrol_b200.optcon import NewtonMethod, ltv_LQR
    from aircraftoptimalcontrol_b200.lqr_tracking import lqr_tracking
    from aircraftoptimalcontrol_b200 import BatchedNewton                          # the batched solver

All arithmetic runs in hand-written sm_100a CUDA kernels inside libacoc.so (C ABI: include/acoc.h); there is
no CPU fallback.
"""
from ._lib import AcocError, device_count, device_info  # noqa: F401
from .batch import BatchedNewton, PipelinedNewton  # noqa: F401

__all__ = ["BatchedNewton", "PipelinedNewton", "AcocError", "device_count", "device_info"]
__version__ = "0.1.0"
