"""biolith_b200 -- B200-native log-density + gradient of biolith's occupancy likelihoods.

Only the hot path is here: the packed dataset handle, the fused sm_100a kernels
behind a C ABI (include/biolith_b200.h), and thin Python host code over ctypes.  No torch, no
triton, no CPU fallback: importing works anywhere, computing needs the built .so and a B200.
"""

from . import diagnostics, models
from ._lib import LIB_PATH, BiolithB200Error
from .fit import FitResult, fit
from .likelihood import DeviceBuffer, OccupancyLikelihood
from .nuts import NutsSampler
from .simulate import simulate_occupancy

__all__ = ["OccupancyLikelihood", "NutsSampler", "fit", "FitResult", "models", "diagnostics",
           "simulate_occupancy", "DeviceBuffer", "BiolithB200Error", "LIB_PATH"]
__version__ = "0.1.0"
