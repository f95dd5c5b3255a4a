"""biolith_b200 -- B200-native log-density + gradient of biolith's occupancy likelihoods.

Only the hot path is here: the packed dataset handle, the fused sm_100a kernels
behind a C ABI (include/biolith_b200.h), and thin Python host code over ctypes.  No torch, no
triton, no CPU fallback: importing works anywhere, computing needs the built .so and a B200.
"""

from ._lib import LIB_PATH, BiolithB200Error
from .likelihood import DeviceBuffer, OccupancyLikelihood

__all__ = ["OccupancyLikelihood", "DeviceBuffer", "BiolithB200Error", "LIB_PATH"]
__version__ = "0.1.0"
