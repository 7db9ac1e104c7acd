"""navigation_b200 -- B200-native (sm_100a CUDA) costmap layering/inflation and DWA rollout scoring behind the ROS
navigation stack's plugin seams.  The product is navigation_b200/libnavgpu.so (C ABI in include/navgpu.h); this
package is the thin ctypes binding used by tests and bench.py.  There is no CPU fallback: importing works anywhere,
but every compute call needs a CUDA device and raises NavGpuError otherwise.
"""
from .api import Api, Costmap, Dwa, NavGpuError, load  # noqa: F401
