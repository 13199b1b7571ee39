"""gcm_filters_b200 -- B200-native (sm_100a) drop-in for the gcm-filters iterative Laplacian filter.

Same public names as the reference package (``gcm_filters/__init__.py:11-15``)."""
from .engine import set_devices
from .filter import Filter, FilterShape
from .kernels import GridType, required_grid_vars

__version__ = "0.1.0"
# the reference's four public names, plus the one knob the reference has no counterpart for: which GPUs of this process
# a host-resident batch is sharded over (default: the current device)
__all__ = ["Filter", "FilterShape", "GridType", "required_grid_vars", "set_devices"]
