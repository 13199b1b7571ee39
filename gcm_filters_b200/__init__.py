"""gcm_filters_b200 -- B200-native (sm_100a) drop-in for the gcm-filters iterative Laplacian filter.

Same public names as the reference package (``gcm_filters/__init__.py:11-15``)."""
from .filter import Filter, FilterShape
from .kernels import GridType, required_grid_vars

__version__ = "0.1.0"
__all__ = ["Filter", "FilterShape", "GridType", "required_grid_vars"]
