"""meld_b200 -- Blackwell-native (sm_100a) engine for the MELD graph-filter hot path.

Public surface = the hot path of KrishnaswamyLab/MELD (``meld/__init__.py:3-8``):
``MELD`` (``fit`` / ``transform`` / ``fit_transform`` / ``set_params``) and
``utils.normalize_densities``.  Everything numerical runs in ``libmeld_b200.so``
(hand-written CUDA behind the C-ABI of ``include/meld_b200.h``); importing the
package does not need a GPU, calling it does.
"""

from .meld import MELD
from .graph import DeviceGraph
from .utils import normalize_densities
from . import utils, filter, synthetic  # noqa: F401

__version__ = "0.1.0"
__all__ = ["MELD", "DeviceGraph", "normalize_densities", "utils", "filter", "synthetic"]
