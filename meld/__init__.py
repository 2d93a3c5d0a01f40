"""Drop-in import name: ``import meld; meld.MELD().fit_transform(X, labels)`` runs the B200 engine.

Thin alias of :mod:`meld_b200` covering the hot-path surface of the reference package
(``MELD``, ``utils.normalize_densities``, ``filter.filter``).
"""

from meld_b200 import MELD, DeviceGraph, normalize_densities, utils, filter, __version__  # noqa: F401
