"""``meld.utils`` surface of the hot path (reference ``meld/utils.py``)."""

from __future__ import annotations

import numpy as np
import pandas as pd

from . import _native as nv
from .graph import DeviceGraph


def _check_pygsp_graph(G):
    """Type gate of ``meld/utils.py:11-20``; here the accepted graph is a ``DeviceGraph``."""
    if not isinstance(G, DeviceGraph):
        raise TypeError(
            "Input graph should be of type graphtools.base.BaseGraph."
            " With graphtools, use the `use_pygsp=True` flag."
        )
    return G


def normalize_densities(sample_densities):
    """Row-wise L1 normalisation of sample densities (``meld/utils.py:35-47``).

    Rows whose absolute sum is zero stay zero; DataFrame index / columns are kept.
    Runs on the GPU (``meld_b200_l1_normalize_rows``); CUDA tensors are returned as
    CUDA tensors, everything else as ndarray / DataFrame like the reference.
    """
    torch = nv.require_cuda()
    is_df = isinstance(sample_densities, pd.DataFrame)
    if is_df:
        index, columns = sample_densities.index, sample_densities.columns
    on_device = isinstance(sample_densities, torch.Tensor) and sample_densities.is_cuda
    if on_device:
        x = sample_densities.to(torch.float64).contiguous()
    else:
        arr = np.ascontiguousarray(np.asarray(getattr(sample_densities, "values", sample_densities), dtype=np.float64))
        if arr.ndim != 2:
            raise ValueError("Expected 2D array, got {}D array instead".format(arr.ndim))
        x = nv.from_numpy_readonly(arr).cuda()
    n, p = x.shape
    out = torch.empty_like(x)
    nv.check(nv.lib().meld_b200_l1_normalize_rows(nv.ptr(x), n, p, nv.ptr(out), nv.current_stream_ptr()),
             "l1_normalize_rows")
    if on_device:
        return out
    norm = out.cpu().numpy()
    if is_df:
        norm = pd.DataFrame(norm, index=index, columns=columns, copy=False)  # a fresh array: no second copy
    return norm


def get_meld_cmap():
    raise NotImplementedError("plotting helpers are outside the B200 engine (needs scprep / matplotlib)")
