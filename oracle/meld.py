"""Oracle: label handling and end-to-end order of operations (CPU, float64).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows
``meld/meld.py:143-191`` (indicators), ``:193-250`` (transform),
``:252-274`` (fit_transform) and ``meld/utils.py:35-47`` (normalize_densities).
"""

from __future__ import annotations

import numpy as np
import pandas as pd

from . import graph as _graph
from . import cheby as _cheby


def sample_indicators(sample_labels):
    """``MELD._create_sample_indicators``: one-hot columns in ``np.unique`` order."""
    samples = np.unique(sample_labels)
    labels = getattr(sample_labels, "values", sample_labels)
    labels = np.asarray(labels)
    if labels.ndim > 1:
        if labels.shape[1] == 1:
            labels = labels.reshape(-1)
        else:
            raise ValueError("sample_labels must be a single column. Got" "shape={}".format(labels.shape))
    ind = np.stack([(labels == s) for s in samples], axis=1).astype(int)
    return samples, ind


def transform(L, lmax, sample_labels, beta=60, offset=0, order=1, filter="heat", chebyshev_order=50,
              sample_normalize=True, solver="chebyshev"):
    """``MELD.transform`` given the graph Laplacian and lmax -> DataFrame (N, p)."""
    N = L.shape[0]
    if sample_labels.shape[0] != N:
        raise ValueError(
            "Input data ({}) and input graph ({}) " "are not of the same size".format(sample_labels.shape, N)
        )
    if len(np.unique(sample_labels)) == 1:
        raise ValueError("Found only one unqiue sample label. Cannot estimate density " "of a single sample.")
    index = sample_labels.index if hasattr(sample_labels, "index") else None
    samples, ind = sample_indicators(sample_labels)
    ind = ind.astype(np.float64)
    if sample_normalize:
        ind = ind / ind.sum(axis=0)
    if solver == "chebyshev":
        dens = _cheby.cheby_filter(L, lmax, ind, filter, beta, offset, order, chebyshev_order)
    elif solver == "exact":
        dens = _cheby.exact_filter(L, ind, filter, beta, offset, order)
    else:
        raise ValueError(solver)
    return pd.DataFrame(dens, index=index, columns=samples)


def fit_transform(X, sample_labels, knn=5, decay=40.0, thresh=1e-4, anisotropy=1.0, n_pca=100, random_state=None,
                  lmax=None, data_nu=None, n_jobs=1, **filter_kw):
    """``MELD(...).fit_transform(X, labels)`` on the kNN + Chebyshev path.

    Returns (densities DataFrame, graph dict, lmax).  ``lmax`` may be injected
    (parity runs share one value between oracle and GPU engine).
    """
    g = _graph.build_graph(X, knn=knn, decay=decay, thresh=thresh, anisotropy=anisotropy, n_pca=n_pca,
                           random_state=random_state, n_jobs=n_jobs, data_nu=data_nu)
    if lmax is None:
        lmax = _graph.estimate_lmax(g["L"], g["dw"])
    dens = transform(g["L"], lmax, sample_labels, **filter_kw)
    return dens, g, lmax


def normalize_densities(sample_densities):
    """``meld.utils.normalize_densities``: row-wise L1 normalisation (zero rows stay zero)."""
    from sklearn.preprocessing import normalize

    out = normalize(sample_densities, norm="l1")
    if isinstance(sample_densities, pd.DataFrame):
        out = pd.DataFrame(out, index=sample_densities.index, columns=sample_densities.columns)
    return out
