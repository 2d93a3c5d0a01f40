"""Oracle: MELD filter kernels + PyGSP Chebyshev approximation (CPU, float64).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows
``meld/filter.py:39-59`` (kernel definitions and the three PyGSP calls) and the
published PyGSP 0.5.1 algorithms behind them
(``pygsp/filters/approximations.py``: ``compute_cheby_coeff``, ``cheby_op``;
``pygsp/filters/filter.py``: ``Filter.filter`` with ``method='exact'``).
"""

from __future__ import annotations

import numpy as np
from scipy import sparse


def filter_kernel(name, beta, offset=0, order=1):
    """h(x) on normalised eigenvalues x = lambda / lmax  (``meld/filter.py:42-53``)."""
    name = name.lower()
    if name == "laplacian":
        return lambda x: 1 / (1 + (beta * np.abs(x - offset)) ** order)
    if name == "heat":
        return lambda x: np.exp(-beta * np.abs(x - offset) ** order)
    raise NotImplementedError


def cheby_coeff(h, lmax, m):
    """PyGSP ``compute_cheby_coeff(f, m)``: m+1 coefficients from N = m+1 nodes.

    ``h`` takes normalised eigenvalues; the reference closure divides by
    ``graph.lmax`` itself (``meld/filter.py:45,50``).
    """
    N = m + 1
    a1 = (lmax - 0) / 2
    a2 = (lmax + 0) / 2
    c = np.zeros(m + 1)
    tmpN = np.arange(N)
    num = np.cos(np.pi * (tmpN + 0.5) / N)
    for o in range(m + 1):
        c[o] = 2.0 / N * np.dot(h((a1 * num + a2) / lmax), np.cos(np.pi * o * (tmpN + 0.5) / N))
    return c


def cheby_op(L, lmax, c, signal):
    """PyGSP ``cheby_op(G, c, signal)`` for a single filter (three-term recurrence).

    The reference materialises ``2/a1 * (L - a2 I)`` as a second sparse matrix
    and multiplies with scipy's single-threaded ``csc_matvecs``; same here.
    """
    c = np.asarray(c, dtype=np.float64)
    M = c.shape[0]
    if M < 2:
        raise TypeError("The coefficients have an invalid shape")
    signal = np.asarray(signal, dtype=np.float64)
    L = L.tocsc()
    N = L.shape[0]
    a1 = float(lmax - 0) / 2.0
    a2 = float(lmax + 0) / 2.0
    twf_old = signal
    twf_cur = (L.dot(signal) - a2 * signal) / a1
    r = 0.5 * c[0] * twf_old + c[1] * twf_cur
    factor = 2 / a1 * (L - a2 * sparse.eye(N))
    for k in range(2, M):
        twf_new = factor * twf_cur - twf_old
        r = r + c[k] * twf_new
        twf_old = twf_cur
        twf_cur = twf_new
    return r


def cheby_filter(L, lmax, signal, filter="heat", beta=60, offset=0, order=1, chebyshev_order=50):
    """``meld.filter.filter(..., solver='chebyshev')`` given L and lmax."""
    h = filter_kernel(filter, beta, offset, order)
    c = cheby_coeff(h, lmax, chebyshev_order)
    return cheby_op(L, lmax, c, signal)


def dense_spectral_filter(L, lmax, signal, filter="heat", beta=60, offset=0, order=1):
    """U h(Lambda / lmax) U^T s with an explicit lmax (SURVEY 8c iv)."""
    e, U = np.linalg.eigh(L.toarray() if sparse.issparse(L) else np.asarray(L))
    h = filter_kernel(filter, beta, offset, order)
    return U @ (h(e / lmax)[:, None] * (U.T @ np.asarray(signal, dtype=np.float64)))


def exact_filter(L, signal, filter="heat", beta=60, offset=0, order=1):
    """``solver='exact'``: PyGSP ``compute_fourier_basis`` overwrites lmax with the
    true largest eigenvalue (no 1.01 factor) before the kernel is evaluated."""
    e, U = np.linalg.eigh(L.toarray() if sparse.issparse(L) else np.asarray(L))
    if -1e-12 < e[0] < 1e-12:  # compute_fourier_basis: "smallest eigenvalue should be zero: correct numerical errors"
        e[0] = 0
    lmax = e[-1]
    h = filter_kernel(filter, beta, offset, order)
    return U @ (h(e / lmax)[:, None] * (U.T @ np.asarray(signal, dtype=np.float64)))
