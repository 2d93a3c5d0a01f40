"""MELD graph filter on the GPU: same signature as the reference ``meld.filter.filter``.

Reference: ``meld/filter.py:5-61`` -- estimate lmax, build the heat / laplacian
kernel h(lambda / lmax), apply it with PyGSP's Chebyshev approximation
(``compute_cheby_coeff`` + ``cheby_op``).  Here the m+1 coefficients are computed
on the host (a few dozen numbers) and the three-term recurrence runs in
libmeld_b200 (``meld_b200_cheby_filter``).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nv
from .graph import DeviceGraph, _as_device_f64

MAX_P = 8  # signal columns per recurrence launch


def filter_kernel(name, beta, offset=0, order=1):
    """h(x) on normalised eigenvalues x = lambda / lmax (``meld/filter.py:42-53``)."""
    name = name.lower()
    if name == "laplacian":
        return lambda x: 1 / (1 + (beta * np.abs(x - offset)) ** order)
    if name == "heat":
        return lambda x: np.exp(-beta * np.abs(x - offset) ** order)
    raise NotImplementedError


def cheby_coefficients(h, lmax, m):
    """Chebyshev coefficients of h(lambda/lmax) on [0, lmax]: m+1 values from m+1 nodes.

    Same quadrature as PyGSP ``compute_cheby_coeff`` (N = m+1 nodes, c_0 used halved).
    """
    n = m + 1
    q = np.arange(n)
    a = lmax / 2.0
    hv = h((a * np.cos(np.pi * (q + 0.5) / n) + a) / lmax)
    # one dot product per order, arguments formed as pi*o*(q+1/2)/n, so the few dozen
    # coefficients round exactly like PyGSP's loop
    return np.array([2.0 / n * np.dot(hv, np.cos(np.pi * o * (q + 0.5) / n)) for o in range(m + 1)])


def filter(signal, graph, filter, beta, offset=0, order=1, solver="chebyshev", chebyshev_order=None, apply=None):
    """Low-pass filter ``signal`` (N, p) over ``graph``; returns an ndarray (or a CUDA tensor
    when ``signal`` is one).  ``apply(lmax, coeffs, S)`` replaces the single-GPU recurrence (row-partitioned
    multi-GPU filter, ``meld_b200.distributed.ShardedFilter.apply``)."""
    if not isinstance(graph, DeviceGraph):
        raise TypeError("graph must be a meld_b200.DeviceGraph")
    h = filter_kernel(filter, beta, offset, order)  # NotImplementedError for unknown kernels
    torch = nv.require_cuda()
    if solver == "exact":
        on_device = isinstance(signal, torch.Tensor) and signal.is_cuda
        S = _as_device_f64(torch, signal)
        R = exact_apply(graph, h, S[:, None] if S.dim() == 1 else S)
        R = R[:, 0] if S.dim() == 1 else R
        return R if on_device else R.cpu().numpy()
    if solver != "chebyshev":
        raise ValueError("solver value {} not recognized. Choose from ['chebyshev', 'exact']".format(solver))
    lmax = graph.estimate_lmax()
    coeffs = np.ascontiguousarray(cheby_coefficients(h, lmax, int(chebyshev_order)), dtype=np.float64)
    on_device = isinstance(signal, torch.Tensor) and signal.is_cuda
    S = _as_device_f64(torch, signal)
    squeeze = S.dim() == 1
    if squeeze:
        S = S[:, None]
    R = cheby_apply(graph, lmax, coeffs, S) if apply is None else apply(lmax, coeffs, S)
    if squeeze:
        R = R[:, 0]
    return R if on_device else R.cpu().numpy()


def cheby_apply(graph, lmax, coeffs, S):
    """R = sum_k c_k T_k(L) S on the device; S is an (N, p) float64 CUDA tensor."""
    torch = nv.require_cuda()
    N, p = S.shape
    if N != graph.N:
        raise ValueError("signal has {} rows, graph has {} nodes".format(N, graph.N))
    cptr = coeffs.ctypes.data_as(C.POINTER(C.c_double))
    if p <= MAX_P:
        S = S.contiguous()
        R = torch.empty_like(S)
        nv.check(
            nv.lib().meld_b200_cheby_filter(graph._h, float(lmax), cptr, len(coeffs), nv.ptr(S), p, nv.ptr(R),
                                            nv.current_stream_ptr()),
            "cheby_filter",
        )
        return R
    R = torch.empty((N, p), dtype=torch.float64, device=S.device)
    for j in range(0, p, MAX_P):
        Sj = S[:, j:j + MAX_P].contiguous()
        Rj = torch.empty_like(Sj)
        nv.check(
            nv.lib().meld_b200_cheby_filter(graph._h, float(lmax), cptr, len(coeffs), nv.ptr(Sj), Sj.shape[1],
                                            nv.ptr(Rj), nv.current_stream_ptr()),
            "cheby_filter",
        )
        R[:, j:j + MAX_P] = Rj
    return R


def cheby_sweep(graph, lmax, coeff_matrix, S):
    """Shared-basis sweep: ``coeff_matrix`` is (F, m+1) -- one Chebyshev coefficient vector per filter -- and
    ``S`` an (N, p) float64 CUDA tensor.  The recurrence runs once per chunk of 8 signal columns; returns the
    (F, N, p) densities ``R[f] = sum_k c[f, k] T_k(L) S`` (reference pattern ``meld/benchmark.py:186-200``:
    one graph, many ``MELD(beta=b).transform(labels)``)."""
    torch = nv.require_cuda()
    N, p = S.shape
    if N != graph.N:
        raise ValueError("signal has {} rows, graph has {} nodes".format(N, graph.N))
    cm = np.ascontiguousarray(coeff_matrix, dtype=np.float64)
    if cm.ndim != 2 or cm.shape[1] < 2:
        raise TypeError("The coefficients have an invalid shape")
    F, nc = cm.shape
    cptr = cm.ctypes.data_as(C.POINTER(C.c_double))
    out = torch.empty((F, N, p), dtype=torch.float64, device=S.device)
    for j in range(0, p, MAX_P):
        Sj = S[:, j:j + MAX_P].contiguous()
        pj = Sj.shape[1]
        Rj = out if pj == p else torch.empty((F, N, pj), dtype=torch.float64, device=S.device)
        nv.check(
            nv.lib().meld_b200_cheby_sweep(graph._h, float(lmax), cptr, F, nc, nv.ptr(Sj), pj, nv.ptr(Rj),
                                           nv.current_stream_ptr()),
            "cheby_sweep",
        )
        if Rj is not out:
            out[:, :, j:j + pj] = Rj
    return out


def exact_apply(graph, h, S):
    """``solver="exact"`` (PyGSP ``Filter.filter(method="exact")`` reached from ``meld/filter.py:59``):
    ``U h(e / lmax) U^T S`` with the full eigendecomposition of L.  PyGSP's ``compute_fourier_basis`` clears the
    rounding of the zero eigenvalue and OVERWRITES lmax with the largest eigenvalue (no 1.01 factor), which is what
    the reference's filter closure then reads (SURVEY finding 7).  Test-scale only (N <= 16384): the dense ``eigh``
    and the two GEMMs are library calls (cuSOLVER / cuBLAS through torch), not kernels of this engine."""
    torch = nv.require_cuda()
    Ld = graph.to_dense_L()
    e, U = torch.linalg.eigh(Ld)
    if abs(float(e[0])) < 1e-12:
        e[0] = 0.0
    lmax = float(e[-1])
    graph.lmax = lmax  # like PyGSP, the graph's lmax is now the exact one
    hv = torch.from_numpy(np.asarray(h(e.cpu().numpy() / lmax), dtype=np.float64)).to(S.device)
    return U @ (hv[:, None] * (U.T @ S))
