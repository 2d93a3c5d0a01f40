"""Row-partitioned Chebyshev filter across ranks (one process per GPU, ``torch.distributed``).

SURVEY 8e: the recurrence shards by rows of L with ONE exchange per term -- every rank needs the
whole ``T_{k-1}`` (its rows reference arbitrary columns), so after each term the ranks all-gather
their row slices.  north_star asks for this only "where N outgrows one GPU"; on a single GPU the
full-operator ``meld_b200_cheby_filter`` is used instead.  The driver below is backend-agnostic:
on GPUs ``step`` is ``meld_b200_cheby_step`` on the rank's row slice (``DeviceGraph.from_scipy(L[a:b],
row0=a, n_cols=N)``) and the collective is NCCL; the CPU test runs it over gloo with a reference step.
"""

from __future__ import annotations

import numpy as np


def row_partition(n, world):
    """Contiguous, balanced row ranges: rank r owns [bounds[r], bounds[r+1])."""
    base, rem = divmod(int(n), int(world))
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def cheby_recurrence(step, allgather, T0_full, row_range, lmax, coeffs):
    """Three-term recurrence on this rank's rows (PyGSP ``cheby_op`` order of operations).

    step(T_cur_full, T_old_local, alpha, shift, gamma, c, c_cur, R_local, accumulate) -> (T_new_local, R_local)
    allgather(local_rows) -> full array assembled from every rank's rows, in rank order.
    Returns this rank's rows of R.
    """
    a, b = row_range
    coeffs = np.asarray(coeffs, dtype=np.float64)
    if coeffs.shape[0] < 2:
        raise TypeError("The coefficients have an invalid shape")
    a1 = float(lmax) / 2.0
    T_cur_full = T0_full
    T_old_local = None
    T_new_local, R_local = step(T_cur_full, None, 1.0 / a1, a1, 0.0, coeffs[1], 0.5 * coeffs[0], None, False)
    for k in range(2, coeffs.shape[0]):
        T_old_local = T_cur_full[a:b]
        T_cur_full = allgather(T_new_local)
        T_new_local, R_local = step(T_cur_full, T_old_local, 2.0 / a1, a1, 1.0, coeffs[k], 0.0, R_local, True)
    return R_local


def make_device_step(graph, p):
    """``step`` for :func:`cheby_recurrence` driving ``meld_b200_cheby_step`` on ``graph`` (a row slice)."""
    import torch

    from . import _native as nv

    lib = nv.lib()

    def step(T_cur_full, T_old_local, alpha, shift, gamma, c, c_cur, R_local, accumulate):
        T_new = torch.empty((graph.n_rows, p), dtype=torch.float64, device=T_cur_full.device)
        if R_local is None:
            R_local = torch.empty_like(T_new)
        nv.check(
            lib.meld_b200_cheby_step(graph._h, nv.ptr(T_cur_full.contiguous()),
                                     nv.ptr(None if T_old_local is None else T_old_local.contiguous()),
                                     nv.ptr(T_new), nv.ptr(R_local), p, float(alpha), float(shift), float(gamma),
                                     float(c), float(c_cur), int(bool(accumulate)), nv.current_stream_ptr()),
            "cheby_step",
        )
        return T_new, R_local

    return step


def make_torch_allgather(bounds, p, group=None):
    """All-gather of variable-length row slices (pads to the longest slice; NCCL or gloo)."""
    import torch
    import torch.distributed as dist

    world = len(bounds) - 1
    longest = max(bounds[r + 1] - bounds[r] for r in range(world))

    def allgather(local):
        pad = torch.zeros((longest, p), dtype=local.dtype, device=local.device)
        pad[: local.shape[0]] = local
        out = torch.empty((world * longest, p), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, pad, group=group)
        return torch.cat([out[r * longest: r * longest + (bounds[r + 1] - bounds[r])] for r in range(world)])

    return allgather
