"""Row-partitioned Chebyshev filter across the GPUs of one box (one process per GPU, ``torch.distributed``).

SURVEY 8e: the recurrence shards by rows of L with ONE exchange per term -- every rank needs the whole
``T_{k-1}`` (its rows reference arbitrary columns).  The reference has nothing to match (no distributed code,
reference ``setup.py:45``); the operation being sharded is the recurrence behind ``meld/filter.py:59``.

Two data paths over the same row partition (rank r owns rows ``[r * chunk, (r + 1) * chunk)`` of L in the graph's
INTERNAL cell order, ``chunk = ceil(N / world)``):

* ``mode="p2p"`` (default): ``meld_b200_cheby_filter_dist`` -- the SpMM kernel stores its rows of ``T_k``
  straight into every peer's HBM over NVLink (CUDA IPC mapped buffers) and the ranks meet through flag words
  in each other's memory; no NCCL call and no host on the data path.  torch.distributed only carries the
  64-byte IPC handles once.
* ``mode="nccl"``: north_star's wording -- ``meld_b200_cheby_step`` on the rank's rows followed by an in-place
  NCCL all-gather of the ``T_k`` slices per term (pre-allocated buffers, one stream).  Kept as the baseline
  the fused path is measured against, and it is what the CPU (gloo) test drives with a numpy step.

``cheby_recurrence`` is the backend-agnostic statement of the order of operations (PyGSP ``cheby_op``).
"""

from __future__ import annotations

import ctypes as C

import numpy as np


def row_partition(n, world):
    """Contiguous, balanced row ranges: rank r owns [bounds[r], bounds[r+1])."""
    base, rem = divmod(int(n), int(world))
    bounds = [0]
    for r in range(world):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


def chunk_partition(n, world):
    """Equal chunks (what an all-gather needs): rank r owns [min(n, r*chunk), min(n, (r+1)*chunk))."""
    chunk = -(-int(n) // int(world))
    return chunk, [min(int(n), r * chunk) for r in range(world + 1)]


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


class PeerContext:
    """This rank's peer-memory block (``meld_b200_dist_t``), connected to the blocks of all ranks of the group."""

    def __init__(self, n_rows_total, p_max=8, group=None):
        import torch.distributed as dist

        from . import _native as nv

        lib = nv.lib()
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        h = C.c_void_p()
        nv.check(lib.meld_b200_dist_create(self.rank, self.world, int(n_rows_total), int(p_max),
                                           nv.current_stream_ptr(), C.byref(h)), "dist_create")
        self._h = h
        self.n, self.p_max = int(n_rows_total), int(p_max)
        if self.world > 1:
            nb = lib.meld_b200_dist_handle_bytes()
            blob = (C.c_ubyte * nb)()
            nv.check(lib.meld_b200_dist_export(self._h, blob), "dist_export")
            blobs = [None] * self.world
            dist.all_gather_object(blobs, bytes(blob), group=group)  # plumbing: 64 bytes per rank, once
            raw = b"".join(blobs)
            buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
            nv.check(lib.meld_b200_dist_connect(self._h, buf), "dist_connect")
            dist.barrier(group=group)  # every rank has mapped every block before anyone stores into one

    def error(self):
        from . import _native as nv

        e = C.c_int(0)
        nv.check(nv.lib().meld_b200_dist_error(self._h, C.byref(e)), "dist_error")
        return e.value

    def close(self):
        from . import _native as nv

        if getattr(self, "_h", None) is not None and self._h.value:
            nv.lib().meld_b200_dist_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_CTX_CACHE = {}


def shared_context(n_rows_total, p_max=8, group=None):
    """One PeerContext per (N, p_max, group) and process: mapping peer memory costs a cudaMalloc, IPC opens and
    a barrier, far more than a filter; successive graphs of the same size reuse the buffers."""
    key = (int(n_rows_total), int(p_max), id(group))
    ctx = _CTX_CACHE.get(key)
    if ctx is None:
        ctx = _CTX_CACHE[key] = PeerContext(n_rows_total, p_max, group)
    return ctx


class ShardedFilter:
    """Row-partitioned ``cheby_apply`` for one graph: every rank passes the same (N, p) signal in the caller's cell
    order and gets the same (N, p) result.  ``graph`` is the FULL DeviceGraph (every rank holds one after a
    sharded build); its rows are sliced here."""

    def __init__(self, graph, group=None, mode="p2p", p_max=8, row_slice=None, bounds=None):
        """``graph``: the FULL DeviceGraph (its rows are sliced here in equal chunks), or -- p2p mode only -- pass the
        rank's ready ``row_slice`` with the ranks' row ``bounds`` (what ``DeviceGraph.from_data_sharded_rows`` returns):
        then no rank ever holds the whole graph."""
        import torch.distributed as dist

        if mode not in ("p2p", "nccl"):
            raise ValueError("mode value {} not recognized. Choose from ['p2p', 'nccl']".format(mode))
        self.graph, self.group, self.mode, self.p_max = graph, group, mode, int(p_max)
        if dist.is_available() and dist.is_initialized():
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self._nccl_bufs = {}
        self.halo_fraction = 1.0
        self._own_slice = row_slice is None
        if row_slice is not None:
            if mode != "p2p":
                raise ValueError("a ready row slice needs mode='p2p' (the NCCL all-gather wants equal chunks)")
            self.N = row_slice.n_cols
            self.bounds = list(bounds)
            self.chunk = max(self.bounds[r + 1] - self.bounds[r] for r in range(self.world))
            self.row_range = (self.bounds[self.rank], self.bounds[self.rank + 1])
            assert self.row_range == (row_slice.row0, row_slice.row0 + row_slice.n_rows)
            self.slice = row_slice
            self.ctx = shared_context(self.N, self.p_max, self.group)
            if self.world > 1:
                self._exchange_halo()
            return
        self.N = graph.N
        self.chunk, self.bounds = chunk_partition(self.N, self.world)
        self.row_range = (self.bounds[self.rank], self.bounds[self.rank + 1])
        self._attach()

    def _attach(self):
        """Native state: this rank's row slice and (p2p) the peer-memory context."""
        self.slice = self.graph.row_slice(*self.row_range)
        self.ctx = shared_context(self.N, self.p_max, self.group) if self.mode == "p2p" else None
        if self.mode == "p2p" and self.world > 1:
            self._exchange_halo()

    def _exchange_halo(self):
        """Which of this rank's rows does each peer gather?  One all-to-all of byte marks (plumbing, once per graph);
        afterwards a row of T_k is only stored into the peers that reference it."""
        import torch
        import torch.distributed as dist

        from . import _native as nv

        lib = nv.lib()
        dev = self.slice.device
        a, b = self.row_range
        nloc = b - a
        rows = [self.bounds[r + 1] - self.bounds[r] for r in range(self.world)]
        ref = torch.zeros(max(self.N, 1), dtype=torch.uint8, device=dev)  # my marks of ALL columns, in row order
        nv.check(lib.meld_b200_graph_mark_columns(self.slice._h, nv.ptr(ref), nv.current_stream_ptr()), "graph_mark_columns")
        recv = torch.empty(max(self.world * nloc, 1), dtype=torch.uint8, device=dev)
        # rank w's marks of MY rows arrive at recv[w * nloc : (w + 1) * nloc]
        dist.all_to_all_single(recv[: self.world * nloc], ref[: self.N], output_split_sizes=[nloc] * self.world,
                               input_split_sizes=rows, group=self.group)
        nv.check(lib.meld_b200_graph_set_halo(self.slice._h, nv.ptr(recv), max(nloc, 1), self.world, self.rank,
                                              nv.current_stream_ptr()), "graph_set_halo")
        peers = [w for w in range(self.world) if w != self.rank]
        # fraction of (row, peer) pairs that actually travel per term (1.0 = plain all-gather)
        self.halo_fraction = float(recv[: self.world * nloc].view(self.world, nloc)[peers].float().mean()) if nloc else 0.0
        torch.cuda.current_stream().synchronize()  # recv / ref die here

    # ---- the two native operations of the NCCL path (overridden by the CPU / gloo test) -----------------
    def _permute(self, src, p, to_internal, dst):
        from . import _native as nv

        nv.check(nv.lib().meld_b200_graph_permute_signal(self.graph._h, nv.ptr(src), p, int(to_internal), nv.ptr(dst),
                                                         nv.current_stream_ptr()), "graph_permute_signal")

    def _step(self, cur, told, tnew, Rloc, p, alpha, shift, gamma, c, c_cur, accumulate):
        from . import _native as nv

        nv.check(
            nv.lib().meld_b200_cheby_step(self.slice._h, nv.ptr(cur), nv.ptr(told), nv.ptr(tnew), nv.ptr(Rloc), p,
                                          float(alpha), float(shift), float(gamma), float(c), float(c_cur),
                                          int(bool(accumulate)), nv.current_stream_ptr()),
            "cheby_step",
        )

    def estimate_lmax(self, max_iters=0, rel_tol=0.0):
        """1.01 x lambda_max by the row-partitioned Lanczos (p2p mode; every rank returns the same value bit for bit).
        NCCL mode: the full graph's own (replicated) estimate."""
        if self.mode != "p2p":
            return self.graph.estimate_lmax(max_iters, rel_tol)
        from . import _native as nv

        lmax, iters = C.c_double(), C.c_int()
        nv.check(
            nv.lib().meld_b200_estimate_lmax_dist(self.slice._h, self.ctx._h, int(max_iters), float(rel_tol),
                                                  nv.current_stream_ptr(), C.byref(lmax), C.byref(iters)),
            "estimate_lmax_dist",
        )
        self.lmax_iters = iters.value
        return lmax.value

    # ---- p2p ---------------------------------------------------------------------------------------------
    def _apply_p2p(self, lmax, coeffs, S):
        import torch

        from . import _native as nv

        R = torch.empty_like(S)
        cptr = coeffs.ctypes.data_as(C.POINTER(C.c_double))
        nv.check(
            nv.lib().meld_b200_cheby_filter_dist(self.slice._h, self.ctx._h, float(lmax), cptr, len(coeffs), nv.ptr(S),
                                                 int(S.shape[1]), nv.ptr(R), nv.current_stream_ptr()),
            "cheby_filter_dist",
        )
        return R

    # ---- nccl --------------------------------------------------------------------------------------------
    def _apply_nccl(self, lmax, coeffs, S):
        import torch
        import torch.distributed as dist

        p = int(S.shape[1])
        a, b = self.row_range
        nloc = b - a
        rows = self.world * self.chunk
        if p not in self._nccl_bufs:
            z = lambda r: torch.zeros((r, p), dtype=torch.float64, device=S.device)  # noqa: E731
            self._nccl_bufs[p] = (z(rows), z(rows), z(self.chunk))
        bufA, bufB, Rloc = self._nccl_bufs[p]
        self._permute(S, p, True, bufA)  # T_0 = S in graph order (every rank the whole signal)
        cur, old = bufA, bufB
        a1 = float(lmax) / 2.0
        m = len(coeffs) - 1
        lo = self.rank * self.chunk
        for k in range(1, m + 1):
            mine = old[lo:lo + self.chunk]  # T_k over T_{k-2}, in place; the in-place all-gather publishes it
            if nloc > 0:
                self._step(cur, mine if k >= 2 else None, mine if k < m else None, Rloc, p,
                           (1.0 if k == 1 else 2.0) / a1, a1, 1.0 if k >= 2 else 0.0, coeffs[k],
                           0.5 * coeffs[0] if k == 1 else 0.0, k >= 2)
            if k < m:
                if self.world > 1:
                    dist.all_gather_into_tensor(old, mine, group=self.group)
                cur, old = old, cur
        if self.world > 1:
            dist.all_gather_into_tensor(old, Rloc, group=self.group)
        else:
            old[: self.chunk] = Rloc
        R = torch.empty_like(S)
        self._permute(old, p, False, R)
        return R

    def apply(self, lmax, coeffs, S):
        import torch

        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        if coeffs.shape[0] < 2:
            raise TypeError("The coefficients have an invalid shape")
        S = S.contiguous()
        N, p = S.shape
        if N != self.N:
            raise ValueError("signal has {} rows, graph has {} nodes".format(N, self.N))
        fn = self._apply_p2p if self.mode == "p2p" else self._apply_nccl
        if p <= self.p_max:
            return fn(lmax, coeffs, S)
        R = torch.empty_like(S)
        for j in range(0, p, self.p_max):
            R[:, j:j + self.p_max] = fn(lmax, coeffs, S[:, j:j + self.p_max].contiguous())
        return R

    def close(self):
        if self._own_slice:
            self.slice.close()  # the peer context is shared (shared_context) and lives until the process ends

    def gather_scipy_L(self):
        """COLLECTIVE: the whole Laplacian in the caller's cell order as a scipy CSR matrix on rank 0 (None elsewhere)
        -- parity checks and debugging only; the data path never assembles it."""
        import torch.distributed as dist
        from scipy import sparse

        local = self.slice.to_scipy_L()  # rows [a, b) in the internal order, global columns
        if self.world == 1:
            parts = [local]
        else:
            parts = [None] * self.world if self.rank == 0 else None
            dist.gather_object((local.indptr, local.indices, local.data), parts, dst=0, group=self.group)
            if self.rank != 0:
                return None
            parts = [sparse.csr_matrix((d, i, p), shape=(len(p) - 1, self.N)) for p, i, d in parts]
        M = sparse.vstack(parts).tocsr()
        perm = self.slice.permutation()
        if perm is not None:
            coo = M.tocoo()
            M = sparse.csr_matrix((coo.data, (perm[coo.row], perm[coo.col])), shape=M.shape)
        M.sort_indices()
        return M
