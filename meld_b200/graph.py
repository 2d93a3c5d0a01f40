"""Device-resident graph handle (the object ``MELD.graph`` holds).

Stands in for the ``graphtools`` kNN + PyGSP graph that ``MELD.fit`` builds by
inheritance (reference ``meld/meld.py:13,117-118,273``) and that
``meld/filter.py:39-59`` consumes through the duck-typed protocol
``graph.estimate_lmax()`` / ``graph.lmax`` / ``graph.L`` / ``graph.N``.  The
Laplacian lives in HBM as CSR inside libmeld_b200; this class only owns the
opaque handle and the parameters the graph was built with.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nv


class DeviceGraph:
    """kNN alpha-decay graph / combinatorial Laplacian resident on one GPU."""

    def __init__(self, handle, params=None, device=None):
        self._h = C.c_void_p(handle)
        self.params = dict(params or {})
        self.device = device
        self._lmax = None
        self.lmax_iters = None
        n_rows, n_cols, row0, nnz = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        nv.check(
            nv.lib().meld_b200_graph_info(self._h, C.byref(n_rows), C.byref(n_cols), C.byref(row0), C.byref(nnz)),
            "graph_info",
        )
        self.n_rows, self.n_cols, self.row0, self.nnz = n_rows.value, n_cols.value, row0.value, nnz.value

    # ---- construction ----------------------------------------------------------------
    @classmethod
    def from_data(cls, data_nu, knn=5, decay=40.0, thresh=1e-4, anisotropy=1.0, bandwidth_scale=1.0,
                  keep_knn_kernel=False, simt_search=False):
        """Build the graph from (N, d) float64 data already reduced (``data_nu``).

        ``data_nu`` may be a numpy array (copied to the current device) or a CUDA
        torch tensor.  Mirrors ``graphtools.Graph(data, knn, decay, thresh,
        anisotropy, use_pygsp=True)`` on its kNN branch.
        """
        torch = nv.require_cuda()
        X = _as_device_f64(torch, data_nu)
        if X.dim() != 2:
            raise ValueError("Expected a 2D matrix. Got shape {}".format(tuple(X.shape)))
        N, d = X.shape
        knn = _clamp_knn(knn, N)
        decay = 0.0 if decay is None else decay  # C-ABI: 0 = the reference's decay=None (binary kNN kernel)
        flags = (nv.FLAG_KEEP_KNN_KERNEL if keep_knn_kernel else 0) | (nv.FLAG_SIMT_SEARCH if simt_search else 0)
        out = C.c_void_p()
        nv.check(
            nv.lib().meld_b200_knn_graph_build(
                nv.ptr(X), N, d, int(knn), float(decay), float(thresh), float(anisotropy), float(bandwidth_scale),
                flags, nv.current_stream_ptr(), C.byref(out),
            ),
            "knn_graph_build",
        )
        params = dict(knn=knn, decay=decay, thresh=thresh, anisotropy=anisotropy, bandwidth_scale=bandwidth_scale)
        return cls(out.value, params, device=X.device)

    @classmethod
    def from_data_dense(cls, data_nu, knn=5, decay=40.0, anisotropy=1.0, bandwidth_scale=1.0, keep_knn_kernel=False):
        """``graphtools.Graph(..., thresh=0)``: the dense "exact" graph (TraditionalGraph) -- the same kernel for
        every pair of cells whose value has not underflowed to zero.  Test-scale only (N <= 16384); it is what the
        reference's own known-answer test builds (``test/test_meld.py:58-67``)."""
        torch = nv.require_cuda()
        X = _as_device_f64(torch, data_nu)
        if X.dim() != 2:
            raise ValueError("Expected a 2D matrix. Got shape {}".format(tuple(X.shape)))
        N, d = X.shape
        knn = _clamp_knn(knn, N)
        if decay is None:
            raise ValueError("the dense exact graph needs a decay (decay=None selects the binary kNN kernel)")
        out = C.c_void_p()
        nv.check(
            nv.lib().meld_b200_dense_graph_build(nv.ptr(X), N, d, int(knn), float(decay), float(anisotropy),
                                                 float(bandwidth_scale), nv.FLAG_KEEP_KNN_KERNEL if keep_knn_kernel else 0,
                                                 nv.current_stream_ptr(), C.byref(out)),
            "dense_graph_build",
        )
        params = dict(knn=knn, decay=decay, thresh=0, anisotropy=anisotropy, bandwidth_scale=bandwidth_scale)
        return cls(out.value, params, device=X.device)

    def to_dense_L(self):
        """L as a dense (N, N) float64 CUDA tensor in the CALLER's cell order (exact solver, N <= 16384)."""
        torch = nv.require_cuda()
        if self.n_rows != self.n_cols or self.row0 != 0:
            raise ValueError("needs the full operator")
        if self.n_cols > 16384:
            raise NotImplementedError(
                "solver='exact' needs the dense eigendecomposition of L (O(N^3)); it is limited to N <= 16384 "
                "(got {}); use solver='chebyshev'".format(self.n_cols))
        dev = self.device
        indptr = torch.empty(self.n_rows + 1, dtype=torch.int64, device=dev)
        indices = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)
        data = torch.empty(max(self.nnz, 1), dtype=torch.float64, device=dev)
        nv.check(nv.lib().meld_b200_graph_export_csr(self._h, nv.ptr(indptr), nv.ptr(indices), nv.ptr(data),
                                                     nv.current_stream_ptr()), "graph_export_csr")
        rows = torch.repeat_interleave(torch.arange(self.n_rows, device=dev), indptr[1:] - indptr[:-1])
        cols = indices[: self.nnz].long()
        ident = C.c_int(1)
        perm = torch.empty(self.n_rows, dtype=torch.int32, device=dev)
        nv.check(nv.lib().meld_b200_graph_permutation(self._h, nv.ptr(perm), C.byref(ident), nv.current_stream_ptr()),
                 "graph_permutation")
        if not ident.value:
            rows, cols = perm.long()[rows], perm.long()[cols]
        Ld = torch.zeros((self.n_rows, self.n_cols), dtype=torch.float64, device=dev)
        Ld[rows, cols] = data[: self.nnz]
        return Ld

    # ---- sharded construction (one process per GPU) ------------------------------------------------
    @staticmethod
    def shard_bounds(n, world, align=512):
        """Row ranges of the candidate search per rank: contiguous runs of ``align``-row tiles."""
        tiles = (n + align - 1) // align
        base, rem = divmod(tiles, world)
        bounds, t = [0], 0
        for r in range(world):
            t += base + (1 if r < rem else 0)
            bounds.append(min(n, t * align))
        return bounds

    @staticmethod
    def candidates(X, row_begin, row_end, knn, decay, thresh, bandwidth_scale=1.0, simt_search=False):
        """Stage 1 of a build for query rows [row_begin, row_end) in the internal cell order: returns
        CUDA tensors (counts int64, cand int32, d2 float64, eps float64, perm int32 or None)."""
        torch = nv.require_cuda()
        N, d = X.shape
        dev = X.device
        if row_end <= row_begin:
            e = lambda dt: torch.empty(0, dtype=dt, device=dev)  # noqa: E731
            return e(torch.int64), e(torch.int32), e(torch.float64), e(torch.float64), None
        h = C.c_void_p()
        nv.check(
            nv.lib().meld_b200_knn_candidates(nv.ptr(X), N, d, int(knn), float(decay), float(thresh),
                                              float(bandwidth_scale), int(row_begin), int(row_end),
                                              nv.FLAG_SIMT_SEARCH if simt_search else 0, nv.current_stream_ptr(),
                                              C.byref(h)),
            "knn_candidates",
        )
        try:
            nloc, total, has_perm, mx = C.c_int64(), C.c_int64(), C.c_int(), C.c_int64()
            nv.check(nv.lib().meld_b200_cands_info(h, C.byref(nloc), C.byref(total), C.byref(has_perm), C.byref(mx)),
                     "cands_info")
            counts = torch.empty(nloc.value, dtype=torch.int64, device=dev)
            cand = torch.empty(total.value, dtype=torch.int32, device=dev)
            d2 = torch.empty(total.value, dtype=torch.float64, device=dev)
            eps = torch.empty(nloc.value, dtype=torch.float64, device=dev)
            perm = torch.empty(N, dtype=torch.int32, device=dev) if has_perm.value else None
            nv.check(nv.lib().meld_b200_cands_export(h, nv.ptr(counts), nv.ptr(cand), nv.ptr(d2), nv.ptr(eps),
                                                     nv.ptr(perm), nv.current_stream_ptr()), "cands_export")
            torch.cuda.current_stream().synchronize()
        finally:
            nv.lib().meld_b200_cands_destroy(h)
        return counts, cand, d2, eps, perm

    @classmethod
    def from_candidates(cls, N, counts, cand, d2, eps, perm, knn, decay, thresh, anisotropy, bandwidth_scale=1.0,
                        keep_knn_kernel=False, device=None):
        """Stage 2 of a build from the candidates of ALL rows (row order)."""
        out = C.c_void_p()
        nv.check(
            nv.lib().meld_b200_graph_from_candidates(
                int(N), nv.ptr(counts.contiguous()), nv.ptr(cand.contiguous()), nv.ptr(d2.contiguous()),
                int(cand.shape[0]), nv.ptr(eps.contiguous()), nv.ptr(perm), int(knn), float(decay), float(thresh),
                float(anisotropy), float(bandwidth_scale), nv.FLAG_KEEP_KNN_KERNEL if keep_knn_kernel else 0,
                nv.current_stream_ptr(), C.byref(out)),
            "graph_from_candidates",
        )
        params = dict(knn=knn, decay=decay, thresh=thresh, anisotropy=anisotropy, bandwidth_scale=bandwidth_scale)
        return cls(out.value, params, device=device if device is not None else cand.device)

    @classmethod
    def from_data_sharded(cls, data_nu, knn=5, decay=40.0, thresh=1e-4, anisotropy=1.0, bandwidth_scale=1.0, group=None):
        """Build the graph with the candidate search (the dominant stage) sharded over the ranks of
        ``torch.distributed`` by query rows; one NCCL all-gather of the per-row candidate lists follows,
        then every rank assembles the same Laplacian.  Every rank passes the same data."""
        torch = nv.require_cuda()
        import torch.distributed as dist

        import os
        import time

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        timing = os.environ.get("MELD_B200_TIMING") and rank == 0
        t0 = time.perf_counter()

        def lap(what):
            nonlocal t0
            if timing:
                torch.cuda.synchronize()
                print("[meld_b200 timing] sharded: {:24s} {:9.3f} ms".format(what, 1e3 * (time.perf_counter() - t0)),
                      flush=True)
                t0 = time.perf_counter()

        X = _as_device_f64(torch, data_nu)
        N, d = X.shape
        knn = _clamp_knn(knn, N)
        decay = 0.0 if decay is None else decay
        bounds = cls.shard_bounds(N, world)
        lap("input to device")
        dev = X.device
        lib = nv.lib()
        a, b = bounds[rank], bounds[rank + 1]
        # ---- stage 1 on this rank's query rows
        h = C.c_void_p()
        nloc = total = 0
        has_perm = 0
        if b > a:
            nv.check(lib.meld_b200_knn_candidates(nv.ptr(X), N, d, int(knn), float(decay), float(thresh),
                                                  float(bandwidth_scale), int(a), int(b), 0, nv.current_stream_ptr(),
                                                  C.byref(h)), "knn_candidates")
            c_n, c_t, c_p, c_m = C.c_int64(), C.c_int64(), C.c_int(), C.c_int64()
            nv.check(lib.meld_b200_cands_info(h, C.byref(c_n), C.byref(c_t), C.byref(c_p), C.byref(c_m)), "cands_info")
            nloc, total, has_perm = c_n.value, c_t.value, c_p.value
        lap("stage 1 (local rows)")
        try:
            # ---- one exchange: every rank's [d2 | eps | counts | cand] block, padded to the longest, all-gathered
            sizes = torch.tensor([nloc, total, has_perm], dtype=torch.int64, device=dev)
            all_sizes = torch.empty(3 * world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(all_sizes, sizes, group=group)
            all_sizes = all_sizes.cpu().view(world, 3)
            rows, tots = all_sizes[:, 0].tolist(), all_sizes[:, 1].tolist()
            any_perm = bool(all_sizes[:, 2].max().item())
            blk = max(12 * t + 16 * r for r, t in zip(rows, tots))
            blk = (blk + 15) // 16 * 16

            def views(buf, r, t):  # typed views of one rank's block
                o = 0
                v_d2 = buf[o:o + 8 * t].view(torch.float64)
                o += 8 * t
                v_eps = buf[o:o + 8 * r].view(torch.float64)
                o += 8 * r
                v_cnt = buf[o:o + 8 * r].view(torch.int64)
                o += 8 * r
                v_cand = buf[o:o + 4 * t].view(torch.int32)
                return v_d2, v_eps, v_cnt, v_cand

            recv = torch.empty(world * blk, dtype=torch.uint8, device=dev)
            mine = recv[rank * blk:(rank + 1) * blk]  # in-place all-gather: this rank's block sits in its slot
            perm = torch.empty(N, dtype=torch.int32, device=dev) if any_perm else None
            if nloc > 0:
                v_d2, v_eps, v_cnt, v_cand = views(mine, nloc, total)
                nv.check(lib.meld_b200_cands_export(h, nv.ptr(v_cnt), nv.ptr(v_cand), nv.ptr(v_d2), nv.ptr(v_eps),
                                                    nv.ptr(perm) if has_perm else C.c_void_p(0),
                                                    nv.current_stream_ptr()), "cands_export")
            if world > 1:
                dist.all_gather_into_tensor(recv, mine, group=group)
            parts = [views(recv[r * blk:(r + 1) * blk], rows[r], tots[r]) for r in range(world)]
            d2 = torch.cat([p_[0] for p_ in parts])
            eps = torch.cat([p_[1] for p_ in parts])
            counts = torch.cat([p_[2] for p_ in parts])
            cand = torch.cat([p_[3] for p_ in parts])
            if any_perm and world > 1:  # every rank with rows derived the same order; take the first one's
                src = next(r for r in range(world) if rows[r] > 0)
                dist.broadcast(perm, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
            lap("all-gather")
        finally:
            if h.value:
                torch.cuda.current_stream().synchronize()
                lib.meld_b200_cands_destroy(h)
        g = cls.from_candidates(N, counts, cand, d2, eps, perm, knn, decay, thresh, anisotropy, bandwidth_scale,
                                device=dev)
        lap("stage 2 (replicated)")
        return g

    @classmethod
    def from_data_sharded_rows(cls, data_nu, knn=5, decay=40.0, thresh=1e-4, anisotropy=1.0, bandwidth_scale=1.0,
                               group=None):
        """Fully row-partitioned build: every rank ends with ONLY its rows of L (a slice handle, internal cell order) --
        nothing O(nnz) is replicated or exchanged.  Stage 1 as in :meth:`from_data_sharded`; then an
        all-gather of eps (8 N bytes), an all-to-all-v of the mirrored entries (~10 x 16 bytes per row) and an
        all-gather of the kernel row sums (8 N bytes) around the three ``meld_b200_stage2_*`` calls.  Returns
        ``(slice_graph, bounds)``; every rank passes the same data."""
        torch = nv.require_cuda()
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        X = _as_device_f64(torch, data_nu)
        N, d = X.shape
        knn = _clamp_knn(knn, N)
        decay = 0.0 if decay is None else decay
        bounds = cls.shard_bounds(N, world)
        if any(bounds[r + 1] <= bounds[r] for r in range(world)):
            raise ValueError("{} cells are too few for {} row-partitioned ranks (512-row tiles)".format(N, world))
        dev, lib = X.device, nv.lib()
        a, b = bounds[rank], bounds[rank + 1]
        nloc = b - a
        rows = [bounds[r + 1] - bounds[r] for r in range(world)]
        rows_max = max(rows)
        h, st = C.c_void_p(), C.c_void_p()
        nv.check(lib.meld_b200_knn_candidates(nv.ptr(X), N, d, int(knn), float(decay), float(thresh),
                                              float(bandwidth_scale), int(a), int(b), 0, nv.current_stream_ptr(),
                                              C.byref(h)), "knn_candidates")
        try:
            c_n, c_t, c_p, c_m = C.c_int64(), C.c_int64(), C.c_int(), C.c_int64()
            nv.check(lib.meld_b200_cands_info(h, C.byref(c_n), C.byref(c_t), C.byref(c_p), C.byref(c_m)), "cands_info")

            def allgather_rows(local):  # (rows_max,) padded slices -> the N-vector in row order
                out = torch.empty(world * rows_max, dtype=local.dtype, device=dev)
                dist.all_gather_into_tensor(out, local, group=group)
                if all(r == rows_max for r in rows):
                    return out
                return torch.cat([out[r * rows_max: r * rows_max + rows[r]] for r in range(world)])

            eps_loc = torch.zeros(rows_max, dtype=torch.float64, device=dev)
            nv.check(lib.meld_b200_cands_export(h, C.c_void_p(0), C.c_void_p(0), C.c_void_p(0), nv.ptr(eps_loc),
                                                C.c_void_p(0), nv.current_stream_ptr()), "cands_export")
            eps_full = allgather_rows(eps_loc)
            hb = (C.c_int64 * (world + 1))(*bounds)
            hc = (C.c_int64 * world)()
            nv.check(lib.meld_b200_stage2_begin(h, nv.ptr(eps_full), hb, world, int(knn), float(decay), float(thresh),
                                                float(anisotropy), float(bandwidth_scale), nv.current_stream_ptr(),
                                                C.byref(st), hc), "stage2_begin")
            send_counts = [int(v) for v in hc]
            sc = torch.tensor(send_counts, dtype=torch.int64, device=dev)
            rc = torch.empty_like(sc)
            dist.all_to_all_single(rc, sc, group=group)
            recv_counts = rc.cpu().tolist()
            send = torch.empty(max(sum(send_counts), 1) * 16, dtype=torch.uint8, device=dev)
            nv.check(lib.meld_b200_stage2_records(st, nv.ptr(send), nv.current_stream_ptr()), "stage2_records")
            recv = torch.empty(max(sum(recv_counts), 1) * 16, dtype=torch.uint8, device=dev)
            dist.all_to_all_single(recv[: sum(recv_counts) * 16], send[: sum(send_counts) * 16],
                                   output_split_sizes=[16 * v for v in recv_counts],
                                   input_split_sizes=[16 * v for v in send_counts], group=group)
            q_loc = torch.zeros(rows_max, dtype=torch.float64, device=dev)
            nv.check(lib.meld_b200_stage2_assemble(st, nv.ptr(recv), int(sum(recv_counts)), nv.current_stream_ptr(),
                                                   nv.ptr(q_loc)), "stage2_assemble")
            q_full = allgather_rows(q_loc)
            out = C.c_void_p()
            nv.check(lib.meld_b200_stage2_finish(st, nv.ptr(q_full), nv.current_stream_ptr(), C.byref(out)),
                     "stage2_finish")
            torch.cuda.current_stream().synchronize()
        finally:
            if st.value:
                lib.meld_b200_stage2_destroy(st)
            lib.meld_b200_cands_destroy(h)
        params = dict(knn=knn, decay=decay, thresh=thresh, anisotropy=anisotropy, bandwidth_scale=bandwidth_scale)
        return cls(out.value, params, device=dev), bounds

    @classmethod
    def from_scipy(cls, L, row0=0, n_cols=None, params=None):
        """Adopt a Laplacian (or a row slice of one) built elsewhere, e.g. by the test oracle."""
        torch = nv.require_cuda()
        L = L.tocsr()
        L.sort_indices()
        n_rows = L.shape[0]
        n_cols = L.shape[1] if n_cols is None else n_cols
        dev = torch.device("cuda", torch.cuda.current_device())
        indptr = torch.from_numpy(np.ascontiguousarray(L.indptr, dtype=np.int64)).to(dev)
        indices = torch.from_numpy(np.ascontiguousarray(L.indices, dtype=np.int32)).to(dev)
        data = torch.from_numpy(np.ascontiguousarray(L.data, dtype=np.float64)).to(dev)
        out = C.c_void_p()
        nv.check(
            nv.lib().meld_b200_graph_from_csr(
                n_rows, n_cols, row0, L.nnz, nv.ptr(indptr), nv.ptr(indices), nv.ptr(data), nv.current_stream_ptr(),
                C.byref(out),
            ),
            "graph_from_csr",
        )
        torch.cuda.current_stream().synchronize()
        return cls(out.value, params, device=dev)

    def row_slice(self, row_begin, row_end):
        """Rows [row_begin, row_end) of L in the INTERNAL cell order as their own handle (what one rank of a
        row-partitioned filter holds); it keeps the full graph's cell order for signal conversion."""
        out = C.c_void_p()
        nv.check(nv.lib().meld_b200_graph_row_slice(self._h, int(row_begin), int(row_end), nv.current_stream_ptr(),
                                                    C.byref(out)), "graph_row_slice")
        g = DeviceGraph(out.value, self.params, device=self.device)
        g._lmax = self._lmax
        return g

    def permute_signal(self, x, to_internal=True):
        """(N, p) float64 CUDA tensor between the caller's cell order and the graph's internal one."""
        torch = nv.require_cuda()
        x = x.contiguous()
        out = torch.empty_like(x)
        nv.check(nv.lib().meld_b200_graph_permute_signal(self._h, nv.ptr(x), int(x.shape[1]), int(bool(to_internal)),
                                                         nv.ptr(out), nv.current_stream_ptr()), "graph_permute_signal")
        return out

    # ---- the protocol meld/filter.py relies on ----------------------------------------
    @property
    def N(self):
        return self.n_cols

    @property
    def knn(self):
        return self.params.get("knn")

    def estimate_lmax(self, max_iters=0, rel_tol=0.0):
        """1.01 x largest Laplacian eigenvalue by Lanczos on the GPU (cached, like PyGSP)."""
        if self._lmax is None:
            lmax, iters = C.c_double(), C.c_int()
            nv.check(
                nv.lib().meld_b200_estimate_lmax(
                    self._h, int(max_iters), float(rel_tol), nv.current_stream_ptr(), C.byref(lmax), C.byref(iters)
                ),
                "estimate_lmax",
            )
            self._lmax, self.lmax_iters = lmax.value, iters.value
        return self._lmax

    @property
    def lmax(self):
        return self.estimate_lmax()

    @lmax.setter
    def lmax(self, value):
        """Inject lmax (parity runs share the oracle's ARPACK value)."""
        self._lmax = None if value is None else float(value)

    # ---- exporters (debugging / parity) -------------------------------------------------
    def _export(self, nnz, fn, what):
        torch = nv.require_cuda()
        dev = self.device
        indptr = torch.empty(self.n_rows + 1, dtype=torch.int64, device=dev)
        indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        data = torch.empty(max(nnz, 1), dtype=torch.float64, device=dev)
        nv.check(fn(self._h, nv.ptr(indptr), nv.ptr(indices), nv.ptr(data), nv.current_stream_ptr()), what)
        from scipy import sparse

        M = sparse.csr_matrix(
            (data[:nnz].cpu().numpy(), indices[:nnz].cpu().numpy(), indptr.cpu().numpy()),
            shape=(self.n_rows, self.n_cols),
        )
        if self.n_rows != self.n_cols:  # a row slice: rows [row0, row0 + n_rows) in the INTERNAL cell order
            M.sort_indices()
            return M
        perm = self.permutation()
        if perm is not None:  # rows / columns back to the caller's cell order
            coo = M.tocoo()
            M = sparse.csr_matrix((coo.data, (perm[coo.row], perm[coo.col])), shape=M.shape)
        M.sort_indices()
        return M

    def permutation(self):
        """Internal cell order: graph row a is the caller's cell ``perm[a]`` (None = identity)."""
        torch = nv.require_cuda()
        ident = C.c_int(1)
        perm = torch.empty(self.n_cols, dtype=torch.int32, device=self.device)
        nv.check(nv.lib().meld_b200_graph_permutation(self._h, nv.ptr(perm), C.byref(ident), nv.current_stream_ptr()),
                 "graph_permutation")
        if ident.value:
            return None
        return perm.cpu().numpy().astype(np.int64)

    def to_scipy_L(self):
        return self._export(self.nnz, nv.lib().meld_b200_graph_export_csr, "graph_export_csr")

    @property
    def L(self):
        return self.to_scipy_L()

    def to_scipy_W(self):
        L = self.to_scipy_L().tolil()
        L.setdiag(0)
        W = (-L).tocsr()
        W.eliminate_zeros()
        return W

    def to_scipy_knn_kernel(self):
        nnz = C.c_int64()
        nv.check(nv.lib().meld_b200_graph_knn_kernel_nnz(self._h, C.byref(nnz)), "graph_knn_kernel_nnz")
        return self._export(nnz.value, nv.lib().meld_b200_graph_export_knn_kernel, "graph_export_knn_kernel")

    def build_stats(self):
        arr = (C.c_int64 * 8)()
        nv.check(nv.lib().meld_b200_graph_build_stats(self._h, arr), "graph_build_stats")
        keys = ["search_passes", "max_candidates", "candidate_cap", "overflow_rows", "search_impl", "reserved5",
                "reserved6", "row_blocks"]
        return {k: int(arr[i]) for i, k in enumerate(keys)}

    def build_times(self):
        arr = (C.c_double * 8)()
        nv.check(nv.lib().meld_b200_graph_build_times(self._h, arr), "graph_build_times")
        return {"pass1_ms": arr[0], "pass2_ms": arr[1], "flops_per_pass": arr[2], "flops_pass1": arr[3] or arr[2],
                "flops_unpruned_pass": arr[4] or arr[2]}

    # ---- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            nv.lib().meld_b200_graph_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __getstate__(self):
        L = self.to_scipy_L()
        return dict(L=L, row0=self.row0, n_cols=self.n_cols, params=self.params, lmax=self._lmax)

    def __setstate__(self, st):
        g = DeviceGraph.from_scipy(st["L"], row0=st["row0"], n_cols=st["n_cols"], params=st["params"])
        self.__dict__.update(g.__dict__)
        g._h = C.c_void_p(0)
        self._lmax = st["lmax"]


def _clamp_knn(knn, N):
    """graphtools ``kNNGraph.__init__``: knn above n_samples - 2 is clamped with a warning, not an error."""
    if knn > N - 2:
        import warnings

        warnings.warn("Cannot set knn ({k}) to be greater than  n_samples - 2 ({n}). Setting knn={n}".format(
            k=knn, n=N - 2))
        knn = N - 2
    if knn < 1:
        raise ValueError("Expected at least 3 cells to build a kNN graph, got {}".format(N))
    return int(knn)


_STAGE_BYTES = 32 << 20  # pinned staging buffer of the pageable-input copy pipeline
_stage_bufs = []


def _pageable_to_device(torch, t):
    """Host tensor in PAGEABLE memory -> device, through a 3-deep ring of pinned staging buffers: the (multi-threaded)
    host copy of chunk i + 1 into a staging buffer overlaps the DMA of chunk i.  A plain ``.cuda()`` of pageable
    memory lets the driver stage it on one thread (~15 GB/s); this is the drop-in case -- users hand over ordinary
    numpy arrays."""
    flat = t.reshape(-1)
    n = flat.numel()
    per = max(1, _STAGE_BYTES // flat.element_size())
    out = torch.empty(n, dtype=flat.dtype, device="cuda")
    while len(_stage_bufs) < 3:
        _stage_bufs.append([torch.empty(_STAGE_BYTES, dtype=torch.uint8).pin_memory(), None])
    stream = torch.cuda.current_stream()
    for i, s in enumerate(range(0, n, per)):
        e = min(n, s + per)
        slot = _stage_bufs[i % 3]
        if slot[1] is not None:
            slot[1].synchronize()  # the DMA that last read this staging buffer has finished
        stage = slot[0].view(flat.dtype)[: e - s]
        stage.copy_(flat[s:e])
        out[s:e].copy_(stage, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(stream)
        slot[1] = ev
    return out.reshape(t.shape)


def _as_device_f64(torch, x):
    """numpy / DataFrame / torch tensor -> contiguous float64 CUDA tensor."""
    if isinstance(x, torch.Tensor):
        t = x
    else:
        arr = np.asarray(getattr(x, "values", x))
        if arr.dtype != np.float64 and arr.dtype != np.float32:
            arr = arr.astype(np.float64)
        t = nv.from_numpy_readonly(np.ascontiguousarray(arr))
    if not t.is_cuda:
        if t.is_contiguous() and t.numel() * t.element_size() >= (64 << 20) and not t.is_pinned():
            t = _pageable_to_device(torch, t)
        else:
            t = t.cuda(non_blocking=True)
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    return t.contiguous()
