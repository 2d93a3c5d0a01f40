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
        if knn + 1 > N:
            raise ValueError("knn + 1 = {} exceeds the number of cells {}".format(knn + 1, N))
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
        perm = torch.empty(self.n_rows, dtype=torch.int32, device=self.device)
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
        keys = ["search_passes", "max_candidates", "candidate_cap", "overflow_rows", "search_impl", "dict_total",
                "direct_blocks", "row_blocks"]
        return {k: int(arr[i]) for i, k in enumerate(keys)}

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


def _as_device_f64(torch, x):
    """numpy / DataFrame / torch tensor -> contiguous float64 CUDA tensor."""
    if isinstance(x, torch.Tensor):
        t = x
        if not t.is_cuda:
            t = t.cuda(non_blocking=True)
    else:
        arr = np.asarray(getattr(x, "values", x))
        if arr.dtype != np.float64 and arr.dtype != np.float32:
            arr = arr.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(arr)).cuda(non_blocking=True)
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    return t.contiguous()
