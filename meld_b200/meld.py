"""``MELD`` estimator with the reference's public API, driving the B200 engine.

Mirrors ``meld/meld.py`` of KrishnaswamyLab/MELD (constructor defaults ``:94-107``,
``set_params`` reset semantics ``:127-141``, label handling ``:143-191``,
``transform`` ``:193-250``, ``fit_transform`` ``:252-274``) and the part of
``graphtools.estimator.GraphEstimator`` it inherits (``fit``, graph parameters
``knn=5, decay=40, n_pca=100, thresh=1e-4``).  Argument names, defaults, returned
DataFrame layout and exception texts are the reference's; the graph build and the
Chebyshev filter run in libmeld_b200 on the GPU.  There is no CPU path.
"""

from __future__ import annotations

import numbers
import os
import threading
import time
import weakref

import numpy as np
import pandas as pd

from . import _native as nv
from . import filter as _filter
from . import utils
from .graph import DeviceGraph, _as_device_f64

_FILTER_PARAMS = ["beta", "offset", "order", "solver", "chebyshev_order", "lap_type", "filter"]
_GRAPH_PARAMS = ["knn", "decay", "thresh", "n_pca", "distance", "anisotropy", "random_state", "bandwidth_scale"]


def _check_positive(**params):
    for p in params:
        if not isinstance(params[p], numbers.Number) or params[p] <= 0:
            raise ValueError("Expected {} > 0, got {}".format(p, params[p]))


def _check_int(**params):
    for p in params:
        if not isinstance(params[p], numbers.Integral):
            raise ValueError("Expected {} integer, got {}".format(p, params[p]))


def _check_in(choices, **params):
    for p in params:
        if params[p] not in choices:
            raise ValueError("{} value {} not recognized. Choose from {}".format(p, params[p], choices))


class MELD(object):
    """MELD operator for filtering sample-indicator signals over a cell-similarity graph.

    Parameters
    ----------
    beta : int, optional, Default: 60
        Amount of smoothing to apply.
    offset : float, optional, Default: 0
        Shift of the filter in the (normalised) eigenvalue spectrum, in [0, 1].
    order : int, optional, Default: 1
        Falloff / smoothness of the filter.
    filter : str, optional, Default: 'heat'
        'heat' or 'laplacian'.
    solver : str, optional, Default: 'chebyshev'
        'chebyshev' is the engine's hot path; 'exact' (dense eigendecomposition, library ``eigh``) is
        available for test-scale graphs (N <= 16384), e.g. the reference's own known-answer test.
    chebyshev_order : int, optional, Default: 50
    lap_type : ('combinatorial', 'normalized'), Default: 'combinatorial'
        Validated and stored; like the reference it is never forwarded to the graph,
        whose Laplacian is always combinatorial.
    sample_normalize : bool, optional, Default: True
        Column-normalise the indicator vectors to sum 1.
    anisotropy : float, optional, Default: 1
    **kwargs : graph parameters (``knn=5, decay=40, n_pca=100, thresh=1e-4,
        distance='euclidean', n_jobs=1, random_state=None, verbose=1,
        bandwidth_scale=1.0``).
    """

    def __init__(self, beta=60, offset=0, order=1, filter="heat", solver="chebyshev", chebyshev_order=50,
                 lap_type="combinatorial", sample_normalize=True, anisotropy=1, n_landmark=None, **kwargs):
        self.graph = None
        self.X = None
        self.sample_densities = None
        self.filt = None
        self.beta = beta
        self.offset = offset
        self.order = order
        self.solver = solver
        self.chebyshev_order = chebyshev_order
        self.lap_type = lap_type
        self.filter = filter
        self.sample_normalize = sample_normalize

        kwargs.pop("use_pygsp", None)
        self.knn = kwargs.pop("knn", 5)
        self.decay = kwargs.pop("decay", 40)
        self.n_pca = kwargs.pop("n_pca", 100)
        self.thresh = kwargs.pop("thresh", 1e-4)
        self.distance = kwargs.pop("distance", "euclidean")
        self.n_jobs = kwargs.pop("n_jobs", 1)
        self.random_state = kwargs.pop("random_state", None)
        self.verbose = kwargs.pop("verbose", 1)
        # engine extension: shard the graph build over torch.distributed ranks (one process per GPU)
        self.distributed = bool(kwargs.pop("distributed", False))
        # how a distributed estimator filters: "p2p" = row-partitioned recurrence with peer stores over NVLink,
        # "nccl" = row-partitioned with an NCCL all-gather per term, "replicated" = every rank filters alone
        self.dist_mode = kwargs.pop("dist_mode", "p2p")
        _check_in(["p2p", "nccl", "replicated"], dist_mode=self.dist_mode)
        # how a distributed estimator assembles the graph: "rows" = every rank only its rows (nothing O(nnz) replicated;
        # needs dist_mode="p2p"), "replicated" = candidate lists all-gathered, every rank assembles the whole graph
        self.dist_build = kwargs.pop("dist_build", "rows" if self.dist_mode == "p2p" else "replicated")
        _check_in(["rows", "replicated"], dist_build=self.dist_build)
        self._sharded = None
        self.anisotropy = anisotropy
        self.n_landmark = n_landmark
        self.kwargs = kwargs  # remaining graphtools.Graph keywords, checked at fit time
        self.timings_ = {}
        self.profile_events = None  # bench hook: list receiving (start, stop, n_launches) CUDA events of the filter

    # ---- validated parameters (graphtools.estimator.attribute equivalents) ---------------
    beta = property(lambda self: self._beta)

    @beta.setter
    def beta(self, v):
        _check_positive(beta=v)
        self._beta = v

    filter = property(lambda self: self._filter)

    @filter.setter
    def filter(self, v):
        _check_in(["heat", "laplacian"], filter=v)
        self._filter = v

    solver = property(lambda self: self._solver)

    @solver.setter
    def solver(self, v):
        _check_in(["chebyshev", "exact"], solver=v)
        self._solver = v

    chebyshev_order = property(lambda self: self._chebyshev_order)

    @chebyshev_order.setter
    def chebyshev_order(self, v):
        _check_int(chebyshev_order=v)
        _check_positive(chebyshev_order=v)
        self._chebyshev_order = v

    lap_type = property(lambda self: self._lap_type)

    @lap_type.setter
    def lap_type(self, v):
        _check_in(["combinatorial", "normalized"], lap_type=v)
        self._lap_type = v

    knn = property(lambda self: self._knn)

    @knn.setter
    def knn(self, v):
        _check_positive(knn=v)
        _check_int(knn=v)
        self._knn = v

    decay = property(lambda self: self._decay)

    @decay.setter
    def decay(self, v):
        if v is not None:
            _check_positive(decay=v)
        self._decay = v

    # ---- cache invalidation ------------------------------------------------------------------
    def _reset_graph(self):
        self._reset_filter()

    def _reset_filter(self):
        self.filt = None
        self.sample_densities = None

    def set_params(self, **params):
        """Set parameters; changing a filter parameter drops the cached densities, changing a
        graph parameter drops the graph as well (reference ``meld/meld.py:127-141``)."""
        for p in _FILTER_PARAMS:
            if p in params and params[p] != getattr(self, p):
                self._reset_filter()
                setattr(self, p, params[p])
                del params[p]
            elif p in params:
                del params[p]
        reset_graph = False
        for p in list(params):
            if p in _GRAPH_PARAMS:
                cur = getattr(self, p) if hasattr(self, p) else self.kwargs.get(p)
                if params[p] != cur:
                    reset_graph = True
                    if hasattr(self, p):
                        setattr(self, p, params[p])
                    else:
                        self.kwargs[p] = params[p]
                del params[p]
            elif p in ("n_jobs", "verbose", "sample_normalize", "n_landmark"):
                setattr(self, p, params.pop(p))
        if params:
            raise ValueError("Invalid parameter(s) for MELD: {}".format(sorted(params)))
        if reset_graph:
            self.graph = None
            self._reset_graph()
        return self

    # ---- fit ----------------------------------------------------------------------------------
    def _log(self, msg):
        if self.verbose:
            print(msg, flush=True)

    def _check_supported(self, extra):
        if self.distance != "euclidean":
            raise NotImplementedError("distance='{}': only 'euclidean' runs on the B200 engine".format(self.distance))
        if self.thresh < 0:
            raise ValueError("Expected thresh >= 0, got {}".format(self.thresh))
        if self.n_landmark is not None:
            raise NotImplementedError("landmark graphs are not available in the B200 engine")
        known = {"bandwidth_scale"}
        unsupported = sorted(k for k in extra if k not in known)
        if unsupported:
            raise NotImplementedError("graph keyword(s) {} are not available in the B200 engine".format(unsupported))

    def _reduce_data(self, X):
        """graphtools ``Data._reduce_data``: randomised PCA when n_pca < min(X.shape) (SURVEY 8a row B').
        Runs on the device (``meld_b200/pca.py``: scikit-learn's randomized-SVD algorithm with the same
        RandomState draw, so ``data_nu`` equals the reference's to rounding)."""
        n_pca = self.n_pca
        if n_pca is None or n_pca >= min(X.shape):
            return X
        torch = nv.require_cuda()
        from . import pca as _pca

        t0 = time.perf_counter()
        self._log("Calculating PCA...")
        Xd = _as_device_f64(torch, X)
        out, self.data_pca = _pca.randomized_pca(Xd, n_pca, random_state=self.random_state)
        torch.cuda.current_stream(out.device).synchronize()
        self.timings_["pca"] = time.perf_counter() - t0
        self._log("Calculated PCA in {:.2f} seconds.".format(self.timings_["pca"]))
        return out

    def _distributed_upload(self, torch, X):
        """Distributed estimators get the same HOST matrix on every rank; every rank needs all of it on its GPU (the
        candidate search reads every column).  Instead of N ranks pulling the whole matrix through their PCIe links at
        once (8 x 800 MB at config 5: 175 ms), every rank uploads its 1 / N of the rows and the pieces are all-gathered
        over NVLink (plumbing: one NCCL call)."""
        if not self.distributed or isinstance(X, torch.Tensor):
            return X
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() < 2:
            return X
        arr = np.asarray(getattr(X, "values", X))
        if arr.ndim != 2 or arr.dtype not in (np.float64, np.float32) or arr.nbytes < (64 << 20):
            return X
        world, rank = dist.get_world_size(), dist.get_rank()
        n, d = arr.shape
        chunk = -(-n // world)
        lo, hi = min(n, rank * chunk), min(n, (rank + 1) * chunk)
        full = torch.empty((world * chunk, d), dtype=torch.from_numpy(arr[:1]).dtype, device="cuda")
        mine = full[rank * chunk:(rank + 1) * chunk]
        if hi > lo:
            src = torch.from_numpy(np.ascontiguousarray(arr[lo:hi]))
            if src.is_pinned() or src.numel() * src.element_size() < (64 << 20):
                mine[: hi - lo].copy_(src, non_blocking=True)
            else:  # pageable: through the pinned staging ring
                from .graph import _pageable_to_device

                mine[: hi - lo].copy_(_pageable_to_device(torch, src))
        dist.all_gather_into_tensor(full, mine)
        return full[:n]

    def fit(self, X, **kwargs):
        """Build the kNN alpha-decay graph on ``X`` (array-like (N, D), CUDA tensor, or a prebuilt
        ``DeviceGraph``).  Stands in for the inherited ``GraphEstimator.fit``."""
        torch = nv.require_cuda()
        nv.lib()
        if isinstance(X, DeviceGraph):
            self.graph = X
            self.X = None
            if self._sharded is not None:
                self._sharded.close()
                self._sharded = None
            self._reset_graph()
            return self
        if hasattr(X, "tocsr") or hasattr(X, "todense"):
            raise NotImplementedError("sparse input is not available in the B200 engine; pass a dense matrix")
        extra = dict(self.kwargs)
        extra.update(kwargs)
        if extra.pop("sample_idx", None) is not None:
            raise NotImplementedError("sample_idx (MNN graphs) is not available in the B200 engine")
        extra.pop("use_pygsp", None)
        self._check_supported(extra)
        build_key = (self.knn, self.decay, self.thresh, self.anisotropy, self.n_pca, self.random_state, self.distributed,
                     self.dist_mode, self.dist_build, tuple(sorted(extra.items())))
        if (self.graph is not None and self.X is not None and getattr(self, "_build_key", None) == build_key
                and _same_data(torch, X, self.X)):
            return self  # same data AND same effective build parameters: keep the graph
        shape = tuple(X.shape)
        if len(shape) != 2:
            raise ValueError("Expected a 2D matrix. Got shape {}".format(shape))
        self._log("Building graph on {} samples and {} features.".format(shape[0], shape[1]))
        t0 = time.perf_counter()
        data_nu = self._reduce_data(self._distributed_upload(torch, X))
        self._log("Calculating graph and diffusion operator...")
        t1 = time.perf_counter()
        slice_bounds = None
        if self.thresh == 0 and self.decay is not None:
            # graphtools.api.Graph: thresh == 0 with a decay selects the dense "exact" graph (TraditionalGraph)
            self.graph = DeviceGraph.from_data_dense(data_nu, knn=self.knn, decay=self.decay, anisotropy=self.anisotropy,
                                                     bandwidth_scale=extra.get("bandwidth_scale", 1.0))
        else:
            build = DeviceGraph.from_data
            rows_only = self.distributed and self.dist_mode == "p2p" and self.dist_build == "rows"
            if rows_only:  # every rank assembles only its rows of L
                build = DeviceGraph.from_data_sharded_rows
            elif self.distributed:  # candidate search sharded over the ranks (same data everywhere), full graph on each
                build = DeviceGraph.from_data_sharded
            built = build(
                data_nu, knn=self.knn, decay=0.0 if self.decay is None else self.decay, thresh=self.thresh,
                anisotropy=self.anisotropy,
                bandwidth_scale=extra.get("bandwidth_scale", 1.0),
            )
            self.graph, slice_bounds = built if rows_only else (built, None)
        if self._sharded is not None:
            self._sharded.close()
            self._sharded = None
        if self.distributed and self.dist_mode != "replicated" and not (self.thresh == 0 and self.decay is not None):
            from .distributed import ShardedFilter

            if slice_bounds is not None:
                self._sharded = ShardedFilter(None, mode="p2p", row_slice=self.graph, bounds=slice_bounds)
            else:
                self._sharded = ShardedFilter(self.graph, mode=self.dist_mode)
        self.timings_["graph"] = time.perf_counter() - t1
        self._log("Calculated graph and diffusion operator in {:.2f} seconds.".format(time.perf_counter() - t0))
        self.X = X
        self._build_key = build_key
        self._reset_graph()
        return self

    # ---- labels -----------------------------------------------------------------------------
    def _label_codes(self, sample_labels):
        """Sorted unique labels (``np.unique`` order) and an int32 code per cell."""
        labels = getattr(sample_labels, "values", sample_labels)
        labels = np.asarray(labels)
        if labels.ndim > 1:
            if labels.shape[1] == 1:
                labels = labels.reshape(-1)
            else:
                raise ValueError("sample_labels must be a single column. Got" "shape={}".format(labels.shape))
        codes, uniques = _factorize(labels)  # hash pass; np.unique would sort N strings
        if len(codes) and int(codes.min()) < 0:  # NaN / None: np.unique (the reference) cannot sort them with strings,
            raise ValueError("sample_labels contains missing values (NaN / None)")  # and never merges them silently
        uniques = np.asarray(uniques)
        order = np.argsort(uniques, kind="stable")
        rank = np.empty(len(order), dtype=np.int32)
        rank[order] = np.arange(len(order), dtype=np.int32)
        samples = uniques[order]
        if labels.dtype.kind in "US":
            samples = samples.astype(labels.dtype)
        return samples, rank[codes]

    def _create_sample_indicators(self, sample_labels):
        """One-hot indicator DataFrame, columns in ``np.unique`` order (``meld/meld.py:143-191``)."""
        self.sample_labels_ = sample_labels
        self.samples, codes = self._label_codes(sample_labels)
        self._codes = codes
        p = len(self.samples)
        ind = np.zeros((len(codes), p), dtype=int)
        ind[np.arange(len(codes)), codes] = 1
        index = getattr(self, "_labels_index", None) if p == 2 else None
        self._indicators = pd.DataFrame(ind, index=index, columns=self.samples, copy=False)
        return self._indicators

    @property
    def sample_indicators(self):
        """Indicator DataFrame (column-normalised when ``sample_normalize``), built on first use."""
        if getattr(self, "_indicators", None) is None and getattr(self, "_codes", None) is not None:
            if not isinstance(self._codes, np.ndarray):  # factorised on the device
                self._codes = self._codes.cpu().numpy()
            p = len(self.samples)
            ind = np.zeros((len(self._codes), p), dtype=int)
            ind[np.arange(len(self._codes)), self._codes] = 1
            df = pd.DataFrame(ind, index=self._labels_index if p == 2 else None, columns=self.samples, copy=False)
            if self.sample_normalize:
                df = df / df.sum(axis=0)
            self._indicators = df
        return getattr(self, "_indicators", None)

    @sample_indicators.setter
    def sample_indicators(self, value):
        self._indicators = value

    # ---- transform --------------------------------------------------------------------------
    def transform(self, sample_labels, _codes=None):
        """Filter the sample indicators of ``sample_labels`` over the graph.

        Returns a DataFrame (N, p) of sample densities, columns = sorted unique labels,
        index = ``sample_labels.index`` when it has one.
        """
        torch = nv.require_cuda()
        self.graph = utils._check_pygsp_graph(self.graph)
        self._sample_labels = sample_labels

        if sample_labels.shape[0] != self.graph.N:
            raise ValueError(
                "Input data ({}) and input graph ({}) "
                "are not of the same size".format(sample_labels.shape, self.graph.N)
            )
        flat_check = getattr(sample_labels, "values", sample_labels)
        if np.asarray(flat_check).ndim == 1 or np.asarray(flat_check).shape[1] == 1:
            # fit_transform factorises the labels on a host thread while the GPU builds the graph
            samples, codes = _codes if _codes is not None else self._label_codes(sample_labels)
            n_unique = len(samples)
        else:
            samples, codes = None, None
            n_unique = len(np.unique(flat_check))
        if n_unique == 1:
            raise ValueError("Found only one unqiue sample label. Cannot estimate density " "of a single sample.")

        if hasattr(sample_labels, "index"):
            self._labels_index = sample_labels.index
        else:
            self._labels_index = None

        if samples is None:
            self._create_sample_indicators(sample_labels)  # raises the single-column ValueError
        self.sample_labels_ = sample_labels
        self.samples, self._codes = samples, codes
        self._indicators = None  # rebuilt lazily by the sample_indicators property

        _filter.filter_kernel(self.filter, self.beta, self.offset, self.order)
        t0 = time.perf_counter()
        dev = self.graph.device
        d_codes = codes if isinstance(codes, torch.Tensor) else torch.from_numpy(codes).to(dev, non_blocking=True)
        densities = self.transform_device(d_codes, len(samples))
        # Read-back into pinned host memory (a pageable 16 MB read-back costs ~3x as long).  The DataFrame is built
        # straight on a block of a small pool of pinned buffers that is only reused once nothing refers to the array
        # any more (_pinned_result); when the caller keeps more results than the pool holds, through one staging
        # buffer and a host copy.  copy=False: pandas >= 3 otherwise copies (and transposes) the array, ~4 ms for 16 MB.
        host = _pinned_result(torch, densities)
        if host is None:
            staged = _pinned_staging(torch, densities)
            staged.copy_(densities, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            host = torch.empty(densities.shape, dtype=densities.dtype).copy_(staged).numpy()
        self.timings_["transform"] = time.perf_counter() - t0
        self.sample_densities = pd.DataFrame(host, index=self._labels_index, columns=self.samples, copy=False)
        self.timings_["transform_total"] = time.perf_counter() - t0
        return self.sample_densities

    def transform_device(self, codes, n_samples):
        """Engine-level entry (extension of the reference API): ``codes`` is an int32 CUDA tensor of
        label codes in ``[0, n_samples)``; returns the (N, n_samples) float64 densities as a CUDA
        tensor without touching the host.  ``transform`` is this plus label handling and the
        DataFrame wrap."""
        torch = nv.require_cuda()
        self.graph = utils._check_pygsp_graph(self.graph)
        if codes.shape[0] != self.graph.N:
            raise ValueError(
                "Input data ({}) and input graph ({}) "
                "are not of the same size".format(tuple(codes.shape), self.graph.N)
            )
        _filter.filter_kernel(self.filter, self.beta, self.offset, self.order)
        codes = codes.to(torch.int32).contiguous()
        S = torch.empty((codes.shape[0], int(n_samples)), dtype=torch.float64, device=codes.device)
        nv.check(
            nv.lib().meld_b200_indicator_matrix(nv.ptr(codes), codes.shape[0], int(n_samples),
                                                int(bool(self.sample_normalize)), nv.ptr(S), nv.current_stream_ptr()),
            "indicator_matrix",
        )
        if self._sharded is not None and self.solver == "chebyshev" and self.graph._lmax is None:
            # row-partitioned Lanczos on the ranks' slices (identical result on every rank)
            self.graph._lmax = self._sharded.estimate_lmax()
            self.graph.lmax_iters = getattr(self._sharded, "lmax_iters", None)
        events = self.profile_events
        if events is not None and self.solver == "chebyshev":
            self.graph.estimate_lmax()  # keep the Lanczos launches out of the filter bracket
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        densities = _filter.filter(
            signal=S, graph=self.graph, filter=self.filter, beta=self.beta, offset=self.offset, order=self.order,
            solver=self.solver, chebyshev_order=self.chebyshev_order,
            apply=self._sharded.apply if (self._sharded is not None and self.solver == "chebyshev") else None,
        )
        if events is not None:
            e1.record()
            events.append((e0, e1, int(self.chebyshev_order)))
        self.sample_densities_device = densities
        return densities

    def transform_sweep(self, sample_labels, betas=None, filter_params=None, as_tensor=False):
        """Densities of one or several label vectors under MANY filter settings on the fitted graph.

        The reference's dominant workload is a parameter search: one graph, thousands of
        ``MELD(beta=b).fit(graph).transform(labels)`` calls (``meld/benchmark.py:186-200``,
        ``notebooks/MELD_Quickstart.ipynb:729-747``: 25 label draws x 199 beta per graph).  Filters that differ
        only in ``beta / offset / order / filter`` share the Chebyshev basis ``T_k(L) S``, so the recurrence
        runs once per 8 signal columns and every setting only costs its coefficient vector
        (``meld_b200_cheby_sweep``).

        sample_labels : one label vector, or a list of label vectors (each as for ``transform``).
        betas : iterable of beta values (other filter parameters are the estimator's), or
        filter_params : list of dicts with any of ``beta, offset, order, filter``.
        as_tensor : return ``(R, columns)`` with ``R`` a CUDA tensor (n_settings, N, total columns) and
            ``columns`` a list of (label-vector index, sample name) instead of DataFrames on the host.

        Returns ``out[i][f]`` = DataFrame (N, p_i) of label vector ``i`` under setting ``f``; each equals
        what ``transform`` returns for that setting.
        """
        torch = nv.require_cuda()
        self.graph = utils._check_pygsp_graph(self.graph)
        if self.solver != "chebyshev":
            raise NotImplementedError(
                "solver='{}' is not available in the B200 engine; use solver='chebyshev'".format(self.solver)
            )
        if (betas is None) == (filter_params is None):
            raise ValueError("pass exactly one of betas / filter_params")
        settings = [dict(beta=b) for b in betas] if betas is not None else [dict(fp) for fp in filter_params]
        if not settings:
            raise ValueError("empty parameter sweep")
        single = not isinstance(sample_labels, (list, tuple))
        label_sets = [sample_labels] if single else list(sample_labels)
        dev = self.graph.device
        cols, blocks, meta = [], [], []
        for i, labels in enumerate(label_sets):
            if labels.shape[0] != self.graph.N:
                raise ValueError(
                    "Input data ({}) and input graph ({}) "
                    "are not of the same size".format(labels.shape, self.graph.N)
                )
            samples, codes = self._label_codes(labels)
            if len(samples) == 1:
                raise ValueError("Found only one unqiue sample label. Cannot estimate density " "of a single sample.")
            d_codes = torch.from_numpy(np.ascontiguousarray(codes, dtype=np.int32)).to(dev)
            S = torch.empty((len(codes), len(samples)), dtype=torch.float64, device=dev)
            nv.check(
                nv.lib().meld_b200_indicator_matrix(nv.ptr(d_codes), len(codes), len(samples),
                                                    int(bool(self.sample_normalize)), nv.ptr(S),
                                                    nv.current_stream_ptr()),
                "indicator_matrix",
            )
            blocks.append(S)
            meta.append((samples, getattr(labels, "index", None)))
            cols.extend((i, s) for s in samples)
        S_all = torch.cat(blocks, dim=1) if len(blocks) > 1 else blocks[0]
        lmax = self.graph.estimate_lmax()
        m = int(self.chebyshev_order)
        coeff = np.empty((len(settings), m + 1), dtype=np.float64)
        for f, st in enumerate(settings):
            unknown = set(st) - {"beta", "offset", "order", "filter"}
            if unknown:
                raise ValueError("Invalid sweep parameter(s): {}".format(sorted(unknown)))
            h = _filter.filter_kernel(st.get("filter", self.filter), st.get("beta", self.beta),
                                      st.get("offset", self.offset), st.get("order", self.order))
            coeff[f] = _filter.cheby_coefficients(h, lmax, m)
        R = _filter.cheby_sweep(self.graph, lmax, coeff, S_all)  # (F, N, sum p_i)
        if as_tensor:
            return R, cols
        host = R.cpu().numpy()
        out, c0 = [], 0
        for samples, index in meta:
            pi = len(samples)
            out.append([pd.DataFrame(host[f, :, c0:c0 + pi], index=index, columns=samples, copy=False)  # views of `host`
                        for f in range(len(settings))])
            c0 += pi
        return out[0] if single else out

    def fit_transform(self, X, sample_labels, **kwargs):
        """Build the graph on ``X`` and estimate the density of each sample in ``sample_labels``."""
        # The label codes do not depend on the graph: factorise them on a host thread while the build runs
        # (the C call releases the GIL).  Any error is left for transform to raise in the reference's order.
        pre = {}

        def _prefetch():
            # let the caller's thread get into the (GIL-free) native build first: factorising right away competes for
            # the GIL with the few Python steps in front of it and delays the host-to-device copy by milliseconds
            time.sleep(0.002)
            t_thr = time.perf_counter()
            try:
                flat = np.asarray(getattr(sample_labels, "values", sample_labels))
                if flat.ndim == 1 or (flat.ndim == 2 and flat.shape[1] == 1):
                    pre["codes"] = self._label_codes(sample_labels)
            except Exception:  # noqa: BLE001 - recomputed (and raised) by transform
                pre.clear()
            self.timings_["labels_thread"] = time.perf_counter() - t_thr

        import torch as _torch

        # None without CUDA: fit raises the engine's "no CPU fallback" error; the labels are still factorised on the host
        dev_index = _torch.cuda.current_device() if _torch.cuda.is_available() else None
        worker = threading.Thread(target=_prefetch, name="meld_b200-labels", daemon=True)
        # Single GPU: the host factorisation (12-18 ms for 500k string labels) hides behind the ~30 ms build on a thread.
        # Distributed: the build is only ~10 ms per rank and every rank waits for the slowest host, so the labels are
        # factorised on the GPU first (~3 ms, _factorize_device).  (Running that on a side stream beside the build was
        # measured too: the persistent search kernels hold every SM for milliseconds, the small kernels queue behind
        # them -- 30 ms -- and the end-to-end step got slower and noisier.)
        device_labels = (dev_index is not None and not os.environ.get("MELD_B200_HOST_LABELS")
                         and (self.distributed or os.environ.get("MELD_B200_DEVICE_LABELS")))
        if device_labels:
            t_lab = time.perf_counter()
            try:
                flat = np.asarray(getattr(sample_labels, "values", sample_labels))
                if flat.ndim == 1 or (flat.ndim == 2 and flat.shape[1] == 1):
                    got = _factorize_device(_torch, flat.reshape(-1), _torch.device("cuda", dev_index))
                    if got is not None:
                        pre["codes"] = got
            except Exception:  # noqa: BLE001 - the host path below (and transform) deal with it
                pre.clear()
            self.timings_["labels_device"] = time.perf_counter() - t_lab
        if "codes" in pre:
            pass
        elif os.environ.get("MELD_B200_NO_LABEL_THREAD"):
            _prefetch()
        else:
            worker.start()
        t_fit = time.perf_counter()
        try:
            self.fit(X, **kwargs)
        finally:
            self.timings_["fit_call"] = time.perf_counter() - t_fit
            if worker.ident is not None:
                worker.join()
        self.timings_["fit_and_join"] = time.perf_counter() - t_fit
        return self.transform(sample_labels, _codes=pre.get("codes"))


_STAGING = {}


def _pinned_staging(torch, like):
    key = (tuple(like.shape), like.dtype)
    buf = _STAGING.get(key)
    if buf is None:
        if len(_STAGING) >= 4:
            _STAGING.clear()
        buf = _STAGING[key] = torch.empty(like.shape, dtype=like.dtype, device="cpu", pin_memory=True)
    return buf


def _alloc_pinned(torch, shape, dtype):
    return torch.empty(shape, dtype=dtype, device="cpu", pin_memory=True)


def _sync_stream(torch, device):
    torch.cuda.current_stream(device).synchronize()


_RESULT_POOL = []  # [pinned tensor, weakref to the ndarray handed out last (None: never used)]
_RESULT_POOL_MAX = 4
_RESULT_POOL_BYTES = 1 << 28
_RESULT_POOL_LOCK = threading.Lock()


def _pinned_result(torch, dev_tensor):
    """Copy a CUDA tensor into a pinned host block and return it as an ndarray WITHOUT a second host copy.

    A block is handed out again only when the array made from it last time is gone (every view of that array -- the
    DataFrame's blocks, ``df.values`` -- keeps it alive, so a live result is never overwritten).  At most
    ``_RESULT_POOL_MAX`` blocks / ``_RESULT_POOL_BYTES`` stay pinned; a caller that holds more results than that, or
    asks for a bigger one, gets None and the staged copy.  Allocating a pinned block costs milliseconds
    (cudaHostAlloc), which is why they are kept."""
    nbytes = dev_tensor.numel() * dev_tensor.element_size()
    if nbytes == 0 or nbytes > _RESULT_POOL_BYTES:
        return None
    shape, dtype = tuple(dev_tensor.shape), dev_tensor.dtype

    def is_free(e):
        return e[1] is None or e[1]() is None

    with _RESULT_POOL_LOCK:  # two threads must not pick the same free block
        entry = next((e for e in _RESULT_POOL if is_free(e) and tuple(e[0].shape) == shape and e[0].dtype == dtype), None)
        if entry is None:
            busy = [e for e in _RESULT_POOL if not is_free(e)]
            held = sum(e[0].numel() * e[0].element_size() for e in busy)
            if len(busy) >= _RESULT_POOL_MAX or held + nbytes > _RESULT_POOL_BYTES:
                return None  # the caller holds every block: do not pin more
            _RESULT_POOL[:] = busy  # blocks of other shapes nobody refers to go
            entry = [_alloc_pinned(torch, shape, dtype), None]
            _RESULT_POOL.append(entry)
        placeholder = np.empty(0)
        entry[1] = weakref.ref(placeholder)  # taken: busy until the result array replaces this below
    entry[0].copy_(dev_tensor, non_blocking=True)
    _sync_stream(torch, dev_tensor.device)
    host = entry[0].numpy()
    entry[1] = weakref.ref(host)
    del placeholder
    return host


_HASH_MULT = np.random.default_rng(0x5EED).integers(1, 2**63 - 1, size=64, dtype=np.int64).astype(np.uint64) | np.uint64(1)


def _factorize_device(torch, labels, dev):
    """Label codes on the GPU for fixed-width numpy labels (strings, integers, booleans): the raw bytes go up once
    (16 MB for 500k '<U8' labels), 64-bit words are hashed, ``torch.unique`` (plumbing) gives the codes, and an exact
    comparison of every label with its representative rules out hash collisions.  Returns (sorted unique labels --
    ``np.unique`` order --, int32 CUDA tensor of codes) or None when the labels do not qualify.  Runs on the label
    thread's own stream beside the graph build: with several ranks on one host the Python-side factorisation (12-80 ms,
    and every rank waits for the slowest) is what an end-to-end step spends its time on."""
    if not (isinstance(labels, np.ndarray) and labels.ndim == 1 and labels.dtype.kind in "USiub"
            and 0 < labels.dtype.itemsize <= 512 and len(labels) > 4096):
        return None
    n, w = len(labels), labels.dtype.itemsize
    raw = np.ascontiguousarray(labels).view(np.uint8).reshape(n, w)
    w8 = (w + 7) // 8
    if w8 * 8 != w:
        pad = np.zeros((n, w8 * 8), dtype=np.uint8)
        pad[:, :w] = raw
        raw = pad
    words = torch.from_numpy(raw.view(np.int64).reshape(n, w8)).to(dev)
    mult = torch.from_numpy(_HASH_MULT[:w8].view(np.int64).copy()).to(dev)
    h = (words * mult).sum(dim=1)  # wraps modulo 2^64
    uniq, inv = torch.unique(h, return_inverse=True)
    k = int(uniq.numel())
    if k > 4096:
        return None
    first = torch.full((k,), n, dtype=torch.int64, device=dev)
    first.scatter_reduce_(0, inv, torch.arange(n, dtype=torch.int64, device=dev), reduce="amin")
    if not bool(torch.equal(words, words[first][inv])):  # two different labels shared a hash: host path
        return None
    reps = labels[first.cpu().numpy()]
    order = np.argsort(reps, kind="stable")
    rank = np.empty(k, dtype=np.int32)
    rank[order] = np.arange(k, dtype=np.int32)
    codes = torch.from_numpy(rank).to(dev)[inv]
    return reps[order], codes


def _factorize(labels):
    """``pd.factorize`` (codes in order of first appearance, uniques), with a fast exact path for fixed-width
    numpy string labels: hashing 64-bit words of the raw bytes avoids creating one Python string per cell."""
    if labels.dtype.kind in "US" and labels.ndim == 1 and 0 < labels.dtype.itemsize <= 512 and len(labels) > 4096:
        n, w = len(labels), labels.dtype.itemsize
        raw = np.ascontiguousarray(labels).view(np.uint8).reshape(n, w)
        w8 = (w + 7) // 8
        if w8 * 8 != w:
            pad = np.zeros((n, w8 * 8), dtype=np.uint8)
            pad[:, :w] = raw
            raw = pad
        words = raw.view(np.uint64).reshape(n, w8)
        h = words @ _HASH_MULT[:w8]  # wraps modulo 2^64
        codes, _ = pd.factorize(h)
        k = int(codes.max()) + 1
        first = np.empty(k, dtype=np.int64)
        first[codes[::-1]] = np.arange(n - 1, -1, -1, dtype=np.int64)  # last write wins = first occurrence
        rep = words[first]
        # exact check that no two different labels shared a hash (else fall through to the generic path)
        if k <= 4096 and all(np.array_equal(words[:, j], rep[:, j][codes]) for j in range(w8)):
            return codes, labels[first]
    return pd.factorize(labels)


def _same_data(torch, a, b):
    if a is b:
        return True
    if isinstance(a, torch.Tensor) or isinstance(b, torch.Tensor):
        return (
            isinstance(a, torch.Tensor)
            and isinstance(b, torch.Tensor)
            and a.shape == b.shape
            and a.device == b.device
            and bool(torch.equal(a, b))
        )
    try:
        a_, b_ = np.asarray(getattr(a, "values", a)), np.asarray(getattr(b, "values", b))
        return a_.shape == b_.shape and bool(np.array_equal(a_, b_))
    except Exception:
        return False
