"""Oracle: kNN alpha-decay graph + combinatorial Laplacian (CPU, float64).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Restates what
``graphtools.Graph(X, knn, decay, thresh, anisotropy=1, use_pygsp=True)`` builds
when reached from ``MELD.fit`` (reference ``meld/meld.py:117-118,273``; the class
hierarchy is upstream graphtools 1.5.x: ``graphs.kNNGraph`` ->
``base.BaseGraph`` / ``base.PyGSPGraph``) and what ``pygsp.graphs.Graph`` derives
from the weight matrix (``compute_laplacian``, ``estimate_lmax``; reference call
site ``meld/filter.py:39``).  It issues the same scikit-learn / scipy calls as
the upstream packages: ball-tree ``kneighbors`` / ``radius_neighbors``,
``scipy.sparse`` arithmetic and ARPACK ``eigsh``.
"""

from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.sparse.linalg import eigsh, ArpackNoConvergence
from scipy.spatial.distance import pdist, squareform
from sklearn.neighbors import NearestNeighbors

EPS = np.finfo(float).eps
SEARCH_MULTIPLIER = 6  # graphtools kNNGraph.__init__ default search_multiplier


def reduce_data(X, n_pca, random_state=None):
    """graphtools ``Data._reduce_data`` for dense input (SURVEY 8a row B').

    ``n_pca`` is disabled by ``GraphEstimator._parse_input`` when
    ``min(X.shape) <= n_pca``; otherwise randomized PCA, centred.
    """
    X = np.asarray(X, dtype=np.float64)
    if n_pca is None or n_pca >= min(X.shape):
        return X
    from sklearn.decomposition import PCA

    pca = PCA(n_pca, svd_solver="randomized", random_state=random_state)
    pca.fit(X)  # graphtools fits, then transforms: for the randomized solver transform(X) != fit_transform(X)
    return pca.transform(X)


def knn_kernel(
    data,
    knn=5,
    decay=40.0,
    thresh=1e-4,
    bandwidth_scale=1.0,
    n_jobs=1,
    algorithm="ball_tree",
    return_stats=False,
    timings=None,
):
    """``kNNGraph.build_kernel`` -> ``build_kernel_to_data(data, knn=knn + 1)``.

    Sparse fast alpha-decay branch (``decay`` not None, ``0 < thresh < 1``):
    search ``6*(knn+1)`` neighbours, take the bandwidth from the k-th non-self
    neighbour, widen the search for rows whose window does not reach the
    radius ``eps_i * (-ln thresh)^(1/decay)``, evaluate
    ``exp(-(d/eps_i)^decay)`` and drop entries below ``thresh``.
    Returns the un-symmetrised CSR kernel (self loops = 1).
    """
    data = np.ascontiguousarray(data, dtype=np.float64)
    N = data.shape[0]
    knn_arg = knn + 1  # build_kernel passes knn + 1 (self is its own neighbour)
    if knn_arg > N:
        raise ValueError("knn + 1 = {} exceeds the number of cells {}".format(knn_arg, N))
    if decay is None or thresh == 1:
        # kNNGraph.build_kernel_to_data, binary branch: unweighted connectivity of the knn + 1 nearest (self included)
        tree = NearestNeighbors(n_neighbors=knn_arg, algorithm=algorithm, n_jobs=n_jobs).fit(data)
        K = tree.kneighbors_graph(data, n_neighbors=knn_arg, mode="connectivity").tocsr()
        K.sort_indices()
        return (K, {}) if return_stats else K
    thresh = max(thresh, EPS)  # kNNGraph.__init__: thresh == 0 with decay -> eps
    knn_max = N  # knn_max=None upstream means "no cap"

    import time as _time

    t_search = _time.perf_counter()
    tree = NearestNeighbors(n_neighbors=knn_arg, algorithm=algorithm, n_jobs=n_jobs).fit(data)
    search_knn = min(knn_arg * SEARCH_MULTIPLIER, knn_max)
    distances, indices = tree.kneighbors(data, n_neighbors=search_knn)
    if timings is not None:  # "KNN search" vs "affinities" of the reference's own log (bench.py stage split)
        timings["knn_first_search_s"] = _time.perf_counter() - t_search

    bandwidth = distances[:, knn_arg - 1] * bandwidth_scale
    bandwidth = np.maximum(bandwidth, EPS)
    radius = bandwidth * np.power(-1 * np.log(thresh), 1 / decay)
    update_idx = np.argwhere(np.max(distances, axis=1) < radius).reshape(-1)
    n_overflow_first = len(update_idx)

    if len(update_idx) > 0:
        distances = [d for d in distances]
        indices = [i for i in indices]

    search_knn = min(search_knn * SEARCH_MULTIPLIER, knn_max)
    while len(update_idx) > N // 10 and search_knn < N / 2 and search_knn < knn_max:
        dist_new, ind_new = tree.kneighbors(data[update_idx], n_neighbors=search_knn)
        for i, idx in enumerate(update_idx):
            distances[idx] = dist_new[i]
            indices[idx] = ind_new[i]
        keep = np.max(dist_new, axis=1) < radius[update_idx]
        update_idx = update_idx[keep]
        search_knn = min(search_knn * SEARCH_MULTIPLIER, knn_max)

    if len(update_idx) > 0:
        # remaining rows: everything inside the row's own radius
        dist_new, ind_new = tree.radius_neighbors(data[update_idx], radius=radius[update_idx].max())
        for i, idx in enumerate(update_idx):
            distances[idx] = dist_new[i]
            indices[idx] = ind_new[i]

    if isinstance(distances, list):
        lens = np.array([len(d) for d in distances])
        vals = np.concatenate(distances) / np.repeat(bandwidth, lens)
        cols = np.concatenate(indices)
        indptr = np.concatenate([[0], np.cumsum(lens)])
    else:
        vals = (distances / bandwidth[:, None]).reshape(-1)
        cols = indices.reshape(-1)
        indptr = np.arange(0, N * distances.shape[1] + 1, distances.shape[1])

    vals = np.exp(-1 * np.power(vals, decay))
    vals = np.where(np.isnan(vals), 1, vals)
    vals[vals < thresh] = 0
    K = sparse.csr_matrix((vals, cols, indptr), shape=(N, N))
    K.eliminate_zeros()
    K.sort_indices()
    if return_stats:
        return K, {"bandwidth": bandwidth, "radius": radius, "n_overflow_first": n_overflow_first}
    return K


def knn_kernel_bruteforce(data, knn=5, decay=40.0, thresh=1e-4, bandwidth_scale=1.0):
    """Definition the search/expansion loop must be equivalent to (SURVEY 8c vii).

    ``K_ij = exp(-(d_ij/eps_i)^decay)`` for every j with ``K_ij >= thresh``,
    ``eps_i`` = distance to the k-th non-self neighbour.  O(N^2) dense.
    """
    data = np.asarray(data, dtype=np.float64)
    D = squareform(pdist(data))
    Ds = np.sort(D, axis=1)
    bandwidth = np.maximum(Ds[:, knn] * bandwidth_scale, EPS)
    thresh = max(thresh, EPS)
    K = np.exp(-1 * np.power(D / bandwidth[:, None], decay))
    K[K < thresh] = 0
    return sparse.csr_matrix(K)


def knn_kernel_rows_bruteforce(data, rows, knn=5, decay=40.0, thresh=1e-4, bandwidth_scale=1.0, chunk=256):
    """Rows ``rows`` of the un-symmetrised kernel by its DEFINITION (``knn_kernel_bruteforce`` restricted to a
    sample of rows, usable at 500k - 2M cells where neither the ball tree nor a dense N x N matrix is an
    option): all j with ``exp(-(d_ij / eps_i)^decay) >= thresh``, ``eps_i`` = distance to the knn-th non-self
    neighbour.  A float64 GEMM expansion only pre-selects candidates (with a relative safety margin); every
    kept distance is then recomputed as ``sqrt(sum_k (x_ik - x_jk)^2)`` in feature order like the ball tree.
    Returns a list of (cols ascending, values) per requested row."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    rows = np.asarray(rows, dtype=np.int64)
    N = data.shape[0]
    thresh = max(thresh, EPS)
    rho = np.power(-np.log(thresh), 1.0 / decay)
    sq = np.einsum("ij,ij->i", data, data)
    out = []
    for s in range(0, len(rows), chunk):
        r = rows[s:s + chunk]
        d2 = sq[r][:, None] + sq[None, :] - 2.0 * (data[r] @ data.T)  # approximate, selection only
        np.maximum(d2, 0.0, out=d2)
        kth = np.partition(d2, knn, axis=1)[:, knn]  # ~ squared distance of the knn-th non-self neighbour
        slack = 1e-6 * (sq[r] + sq.max()) + 1e-300
        for t, i in enumerate(r):
            lim = (kth[t] + slack[t]) * (rho * bandwidth_scale) ** 2 * (1.0 + 1e-6) + slack[t]
            cand = np.flatnonzero(d2[t] <= max(lim, kth[t] + slack[t]))
            diff = data[cand] - data[i]
            acc = np.zeros(len(cand))
            for k in range(data.shape[1]):  # sequential, unfused: sklearn's rdist order
                acc = acc + diff[:, k] * diff[:, k]
            dist = np.sqrt(acc)
            eps = max(np.sort(dist)[knn] * bandwidth_scale, EPS)
            val = np.exp(-1 * np.power(dist / eps, decay))
            val = np.where(np.isnan(val), 1, val)
            keep = val >= thresh
            order = np.argsort(cand[keep], kind="stable")
            out.append((cand[keep][order], val[keep][order]))
    return out


def traditional_kernel(data, knn=5, decay=40.0, thresh=0.0, bandwidth_scale=1.0):
    """graphtools ``TraditionalGraph.build_kernel`` (dense, exact) -- chosen by
    ``graphtools.api.Graph`` when ``thresh == 0``; only the reference's KAT
    (``test/test_meld.py:58-67``) reaches it.  Out of product scope."""
    data = np.asarray(data, dtype=np.float64)
    D = squareform(pdist(data))
    knn_dist = np.partition(D, knn + 1, axis=1)[:, : knn + 1]
    bandwidth = np.max(knn_dist, axis=1) * bandwidth_scale
    D = (D.T / bandwidth).T
    K = np.exp(-1 * np.power(D, decay))
    K = np.where(np.isnan(K), 1, K)
    K[K < thresh] = 0
    return sparse.csr_matrix(K)


def symmetrize(K):
    """``BaseGraph.symmetrize_kernel`` with ``kernel_symm='+'``: (K + K^T)/2."""
    return ((K + K.T) / 2).tocsr()


def apply_anisotropy(K, anisotropy=1.0):
    """``BaseGraph.apply_anisotropy``: K_ij / (q_i q_j)^a, q = row sums incl. diagonal."""
    if anisotropy == 0:
        return K
    d = np.array(K.sum(1)).flatten()
    K = K.tocoo()
    K.data = ((d[K.row] * d[K.col]) ** -anisotropy) * K.data
    return K.tocsr()


def weights_from_kernel(K):
    """``PyGSPGraph._build_weight_from_kernel``: W = K with the diagonal removed."""
    W = K.tolil(copy=True)
    W.setdiag(0)
    W = W.tocsr()
    W.eliminate_zeros()
    return W


def laplacian(W):
    """pygsp ``Graph.compute_laplacian('combinatorial')``: L = diag(W 1) - W."""
    dw = np.ravel(W.sum(1))
    L = (sparse.diags(dw, 0) - W).tocsr()
    L.sort_indices()
    return L


def estimate_lmax(L, dw=None):
    """pygsp 0.5.1 ``Graph.estimate_lmax``: 1.01 * ARPACK(k=1, tol=5e-3, ncv=min(N,10))."""
    N = L.shape[0]
    try:
        lmax = eigsh(L.tocsc(), k=1, tol=5e-3, ncv=min(N, 10), return_eigenvectors=False)[0]
        return float(lmax) * 1.01
    except ArpackNoConvergence:
        if dw is None:
            dw = L.diagonal()
        return float(2 * np.max(dw))


def build_graph(
    X,
    knn=5,
    decay=40.0,
    thresh=1e-4,
    anisotropy=1.0,
    n_pca=100,
    random_state=None,
    bandwidth_scale=1.0,
    n_jobs=1,
    data_nu=None,
    timings=None,
):
    """Stages B..F of SURVEY 8a.  Returns a dict with data_nu, K, W, L, dw."""
    if data_nu is None:
        data_nu = reduce_data(X, n_pca, random_state)
    K0 = knn_kernel(data_nu, knn=knn, decay=decay, thresh=thresh, bandwidth_scale=bandwidth_scale, n_jobs=n_jobs,
                    timings=timings)
    K = apply_anisotropy(symmetrize(K0), anisotropy)
    K.sort_indices()
    W = weights_from_kernel(K)
    L = laplacian(W)
    return {"data_nu": data_nu, "K_knn": K0, "K": K, "W": W, "L": L, "dw": np.ravel(W.sum(1)), "N": L.shape[0]}
