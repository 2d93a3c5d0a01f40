"""GPU parity of the graph build (kNN alpha-decay kernel -> symmetrise -> anisotropy -> Laplacian)
against the golden vectors written by the CPU oracle, and end-to-end fit_transform parity.

Graph gate: identical sparsity pattern and |dL| <= 1e-10 max|L|.
Density gate: north_star's 1e-5 relative with the oracle's lmax injected."""

import numpy as np
import pytest
from scipy import sparse

from conftest import GOLDEN_CASES, density_parity, load_golden

pytestmark = pytest.mark.gpu

SEARCHES = [pytest.param(True, id="simt"), pytest.param(False, id="tcgen05")]


@pytest.fixture(scope="module")
def mb():
    import torch

    assert torch.cuda.is_available()
    import meld_b200

    return meld_b200


def _same_pattern(A, B):
    A = A.tocsr().copy()
    B = B.tocsr().copy()
    A.sort_indices()
    B.sort_indices()
    return A.shape == B.shape and np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)


def _graph_kwargs(g):
    kw = dict(knn=5, decay=40.0, thresh=1e-4, anisotropy=1.0)
    kw.update(g["graph_kwargs"])
    return kw


@pytest.mark.parametrize("simt", SEARCHES)
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_graph_matches_golden(mb, name, simt):
    g = load_golden(name)
    graph = mb.DeviceGraph.from_data(g["X"], keep_knn_kernel=True, simt_search=simt, **_graph_kwargs(g))
    K = graph.to_scipy_knn_kernel()
    assert _same_pattern(K, g["K"]), (name, K.nnz, g["K"].nnz)
    K.sort_indices()
    assert np.abs(K.data - g["K"].data).max() <= 1e-12
    L = graph.to_scipy_L()
    assert _same_pattern(L, g["L"]), (name, L.nnz, g["L"].nnz)
    assert L.has_sorted_indices or True
    scale = np.abs(g["L"].data).max()
    assert np.abs(L.data - g["L"].data).max() <= 1e-10 * scale
    # invariants of the reference construction
    assert abs(L - L.T).max() == 0.0
    assert np.abs(L @ np.ones(L.shape[0])).max() <= 1e-13 * scale * 50
    stats = graph.build_stats()
    assert stats["search_impl"] == (1 if simt else 0)


@pytest.mark.parametrize("simt", SEARCHES)
@pytest.mark.parametrize("name", ["readme_toy", "blobs2k_k15", "blobs1k5_aniso0"])
def test_fit_transform_matches_golden(mb, name, simt, monkeypatch):
    g = load_golden(name)
    if simt:
        orig = mb.DeviceGraph.from_data.__func__
        monkeypatch.setattr(mb.DeviceGraph, "from_data",
                            classmethod(lambda cls, *a, **k: orig(cls, *a, simt_search=True, **k)))
    gk = g["graph_kwargs"]
    op = mb.MELD(verbose=0, n_pca=None, **gk, **g["filter_kwargs"])
    op.fit(g["X"])
    own_lmax = op.graph.estimate_lmax()
    # the reference's ARPACK estimate (tol 5e-3) and the GPU Lanczos value differ by < 2e-4 relative
    assert abs(own_lmax - g["lmax"]) <= 3e-4 * g["lmax"]
    op.graph.lmax = g["lmax"]  # parity mode: share the oracle's lmax
    dens = op.transform(g["labels"])
    normwise, ok = density_parity(dens.values, g["densities"], 1e-5)
    assert ok and normwise < 1e-9, (name, normwise)
    # and with the engine's own lmax the change stays well inside what lmax noise explains
    op.graph.lmax = own_lmax
    op.set_params(beta=op.beta)  # no-op
    dens2 = op.transform(g["labels"])
    normwise2, _ = density_parity(dens2.values, g["densities"], 1e-5)
    assert normwise2 < 2e-4


def test_permutation_equivariance(mb):
    """Permuting cells permutes the densities (SURVEY 8c v) -- independent of the oracle."""
    g = load_golden("blobs1k5_aniso0")
    rng = np.random.default_rng(0)
    perm = rng.permutation(g["X"].shape[0])
    kw = dict(verbose=0, n_pca=None, **g["graph_kwargs"], **g["filter_kwargs"])
    a = mb.MELD(**kw)
    a.fit(g["X"])
    a.graph.lmax = g["lmax"]
    da = a.transform(g["labels"])
    b = mb.MELD(**kw)
    b.fit(g["X"][perm])
    b.graph.lmax = g["lmax"]
    db = b.transform(g["labels"][perm])
    assert np.abs(db.values - da.values[perm]).max() <= 1e-11 * np.abs(da.values).max()


def test_uncentred_far_from_origin_data(mb):
    """Large common offset: the low-precision search must still nominate the exact neighbourhood."""
    from oracle import graph as og

    rng = np.random.default_rng(21)
    X = rng.normal(size=(1200, 24)) * np.exp(-np.arange(24) / 6.0) + 1000.0
    ref = og.build_graph(X, knn=9, n_pca=None)
    for simt in (True, False):
        graph = mb.DeviceGraph.from_data(X, knn=9, simt_search=simt)
        L = graph.to_scipy_L()
        assert _same_pattern(L, ref["L"])
        assert np.abs(L.data - ref["L"].data).max() <= 1e-10 * np.abs(ref["L"].data).max()


def test_duplicate_points_and_small_n(mb):
    from oracle import graph as og

    rng = np.random.default_rng(22)
    X = rng.normal(size=(300, 5))
    X[10:20] = X[0]  # exact duplicates: zero distances, eps floor
    ref = og.build_graph(X, knn=3, n_pca=None)
    graph = mb.DeviceGraph.from_data(X, knn=3, simt_search=True)
    L = graph.to_scipy_L()
    assert _same_pattern(L, ref["L"])
    assert np.abs(L.data - ref["L"].data).max() <= 1e-10 * np.abs(ref["L"].data).max()


def test_two_stage_build_over_row_ranges_matches_single_call(mb):
    """The sharded build's C-ABI stages, driven on one GPU: stage 1 over two row ranges, concatenated,
    then stage 2, reproduces meld_b200_knn_graph_build exactly (same graph, bit for bit)."""
    import torch
    from meld_b200.graph import DeviceGraph, _as_device_f64

    X, _ = mb.synthetic.make_blobs(6000, 30, 6, 3, 8.0, seed=5)  # >= 4096 cells: Morton order is active
    kw = dict(knn=9, decay=40.0, thresh=1e-4)
    ref = DeviceGraph.from_data(X, anisotropy=1.0, **kw).to_scipy_L()
    Xd = _as_device_f64(torch, X)
    bounds = DeviceGraph.shard_bounds(X.shape[0], 3)
    assert bounds == [0, 2048, 4096, 6000]
    parts = [DeviceGraph.candidates(Xd, a, b, **kw) for a, b in zip(bounds[:-1], bounds[1:])]
    counts = torch.cat([p[0] for p in parts])
    cand = torch.cat([p[1] for p in parts])
    d2 = torch.cat([p[2] for p in parts])
    eps = torch.cat([p[3] for p in parts])
    perm = parts[0][4]
    assert perm is not None and all(torch.equal(p[4], perm) for p in parts)
    g = DeviceGraph.from_candidates(X.shape[0], counts, cand, d2, eps, perm, anisotropy=1.0, **kw)
    L = g.to_scipy_L()
    assert _same_pattern(L, ref)
    assert np.array_equal(L.data, ref.data)


@pytest.mark.parametrize("tuning", [dict(tc_multicast=1), dict(tc_multicast=4)])
def test_search_kernel_variants_build_the_same_graph(mb, tuning):
    """No multicast and 4-CTA multicast clusters give the same golden graph as the default CTA pairs."""
    from meld_b200 import _native as nv

    g = load_golden("blobs2k_k15")
    try:
        nv.set_tuning(**tuning)
        graph = mb.DeviceGraph.from_data(g["X"], **_graph_kwargs(g))
        L = graph.to_scipy_L()
        assert _same_pattern(L, g["L"])
        assert np.abs(L.data - g["L"].data).max() <= 1e-10 * np.abs(g["L"].data).max()
    finally:
        nv.set_tuning(tc_multicast=2)


def test_pruned_search_matches_oracle_and_unpruned(mb):
    """k-means cell order + bounding-ball tile pruning (forced on at a test-sized n): the graph equals the
    oracle's and the unpruned search's bit for bit, and tile pairs really were skipped."""
    from meld_b200 import _native as nv
    from oracle import graph as og

    X, _ = mb.synthetic.make_blobs(20000, 30, 6, 3, 8.0, seed=9)
    kw = dict(knn=9, decay=40.0, thresh=1e-4)
    ref = og.build_graph(X, n_pca=None, **kw)["L"]
    try:
        nv.set_tuning(cluster_cells=512, clusters=32)  # 20000 cells -> 32 clusters of ~2.4 tiles
        g1 = mb.DeviceGraph.from_data(X, anisotropy=1.0, **kw)
        bt = g1.build_times()
        L1 = g1.to_scipy_L()
        nv.set_tuning(prune=0)
        L0 = mb.DeviceGraph.from_data(X, anisotropy=1.0, **kw).to_scipy_L()
        nv.set_tuning(prune=1, tc_multicast=1)
        L2 = mb.DeviceGraph.from_data(X, anisotropy=1.0, **kw).to_scipy_L()
    finally:
        nv.set_tuning(cluster_cells=1024, clusters=64, prune=1, tc_multicast=2)
    assert bt["flops_per_pass"] < 0.8 * bt["flops_unpruned_pass"], bt
    for L in (L1, L2):
        assert _same_pattern(L, ref)
        assert np.abs(L.data - ref.data).max() <= 1e-10 * np.abs(ref.data).max()
        assert _same_pattern(L, L0) and np.array_equal(L.data, L0.data)


def test_device_pca_matches_sklearn(mb):
    """Row B' on the device: same algorithm and RandomState draw as sklearn's randomized PCA -> same data_nu
    (1e-9 relative), for N > D and N < D (transposed branch), and fit_transform through it matches the oracle."""
    import torch
    from sklearn.decomposition import PCA

    from meld_b200 import pca
    from oracle import meld as omeld

    for n, D, k, seed in ((3000, 300, 50, 0), (400, 1000, 60, 3)):
        X, _ = mb.synthetic.make_blobs(n, D, 6, 3, tau=D / 10.0, seed=5)
        ref = PCA(k, svd_solver="randomized", random_state=seed).fit(X).transform(X)  # graphtools: fit, then transform
        out, obj = pca.randomized_pca(torch.from_numpy(X).cuda(), k, random_state=seed)
        assert np.abs(out.cpu().numpy() - ref).max() <= 1e-9 * np.abs(ref).max()
        assert obj.components_.shape == (k, D)
    X, labels = mb.synthetic.make_blobs(2500, 400, 5, 3, tau=40.0, seed=11)
    ref, g, lmax = omeld.fit_transform(X, labels, n_pca=50, random_state=0, knn=7)
    op = mb.MELD(verbose=0, n_pca=50, random_state=0, knn=7)
    op.fit(X)
    L = op.graph.to_scipy_L()
    assert _same_pattern(L, g["L"])
    op.graph.lmax = lmax
    dens = op.transform(labels)
    normwise, ok = density_parity(dens.values, ref.values, 1e-5)
    assert ok and normwise < 1e-7, normwise


@pytest.mark.parametrize("world", [2, 3])
def test_row_partitioned_stage2_matches_single_call(mb, world):
    """The fully sharded build (meld_b200_stage2_*: every rank assembles only its rows of L), all ranks driven on one
    GPU with the collectives done by hand: eps all-gather = concatenation, all-to-all-v of the mirror records =
    re-grouping the send buffers, row-sum all-gather = concatenation.  The stacked slices equal the single-call
    Laplacian bit for bit."""
    import ctypes as C
    import torch
    from scipy import sparse
    from meld_b200 import _native as nv
    from meld_b200.graph import DeviceGraph, _as_device_f64

    lib = nv.lib()
    X, _ = mb.synthetic.make_blobs(6000, 30, 6, 3, 8.0, seed=5)  # >= 4096 cells: the internal cell order is active
    kw = dict(knn=9, decay=40.0, thresh=1e-4)
    full = DeviceGraph.from_data(X, anisotropy=1.0, **kw)
    ref = full.to_scipy_L()
    Xd = _as_device_f64(torch, X)
    N = X.shape[0]
    bounds = DeviceGraph.shard_bounds(N, world)
    sp = nv.current_stream_ptr
    hs, eps = [], []
    for r in range(world):
        h = C.c_void_p()
        nv.check(lib.meld_b200_knn_candidates(nv.ptr(Xd), N, X.shape[1], 9, 40.0, 1e-4, 1.0, bounds[r], bounds[r + 1], 0,
                                              sp(), C.byref(h)), "knn_candidates")
        e = torch.zeros(bounds[r + 1] - bounds[r], dtype=torch.float64, device="cuda")
        nv.check(lib.meld_b200_cands_export(h, None, None, None, nv.ptr(e), None, sp()), "cands_export")
        hs.append(h)
        eps.append(e)
    eps_full = torch.cat(eps)
    hb = (C.c_int64 * (world + 1))(*bounds)
    sts, counts, sends = [], [], []
    for r in range(world):
        st, hc = C.c_void_p(), (C.c_int64 * world)()
        nv.check(lib.meld_b200_stage2_begin(hs[r], nv.ptr(eps_full), hb, world, 9, 40.0, 1e-4, 1.0, 1.0, sp(), C.byref(st),
                                            hc), "stage2_begin")
        cnt = [int(v) for v in hc]
        assert cnt[r] == 0  # a rank never sends records to itself
        buf = torch.empty(max(sum(cnt), 1) * 16, dtype=torch.uint8, device="cuda")
        nv.check(lib.meld_b200_stage2_records(st, nv.ptr(buf), sp()), "stage2_records")
        sts.append(st)
        counts.append(cnt)
        sends.append(buf)
    assert sum(sum(c) for c in counts) > 0  # rows of different ranks do mirror into each other
    qs = []
    for r in range(world):  # the all-to-all-v: what every source grouped for destination r
        parts = []
        for s_ in range(world):
            off = 16 * sum(counts[s_][:r])
            parts.append(sends[s_][off: off + 16 * counts[s_][r]])
        recv = torch.cat(parts) if parts else torch.empty(0, dtype=torch.uint8, device="cuda")
        n_recv = recv.numel() // 16
        if n_recv == 0:
            recv = torch.empty(16, dtype=torch.uint8, device="cuda")
        q = torch.zeros(bounds[r + 1] - bounds[r], dtype=torch.float64, device="cuda")
        nv.check(lib.meld_b200_stage2_assemble(sts[r], nv.ptr(recv), n_recv, sp(), nv.ptr(q)), "stage2_assemble")
        torch.cuda.synchronize()
        qs.append(q)
    q_full = torch.cat(qs)
    slices = []
    for r in range(world):
        out = C.c_void_p()
        nv.check(lib.meld_b200_stage2_finish(sts[r], nv.ptr(q_full), sp(), C.byref(out)), "stage2_finish")
        slices.append(DeviceGraph(out.value, device=Xd.device))
    torch.cuda.synchronize()
    for r in range(world):
        lib.meld_b200_stage2_destroy(sts[r])
        lib.meld_b200_cands_destroy(hs[r])
    assert [(g.row0, g.n_rows, g.n_cols) for g in slices] == [(bounds[r], bounds[r + 1] - bounds[r], N) for r in range(world)]
    M = sparse.vstack([g.to_scipy_L() for g in slices]).tocsr()
    perm = slices[0].permutation()
    coo = M.tocoo()
    M = sparse.csr_matrix((coo.data, (perm[coo.row], perm[coo.col])), shape=M.shape)
    M.sort_indices()
    assert _same_pattern(M, ref)
    assert np.array_equal(M.data, ref.data)
    # the slices drive the row-partitioned filter like slices cut from a full graph
    lmax = full.estimate_lmax()
    for g, r in zip(slices, range(world)):
        cut = full.row_slice(bounds[r], bounds[r + 1]).to_scipy_L()
        mine = g.to_scipy_L()
        assert _same_pattern(mine, cut) and np.array_equal(mine.data, cut.data)
    assert lmax > 0
