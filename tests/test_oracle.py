"""CPU tests that pin the oracle: the reference's own
known-answer test, the derived invariants, and the committed golden vectors."""

import numpy as np
import pandas as pd
import pytest
from scipy import sparse

from oracle import graph as og, cheby as oc, meld as om
from meld_b200 import synthetic
from conftest import density_parity


def _kat_inputs():
    # reference test/test_meld.py:46-57 (legacy global-seed RNG on purpose)
    np.random.seed(42)

    def norm(x):
        x = x.copy()
        x = x - np.min(x)
        x = x / np.max(x)
        return x

    data = np.random.normal(0, 2, (1000, 2))
    sample_labels = np.random.binomial(1, norm(data[:, 0]), 1000)
    return data, np.array(["treat" if val else "ctrl" for val in sample_labels])


@pytest.mark.parametrize("filt", ["heat", "laplacian"])
def test_kat_532(filt):
    """reference test/test_meld.py:59-81: knn=20, decay=10, thresh=0, anisotropy=0, exact solver."""
    data, labels = _kat_inputs()
    K = og.apply_anisotropy(og.symmetrize(og.traditional_kernel(data, knn=20, decay=10, thresh=0)), 0)
    L = og.laplacian(og.weights_from_kernel(K))
    dens = om.transform(L, None, labels, filter=filt, solver="exact", sample_normalize=False)
    assert list(dens.columns) == ["ctrl", "treat"]
    np.testing.assert_allclose(np.sum(dens.iloc[:, 1]), 532)
    np.testing.assert_allclose(np.sum(dens.iloc[:, 0]), 468)


def test_graph_matches_bruteforce_definition():
    """(vii) the search/expansion loop == 'all j with K_ij >= thresh' (expansion fires on the README toy)."""
    X, _ = synthetic.make_readme_toy(1)
    K, stats = og.knn_kernel(X, knn=5, return_stats=True)
    assert stats["n_overflow_first"] > 0
    Kb = og.knn_kernel_bruteforce(X, knn=5)
    assert (K != Kb).nnz == 0
    X, _ = synthetic.make_blobs(1500, 10, 5, 3, 3.0, seed=5)
    for knn, decay, thresh in [(3, 40.0, 1e-4), (10, 10.0, 1e-4), (15, 2.0, 1e-2)]:
        K = og.knn_kernel(X, knn=knn, decay=decay, thresh=thresh)
        Kb = og.knn_kernel_bruteforce(X, knn=knn, decay=decay, thresh=thresh)
        assert (K != Kb).nnz == 0
        np.testing.assert_allclose(K.diagonal(), 1.0)
        assert np.diff(K.indptr).min() >= knn + 1


def test_laplacian_invariants():
    """(iii) L symmetric, L 1 = 0, pattern(L) = pattern(K)."""
    X, y = synthetic.make_blobs(2000, 20, 6, 3, 5.0, seed=2)
    g = og.build_graph(X, knn=7, n_pca=None)
    L, K = g["L"], g["K"]
    assert abs(L - L.T).max() == 0
    assert np.abs(L @ np.ones(L.shape[0])).max() <= 1e-15 * abs(L).max() * 64
    assert (L != 0).nnz == (K != 0).nnz
    assert np.all(L.diagonal() > 0)


@pytest.mark.parametrize("filt,kw", [("heat", dict(beta=60)), ("laplacian", dict(beta=20, order=2)), ("heat", dict(beta=150))])
def test_chebyshev_matches_dense_spectral(filt, kw):
    """(iv) Chebyshev recurrence == U h(Lambda/lmax) U^T S with the same lmax."""
    X, y = synthetic.make_blobs(800, 10, 4, 3, 3.0, seed=3)
    g = og.build_graph(X, knn=5, n_pca=None)
    lmax = og.estimate_lmax(g["L"])
    _, S = om.sample_indicators(y)
    S = S / S.sum(axis=0)
    R = oc.cheby_filter(g["L"], lmax, S, filter=filt, chebyshev_order=120, **kw)
    Rd = oc.dense_spectral_filter(g["L"], lmax, S, filter=filt, **kw)
    assert np.abs(R - Rd).max() <= 1e-9 * np.abs(Rd).max()


def test_mass_conservation():
    """(ii) column sums are multiplied by h~(0) = c0/2 + sum_k (-1)^k c_k ~= 1."""
    X, y = synthetic.make_blobs(1500, 10, 4, 4, 3.0, seed=4)
    dens, g, lmax = om.fit_transform(X, y, n_pca=None, sample_normalize=False)
    _, S = om.sample_indicators(y)
    c = oc.cheby_coeff(oc.filter_kernel("heat", 60), lmax, 50)
    h0 = 0.5 * c[0] + sum((-1) ** k * c[k] for k in range(1, 51))
    np.testing.assert_allclose(dens.values.sum(axis=0), h0 * S.sum(axis=0), rtol=1e-10)
    assert abs(h0 - 1) < 1e-10


def test_permutation_equivariance():
    """(v) permuting cells permutes outputs; renaming samples permutes columns."""
    X, y = synthetic.make_blobs(1200, 8, 4, 3, 3.0, seed=6)
    dens, g, lmax = om.fit_transform(X, y, n_pca=None)
    perm = np.random.default_rng(0).permutation(len(X))
    dens_p, _, _ = om.fit_transform(X[perm], y[perm], n_pca=None, lmax=lmax)
    nw, ok = density_parity(dens_p.values, dens.values[perm], rtol=1e-9)
    assert nw < 1e-10 and ok


def test_readme_toy_contract():
    """(vi) README.md:50-60 shapes/columns; normalize_densities rows sum to 1."""
    X, y = synthetic.make_readme_toy(1)
    dens, g, lmax = om.fit_transform(X, y)
    assert dens.shape == (500, 2)
    assert list(dens.columns) == ["control", "treatment"]
    lik = om.normalize_densities(dens)
    assert isinstance(lik, pd.DataFrame)
    np.testing.assert_allclose(lik.values.sum(axis=1), 1.0)
    assert 40 < g["L"].nnz / 500 < 60


def test_golden_reproducible(golden):
    """The committed golden vectors are what the oracle produces today (lmax injected)."""
    g = og.build_graph(golden["X"], n_pca=None, **golden["graph_kwargs"])
    assert (g["K_knn"] != golden["K"]).nnz == 0
    assert abs(g["L"] - golden["L"]).max() <= 1e-13 * abs(golden["L"]).max()
    dens = om.transform(g["L"], golden["lmax"], golden["labels"], **golden["filter_kwargs"])
    assert list(dens.columns) == list(golden["samples"])
    nw, ok = density_parity(dens.values, golden["densities"], rtol=1e-9)
    assert nw < 1e-10 and ok


def test_indicator_errors():
    """error strings of meld/meld.py:164-167, 209-219."""
    with pytest.raises(ValueError, match="sample_labels must be a single column. Gotshape="):
        om.sample_indicators(np.ones((10, 2)))
    L = sparse.identity(5, format="csr")
    with pytest.raises(ValueError, match="are not of the same size"):
        om.transform(L, 1.0, np.ones((6, 2), dtype=str))
    with pytest.raises(ValueError, match="Found only one unqiue sample label"):
        om.transform(L, 1.0, np.ones(5))


def test_tile_pruning_bounds_are_lower_bounds():
    """The two inequalities the pruned candidate search relies on, restated in numpy and checked by
    brute force: for cells x of one tile segment and y of another,
      |x - y| >= |c_R - c_C| - rho_R - rho_C                      (balls around the segments' own centroids)
      |x - y| >= -max_x p_AB(x) - max_y p_BA(y), p_AB(x) = w.(x - m)   (axis between ANY two centroids A != B)
    hold for arbitrary centroids -- the k-means clustering is only a heuristic and cannot break exactness."""
    rng = np.random.default_rng(42)
    d, C = 12, 5
    cent = rng.normal(size=(C, d)) * 3.0  # "k-means centroids": any vectors will do
    pts = [cent[k] + rng.normal(size=(rng.integers(20, 60), d)) * rng.uniform(0.3, 1.5) for k in range(C)]
    cnorm = (cent ** 2).sum(1)
    worst_ball = worst_proj = np.inf
    for A in range(C):
        for B in range(C):
            X, Y = pts[A], pts[B]
            D = np.sqrt(((X[:, None, :] - Y[None, :, :]) ** 2).sum(-1))
            cR, cC = X.mean(0), Y.mean(0)
            rhoR = np.sqrt(((X - cR) ** 2).sum(1)).max()
            rhoC = np.sqrt(((Y - cC) ** 2).sum(1)).max()
            lb_ball = np.linalg.norm(cR - cC) - rhoR - rhoC
            worst_ball = min(worst_ball, (D - lb_ball).min())
            if A != B:
                dist = np.linalg.norm(cent[B] - cent[A])
                # the kernel's form: p_AB(x) = (g_B - g_A - (|c_B|^2 - |c_A|^2)/2) / |c_B - c_A|, g_K = c_K . x
                pAB = (X @ cent[B] - X @ cent[A] - 0.5 * (cnorm[B] - cnorm[A])) / dist
                pBA = (Y @ cent[A] - Y @ cent[B] - 0.5 * (cnorm[A] - cnorm[B])) / dist
                lb_proj = -pAB.max() - pBA.max()
                worst_proj = min(worst_proj, (D - lb_proj).min())
    assert worst_ball >= -1e-9 and worst_proj >= -1e-9, (worst_ball, worst_proj)


def test_sampled_bruteforce_rows_equal_the_ball_tree_kernel():
    """oracle.graph.knn_kernel_rows_bruteforce (the checker of the 500k / 2M-cell GPU tests) reproduces the rows of
    the ball-tree kernel exactly on a size where both run."""
    from meld_b200 import synthetic

    X, _ = synthetic.make_blobs(5000, 40, 6, 3, 8.0, seed=3)
    for knn, decay in ((9, 40.0), (5, 10.0)):
        K = og.knn_kernel(X, knn=knn, decay=decay)
        rows = np.random.default_rng(0).choice(X.shape[0], 200, replace=False)
        for (cols, vals), i in zip(og.knn_kernel_rows_bruteforce(X, rows, knn=knn, decay=decay), rows):
            a, b = K.indptr[i], K.indptr[i + 1]
            assert np.array_equal(K.indices[a:b], cols)
            assert np.abs(K.data[a:b] - vals).max() <= 1e-12


def test_decay_none_is_the_binary_knn_kernel():
    X = np.random.default_rng(5).normal(size=(400, 6))
    K = og.knn_kernel(X, knn=4, decay=None)
    assert K.shape == (400, 400) and np.all(K.data == 1.0) and np.all(np.diff(K.indptr) == 5)
    assert np.all(K.diagonal() == 1.0)
    g = og.build_graph(X, knn=4, decay=None, n_pca=None)
    assert abs(g["L"] - g["L"].T).max() == 0.0
