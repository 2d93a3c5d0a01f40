"""The reference's own API tests (``test/test_meld.py``, ``test/test_utils.py``), run against this package through
the drop-in import name ``meld``.  Same data recipes, sizes, calls and expected messages; each test cites the lines it
mirrors.  What the reference tests that is outside the hot path is stated where it is left out: VertexFrequencyCluster
(``test/test_meld.py:203-385``), the Benchmarker (``test/test_benchmark.py``), ``get_meld_cmap`` and MNN graphs
(``sample_idx=``: an explicit NotImplementedError here, asserted below).  The numerical known-answer test of the
reference (``test_meld``, sum of the density = 532) is ``tests/test_gpu_filter.py::test_reference_kat_532_on_the_product``.
"""

import re

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

CELLS = [100, 1000]  # the reference's tests use 100 cells x 2 dims; 1000 is the size of its known-answer test


@pytest.fixture(scope="module")
def meld():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import meld as meld_pkg

    return meld_pkg


@pytest.fixture(params=CELLS)
def N_CELLS(request):
    return request.param


def make_batches(n_pts_per_cluster=250, seed=0):
    """test/utils/__init__.py:38-72: two batches of three Gaussian clusters each, labels ctrl / expt."""
    rng = np.random.RandomState(seed)

    def make(x, y, s):
        return np.concatenate([rng.normal(x, s, (n_pts_per_cluster, 1)), rng.normal(y, s, (n_pts_per_cluster, 1))], axis=1)

    data = np.concatenate([make(0, 0, 0.1), make(1, 1, 0.1), make(0, 1, 0.1),
                           make(1, -1, 0.1), make(2, 0, 0.1), make(-2, -1, 0.1)], axis=0)
    labels = np.concatenate([np.zeros(3 * n_pts_per_cluster), np.ones(3 * n_pts_per_cluster)])
    return data, np.array(["expt" if v else "ctrl" for v in labels])


def test_check_pygsp_graph(meld):
    """test/test_meld.py:17-28 (the accepted graph type here is the engine's DeviceGraph)."""
    data = np.random.RandomState(0).normal(0, 2, (10, 2))
    G = meld.DeviceGraph.from_data(data)
    assert meld.utils._check_pygsp_graph(G) is G
    with pytest.raises(TypeError, match=re.escape("Input graph should be of type graphtools.base.BaseGraph. "
                                                  "With graphtools, use the `use_pygsp=True` flag.")):
        meld.utils._check_pygsp_graph(G="hello world")


def test_mnn_is_refused_explicitly(meld):
    """test/test_meld.py:31-40 builds an MNN graph (sample_idx=labels); this engine refuses it before touching the GPU."""
    data, labels = make_batches(n_pts_per_cluster=50)
    with pytest.raises(NotImplementedError):
        meld.MELD(verbose=0).fit_transform(data, labels, sample_idx=labels)


def test_meld_invalid_lap_type(meld, N_CELLS):
    """test/test_meld.py:96-105."""
    data = np.random.RandomState(1).normal(0, 2, (10 * N_CELLS, 2))
    lap_type = "hello world"
    with pytest.raises(ValueError, match=re.escape("lap_type value {} not recognized. "
                                                   "Choose from ['combinatorial', 'normalized']".format(lap_type))):
        meld.MELD(verbose=0, lap_type=lap_type).fit(data)


def test_meld_labels_wrong_shape(meld, N_CELLS):
    """test/test_meld.py:108-120."""
    data = np.random.RandomState(2).normal(0, 2, (N_CELLS, 2))
    sample_labels = np.ones([N_CELLS + 1, 2], dtype=str)
    with pytest.raises(ValueError, match=re.escape("Input data ({}) and input graph ({}) "
                                                   "are not of the same size".format(sample_labels.shape, data.shape[0]))):
        meld.MELD(verbose=0).fit_transform(X=data, sample_labels=sample_labels)


def test_meld_label_2d(meld, N_CELLS):
    """test/test_meld.py:123-139: an (N, 1) DataFrame of string labels with an index."""
    data = np.random.RandomState(3).normal(0, 2, (N_CELLS, 2))
    index = pd.Index(["cell_{}".format(i) for i in range(N_CELLS)])
    half = N_CELLS // 2
    sample_labels = pd.DataFrame(np.concatenate([np.zeros((half, 1)), np.ones((N_CELLS - half, 1))]), index=index,
                                 columns=pd.Index(["A"]), dtype=str)
    dens = meld.MELD(verbose=0).fit_transform(X=data, sample_labels=sample_labels)
    assert dens.shape == (N_CELLS, 2) and np.all(np.isfinite(dens.values))


def test_meld_label_dataframe(meld, N_CELLS):
    """test/test_meld.py:142-159."""
    data = np.random.RandomState(4).normal(0, 2, (N_CELLS, 2))
    index = pd.Index(["cell_{}".format(i) for i in range(N_CELLS)])
    half = N_CELLS // 2
    sample_labels = pd.DataFrame(np.concatenate([np.zeros(half), np.ones(N_CELLS - half)]), index=index,
                                 columns=["sample_labels"], dtype=str)
    sample_densities = meld.MELD(verbose=0).fit_transform(X=data, sample_labels=sample_labels)
    assert np.all(sample_densities.index == index)
    assert np.all(sample_densities.columns == pd.Index(np.unique(sample_labels)))


def test_meld_labels_non_numeric(meld, N_CELLS):
    """test/test_meld.py:162-172; the densities are also checked against the CPU oracle on the same graph."""
    from oracle import meld as omeld

    rng = np.random.RandomState(5)
    data = rng.normal(size=(N_CELLS, 2))
    sample_labels = rng.choice(["A", "B"], size=N_CELLS)
    meld.MELD(verbose=0).fit_transform(data, sample_labels)
    sample_labels = rng.choice(["A", "B", "C"], size=N_CELLS)
    meld_op = meld.MELD(verbose=0)
    sample_densities = meld_op.fit_transform(data, sample_labels)
    assert np.all(sample_densities.columns == ["A", "B", "C"])
    ref = omeld.transform(meld_op.graph.to_scipy_L(), meld_op.graph.lmax, sample_labels)
    assert np.abs(sample_densities.values - ref.values).max() <= 1e-9 * np.abs(ref.values).max()


def test_sample_labels_2d(meld):
    """test/test_meld.py:175-181."""
    labels = np.ones((10, 2))
    with pytest.raises(ValueError, match=re.escape("sample_labels must be a single column. Got"
                                                   "shape={}".format(labels.shape))):
        meld.MELD()._create_sample_indicators(labels)


def test_sample_labels_one_sample(meld, N_CELLS):
    """test/test_meld.py:184-192."""
    data = np.random.RandomState(6).normal(size=(N_CELLS, 2))
    labels = np.ones(N_CELLS)
    with pytest.raises(ValueError, match=re.escape("Found only one unqiue sample label. Cannot estimate density "
                                                   "of a single sample.")):
        meld.MELD(verbose=0).fit_transform(data, labels)


def test_utils(meld):
    """test/test_utils.py:9-31 without the MNN graph, VertexFrequencyCluster and the colour map: a prebuilt graph goes
    into fit_transform, normalize_densities gives rows that sum to one."""
    data, labels = make_batches(n_pts_per_cluster=250)
    G = meld.DeviceGraph.from_data(data)
    meld_op = meld.MELD(verbose=0)
    sample_densities = meld_op.fit_transform(G, labels)
    sample_likelihoods = meld.utils.normalize_densities(sample_densities)
    assert list(sample_likelihoods.columns) == ["ctrl", "expt"] and sample_likelihoods.shape == (1500, 2)
    np.testing.assert_allclose(sample_likelihoods.sum(axis=1).values, 1.0, rtol=1e-12)
    # the batches do not overlap: each cell's likelihood is that of its own batch
    assert (sample_likelihoods["expt"].values[labels == "expt"] > 0.99).all()
    assert (sample_likelihoods["ctrl"].values[labels == "ctrl"] > 0.99).all()
    np.testing.assert_allclose(meld.utils.normalize_densities(sample_densities=np.ones([100, 3])), 1.0 / 3)
    np.testing.assert_allclose(meld.utils.normalize_densities(sample_densities=np.ones([100, 2])), 0.5)
