"""CPU-side tests: host logic, the oracle-independent parts of the API contract, and that the
C-ABI library loads and exports every symbol include/meld_b200.h declares (no compute calls)."""

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from meld_b200 import build

    return build.build(verbose=False)


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "meld_b200.h")).read()
    declared = set(re.findall(r"\b(meld_b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from meld_b200 import _native

    assert set(_native.SIGNATURES) == declared
    assert _native.lib().meld_b200_version() >= 100
    assert _native.lib().meld_b200_last_error() is not None


def test_cuda_sass_is_sm100a_and_uses_tma(built_lib):
    """The distance GEMM of the graph build is tcgen05 + TMA code, the Chebyshev SpMM gathers with 256-bit loads
    (SASS mnemonics of /opt/skills/guides/B200_PROFILING.md)."""
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "cheby_flat2_kernel" in sass and "cheby_step_kernel" not in sass
    assert "UTMALDG" in sass  # cp.async.bulk.tensor (TMA) operand tiles of the candidate search
    assert "UTCHMMA" in sass or "UTCMMA" in sass  # tcgen05.mma
    assert ".256" in sass  # 256-bit gathers of the signal rows


def test_invalid_parameters_raise_reference_messages():
    import meld_b200 as mb

    with pytest.raises(ValueError, match=re.escape(
            "lap_type value hello world not recognized. Choose from ['combinatorial', 'normalized']")):
        mb.MELD(verbose=0, lap_type="hello world")
    with pytest.raises(ValueError, match="Expected beta > 0"):
        mb.MELD(beta=-1)
    with pytest.raises(ValueError, match="Expected chebyshev_order integer"):
        mb.MELD(chebyshev_order=2.5)
    with pytest.raises(ValueError, match="solver value"):
        mb.MELD(solver="cg")
    op = mb.MELD()
    assert (op.beta, op.chebyshev_order, op.knn, op.decay, op.thresh, op.n_pca, op.anisotropy) == (60, 50, 5, 40, 1e-4, 100, 1)


def test_sample_labels_2d_message():
    import meld_b200 as mb

    labels = np.ones((10, 2))
    with pytest.raises(ValueError, match=re.escape(
            "sample_labels must be a single column. Got" "shape={}".format(labels.shape))):
        mb.MELD()._create_sample_indicators(labels)


def test_label_codes_follow_np_unique_order():
    import meld_b200 as mb

    rng = np.random.default_rng(0)
    for labels in [rng.choice(["B", "A", "C"], size=200), rng.integers(0, 5, 100), rng.choice([2.5, -1.0], 50),
                   rng.choice(["treatment", "control"], 64).reshape(-1, 1)]:
        samples, codes = mb.MELD()._label_codes(labels)
        uniq, inv = np.unique(labels, return_inverse=True)
        assert np.array_equal(samples, uniq) and samples.dtype == uniq.dtype
        assert np.array_equal(codes, inv.reshape(-1))
    ind = mb.MELD()._create_sample_indicators(np.array(["x", "y", "x", "z"]))
    assert list(ind.columns) == ["x", "y", "z"] and ind.values.tolist() == [[1, 0, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1]]


def test_cheby_coefficients_match_oracle():
    from meld_b200 import filter as mf
    from oracle import cheby

    for name, kw in [("heat", dict(beta=60)), ("laplacian", dict(beta=20, offset=0.1, order=2))]:
        for m in (1, 7, 50, 64):
            a = mf.cheby_coefficients(mf.filter_kernel(name, **kw), 0.173, m)
            b = cheby.cheby_coeff(cheby.filter_kernel(name, **kw), 0.173, m)
            np.testing.assert_allclose(a, b, rtol=0, atol=1e-15)
    with pytest.raises(NotImplementedError):
        mf.filter_kernel("gaussian", 1)


def test_product_path_fails_loudly_without_cuda():
    import torch
    import meld_b200 as mb
    from meld_b200._native import NativeError

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(NativeError, match="no CPU fallback"):
        mb.MELD(verbose=0).fit_transform(np.zeros((20, 3)), np.arange(20) % 2)
    with pytest.raises(NativeError):
        mb.normalize_densities(np.ones((4, 2)))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "meld_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_shard_bounds_are_tile_aligned_and_cover_all_rows():
    from meld_b200.graph import DeviceGraph

    for n, world in [(500_000, 8), (6000, 3), (1000, 4), (512, 2), (2_000_000, 8), (100, 8)]:
        b = DeviceGraph.shard_bounds(n, world)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1
        assert all(x <= y for x, y in zip(b[:-1], b[1:]))
        assert all(x % 512 == 0 for x, y in zip(b[:-1], b[1:]) if y > x)  # non-empty ranges start on a 512-row tile
    assert DeviceGraph.shard_bounds(6000, 3) == [0, 2048, 4096, 6000]


def test_fast_label_factorize_equals_pandas():
    """The raw-bytes hashing path of MELD._label_codes returns exactly pd.factorize's codes and uniques."""
    import pandas as pd

    from meld_b200.meld import _factorize

    rng = np.random.default_rng(3)
    for names, dtype in ((["sample_%d" % i for i in range(5)], None), (["a", "bb", "ccc"], "S3"),
                         (["treatment", "control"], None)):
        lab = np.array(names)[rng.integers(0, len(names), 20000)]
        if dtype:
            lab = lab.astype(dtype)
        c0, u0 = pd.factorize(lab)
        c1, u1 = _factorize(lab)
        assert np.array_equal(c0, c1) and list(u0) == list(u1)
    few = np.array(["x", "y", "x"])
    c1, u1 = _factorize(few)  # short inputs take the generic path
    assert list(c1) == [0, 1, 0] and list(u1) == ["x", "y"]


def test_device_pca_algorithm_equals_sklearn_on_cpu_tensors(monkeypatch):
    """meld_b200/pca.py restates scikit-learn's randomized PCA (same RandomState draw): run on CPU tensors here
    (the product path only ever hands it CUDA tensors), data_nu must equal sklearn's to rounding."""
    import torch
    from sklearn.decomposition import PCA

    from meld_b200 import _native as nv
    from meld_b200 import pca, synthetic

    monkeypatch.setattr(nv, "require_cuda", lambda: torch)
    for n, D, k, seed in ((1500, 200, 40, 0), (300, 700, 50, 3), (900, 120, 100, 1)):
        X, _ = synthetic.make_blobs(n, D, 5, 3, tau=D / 10.0, seed=7)
        ref = PCA(k, svd_solver="randomized", random_state=seed).fit(X).transform(X)  # graphtools: fit, then transform
        out, obj = pca.randomized_pca(torch.from_numpy(X), k, random_state=seed)
        assert np.abs(out.numpy() - ref).max() <= 1e-9 * np.abs(ref).max()
        assert tuple(obj.components_.shape) == (k, D) and obj.n_components_ == k
    with pytest.raises(ValueError):
        pca.randomized_pca(torch.zeros(10, 5, dtype=torch.float64), 6)


def test_fit_transform_prefetches_label_codes_on_a_thread(monkeypatch):
    """fit_transform factorises the labels while fit runs and hands the codes to transform; labels it cannot
    factorise (multi-column) are left for transform to reject with the reference's own error."""
    import meld_b200

    op = meld_b200.MELD(verbose=0)
    seen = {}
    monkeypatch.setattr(op, "fit", lambda X, **kw: op)

    def fake_transform(labels, _codes=None):
        seen["codes"] = _codes
        return "ok"

    monkeypatch.setattr(op, "transform", fake_transform)
    labels = np.array(["b", "a", "b", "c"] * 2000)
    assert op.fit_transform(np.zeros((8000, 3)), labels) == "ok"
    samples, codes = seen["codes"]
    ref_samples, ref_codes = np.unique(labels, return_inverse=True)
    assert list(samples) == list(ref_samples) and np.array_equal(codes, ref_codes)
    assert op.timings_["labels_thread"] >= 0
    op.fit_transform(np.zeros((4, 3)), np.zeros((4, 2)))
    assert seen["codes"] is None

    def boom(X, **kw):
        raise RuntimeError("build failed")

    monkeypatch.setattr(op, "fit", boom)
    with pytest.raises(RuntimeError, match="build failed"):
        op.fit_transform(np.zeros((8000, 3)), labels)


def test_round2_host_logic_without_a_gpu():
    """Row partitions, knn clamping, missing labels, filter kernels of the sweep -- everything that runs before the GPU."""
    import warnings

    import meld_b200
    from meld_b200 import distributed as mdist
    from meld_b200.graph import DeviceGraph, _clamp_knn

    # 512-row tiles dealt to ranks in contiguous runs; the last rank takes the remainder
    assert DeviceGraph.shard_bounds(6000, 3) == [0, 2048, 4096, 6000]
    assert DeviceGraph.shard_bounds(500000, 8)[-1] == 500000
    b = DeviceGraph.shard_bounds(500000, 8)
    assert all(x % 512 == 0 for x in b[:-1]) and all(b[i] < b[i + 1] for i in range(8))
    assert mdist.chunk_partition(500000, 8) == (62500, [62500 * r for r in range(9)])
    # graphtools clamps knn > n - 2 with a warning instead of failing
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert _clamp_knn(50, 20) == 18
        assert any("Cannot set knn" in str(x.message) for x in w)
    assert _clamp_knn(5, 20) == 5
    with pytest.raises(ValueError):
        _clamp_knn(5, 2)
    # NaN / None labels are refused, never merged into another label
    op = meld_b200.MELD(verbose=0)
    with pytest.raises(ValueError, match="missing values"):
        op._label_codes(np.array(["a", None, "b", "a"], dtype=object))
    # new constructor keywords are validated like the reference's
    with pytest.raises(ValueError, match="dist_mode value bogus not recognized"):
        meld_b200.MELD(verbose=0, dist_mode="bogus")
    assert meld_b200.MELD(verbose=0, dist_mode="nccl").dist_build == "replicated"
    assert meld_b200.MELD(verbose=0).dist_build == "rows"
    assert meld_b200.MELD(verbose=0, decay=None).decay is None  # decay=None is the binary kNN kernel now


def test_sharded_filter_rejects_bad_modes():
    from meld_b200.distributed import ShardedFilter

    with pytest.raises(ValueError, match="mode value tcp not recognized"):
        ShardedFilter(None, mode="tcp")


def test_pinned_result_pool_never_overwrites_a_live_result(monkeypatch):
    """meld.py::_pinned_result hands a block out again only when nothing refers to the array made from it; a caller
    that holds more results than the pool keeps gets None (the staged-copy path).  Pinned allocation and the stream
    synchronisation are replaced by host stand-ins: the bookkeeping is what is tested here."""
    import gc

    import pandas as pd
    import torch

    from meld_b200 import meld as mm

    allocs = []

    def fake_alloc(torch_, shape, dtype):
        allocs.append(shape)
        return torch.empty(shape, dtype=dtype)

    monkeypatch.setattr(mm, "_alloc_pinned", fake_alloc)
    monkeypatch.setattr(mm, "_sync_stream", lambda torch_, device: None)
    monkeypatch.setattr(mm, "_RESULT_POOL", [])
    src = [torch.full((50, 3), float(i), dtype=torch.float64) for i in range(8)]
    a = mm._pinned_result(torch, src[0])
    df = pd.DataFrame(a, columns=list("xyz"), copy=False)
    del a
    b = mm._pinned_result(torch, src[1])  # `df` still views the first block: a second one is allocated
    assert len(allocs) == 2 and float(df.values[0, 0]) == 0.0 and float(b[0, 0]) == 1.0
    del df
    gc.collect()
    c = mm._pinned_result(torch, src[2])  # the first block is free again: no new allocation
    assert len(allocs) == 2 and float(c[0, 0]) == 2.0 and float(b[0, 0]) == 1.0
    held = [b, c]
    while True:
        got = mm._pinned_result(torch, src[len(held)])
        if got is None:
            break
        held.append(got)
    assert len(held) == mm._RESULT_POOL_MAX  # every block is referenced: the pool does not grow past its cap
    assert [float(h[0, 0]) for h in held] == [1.0, 2.0, 2.0, 3.0]  # and none was overwritten
    other = mm._pinned_result(torch, torch.zeros((7, 2), dtype=torch.float64))
    assert other is None  # still capped, whatever the shape
    del held, b, c, got
    gc.collect()
    other = mm._pinned_result(torch, torch.ones((7, 2), dtype=torch.float64))  # unreferenced blocks of another shape go
    assert other is not None and other.shape == (7, 2) and len(mm._RESULT_POOL) == 1
    assert mm._pinned_result(torch, torch.zeros((0, 2), dtype=torch.float64)) is None


def test_read_only_inputs_do_not_warn():
    """DataFrame.values is read-only under pandas' copy-on-write; the engine only reads its inputs, so the torch view
    of them is made without torch's non-writable warning (and without a copy)."""
    import warnings

    from meld_b200 import _native as nv

    arr = np.arange(12.0).reshape(3, 4)
    arr.setflags(write=False)
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        t = nv.from_numpy_readonly(arr)
    assert t.data_ptr() == arr.ctypes.data and tuple(t.shape) == (3, 4)
    w = np.ones(3)
    assert nv.from_numpy_readonly(w).data_ptr() == w.ctypes.data
