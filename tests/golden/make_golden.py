"""Generate the golden vectors under tests/golden/ from the CPU oracle.

Run from the repo root:  python tests/golden/make_golden.py

The reference stack itself (graphtools + pygsp) cannot be imported in this
image, so these fixtures come from ``oracle/`` -- the
restatement that makes the same scikit-learn / scipy calls -- after it passed
the reference's own known-answer test (sum == 532, ``test/test_meld.py:72-81``)
and the invariants in ``tests/test_oracle.py``.  Each ``.npz`` holds the inputs,
the graph (CSR of the un-symmetrised kernel and of L), the ``lmax`` the oracle
used, and the densities, so GPU parity tests need neither the oracle's ARPACK
start vector nor ``/root/reference`` at run time.
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import graph as og, meld as om  # noqa: E402
from meld_b200 import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def two_d_batches(n_per=300, seed=7):
    """Shape of the reference fixture ``test/utils/__init__.py:37-71`` (six 2-D blobs, two batches)."""
    rng = np.random.default_rng(seed)
    cs = [(0, 0), (1, 1), (0, 1), (1, -1), (2, 0), (-2, -1)]
    X = np.concatenate([rng.normal(c, 0.1, (n_per, 2)) for c in cs])
    y = np.array(["ctrl"] * (3 * n_per) + ["expt"] * (3 * n_per))
    return X, y


CASES = {
    # name: (inputs fn, graph kwargs, filter kwargs)
    "readme_toy": (lambda: synthetic.make_readme_toy(1), dict(), dict()),
    "blobs2k_k15": (
        lambda: synthetic.make_blobs(2000, 20, 6, 3, 5.0, seed=11),
        dict(knn=15),
        dict(chebyshev_order=64),
    ),
    "blobs2k5_wagner": (
        lambda: synthetic.make_blobs(2500, 50, 8, 4, 10.0, seed=12),
        dict(knn=7),
        dict(beta=67),
    ),
    "batches2d_laplacian": (lambda: two_d_batches(), dict(knn=5, decay=10.0), dict(filter="laplacian", beta=20, order=2)),
    "blobs1k5_aniso0": (
        lambda: synthetic.make_blobs(1500, 30, 4, 2, 8.0, seed=13),
        dict(knn=10, anisotropy=0.0, thresh=1e-3),
        dict(offset=0.1, order=2, sample_normalize=False),
    ),
}


def main():
    for name, (fn, gkw, fkw) in CASES.items():
        X, y = fn()
        dens, g, lmax = om.fit_transform(X, y, n_pca=None, **gkw, **fkw)
        K0, L = g["K_knn"], g["L"]
        out = dict(
            X=X,
            labels=np.asarray(y).astype("U"),
            graph_kwargs=np.array(repr(gkw)),
            filter_kwargs=np.array(repr(fkw)),
            K_indptr=K0.indptr.astype(np.int64),
            K_indices=K0.indices.astype(np.int32),
            K_data=K0.data,
            L_indptr=L.indptr.astype(np.int64),
            L_indices=L.indices.astype(np.int32),
            L_data=L.data,
            lmax=np.float64(lmax),
            densities=dens.values,
            samples=np.asarray(dens.columns).astype("U"),
        )
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, X.shape, "nnz(L)/row=%.1f" % (L.nnz / L.shape[0]), "lmax=%.6g" % lmax, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
