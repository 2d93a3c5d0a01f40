"""Digests of FULL-SIZE oracle runs for BASELINE.json configs 2 and 3 (tests/golden/scale_*.npz).

Run from the repo root (minutes of CPU per config):  python tests/golden/make_scale_digests.py [c3 c2pca c2raw]

A full oracle result at 50k - 90k cells is tens of MB; the GPU box has no /root/reference and spending GPU-box
minutes on a CPU ball tree buys nothing.  So the oracle (``oracle/``: the same scikit-learn / scipy calls the
reference stack makes) runs HERE on the seeded inputs of ``meld_b200.synthetic`` and only a digest is committed:

* per row of L: an order-independent 64-bit hash of the column pattern and the diagonal value (= weighted degree;
  it sums every off-diagonal value of the row, so a wrong value anywhere moves it),
* nnz(L), lmax (the oracle's ARPACK estimate, injected on the GPU side -- SURVEY H1),
* the densities of 8192 sampled rows plus per-column max / sum of the full density matrix.

``tests/test_gpu_scale.py`` rebuilds the inputs from the seed on the GPU box, runs the engine at full size with
default tuning and checks pattern hash, diagonal (1e-10) and densities (1e-5 gate, asserted far tighter).
"""

import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import graph as og, meld as om  # noqa: E402
from meld_b200 import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
HASH_MULT = np.uint64(0x9E3779B97F4A7C15)

# name -> (synthetic config, n_pca, extra MELD kwargs)
CASES = {
    "c3": ("c3", None, {}),
    "c2pca": ("c2", 100, {}),
    "c2raw": ("c2", None, {}),
}


def row_pattern_hash(L):
    """Order-independent hash of each row's column set: sum of (col + 1) * odd constant, mod 2^64."""
    L = L.tocsr()
    h = (L.indices.astype(np.uint64) + np.uint64(1)) * HASH_MULT
    out = np.zeros(L.shape[0], dtype=np.uint64)
    lens = np.diff(L.indptr)
    nz = lens > 0
    out[nz] = np.add.reduceat(h, L.indptr[:-1][nz])
    return out


def digest(name):
    cfg, n_pca, extra = CASES[name]
    X, y, kw = synthetic.make_config(cfg)
    kw = dict(kw, **extra)
    t0 = time.perf_counter()
    dens, g, lmax = om.fit_transform(X, y, n_pca=n_pca, random_state=0, n_jobs=os.cpu_count(), **kw)
    secs = time.perf_counter() - t0
    L = g["L"]
    N = L.shape[0]
    rows = np.sort(np.random.default_rng(7).choice(N, size=min(8192, N), replace=False))
    vals = dens.values
    out = dict(
        config=np.array(cfg), n_pca=np.array(-1 if n_pca is None else n_pca), meld_kwargs=np.array(repr(kw)),
        nnz=np.int64(L.nnz), lmax=np.float64(lmax),
        row_hash=row_pattern_hash(L), diag=L.diagonal(),
        dens_rows=rows, dens=vals[rows], dens_colmax=np.abs(vals).max(axis=0), dens_colsum=vals.sum(axis=0),
        samples=np.asarray(dens.columns).astype("U"), oracle_seconds=np.float64(secs), oracle_cores=np.int64(os.cpu_count()),
    )
    path = os.path.join(HERE, "scale_{}.npz".format(name))
    np.savez_compressed(path, **out)
    print(name, X.shape, "nnz/row %.1f" % (L.nnz / N), "lmax %.6g" % lmax, "%.0f s" % secs,
          os.path.getsize(path) // 1024, "KiB", flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        digest(nm)
