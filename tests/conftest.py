import os
import sys

# Tests that drive several "ranks" of the peer-store protocol inside ONE process let a kernel of rank A spin on a flag
# that a later kernel of rank B sets.  With CUDA's lazy module loading the first launch of a not-yet-loaded kernel can
# wait for the device to drain, i.e. for A's spin to time out.  Real multi-process runs are not affected.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["readme_toy", "blobs2k_k15", "blobs2k5_wagner", "batches2d_laplacian", "blobs1k5_aniso0"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def load_golden(name):
    """Golden vectors written by tests/golden/make_golden.py (oracle outputs)."""
    from scipy import sparse

    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    N = z["X"].shape[0]
    g = {k: z[k] for k in z.files}
    g["graph_kwargs"] = eval(str(z["graph_kwargs"]))  # written by make_golden.py as repr(dict)
    g["filter_kwargs"] = eval(str(z["filter_kwargs"]))
    g["K"] = sparse.csr_matrix((z["K_data"], z["K_indices"], z["K_indptr"]), shape=(N, N))
    g["L"] = sparse.csr_matrix((z["L_data"], z["L_indices"], z["L_indptr"]), shape=(N, N))
    g["lmax"] = float(z["lmax"])
    return g


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)


def density_parity(R, Rref, rtol=1e-5):
    """SURVEY.md section 8d parity criterion: per column, norm-wise and floored element-wise."""
    R = np.asarray(R, dtype=np.float64)
    Rref = np.asarray(Rref, dtype=np.float64)
    assert R.shape == Rref.shape
    colmax = np.abs(Rref).max(axis=0)
    err = np.abs(R - Rref)
    normwise = (err.max(axis=0) / colmax).max()
    elementwise_ok = np.all(err <= rtol * np.abs(Rref) + 1e-9 * colmax)
    return normwise, bool(elementwise_ok)


def row_pattern_hash(L):
    """Order-independent 64-bit hash of each row's column set (same as tests/golden/make_scale_digests.py)."""
    L = L.tocsr()
    h = (L.indices.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    out = np.zeros(L.shape[0], dtype=np.uint64)
    lens = np.diff(L.indptr)
    nz = lens > 0
    out[nz] = np.add.reduceat(h, L.indptr[:-1][nz])
    return out
