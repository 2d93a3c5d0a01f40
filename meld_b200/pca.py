"""Randomised PCA on the device (SURVEY 8a row B', 8f rank 2).

The reference reduces raw gene matrices through graphtools ``Data._reduce_data`` ->
``sklearn.decomposition.PCA(n_pca, svd_solver="randomized", random_state=...)`` before the graph
is built (``meld/meld.py:117-118`` forwards ``n_pca=100``); on a 50k x 2000 input that host call
costs seconds while the whole GPU graph build + filter takes ~50 ms.

This is the same published algorithm (Halko, Martinsson, Tropp 2011, as scikit-learn runs it for
``svd_solver="randomized"``), restated on float64 CUDA tensors:

  mean-centre; Q0 = N(0,1) of shape (D, k + 10) drawn from the SAME numpy RandomState stream;
  n_iter = 7 if k < 0.1 min(N, D) else 4 power iterations Q <- orth(A Q), Q <- orth(A^T Q)
  (scikit-learn re-normalises with an LU factorisation, here with QR -- scikit-learn's own alternative
  normaliser; both keep span(Q), which is all the result depends on, so the outputs agree to rounding, and
  torch's LU would materialise an N x N permutation); Q = qr(A Q); B = Q^T A; B = Uh S Vt; U = Q Uh;
  signs fixed so that the largest |entry| of every row of Vt is positive;
  data_nu = X V_k^T - mean V_k^T, i.e. ``pca.fit(X); pca.transform(X)`` -- what graphtools runs -- NOT
  ``fit_transform`` (U_k S_k): for the randomized solver the two differ by the approximation residual.

The GEMMs / LU / QR / SVD are library calls (cuBLAS / cuSOLVER through torch) -- plumbing, not a kernel
of this engine.  Because the random draw and every step match, the result equals scikit-learn's
to rounding (tests/test_gpu_graph.py::test_device_pca_matches_sklearn), not merely in distribution.
"""

from __future__ import annotations

from . import _native as nv


class DevicePCA:
    """Fitted projection: ``mean_`` (D,), ``components_`` (k, D), ``singular_values_`` (k,),
    ``explained_variance_`` (k,) as CUDA tensors; ``transform`` maps new rows."""

    def __init__(self, mean, components, singular_values, n_samples):
        self.mean_ = mean
        self.components_ = components
        self.singular_values_ = singular_values
        self.explained_variance_ = singular_values**2 / max(n_samples - 1, 1)
        self.n_components_ = int(components.shape[0])
        self.n_samples_ = int(n_samples)

    def transform(self, Y):
        """scikit-learn ``PCA.transform``: ``Y @ components_.T - mean_ @ components_.T``."""
        import torch

        Y = torch.as_tensor(Y, dtype=torch.float64, device=self.mean_.device)
        return Y @ self.components_.T - (self.mean_[None, :] @ self.components_.T)


def randomized_pca(X, n_components, random_state=None, n_oversamples=10):
    """``pca = PCA(n_components, svd_solver="randomized", random_state=random_state).fit(X)`` followed by
    ``pca.transform(X)`` on the device -- the two calls graphtools ``Data._reduce_data`` makes (SURVEY 8a row
    B').  ``X``: (N, D) float64 CUDA tensor (not modified).  Returns ``(data_nu (N, k), DevicePCA)``."""
    torch = nv.require_cuda()
    from sklearn.utils import check_random_state  # the reference's RandomState handling, host side only

    if X.dtype != torch.float64:
        X = X.to(torch.float64)
    n, d = X.shape
    k = int(n_components)
    if not 1 <= k <= min(n, d):
        raise ValueError("n_components={} must be between 1 and min(n_samples, n_features)={}".format(k, min(n, d)))
    rs = check_random_state(random_state)
    mean = X.mean(dim=0)
    A = X - mean
    n_random = k + int(n_oversamples)
    n_iter = 7 if k < 0.1 * min(n, d) else 4
    transpose = n < d
    M = A.T if transpose else A
    Q = torch.from_numpy(rs.normal(size=(M.shape[1], n_random))).to(X.device)

    def orth(Y):
        return torch.linalg.qr(Y, mode="reduced")[0]

    if n_iter <= 2:
        for _ in range(n_iter):
            Q = M.T @ (M @ Q)
    else:
        for _ in range(n_iter):
            Q = orth(M @ Q)
            Q = orth(M.T @ Q)
    Q, _ = torch.linalg.qr(M @ Q, mode="reduced")
    B = Q.T @ M
    Uh, S, Vt = torch.linalg.svd(B, full_matrices=False)
    U = Q @ Uh
    if transpose:  # results back in the input's convention
        U, Vt = Vt.T, U.T
    # svd_flip(U, Vt, u_based_decision=False): the largest |entry| of each row of Vt becomes positive
    idx = Vt.abs().argmax(dim=1)
    signs = torch.sign(Vt[torch.arange(Vt.shape[0], device=Vt.device), idx])
    signs = torch.where(signs == 0, torch.ones_like(signs), signs)
    U = U * signs[None, :]
    Vt = Vt * signs[:, None]
    obj = DevicePCA(mean, Vt[:k].contiguous(), S[:k].contiguous(), n)
    # graphtools: data_pca.fit(X); data_nu = data_pca.transform(X).  U_k S_k (fit_transform) is NOT the same for
    # the randomized solver: it differs by the projection residual (I - Q Q^T) A V_k, far above rounding.
    return obj.transform(X).contiguous(), obj
