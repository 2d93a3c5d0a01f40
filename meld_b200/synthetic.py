"""Seeded synthetic inputs for the BASELINE.json configs (host-side numpy only).

There is no network for real single-cell datasets, so benchmarks and parity
tests use these generators: Gaussian blobs with a
decaying per-dimension spectrum (PCA-like; isotropic blobs are the stated worst
case for kept-neighbours/row) and sample labels that correlate with geometry
(blob-dependent mixing plus a within-blob logistic gradient along dimension 0),
because uniform-random labels give near-constant densities and vacuous parity.
"""

from __future__ import annotations

import numpy as np

# name -> (N, D, n_blobs, n_samples, spectrum length scale, MELD kwargs)
CONFIGS = {
    "c1": dict(N=500, D=100, kind="readme", n_samples=2, meld=dict()),
    "c2": dict(N=50_000, D=2000, n_blobs=10, n_samples=3, tau=200.0, meld=dict(knn=15)),
    "c3": dict(N=90_000, D=100, n_blobs=20, n_samples=4, tau=10.0, meld=dict()),
    "c4": dict(N=500_000, D=100, n_blobs=20, n_samples=4, tau=10.0, meld=dict(knn=15, chebyshev_order=64)),
    "c5": dict(N=2_000_000, D=50, n_blobs=20, n_samples=6, tau=10.0, meld=dict(knn=10)),
    # adversarial for the pruned search: ONE isotropic Gaussian blob in 100 dimensions -- no cluster structure to prune
    # by and distance concentration (SURVEY finding 9: ~170+ kept neighbours per row)
    "c4iso": dict(N=500_000, D=100, n_blobs=1, n_samples=4, tau=1e12, meld=dict(knn=15, chebyshev_order=64)),
}


def make_blobs(N, D, n_blobs=20, n_samples=4, tau=10.0, seed=0, dtype=np.float64, centre_scale=3.0):
    """Blobs with per-dimension scale ``sigma_j = exp(-j / tau)``.

    Returns ``(X (N, D), labels (N,) of str)``.
    """
    rng = np.random.default_rng(seed)
    sigma = np.exp(-np.arange(D) / float(tau))
    centres = rng.normal(size=(n_blobs, D)) * (centre_scale * sigma)
    blob = rng.integers(0, n_blobs, size=N)
    X = np.empty((N, D), dtype=dtype)
    step = 1 << 16
    for s in range(0, N, step):
        e = min(N, s + step)
        X[s:e] = (rng.normal(size=(e - s, D)) * sigma + centres[blob[s:e]]).astype(dtype, copy=False)
    mix = rng.dirichlet(np.full(n_samples, 0.3), size=n_blobs)  # blob-dependent sample mixing
    grad = rng.normal(size=(n_blobs, n_samples))  # within-blob gradient along dim 0
    logits = np.log(mix[blob] + 1e-3) + grad[blob] * ((X[:, 0] - centres[blob, 0]) / sigma[0])[:, None]
    logits -= logits.max(axis=1, keepdims=True)
    prob = np.exp(logits)
    prob /= prob.sum(axis=1, keepdims=True)
    u = rng.random(N)
    lab = (prob.cumsum(axis=1) < u[:, None]).sum(axis=1).clip(0, n_samples - 1)
    names = np.array(["sample_{}".format(i) for i in range(n_samples)])
    return X, names[lab]


def make_readme_toy(seed=1):
    """The README example (reference ``README.md:50-57``), seeded."""
    rng = np.random.default_rng(seed)
    data = rng.normal(size=(500, 100))
    labels = rng.choice(["treatment", "control"], size=500)
    return data, labels


def make_config(name, seed=None, N=None, dtype=np.float64):
    """Inputs + MELD kwargs for a BASELINE.json config (``c1`` .. ``c5``); ``N`` overrides the size."""
    cfg = CONFIGS[name]
    if seed is None:
        seed = int("".join(ch for ch in name if ch.isdigit()))
    if cfg.get("kind") == "readme":
        X, y = make_readme_toy(seed)
        return X, y, dict(cfg["meld"])
    n = cfg["N"] if N is None else int(N)
    X, y = make_blobs(n, cfg["D"], cfg["n_blobs"], cfg["n_samples"], cfg["tau"], seed=seed, dtype=dtype)
    return X, y, dict(cfg["meld"])
