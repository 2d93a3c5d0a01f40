"""Where the end-to-end step spends the time that MELD.timings_ does not cover (run under gpurun).

    python tools/probe_e2e_gap.py [--config c4] [--steps 8]

Per step: constructor, fit_transform, release of the previous estimator + DataFrame, and fit_transform's own marks.
"""

import argparse
import gc
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meld_b200  # noqa: E402
from meld_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--steps", type=int, default=8)
    args = ap.parse_args()
    X, labels, kw = synthetic.make_config(args.config)
    pin = torch.empty(X.shape, dtype=torch.float64, pin_memory=True)
    pin.copy_(torch.from_numpy(X))
    Xp = pin.numpy()
    kw = {k: v for k, v in kw.items() if k in ("knn", "n_pca", "chebyshev_order")}
    op = dens = None
    for _ in range(3):
        op = meld_b200.MELD(verbose=0, **kw)
        dens = op.fit_transform(Xp, labels)
    gc.collect()
    gc.disable()
    for i in range(args.steps):
        t0 = time.perf_counter()
        new = meld_b200.MELD(verbose=0, **kw)
        t1 = time.perf_counter()
        nd = new.fit_transform(Xp, labels)
        t2 = time.perf_counter()
        op, dens = new, nd  # releases the previous estimator (graph_destroy) and DataFrame
        t3 = time.perf_counter()
        print(json.dumps(dict(step=i, ctor_ms=round(1e3 * (t1 - t0), 3), fit_transform_ms=round(1e3 * (t2 - t1), 3),
                              release_previous_ms=round(1e3 * (t3 - t2), 3),
                              marks_ms={k: round(1e3 * v, 2) for k, v in new.timings_.items()})), flush=True)


if __name__ == "__main__":
    main()
