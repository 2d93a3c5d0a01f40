"""Secondary bench line: the reference's parameter-search pattern on one cached graph.

    python tools/bench_sweep.py [--cells 26827] [--labels 25] [--betas 199]

The reference's dominant real workload is `meld/benchmark.py:186-200` in a loop: one graph, then
`MELD(beta=b).fit(graph).transform(labels)` for 25 label draws x 199 beta per graph
(`notebooks/MELD_Quickstart.ipynb:729-747`; "more than 12 hours on a 36 core server" for 24 graphs).  Here:
`MELD.transform_sweep` (one Chebyshev recurrence per 8 signal columns, every beta only its coefficient vector) against
looping `MELD.transform` on the same GPU and against the reference's CPU arithmetic (oracle: scipy Chebyshev filter on
the SAME graph exported to the host, 1 thread like scipy's matvec) -- which also checks the sweep's numbers.
Prints ONE JSON line.
"""

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meld_b200  # noqa: E402
from meld_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=26827, help="cells of the graph (the notebook's dataset: 26 827)")
    ap.add_argument("--labels", type=int, default=25)
    ap.add_argument("--betas", type=int, default=199)
    ap.add_argument("--knn", type=int, default=7)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    X, _, _ = synthetic.make_config("c3", N=args.cells)
    rng = np.random.default_rng(1)
    graph = meld_b200.DeviceGraph.from_data(X, knn=args.knn)
    op = meld_b200.MELD(verbose=0).fit(graph)
    label_sets = [rng.choice(["ctrl", "expt"], size=args.cells) for _ in range(args.labels)]
    betas = list(range(1, args.betas + 1))
    lmax = graph.estimate_lmax()
    nt = len(label_sets) * len(betas)
    op.transform_sweep(label_sets[:2], betas=betas[:4], as_tensor=True)  # warm-up
    torch.cuda.synchronize()
    times = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        R, cols = op.transform_sweep(label_sets, betas=betas, as_tensor=True)  # labels on the host, densities on the device
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    sweep_s = float(np.median(times))
    # the same work one transform at a time on the GPU (a sample, scaled)
    k, t0 = 0, time.perf_counter()
    for lab in label_sets[:3]:
        for b in betas[:20]:
            meld_b200.MELD(verbose=0, beta=b).fit(graph).transform(lab)
            k += 1
    loop_rate = k / (time.perf_counter() - t0)
    # the reference's CPU arithmetic on the same graph: scipy Chebyshev filter, 1 thread; also the parity check
    from oracle import meld as omeld  # the checker / CPU baseline, outside the timed GPU region

    L = graph.to_scipy_L()
    checks, t0, worst = 0, time.perf_counter(), 0.0
    for li in (0, len(label_sets) - 1):
        for b in (betas[0], betas[len(betas) // 3], betas[-1]):
            ref = omeld.transform(L, lmax, label_sets[li], beta=b)
            got = R[betas.index(b), :, 2 * li: 2 * li + 2].cpu().numpy()
            worst = max(worst, float(np.abs(got - ref.values).max() / np.abs(ref.values).max()))
            checks += 1
    cpu_rate = checks / (time.perf_counter() - t0)
    line = {
        "metric": "transforms/sec on a cached graph (parameter sweep: label draws x beta)", "value": nt / sweep_s,
        "unit": "transforms/s", "n_gpus": 1, "higher_is_better": True, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "{} cells x 100 dims (config 3 generator), knn={}, {} label draws x {} beta = {} transforms, "
                               "chebyshev_order=50".format(args.cells, args.knn, args.labels, args.betas, nt),
                   "nnz_L": graph.nnz, "lmax": lmax},
        "seconds_per_sweep": {"median": sweep_s, "all": [round(t, 4) for t in times]},
        "looping_MELD_transform_on_the_gpu": {"value": loop_rate, "unit": "transforms/s"},
        "cpu_baseline": {"value": cpu_rate, "unit": "transforms/s", "cores": 1, "kind": "port",
                         "sample": "oracle.meld.transform (scipy Chebyshev filter, 1 thread) on the exported graph, {} "
                                   "(label set, beta) pairs".format(checks)},
        "parity": {"normwise_max_over_checked_pairs": worst, "pairs": checks},
        "output": "device tensor ({}, {}, {}) float64 = {:.2f} GB".format(R.shape[0], R.shape[1], R.shape[2],
                                                                         R.numel() * 8 / 1e9),
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
