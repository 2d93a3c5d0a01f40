"""Time the Chebyshev filter on the REAL kNN graph of a BASELINE config for several launch configurations
(development probe, run under gpurun).  The graph is rebuilt per configuration because the block
dictionaries are part of graph finalisation."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meld_b200  # noqa: E402
from meld_b200 import _native as nv, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--configs", default="[{}]")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--no-pca", action="store_true", help="n_pca=None: build the graph on the raw features")
    args = ap.parse_args()
    cfg = synthetic.CONFIGS[args.config]
    n = args.cells or cfg["N"]
    Xh, labels, kw = synthetic.make_config(args.config, N=n)
    if args.no_pca:
        kw = dict(kw, n_pca=None)
    X = torch.from_numpy(Xh).cuda()
    samples, codes = np.unique(labels, return_inverse=True)
    codes = torch.from_numpy(codes.astype(np.int32)).cuda()
    p, m = len(samples), kw.get("chebyshev_order", 50)
    base = dict(blk_chunk=768, stage_cap=1024, dict_cap=768, row_cap=64, n_stage=0, threads=512, gather_warps=3,
                team_warps=4, ctas_per_sm=1, group=0, use_dict=1, reorder=1)
    for c in json.loads(args.configs):
        full = dict(base)
        full.update(c)
        nv.set_tuning(**full)
        op = meld_b200.MELD(verbose=0, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        op.fit(X)
        torch.cuda.synchronize()
        t_fit = time.perf_counter() - t0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        op.graph.estimate_lmax()
        torch.cuda.synchronize()
        t_lmax = time.perf_counter() - t0
        st = op.graph.build_stats()
        nnz = op.graph.nnz
        best = 1e9
        for _ in range(args.reps + 1):
            op.profile_events = []
            op.transform_device(codes, p)
            torch.cuda.synchronize()
            e0, e1, _m = op.profile_events[0]
            best = min(best, e0.elapsed_time(e1))
        bytes_step = nnz * 12 + (n + 1) * 4 + 5 * n * p * 8
        gbs = m * bytes_step / (best * 1e-3) / 1e9
        print(json.dumps(dict(cfg=c, fit_s=round(t_fit, 3), filter_ms=round(best, 3), us_step=round(best * 1e3 / m, 1),
                              GBs=round(gbs, 1), frac=round(gbs / 6538.6, 3), nnz=nnz, direct=st["direct_blocks"],
                              blocks=st["row_blocks"], dict_per_nnz=round(st["dict_total"] / nnz, 3),
                              lmax_iters=op.graph.lmax_iters, lmax_ms=round(1e3 * t_lmax, 2))), flush=True)
        del op


if __name__ == "__main__":
    main()
