"""Round-2 development probe (run under gpurun; prints JSON lines, not a bench line).

  python tools/r02_probe.py [--config c4] [--what spmm,lmax,sweep,lines]

spmm   Chebyshev SpMM launch variants on the REAL graph of the config (built once): us per launch, algorithmic
       GB/s and fraction of the measured HBM peak for p = 4 (bench width), 8 and 1.
lmax   Lanczos time / iterations with the Ritz-residual stop at several tolerances.
sweep  transform_sweep (shared basis) vs looping transform on a c3-sized graph: transforms per second.
lines  how many 128-byte lines of T a row's gathers touch (distinct col >> 2 per row / nnz): what same-line
       coalescing could save at p = 4.
"""

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meld_b200  # noqa: E402
from meld_b200 import _native as nv, synthetic  # noqa: E402


def peak():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    return json.load(open(path))["hbm_gbs"] if os.path.exists(path) else 6650.0


def time_filter(graph, lmax, coeffs, S, reps=4):
    lib = nv.lib()
    R = torch.empty_like(S)
    cptr = coeffs.ctypes.data_as(C.POINTER(C.c_double))
    best = 1e9
    for rep in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nv.check(lib.meld_b200_cheby_filter(graph._h, float(lmax), cptr, len(coeffs), nv.ptr(S), S.shape[1], nv.ptr(R),
                                            nv.current_stream_ptr()), "cheby_filter")
        e1.record()
        torch.cuda.synchronize()
        if rep > 0:
            best = min(best, e0.elapsed_time(e1))
    return best, R


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--what", default="spmm,lmax,lines,sweep")
    args = ap.parse_args()
    what = set(args.what.split(","))
    X, labels, kw = synthetic.make_config(args.config, N=args.cells)
    n = X.shape[0]
    m = kw.get("chebyshev_order", 50)
    t0 = time.perf_counter()
    graph = meld_b200.DeviceGraph.from_data(X, knn=kw.get("knn", 5))
    torch.cuda.synchronize()
    print(json.dumps(dict(config=args.config, n=n, nnz=graph.nnz, nnz_per_row=graph.nnz / n,
                          first_build_s=round(time.perf_counter() - t0, 3))), flush=True)
    pk = peak()
    if "lmax" in what:
        for tol in (0.0, 1e-4, 1e-5, 1e-7, 1e-10):
            graph._lmax = None
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                graph._lmax = None
                t1 = time.perf_counter()
                lm = graph.estimate_lmax(rel_tol=tol)
                best = min(best, time.perf_counter() - t1)
            print(json.dumps(dict(lmax_tol=tol, lmax=lm, iters=graph.lmax_iters, ms=round(1e3 * best, 3))), flush=True)
    graph._lmax = None
    lmax = graph.estimate_lmax()
    coeffs = np.ascontiguousarray(meld_b200.filter.cheby_coefficients(meld_b200.filter.filter_kernel("heat", 60), lmax, m))
    if "spmm" in what:
        rng = np.random.default_rng(0)
        base = dict(flat_gen=0, flat_hint=0, flat_layout=0, flat_threads=1024, flat_group=0, flat_pipe=2)
        variants = [
            ("r01 flat kernel", dict()),
            ("flat2 hint0", dict(flat_gen=1)),
            ("flat2 hint1 (stream no_allocate)", dict(flat_gen=1, flat_hint=1)),
            ("flat2 hint2 (all no_allocate)", dict(flat_gen=1, flat_hint=2)),
            ("flat2 hint3 (stream NA, gathers evict_last)", dict(flat_gen=1, flat_hint=3)),
            ("flat2 layout1 hint0", dict(flat_gen=1, flat_layout=1)),
            ("flat2 layout1 hint1", dict(flat_gen=1, flat_layout=1, flat_hint=1)),
            ("r01 flat G=4", dict(flat_group=4)),
            ("r01 flat G=16", dict(flat_group=16)),
            ("r01 flat 768 threads", dict(flat_threads=768)),
            ("r01 flat pipelined", dict(flat_pipe=1)),
        ]
        for p in (4, 8, 1):
            S = torch.from_numpy(rng.normal(size=(n, p))).cuda()
            bytes_step = graph.nnz * 12 + (n + 1) * 4 + 5 * n * p * 8
            ref = None
            for name, tn in variants:
                if p != 4 and ("G=" in name or "768" in name):
                    continue
                nv.set_tuning(**dict(base, **tn))
                ms, R = time_filter(graph, lmax, coeffs, S)
                if ref is None:
                    ref = R.clone()
                err = float((R - ref).abs().max() / ref.abs().max())
                us = 1e3 * ms / m
                gbs = bytes_step / (us * 1e-6) / 1e9
                print(json.dumps(dict(p=p, variant=name, us_per_launch=round(us, 2), GBs=round(gbs, 1),
                                      frac=round(gbs / pk, 4), rel_diff_vs_first=err)), flush=True)
            nv.set_tuning(**base)
    if "lines" in what:
        L = graph.to_scipy_L() if False else None  # caller order is irrelevant here: use the internal CSR
        ip = torch.empty(n + 1, dtype=torch.int64, device="cuda")
        ci = torch.empty(graph.nnz, dtype=torch.int32, device="cuda")
        va = torch.empty(graph.nnz, dtype=torch.float64, device="cuda")
        nv.check(nv.lib().meld_b200_graph_export_csr(graph._h, nv.ptr(ip), nv.ptr(ci), nv.ptr(va),
                                                     nv.current_stream_ptr()), "export")
        rows = torch.repeat_interleave(torch.arange(n, device="cuda"), (ip[1:] - ip[:-1]))
        for shift, what_ in ((2, "128B lines at p=4"), (1, "128B lines at p=8 / 64B pairs at p=4"), (4, "128B lines at p=1")):
            key = rows * (1 << 40) + (ci.long() >> shift)
            distinct = int(torch.unique(key).numel())
            print(json.dumps(dict(stat="distinct (row, col>>%d)" % shift, what=what_, distinct=distinct,
                                  per_nnz=round(distinct / graph.nnz, 4))), flush=True)
        inrange = 0
        sm = 148
        # fraction of columns inside the owning CTA's contiguous row range (rows cut into 148 equal-nnz pieces)
        cuts = torch.searchsorted(ip, torch.arange(sm + 1, device="cuda") * (graph.nnz // sm))
        cta = torch.bucketize(rows, cuts[1:-1].contiguous(), right=True)
        lo, hi = cuts[cta], cuts[cta + 1]
        inrange = int(((ci.long() >= lo) & (ci.long() < hi)).sum())
        print(json.dumps(dict(stat="columns inside the CTA's own row range", frac=round(inrange / graph.nnz, 4))), flush=True)
        del L
    if "prune" in what:
        # cell-order / pruning knobs of the candidate search: ms of a whole build and of its two GEMM passes
        Xd = torch.from_numpy(X).cuda()
        base = dict(clusters=64, kmeans_iters=1, cluster_cells=1024, prune_window=2, tl_chunks=2, tc_multicast=2)
        base["km_var_pct"] = 99
        trials = [dict(), dict(km_var_pct=0), dict(km_var_pct=95), dict(km_var_pct=90), dict(kmeans_iters=0),
                  dict(clusters=96), dict(clusters=128), dict(clusters=32), dict(kmeans_iters=2),
                  dict(prune_window=1), dict(prune_window=3), dict(tl_chunks=1),
                  dict(tl_chunks=4), dict(tc_multicast=1), dict(tc_multicast=4)]
        if os.environ.get("PROBE_PRUNE_SHORT"):
            trials = trials[:5]
        for tr in trials:
            nv.set_tuning(**dict(base, **tr))
            ts = []
            for rep in range(4):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                gg = meld_b200.DeviceGraph.from_data(Xd, knn=kw.get("knn", 5))
                e1.record()
                torch.cuda.synchronize()
                if rep:
                    ts.append(e0.elapsed_time(e1))
                bt, st = gg.build_times(), gg.build_stats()
                gg.close()
            print(json.dumps(dict(prune=tr, build_ms=round(float(np.median(ts)), 3), pass1_ms=round(bt["pass1_ms"], 3),
                                  pass2_ms=round(bt["pass2_ms"], 3), candidates=st["candidate_cap"],
                                  tile_pairs_kept=round(bt["flops_per_pass"] / max(bt["flops_unpruned_pass"], 1), 4))),
                  flush=True)
        nv.set_tuning(**base)
    if "sweep" in what:
        Xs, ys, kws = synthetic.make_config("c3", N=26827)  # the notebook's dataset size (MELD_Quickstart)
        g3 = meld_b200.DeviceGraph.from_data(Xs, knn=7)
        op = meld_b200.MELD(verbose=0).fit(g3)
        rng = np.random.default_rng(1)
        label_sets = [rng.choice(["ctrl", "expt"], size=len(ys)) for _ in range(25)]
        betas = list(range(1, 200))
        g3.estimate_lmax()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        R, cols = op.transform_sweep(label_sets, betas=betas, as_tensor=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t1
        nt = len(label_sets) * len(betas)
        print(json.dumps(dict(sweep="25 label draws x 199 beta, 26827 cells (one knn of the notebook's search)",
                              transforms=nt, seconds=round(dt, 4), transforms_per_s=round(nt / dt, 1),
                              out_shape=list(R.shape))), flush=True)
        t1 = time.perf_counter()
        k = 0
        for lab in label_sets[:3]:
            for b in betas[:20]:
                o = meld_b200.MELD(verbose=0, beta=b).fit(g3)
                o.transform(lab)
                k += 1
        torch.cuda.synchronize()
        dt2 = time.perf_counter() - t1
        print(json.dumps(dict(loop="MELD(beta=b).fit(graph).transform(labels) one by one", transforms=k,
                              seconds=round(dt2, 4), transforms_per_s=round(k / dt2, 1))), flush=True)


if __name__ == "__main__":
    main()
