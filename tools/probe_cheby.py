"""Sweep launch configurations of the Chebyshev SpMM kernel on a synthetic CSR operator.

Not a bench line: a development probe (run under gpurun) that prints achieved algorithmic
GB/s per configuration.  Matrix: N rows, ~r entries/row, columns either uniformly random
("random": worst case for the gather) or within +-window of the row ("local").
"""

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meld_b200 import _native as nv  # noqa: E402
from meld_b200.graph import DeviceGraph  # noqa: E402


def make_csr(N, r, mode, window, seed=0, ragged=True):
    g = torch.Generator(device="cuda").manual_seed(seed)
    if ragged:
        lens = torch.randint(max(2, r // 3), r + (r - r // 3) + 1, (N,), device="cuda", generator=g)
    else:
        lens = torch.full((N,), r, device="cuda", dtype=torch.int64)
    indptr = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
    indptr[1:] = torch.cumsum(lens, 0)
    nnz = int(indptr[-1])
    rows = torch.repeat_interleave(torch.arange(N, device="cuda"), lens)
    if mode == "random":
        cols = torch.randint(0, N, (nnz,), device="cuda", generator=g)
    else:
        off = torch.randint(-window, window + 1, (nnz,), device="cuda", generator=g)
        cols = (rows + off).clamp_(0, N - 1)
    key = rows * N + cols
    key, _ = torch.sort(key)
    cols = (key % N).to(torch.int32)
    vals = torch.rand(nnz, device="cuda", dtype=torch.float64, generator=g) * 1e-2
    return indptr, cols, vals, nnz


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=500_000)
    ap.add_argument("--r", type=int, default=50)
    ap.add_argument("--p", type=int, default=4)
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--mode", default="local")
    ap.add_argument("--window", type=int, default=4000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--configs", default="")
    args = ap.parse_args()
    lib = nv.lib()
    indptr, cols, vals, nnz = make_csr(args.n, args.r, args.mode, args.window)
    N, p, m = args.n, args.p, args.m
    S = torch.rand(N, p, device="cuda", dtype=torch.float64)
    R = torch.empty_like(S)
    coeffs = (np.random.default_rng(0).normal(size=m + 1) * 0.01).astype(np.float64)
    cptr = coeffs.ctypes.data_as(C.POINTER(C.c_double))
    bytes_step = nnz * 12 + (N + 1) * 4 + 5 * N * p * 8
    print(json.dumps(dict(N=N, nnz=nnz, p=p, m=m, mode=args.mode, MB_step=bytes_step / 1e6)))
    configs = [
        dict(),
        dict(threads=256, ctas_per_sm=2),
        dict(blk_chunk=768, stage_cap=1024, n_stage=4, ctas_per_sm=4),
        dict(blk_chunk=3072, stage_cap=4096, n_stage=2, threads=256, ctas_per_sm=2),
        dict(blk_chunk=3072, stage_cap=4096, n_stage=3, threads=256, ctas_per_sm=1),
        dict(group=16),
        dict(group=4),
        dict(group=16, threads=256, ctas_per_sm=2),
        dict(n_stage=4, ctas_per_sm=2),
        dict(n_stage=2, ctas_per_sm=4),
        dict(ctas_per_sm=4),
        dict(ctas_per_sm=6, n_stage=2, blk_chunk=768, stage_cap=1024),
        dict(threads=64, ctas_per_sm=8, n_stage=2, blk_chunk=768, stage_cap=1024),
    ]
    if args.configs:
        configs = json.loads(args.configs)
    base = dict(blk_chunk=1024, stage_cap=1280, dict_cap=640, row_cap=128, n_stage=0, threads=512, gather_warps=3,
                ctas_per_sm=1, group=0, use_dict=1)
    for cfg in configs:
        full = dict(base)
        full.update(cfg)
        nv.set_tuning(**full)
        out = C.c_void_p()
        nv.check(lib.meld_b200_graph_from_csr(N, N, 0, nnz, nv.ptr(indptr), nv.ptr(cols), nv.ptr(vals),
                                              nv.current_stream_ptr(), C.byref(out)), "from_csr")
        g = DeviceGraph(out.value, device=S.device)
        best = 1e9
        for rep in range(args.reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nv.check(lib.meld_b200_cheby_filter(g._h, 1.0, cptr, m + 1, nv.ptr(S), p, nv.ptr(R),
                                                nv.current_stream_ptr()), "cheby_filter")
            e1.record()
            torch.cuda.synchronize()
            if rep > 0:
                best = min(best, e0.elapsed_time(e1))
        gbs = m * bytes_step / (best * 1e-3) / 1e9
        print(json.dumps(dict(cfg=cfg, ms=round(best, 3), us_step=round(best * 1e3 / m, 1), GBs=round(gbs, 1),
                              frac=round(gbs / 6538.6, 3))), flush=True)
        g.close()


if __name__ == "__main__":
    main()
