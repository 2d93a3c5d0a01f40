"""Multi-GPU check + timing of the strong-scaling path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dist_check.py [--config c4] [--cells N]

Every rank: sharded build of ONE dataset -> full graph; row-partitioned filter with peer stores ("p2p") and with the
NCCL all-gather ("nccl"); both must equal the single-GPU filter of the same graph (1e-12).  Prints per-mode times
(CUDA events, max over ranks)."""

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meld_b200  # noqa: E402
from meld_b200 import synthetic  # noqa: E402
from meld_b200.distributed import ShardedFilter  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c4")
    ap.add_argument("--cells", type=int, default=None)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    X, labels, kw = synthetic.make_config(args.config, N=args.cells)
    n = X.shape[0]
    p = len(np.unique(labels))
    m = kw.get("chebyshev_order", 50)
    Xd = torch.from_numpy(X).cuda()
    codes = torch.from_numpy(np.unique(labels, return_inverse=True)[1].astype(np.int32)).cuda()

    def maxr(v):
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    def say(**kv):
        if rank == 0:
            print(json.dumps(kv), flush=True)

    # reference: this rank alone
    ref_op = meld_b200.MELD(verbose=0, **kw)
    ref_op.fit(Xd)
    ref = ref_op.transform_device(codes, p)
    Lref = ref_op.graph.nnz
    torch.cuda.synchronize()
    # sharded build
    os.environ["MELD_B200_TIMING_PY"] = "1"
    g = meld_b200.DeviceGraph.from_data_sharded(Xd, knn=kw.get("knn", 5))
    say(check="sharded build nnz", nnz=g.nnz, single=Lref, equal=bool(g.nnz == Lref))
    lmax = g.estimate_lmax()
    h = meld_b200.filter.filter_kernel("heat", kw.get("beta", 60))
    coeffs = np.ascontiguousarray(meld_b200.filter.cheby_coefficients(h, lmax, m))
    S = torch.empty((n, p), dtype=torch.float64, device="cuda")
    from meld_b200 import _native as nv

    nv.check(nv.lib().meld_b200_indicator_matrix(nv.ptr(codes), n, p, 1, nv.ptr(S), nv.current_stream_ptr()), "ind")
    full = meld_b200.filter.cheby_apply(g, lmax, coeffs, S)
    say(check="sharded-build graph filter vs single-GPU fit_transform",
        rel=float((full - ref).abs().max() / ref.abs().max()))
    for mode in ("p2p", "nccl"):
        sf = ShardedFilter(g, mode=mode)
        out = sf.apply(lmax, coeffs, S)
        torch.cuda.synchronize()
        err = float((out - full).abs().max() / full.abs().max())
        bad = sf.ctx.error() if mode == "p2p" else 0
        dist.barrier()
        best = 1e9
        for _ in range(args.reps):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = sf.apply(lmax, coeffs, S)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, maxr(e0.elapsed_time(e1)))
        say(mode=mode, world=world, rel_err_vs_full=err, flag_timeout=bad, filter_ms=round(best, 3),
            us_per_term=round(1e3 * best / m, 2), halo_fraction_rank0=round(sf.halo_fraction, 4))
        sf.close()
    # fully row-partitioned build: gather the slices and compare with the single-GPU Laplacian
    gs, bounds = meld_b200.DeviceGraph.from_data_sharded_rows(Xd, knn=kw.get("knn", 5))
    sfr = ShardedFilter(None, mode="p2p", row_slice=gs, bounds=bounds)
    Lg = sfr.gather_scipy_L()
    if rank == 0:
        Ls = ref_op.graph.to_scipy_L()
        same = Lg.nnz == Ls.nnz and np.array_equal(Lg.indptr, Ls.indptr) and np.array_equal(Lg.indices, Ls.indices)
        say(check="row-partitioned build vs single-GPU L", pattern_equal=bool(same),
            max_abs_diff=float(np.abs(Lg.data - Ls.data).max()) if same else None, halo_fraction_rank0=round(sfr.halo_fraction, 4))
    lm = sfr.estimate_lmax()
    outr = sfr.apply(lm, coeffs, S)
    torch.cuda.synchronize()
    say(check="row-partitioned build + lanczos + filter vs single GPU", lmax_rel=abs(lm - lmax) / lmax,
        rel=float((outr - full).abs().max() / full.abs().max()), flag_timeout=sfr.ctx.error())
    # single-GPU filter time for the ratio
    best = 1e9
    for _ in range(args.reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        meld_b200.filter.cheby_apply(g, lmax, coeffs, S)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    say(mode="single GPU (replicated)", filter_ms=round(best, 3), us_per_term=round(1e3 * best / m, 2))
    # whole fit_transform, distributed vs single
    for mode in ("p2p rows", "p2p", "nccl", "replicated"):
        def step():
            dm, db = (mode.split()[0], "rows") if mode.endswith("rows") else (mode, "replicated")
            op = meld_b200.MELD(verbose=0, distributed=True, dist_mode=dm, dist_build=db, **kw)
            op.fit(Xd)
            return op.transform_device(codes, p)
        for _ in range(2):
            out = step()
        torch.cuda.synchronize()
        err = float((out - ref).abs().max() / ref.abs().max())
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        say(fit_transform=mode, world=world, ms_per_step=round(maxr(e0.elapsed_time(e1)) / args.reps, 3), rel_err=err)
    def step1():
        op = meld_b200.MELD(verbose=0, **kw)
        op.fit(Xd)
        return op.transform_device(codes, p)
    step1()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        step1()
    e1.record()
    torch.cuda.synchronize()
    say(fit_transform="single GPU", ms_per_step=round(e0.elapsed_time(e1) / args.reps, 3))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
