#!/usr/bin/env python
"""Headline benchmark: cells/s of MELD.fit_transform on the B200 engine + SpMV roofline.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference ...                     # the reference's CPU path (oracle port), same config

A "step" is one full pass of the hot path (kNN alpha-decay graph build -> Laplacian -> lmax ->
Chebyshev filter of the sample indicators) over one batch of synthetic cells.  Workload at N=1:
BASELINE.json configs[3] -- 500k cells x 100 PCA dims, knn=15, Chebyshev order 64, 4 samples --
the configuration the metric's roofline target is quoted on; it fits one GPU.  Prints ONE JSON line
(rank 0).  See DESIGN.md "Measurement" for every field.
"""

from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cells/sec MELD.fit_transform"
UNIT = "cells/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c4", help="synthetic config of meld_b200.synthetic (c1..c5)")
    ap.add_argument("--cells", type=int, default=None, help="override the number of cells")
    ap.add_argument("--n-pca", default="default", help="'none' disables the PCA of config 2 (distance GEMM over 2000 dims)")
    ap.add_argument("--cpu-cells", type=int, default=10000,
                    help="cells of the sample the CPU path's per-row stages (affinities .. filter) are timed on")
    ap.add_argument("--cpu-queries", type=int, default=1024,
                    help="rows queried against the FULL-size CPU ball tree (kNN time is extrapolated by N / queries)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (CPU filter on the exported graph)")
    ap.add_argument("--multi", default="strong", choices=["strong", "replicas"],
                    help="N > 1: 'strong' = ONE dataset: candidate search sharded by query rows, NCCL all-gather of the "
                         "candidate lists, Chebyshev recurrence row-partitioned with one exchange per term; 'replicas' "
                         "= every rank runs its own dataset (weak scaling, no data-path collective)")
    ap.add_argument("--dist-mode", default="p2p", choices=["p2p", "nccl", "replicated"],
                    help="strong scaling: how the filter exchanges T_k per term (peer stores over NVLink from inside "
                         "the SpMM kernel / NCCL all-gather / no exchange, filter replicated)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled every 200 ms while the timed region runs.

    In-process NVML (one already-open device handle, two cheap queries per sample).  Spawning ``nvidia-smi -lms``
    beside the timed region -- the first version -- initialises NVML for every GPU of the box and takes driver
    locks that stalled this process's own cudaMallocAsync / synchronise calls by 30-700 ms at random; a step is
    ~55 ms, so that distorted the measurement it was meant to vouch for.  ``MELD_BENCH_CLOCKS=smi`` restores the
    subprocess sampler, ``MELD_BENCH_NO_CLOCKS=1`` disables sampling."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.samples = []
        self.nvml = None
        self.handle = None
        if os.environ.get("MELD_BENCH_NO_CLOCKS") or os.environ.get("MELD_BENCH_CLOCKS") == "smi":
            return
        try:  # open NVML and the handle BEFORE the timed region
            import pynvml

            pynvml.nvmlInit()
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                idx = int(vis.split(",")[gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _loop(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.samples.append((sm, reasons))
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def start(self):
        if os.environ.get("MELD_BENCH_NO_CLOCKS"):
            return
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._loop, name="clock-sampler", daemon=True)
            self.thread.start()
            return
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            nv = self.nvml
            names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                     ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                     ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                     ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap"))
            fallback = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
            reasons = set()
            for sm, bits in self.samples:
                for name, attr in names:
                    mask = int(getattr(nv, attr, fallback[name]))
                    if bits & mask:
                        reasons.add(name)
            if self.samples:
                out["sm_mhz"] = float(np.median([x[0] for x in self.samples]))
                out["sm_max_mhz"] = self.max_sm
                out["samples"] = len(self.samples)
            out["reasons"] = sorted(reasons)
            out["sampler"] = "nvml"
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    smax.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(max(smax))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        out["sampler"] = "nvidia-smi"
        return out


def make_inputs(args, seed_offset=0):
    from meld_b200 import synthetic

    cfg = synthetic.CONFIGS[args.config]
    n = args.cells or cfg["N"]
    base_seed = int("".join(ch for ch in args.config if ch.isdigit()))
    X, labels, kw = synthetic.make_config(args.config, seed=base_seed + seed_offset, N=n)
    if args.n_pca == "none":
        kw = dict(kw, n_pca=None)
    return X, labels, kw, cfg


def workload_name(args, cfg, n):
    return "{}: {} cells x {} dims, {} samples, {}{}".format(
        args.config, n, cfg["D"], cfg["n_samples"], ", ".join("{}={}".format(k, v) for k, v in cfg["meld"].items()) or "defaults",
        ", n_pca=None" if args.n_pca == "none" else "")


# --------------------------------------------------------------------------------------------------
# The reference's CPU path at the FULL size of the workload.  A 500k-cell ball-tree
# search is hours of CPU, so a step measures a bounded sample and extrapolates, stage by stage:
#   kNN search     `queries` random rows against the ball tree of ALL n cells (the tree is built once, outside the
#                  steps; cost per query row is what it is at full size), x n / queries
#   affinities, symmetrise + anisotropy + Laplacian, lmax, Chebyshev filter
#                  measured on a `sample`-cell graph of the same generator, x n / sample (all are per-row / per-
#                  nonzero costs; a kNN graph keeps its nonzeros per row when the data are subsampled)
# value = n / (sum of the extrapolated stage times).  Every figure is labelled "extrapolated".
class CpuPath:
    def __init__(self, args, n_jobs):
        from sklearn.neighbors import NearestNeighbors

        from meld_b200 import synthetic

        self.args, self.n_jobs = args, n_jobs
        self.cfg = synthetic.CONFIGS[args.config]
        self.n = args.cells or self.cfg["N"]
        self.X, self.labels, kw = synthetic.make_config(args.config, N=self.n)
        self.kw = kw
        self.graph_kw = {k: kw[k] for k in ("knn", "decay", "thresh", "anisotropy") if k in kw}
        self.filter_kw = {k: v for k, v in kw.items() if k not in self.graph_kw and k != "n_pca"}
        self.knn = self.graph_kw.get("knn", 5)
        self.n_pca = None if (self.cfg["D"] <= 100 or args.n_pca == "none") else 100
        self.sample = min(args.cpu_cells, self.n)
        self.queries = min(args.cpu_queries, self.n)
        t0 = time.perf_counter()
        data_nu = self.X
        self.pca_s = 0.0
        if self.n_pca is not None:  # PCA is part of the path (config 2): measured at full size, once
            from oracle import graph as ograph

            data_nu = ograph.reduce_data(self.X, self.n_pca, random_state=0)
            self.pca_s = time.perf_counter() - t0
            t0 = time.perf_counter()
        self.data_nu = data_nu
        self.tree = NearestNeighbors(n_neighbors=self.knn + 1, algorithm="ball_tree", n_jobs=n_jobs).fit(data_nu)
        self.tree_s = time.perf_counter() - t0
        self.rng = np.random.default_rng(0)

    def step(self):
        from oracle import graph as ograph  # test infrastructure; allowed here as the timed CPU baseline only
        from oracle import meld as omeld

        n, s, q = self.n, self.sample, self.queries
        rows = self.rng.choice(n, size=q, replace=False)
        t0 = time.perf_counter()
        self.tree.kneighbors(self.data_nu[rows], n_neighbors=min(6 * (self.knn + 1), n))
        knn_q = time.perf_counter() - t0
        sub = np.sort(self.rng.choice(n, size=s, replace=False)) if s < n else np.arange(n)
        tm = {}
        t1 = time.perf_counter()
        g = ograph.build_graph(self.data_nu[sub], n_pca=None, n_jobs=self.n_jobs, timings=tm, **self.graph_kw)
        t2 = time.perf_counter()
        lmax = ograph.estimate_lmax(g["L"], g["dw"])
        t3 = time.perf_counter()
        omeld.transform(g["L"], lmax, self.labels[sub], **self.filter_kw)
        t4 = time.perf_counter()
        scale = n / s
        stages = {
            "pca_s": self.pca_s,
            "tree_build_s": self.tree_s,
            "knn_search_s": knn_q * n / q,
            "graph_assembly_s": ((t2 - t1) - tm.get("knn_first_search_s", 0.0)) * scale,
            "lmax_s": (t3 - t2) * scale,
            "filter_s": (t4 - t3) * scale,
        }
        total = sum(stages.values())
        return n / total, stages, dict(knn_query_s=knn_q, sample_total_s=t4 - t1, nnz_per_row=g["L"].nnz / s)

    def describe(self):
        return ("oracle port of the reference CPU path, EXTRAPOLATED to the full {n} cells stage by stage: sklearn "
                "ball tree over all {n} cells (n_jobs={j}), {q} query rows timed and scaled by n/{q}; affinities, "
                "symmetrise/anisotropy/Laplacian, ARPACK lmax and the scipy Chebyshev filter (1 thread) timed on a "
                "{s}-cell sample and scaled by n/{s}").format(n=self.n, j=self.n_jobs, q=self.queries, s=self.sample)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    path = CpuPath(args, cores)
    for _ in range(max(0, min(args.warmup, 1))):
        path.step()
    vals, t_all = [], time.perf_counter()
    for _ in range(args.steps):
        vals.append(path.step())
    wall = time.perf_counter() - t_all
    value = float(np.median([v[0] for v in vals]))
    stages = {k: float(np.median([v[1][k] for v in vals])) for k in vals[-1][1]}
    # the reference's default is n_jobs=1 (SURVEY finding 5): one extra sample with a single-threaded search
    one = CpuPath.__new__(CpuPath)
    one.__dict__.update(path.__dict__)
    one.n_jobs = 1
    one.tree = one.tree.set_params(n_jobs=1)
    v1, st1, _ = one.step()
    path.tree.set_params(n_jobs=cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * path.n / value, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 and args.multi == "strong" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, path.cfg, path.n), "extrapolated": True},
        "extrapolated": True,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": path.describe(),
                         "extrapolated": True, "stages_s_full_size": {k: round(v, 3) for k, v in stages.items()},
                         "n_jobs_1": {"value": v1, "stages_s_full_size": {k: round(v, 3) for k, v in st1.items()}},
                         "sample_wall_s_per_step": wall / max(args.steps, 1)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
def parity_block(args, op, out_dev, labels, kw, cfg, L):
    """Does the bench's own output agree with the reference's CPU arithmetic?  (i) FULL size: the oracle's Chebyshev
    filter (scipy matvecs) on the GPU-built graph copied to the host vs the densities of the last timed step
    (BASELINE.md 3.5); (ii) a 20k-cell sample of the same generator through the whole oracle vs the whole engine
    (graph pattern + densities, oracle's lmax injected)."""
    import meld_b200
    from meld_b200 import synthetic
    from oracle import meld as omeld  # the checker, outside every timed region

    res = {}
    filter_kw = {k: v for k, v in kw.items() if k in ("beta", "offset", "order", "filter", "chebyshev_order")}
    t0 = t1 = time.perf_counter()  # L was exported (or gathered from the ranks' row slices) by the caller
    ref = omeld.transform(L, op.graph.lmax, labels, **filter_kw)
    t2 = time.perf_counter()
    got = out_dev.cpu().numpy()
    colmax = np.abs(ref.values).max(axis=0)
    err = np.abs(got - ref.values)
    res["full_size_filter"] = {
        "normwise": float((err.max(axis=0) / colmax).max()),
        "elementwise_ok": bool(np.all(err <= 1e-5 * np.abs(ref.values) + 1e-9 * colmax)),
        "L_symmetric": bool(abs(L - L.T).max() == 0.0),
        "L_row_sums_max": float(np.abs(L @ np.ones(L.shape[0])).max()),
        "cpu_filter_s": t2 - t1, "export_s": t1 - t0,
    }
    ns = min(20000, cfg["N"] if args.cells is None else args.cells)
    Xs, ys, kws = synthetic.make_config(args.config, N=ns)
    n_pca = None if (cfg["D"] <= 100 or args.n_pca == "none") else 100
    ref_s, g, lmax = omeld.fit_transform(Xs, ys, n_pca=n_pca, random_state=0, n_jobs=os.cpu_count(), **kws)
    ops = meld_b200.MELD(verbose=0, n_pca=n_pca, random_state=0, **kws)
    ops.fit(Xs)
    Ls = ops.graph.to_scipy_L()
    same = (Ls.nnz == g["L"].nnz and np.array_equal(Ls.indptr, g["L"].indptr)
            and np.array_equal(Ls.indices, g["L"].indices))
    ops.graph.lmax = lmax
    ds = ops.transform(ys)
    cm = np.abs(ref_s.values).max(axis=0)
    res["sample_20k_end_to_end"] = {
        "cells": ns, "pattern_equal": bool(same),
        "max_abs_dL_over_maxL": float(np.abs(Ls.data - g["L"].data).max() / np.abs(g["L"].data).max()) if same else None,
        "normwise": float((np.abs(ds.values - ref_s.values).max(axis=0) / cm).max()),
    }
    res["tolerance"] = "north_star 1e-5 relative (normwise per column; elementwise floored at 1e-9 colmax)"
    return res, res["full_size_filter"]["cpu_filter_s"]


def run_b200(args):
    import torch
    import torch.distributed as dist

    import meld_b200
    from meld_b200 import _native as nv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = nv.lib()

    # N > 1, default "strong": ONE dataset.  Stage 1 of the build (candidate search, exact distances, eps) is sharded
    # by query rows, the candidate lists are all-gathered (NCCL), the Laplacian assembly and lmax are replicated, and
    # the Chebyshev recurrence is row-partitioned with one exchange of the T_k slices per term.
    # "replicas": every rank runs the whole path on ITS OWN dataset (weak scaling, no data-path collective).
    strong = world > 1 and args.multi == "strong"
    dist_mode = args.dist_mode
    Xh, labels, kw, cfg = make_inputs(args, seed_offset=0 if (strong or world == 1) else 100 * rank)
    n, d = Xh.shape
    p = cfg["n_samples"]
    m = kw.get("chebyshev_order", 50)
    samples, codes_h = np.unique(labels, return_inverse=True)
    X_dev = torch.from_numpy(Xh).cuda()
    codes_dev = torch.from_numpy(codes_h.astype(np.int32)).cuda()
    X_pin = torch.from_numpy(Xh).pin_memory()
    X_pin_np = X_pin.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make_op():
        if strong:
            return meld_b200.MELD(verbose=0, distributed=True, dist_mode=dist_mode, **kw)
        return meld_b200.MELD(verbose=0, **kw)

    dist_note = None
    if strong and dist_mode == "p2p":
        # self-check of the peer-store path before anything is timed: it must reproduce the replicated filter.  A box
        # without working CUDA IPC / P2P falls back to the NCCL exchange -- still row-partitioned, and reported.
        ok = torch.ones(1, device="cuda")
        try:
            opc = make_op()
            opc.fit(X_dev)
            a = opc.transform_device(codes_dev, p)
            opr = meld_b200.MELD(verbose=0, distributed=True, dist_mode="replicated", **kw)  # whole graph + filter per rank
            opr.fit(X_dev)
            b = opr.transform_device(codes_dev, p)
            torch.cuda.synchronize()
            if opc._sharded.ctx.error() != 0 or float((a - b).abs().max()) > 1e-9 * float(b.abs().max()):
                ok.zero_()
            del opc, opr, a, b
        except Exception as exc:  # noqa: BLE001
            dist_note = "p2p unavailable: {}".format(str(exc)[:200])
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:
            dist_note = dist_note or "p2p self-check failed"
            dist_mode = "nccl"

    fit_marks = []

    def step_device(events=None):
        op = make_op()
        op.profile_events = events
        m0 = torch.cuda.Event(enable_timing=True)
        m1 = torch.cuda.Event(enable_timing=True)
        m0.record()
        op.fit(X_dev)
        m1.record()
        if events is not None:
            fit_marks.append((m0, m1))
        out = op.transform_device(codes_dev, p)
        return op, out

    def step_e2e(Xnp):
        op = make_op()
        dens = op.fit_transform(Xnp, labels)  # host buffers in, DataFrame (host) out
        return op, dens

    # ---- device-resident arm ("value") + live SpMV timing.  The last warm-up step runs in the state the timed steps
    # run in (garbage collected and the collector paused, clock sampler started): collecting the earlier steps'
    # estimators frees their graphs, and the first step after that was consistently ~6 ms slower.
    sampler = ClockSampler(local_rank)
    for w in range(args.warmup):
        if w == args.warmup - 1:
            gc.collect()
            gc.disable()
            sampler.start()
        op, out = step_device()
    if args.warmup == 0:
        gc.collect()
        gc.disable()
        sampler.start()
    barrier()
    events = []
    launches0 = lib.meld_b200_launch_count()
    syncs0 = lib.meld_b200_sync_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    e0.record()
    marks[0].record()
    for i in range(args.steps):
        op, out = step_device(events)
        marks[i + 1].record()  # per-step times for the record (no extra synchronisation)
    e1.record()
    barrier()
    gc.enable()
    launches = lib.meld_b200_launch_count() - launches0
    syncs = lib.meld_b200_sync_count() - syncs0
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    step_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    nnz = op.graph.nnz
    rows_only = strong and op._sharded is not None and op.graph.n_rows != op.graph.n_cols  # this rank holds a row slice
    if rows_only:
        tn = torch.tensor([nnz], dtype=torch.int64, device="cuda")
        dist.all_reduce(tn)
        nnz = int(tn)
    lmax = op.graph.lmax
    lmax_iters = op.graph.lmax_iters
    stats = op.graph.build_stats()
    bt = op.graph.build_times()
    filt_ms = [a.elapsed_time(b) for a, b, _ in events]
    launch_us = 1e3 * float(np.mean(filt_ms)) / m  # average duration of one cheby_step launch
    # algorithmic bytes of ONE launch on ONE GPU: the matrix rows this rank owns + its slices of the vectors
    share = world if (strong and dist_mode != "replicated") else 1
    bytes_step = (nnz * 12 + (n + 1) * 4 + 5 * n * p * 8) / share
    achieved = bytes_step / (launch_us * 1e-6) / 1e9

    # ---- end-to-end arm: host (pinned) inputs, DataFrame back on the host, copies inside the timed region
    for w in range(args.warmup):
        if w == args.warmup - 1:
            gc.collect()
            gc.disable()  # a generation-2 collection in the middle of a 60 ms step is measurement noise, not the engine
        step_e2e(X_pin_np)
    if args.warmup == 0:
        gc.collect()
        gc.disable()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = []
    t_e2e = time.perf_counter()
    f0.record()
    e2e_host = []
    for _ in range(args.steps):
        t_s = time.perf_counter()
        op_e2e, dens = step_e2e(X_pin_np)
        e2e_steps.append(1e3 * (time.perf_counter() - t_s))  # the DataFrame is on the host when the call returns
        e2e_host.append({k: round(1e3 * v, 2) for k, v in op_e2e.timings_.items()})
    f1.record()
    barrier()
    e2e_ms = max(f0.elapsed_time(f1), 1e3 * (time.perf_counter() - t_e2e))
    gc.enable()
    # second e2e figure: the drop-in case, a plain (pageable) numpy array as the user would pass it
    pageable_steps = []
    if world == 1:
        step_e2e(Xh)
        torch.cuda.synchronize()
        gc.disable()
        for _ in range(max(2, min(args.steps, 5))):
            t_s = time.perf_counter()
            step_e2e(Xh)
            pageable_steps.append(1e3 * (time.perf_counter() - t_s))
        gc.enable()

    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device="cuda")
    mine = torch.tensor([min(step_ms), float(np.median(step_ms)), max(step_ms), float(np.median(filt_ms))],
                        dtype=torch.float64, device="cuda")
    per_rank = mine[None, :]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_rank = torch.empty((world, 4), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(per_rank, mine)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    per_rank = per_rank.cpu().numpy()
    jobs = 1 if (strong or world == 1) else world  # replicas: every rank pushed n cells through per step
    value = jobs * n * args.steps / (ms_total * 1e-3)
    e2e_value = jobs * n * args.steps / (e2e_ms * 1e-3)

    # ---- N > 1, strong: the replicas figure as a second, labelled measurement (same steps, own datasets)
    replicas = None
    if strong:
        Xr, labr, kwr, _ = make_inputs(args, seed_offset=100 * rank)
        Xr_dev = torch.from_numpy(Xr).cuda()
        cr = torch.from_numpy(np.unique(labr, return_inverse=True)[1].astype(np.int32)).cuda()

        def step_rep():
            o = meld_b200.MELD(verbose=0, **kwr)
            o.fit(Xr_dev)
            return o.transform_device(cr, p)

        for _ in range(args.warmup):
            step_rep()
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(args.steps):
            step_rep()
        r1.record()
        barrier()
        tr = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        replicas = {"value": world * n * args.steps / (float(tr[0]) * 1e-3), "unit": UNIT, "scaling": "weak",
                    "ms_per_step": float(tr[0]) / args.steps,
                    "what": "{} independent replicas (one {}-cell dataset per GPU), no data-path collective".format(world, n)}
        del Xr_dev

    L_full = None
    if not args.no_parity:  # collective when no rank holds the whole graph
        L_full = op._sharded.gather_scipy_L() if rows_only else (op.graph.to_scipy_L() if rank == 0 else None)
    if rank == 0:
        pk, pk_kind = peaks()
        if world == 1:
            par = "1 GPU"
        elif strong:
            build = ("every rank assembles ONLY its rows of L (NCCL all-gathers of eps and of the kernel row sums, 8 N "
                     "bytes each, + all-to-all-v of the mirrored entries)" if rows_only else
                     "NCCL all-gather of the candidate lists, Laplacian assembly replicated")
            par = ("ONE dataset on {w} GPUs, rows of the internal cell order partitioned x{w}: candidate search / exact "
                   "distances / eps by query rows; {build}; Lanczos and Chebyshev recurrence row-partitioned with one "
                   "exchange of the vector slices per term ({how})").format(
                w=world, build=build,
                how={"p2p": "P2P stores over NVLink from inside the SpMM kernel, halo rows only, flag words in peer memory",
                     "nccl": "NCCL all-gather per term", "replicated": "none: filter replicated"}[dist_mode])
        else:
            par = ("{} independent replicas (one {}-cell dataset per GPU, different seeds), no data-path collective; "
                   "value = cells of all ranks / max-over-ranks time").format(world, n)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "step_ms": {"min": min(step_ms), "median": float(np.median(step_ms)), "max": max(step_ms),
                        "all": [round(v, 2) for v in step_ms],
                        "fit_median": float(np.median([a.elapsed_time(b) for a, b in fit_marks])) if fit_marks else None,
                        "filter_median": float(np.median(filt_ms)) if filt_ms else None,
                        "per_rank_min_median_max_filter": [[round(float(v), 3) for v in row] for row in per_rank]},
            "config": {
                "workload": workload_name(args, cfg, n),
                "parallelism": par, "dist_mode": dist_mode if strong else None, "dist_note": dist_note,
                "halo_fraction_rank0": (getattr(op._sharded, "halo_fraction", None) if strong and op._sharded else None),
                "nnz_L": int(nnz), "nnz_per_row": nnz / n, "lmax": lmax, "lmax_iters": lmax_iters,
                "candidate_cap": stats["candidate_cap"],
                "max_candidates": stats["max_candidates"], "search_passes": stats["search_passes"],
                "row_blocks": stats["row_blocks"],
                "l2_note": "inputs (X {} MB, L {} MB) exceed the 126 MB L2; no explicit flush".format(
                    Xh.nbytes // 2**20, nnz * 12 // 2**20),
            },
            # strong scaling: every rank uploads 1 / world of the matrix (all-gathered over NVLink) + the label codes, and
            # every rank reads back the whole density matrix; replicas: one matrix in, one density matrix out per rank
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(Xh.nbytes + world * 4 * n) if strong else int(jobs * (Xh.nbytes + 4 * n)),
                    "d2h_bytes_per_step": int((world if strong else jobs) * 8 * n * p), "ms_per_step": e2e_ms / args.steps,
                    "step_ms": {"min": min(e2e_steps), "median": float(np.median(e2e_steps)), "max": max(e2e_steps),
                                "all": [round(v, 2) for v in e2e_steps]},
                    "host_timings_ms_slowest_step": e2e_host[int(np.argmax(e2e_steps))],
                    "input": "pinned host numpy array",
                    "pageable_input_step_ms": ({"min": min(pageable_steps), "median": float(np.median(pageable_steps)),
                                                "max": max(pageable_steps)} if pageable_steps else None),
                    "host_timings_ms_last_step": {k: round(1e3 * v, 2) for k, v in op_e2e.timings_.items()}},
            "gpu_launches": int(launches),
            "gpu_launches_per_step": int(launches) / max(args.steps, 1),
            "host_syncs_per_step": int(syncs) / max(args.steps, 1),
            "clocks": clocks,
            "roofline": {
                "kernel": "Chebyshev SpMM + fused three-term update, one launch per term (cheby_flat kernels; "
                          "meld_b200_cheby_filter{})".format("_dist, rows of this rank" if share > 1 else ""),
                "bound": "hbm",
                "achieved": achieved, "peak": pk["hbm_gbs"], "peak_kind": pk_kind, "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"],
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this
                # kernel on this workload (profiles/); other workloads: not captured
                "traffic": NCU_TRAFFIC.get((args.config, n, world)),
                "bytes_per_launch": int(bytes_step),
                "us_per_launch": launch_us, "launches_per_step": m,
                "filter_share_of_step": float(np.mean(filt_ms)) / (ms_total / args.steps),
            },
            # second hot kernel: the tcgen05 distance GEMM of the candidate search (two passes per build),
            # timed with CUDA events inside the library during the last timed step
            "roofline_gemm": None if not bt["pass2_ms"] else {
                "kernel": "tc_search_kernel (bf16-split distance GEMM, fused top-k / emit epilogue)", "bound": "tensor",
                "achieved": bt["flops_per_pass"] / (bt["pass2_ms"] * 1e-3) / 1e12,
                "achieved_pass1": bt["flops_pass1"] / (bt["pass1_ms"] * 1e-3) / 1e12,
                "peak": pk.get("bf16_tflops"), "peak_sustained": pk.get("bf16_tflops_sustained"), "unit": "TFLOP/s",
                "frac": bt["flops_per_pass"] / (bt["pass2_ms"] * 1e-3) / 1e12 / pk.get("bf16_tflops", 1632.2),
                "ms_pass1": bt["pass1_ms"], "ms_pass2": bt["pass2_ms"], "flops_per_pass": bt["flops_per_pass"],
                "flops_pass1": bt["flops_pass1"], "flops_unpruned_pass": bt["flops_unpruned_pass"],
                "tile_pairs_kept": bt["flops_per_pass"] / max(bt["flops_unpruned_pass"], 1.0),
                "note": "flops = 2 x 128 x 256 x K' per (row tile, column tile) product actually issued; pairs whose "
                        "bounding balls are farther apart than the emit radius are skipped (ball-tree style pruning)",
                "share_of_step": (bt["pass1_ms"] + bt["pass2_ms"]) / (ms_total / args.steps),
            },
        }
        if replicas is not None:
            line["replicas"] = replicas
        cpu_filter_s = None
        if not args.no_parity:
            line["parity"], cpu_filter_s = parity_block(args, op, out, labels, kw, cfg, L_full)
        if not args.no_cpu_baseline and world == 1:
            cores = os.cpu_count() or 1
            path = CpuPath(args, cores)
            v, stages, extra = path.step()
            if cpu_filter_s is not None:  # the filter was MEASURED at full size on the exported graph: use that
                stages = dict(stages, filter_s=cpu_filter_s)
                v = n / sum(stages.values())
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": True,
                "sample": path.describe() + ("; the filter is MEASURED at full size on the GPU-built graph copied to "
                                             "the host" if cpu_filter_s is not None else ""),
                "stages_s_full_size": {k: round(s, 3) for k, s in stages.items()},
                "sample_detail": {k: round(float(s), 4) for k, s in extra.items()},
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# dram bytes per launch of the Chebyshev kernel from ncu --set full captures (profiles/), keyed by (config, cells, gpus)
NCU_TRAFFIC = {("c4", 500000, 1): 443832576}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
