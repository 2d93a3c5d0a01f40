#!/usr/bin/env python
"""Headline benchmark: cells/s of MELD.fit_transform on the B200 engine + SpMV roofline.

    python bench.py --gpus N --steps K --warmup W            # this engine
    python bench.py --impl reference ...                     # the reference's CPU path (oracle port)

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
    ap.add_argument("--cpu-cells", type=int, default=20000, help="cells in the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--multi", default="replicas", choices=["replicas", "sharded"],
                    help="N > 1: 'replicas' = every rank runs its own dataset of the config (weak scaling, no data-path "
                         "collective); 'sharded' = ONE dataset, candidate search sharded by query rows + NCCL all-gather "
                         "(strong scaling)")
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
    X, labels, kw = synthetic.make_config(args.config, seed=int(args.config[1:]) + seed_offset, N=n)
    return X, labels, kw, cfg


def workload_name(args, cfg, n):
    return "{}: {} cells x {} dims, {} samples, {}".format(
        args.config, n, cfg["D"], cfg["n_samples"], ", ".join("{}={}".format(k, v) for k, v in cfg["meld"].items()) or "defaults")


# --------------------------------------------------------------------------------------------------
def cpu_reference_leg(args, n_cells, n_jobs):
    """The reference's CPU path (oracle port: same sklearn / scipy calls) on a bounded sample.
    Returns (cells/s, seconds, nnz(L), per-stage seconds) -- the stages are oracle.meld.fit_transform unrolled."""
    from oracle import graph as ograph  # test infrastructure; allowed here as the timed CPU baseline only
    from oracle import meld as omeld
    from meld_b200 import synthetic

    cfg = synthetic.CONFIGS[args.config]
    X, labels, kw = synthetic.make_config(args.config, N=n_cells)
    graph_kw = {k: kw[k] for k in ("knn", "decay", "thresh", "anisotropy") if k in kw}
    filter_kw = {k: v for k, v in kw.items() if k not in graph_kw}
    t0 = time.perf_counter()
    g = ograph.build_graph(X, n_pca=None if cfg["D"] <= 100 else 100, random_state=0, n_jobs=n_jobs, **graph_kw)
    t1 = time.perf_counter()
    lmax = ograph.estimate_lmax(g["L"], g["dw"])
    t2 = time.perf_counter()
    omeld.transform(g["L"], lmax, labels, **filter_kw)
    t3 = time.perf_counter()
    dt = t3 - t0
    stages = {"graph_s": round(t1 - t0, 3), "lmax_s": round(t2 - t1, 3), "filter_s": round(t3 - t2, 3)}
    return n_cells / dt, dt, int(g["L"].nnz), stages


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    from meld_b200 import synthetic

    cfg = synthetic.CONFIGS[args.config]
    n_full = args.cells or cfg["N"]
    n = min(args.cpu_cells, n_full)
    vals = []
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_reference_leg(args, n, cores)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, dt, nnz, stages = cpu_reference_leg(args, n, cores)
        vals.append((v, dt))
    total = time.perf_counter() - t_all
    value = n * args.steps / total
    sample = ("oracle port of the reference CPU path (sklearn ball-tree kNN n_jobs={}, scipy CSC matvecs 1 thread) on a "
              "{}-cell sample of the workload generator; CPU kNN cost grows ~N^2 so cells/s at the full size is lower"
              ).format(cores, n)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, cfg, n_full), "sample_cells": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "stages_last_step": stages},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
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

    # N > 1, default: every rank runs the whole path on ITS OWN dataset of the config (weak scaling; a 500k-cell
    # job takes ~55 ms on one GPU and does not outgrow it, so there is nothing to exchange -- DESIGN.md section 6).
    # --multi sharded: ONE dataset, the candidate search sharded over the ranks by query rows, one NCCL
    # all-gather of the candidate lists, the rest replicated (strong scaling).
    sharded = world > 1 and args.multi == "sharded"
    Xh, labels, kw, cfg = make_inputs(args, seed_offset=0 if (sharded or world == 1) else 100 * rank)
    if sharded:
        kw = dict(kw, distributed=True)
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

    fit_marks = []

    def step_device(events=None):
        op = meld_b200.MELD(verbose=0, **kw)
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

    def step_e2e():
        op = meld_b200.MELD(verbose=0, **kw)
        dens = op.fit_transform(X_pin_np, labels)  # host buffers in, DataFrame (host) out
        return op, dens

    # ---- device-resident arm ("value") + live SpMV timing
    for _ in range(args.warmup):
        op, out = step_device()
    barrier()
    events = []
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.meld_b200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    gc.collect()
    gc.disable()
    e0.record()
    marks[0].record()
    for i in range(args.steps):
        op, out = step_device(events)
        marks[i + 1].record()  # per-step times for the record (no extra synchronisation)
    e1.record()
    barrier()
    gc.enable()
    launches = lib.meld_b200_launch_count() - launches0
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    step_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    nnz = op.graph.nnz
    lmax = op.graph.lmax
    stats = op.graph.build_stats()
    bt = op.graph.build_times()
    filt_ms = [a.elapsed_time(b) for a, b, _ in events]
    launch_us = 1e3 * float(np.mean(filt_ms)) / m  # average duration of one cheby_step launch
    bytes_step = nnz * 12 + (n + 1) * 4 + 5 * n * p * 8
    achieved = bytes_step / (launch_us * 1e-6) / 1e9

    # ---- end-to-end arm: host (pinned) inputs, DataFrame back on the host, copies inside the timed region
    for _ in range(args.warmup):
        step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = []
    gc.collect()
    gc.disable()  # a generation-2 collection in the middle of a 60 ms step is measurement noise, not the engine
    t_e2e = time.perf_counter()
    f0.record()
    for _ in range(args.steps):
        t_s = time.perf_counter()
        op_e2e, dens = step_e2e()
        e2e_steps.append(1e3 * (time.perf_counter() - t_s))  # the DataFrame is on the host when the call returns
    f1.record()
    barrier()
    e2e_ms = max(f0.elapsed_time(f1), 1e3 * (time.perf_counter() - t_e2e))
    gc.enable()

    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    jobs = 1 if (sharded or world == 1) else world  # replicas: every rank pushed n cells through per step
    value = jobs * n * args.steps / (ms_total * 1e-3)
    e2e_value = jobs * n * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        pk, pk_kind = peaks()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "step_ms": {"min": min(step_ms), "median": float(np.median(step_ms)), "max": max(step_ms),
                        "fit_median": float(np.median([a.elapsed_time(b) for a, b in fit_marks])) if fit_marks else None,
                        "filter_median": float(np.median(filt_ms)) if filt_ms else None},
            "config": {
                "workload": workload_name(args, cfg, n),
                "parallelism": "1 GPU" if world == 1 else (
                    "query rows of the candidate search sharded x{} + NCCL all-gather of candidate lists; Laplacian "
                    "assembly and filter replicated".format(world) if sharded else
                    "{} independent replicas (one {}-cell dataset per GPU, different seeds), no data-path "
                    "collective; value = cells of all ranks / max-over-ranks time".format(world, n)),
                "nnz_L": int(nnz), "nnz_per_row": nnz / n, "lmax": lmax, "candidate_cap": stats["candidate_cap"],
                "max_candidates": stats["max_candidates"], "search_passes": stats["search_passes"],
                "row_blocks": stats["row_blocks"], "direct_blocks": stats["direct_blocks"],
                "dict_entries_per_nnz": stats["dict_total"] / max(nnz, 1),
                "l2_note": "inputs (X {} MB, L {} MB) exceed the 126 MB L2; no explicit flush".format(
                    Xh.nbytes // 2**20, nnz * 12 // 2**20),
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(jobs * (Xh.nbytes + 4 * n)),
                    "d2h_bytes_per_step": int(jobs * 8 * n * p), "ms_per_step": e2e_ms / args.steps,
                    "step_ms": {"min": min(e2e_steps), "median": float(np.median(e2e_steps)), "max": max(e2e_steps)},
                    "host_timings_ms_last_step": {k: round(1e3 * v, 2) for k, v in op_e2e.timings_.items()}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {
                "kernel": "cheby_flat_kernel<4,8,1024> (CSR SpMM + fused three-term update; meld_b200_cheby_step)",
                "bound": "hbm",
                "achieved": achieved, "peak": pk["hbm_gbs"], "peak_kind": pk_kind, "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"],
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of this
                # kernel on this workload (profiles/r01b_ncu_cheby_flat_c4.txt); other workloads: not captured
                "traffic": 443832576 if (args.config == "c4" and n == 500000) else None,
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
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ncpu = min(args.cpu_cells, n)
            v, dt, _, cpu_stages = cpu_reference_leg(args, ncpu, cores)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "oracle port (sklearn ball-tree kNN n_jobs={}, scipy matvecs) fit_transform on a {}-cell "
                          "sample of the same generator, {:.1f} s".format(cores, ncpu, dt),
                "stages": cpu_stages,
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
