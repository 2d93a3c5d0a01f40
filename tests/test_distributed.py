"""World-size-2 gloo test (CPU) of the row-partitioned Chebyshev driver: the host logic that shards rows,
exchanges T_k after every term and reassembles R.  The per-rank step here is a scipy restatement (test
infrastructure); on GPUs the same driver calls meld_b200_cheby_step (tests/test_gpu_filter.py covers that)."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from meld_b200 import distributed as mdist, filter as mf

    g = load_golden("blobs2k_k15")
    L, lmax = g["L"].tocsr(), g["lmax"]
    N = L.shape[0]
    p, m = 3, 20
    S = np.random.default_rng(0).normal(size=(N, p))
    coeffs = mf.cheby_coefficients(mf.filter_kernel("heat", 60), lmax, m)
    bounds = mdist.row_partition(N, world)
    a, b = bounds[rank], bounds[rank + 1]
    Lr = L[a:b]

    def step(T_cur_full, T_old_local, alpha, shift, gamma, c, c_cur, R_local, accumulate):
        Tc = T_cur_full.numpy()
        y = Lr @ Tc
        tn = alpha * (y - shift * Tc[a:b])
        if gamma != 0.0:
            tn = tn - gamma * T_old_local.numpy()
        rv = c * tn + c_cur * Tc[a:b]
        if accumulate:
            rv = rv + R_local.numpy()
        return torch.from_numpy(tn), torch.from_numpy(rv)

    allgather = mdist.make_torch_allgather(bounds, p)
    R_local = mdist.cheby_recurrence(step, allgather, torch.from_numpy(S), (a, b), lmax, coeffs)
    np.save(os.path.join(out_dir, "R_{}.npy".format(rank)), R_local.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_row_partitioned_driver_two_ranks_gloo(tmp_path):
    from oracle import cheby
    from meld_b200 import filter as mf, distributed as mdist

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = load_golden("blobs2k_k15")
    S = np.random.default_rng(0).normal(size=(g["L"].shape[0], 3))
    coeffs = mf.cheby_coefficients(mf.filter_kernel("heat", 60), g["lmax"], 20)
    ref = cheby.cheby_op(g["L"], g["lmax"], coeffs, S)
    out = np.concatenate([np.load(os.path.join(str(tmp_path), "R_{}.npy".format(r))) for r in range(world)])
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
    assert mdist.row_partition(10, 3) == [0, 4, 7, 10]
