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


def _worker_sharded_filter(rank, world, port, out_dir):
    """ShardedFilter's NCCL-mode host logic (chunk partition, ping-pong full buffers, in-place all-gather, final
    gather of R, cell-order conversion) over gloo, with the two native operations replaced by numpy ones."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from meld_b200 import distributed as mdist, filter as mf

    g = load_golden("blobs2k5_wagner")
    L, lmax = g["L"].tocsr(), g["lmax"]
    N = L.shape[0]
    perm = np.random.default_rng(4).permutation(N)  # internal row a = caller's cell perm[a]
    Lint = L[perm][:, perm].tocsr()

    class CpuSharded(mdist.ShardedFilter):
        def __init__(self):
            self.graph, self.group, self.mode, self.p_max = None, None, "nccl", 4
            self.rank, self.world, self.N = rank, world, N
            self.chunk, self.bounds = mdist.chunk_partition(N, world)
            self.row_range = (self.bounds[rank], self.bounds[rank + 1])
            self._nccl_bufs = {}
            self.Lr = Lint[self.row_range[0]:self.row_range[1]]

        def _permute(self, src, p, to_internal, dst):
            if to_internal:
                dst[:N] = src[:N][torch.from_numpy(perm)]
            else:
                dst[torch.from_numpy(perm)] = src[:N]

        def _step(self, cur, told, tnew, Rloc, p, alpha, shift, gamma, c, c_cur, accumulate):
            a, b = self.row_range
            Tc = cur.numpy()[:N]
            tn = alpha * (self.Lr @ Tc - shift * Tc[a:b])
            if gamma != 0.0:
                tn = tn - gamma * told.numpy()[: b - a]
            rv = c * tn + c_cur * Tc[a:b]
            if accumulate:
                rv = rv + Rloc.numpy()[: b - a]
            Rloc[: b - a] = torch.from_numpy(rv)
            if tnew is not None:
                tnew[: b - a] = torch.from_numpy(tn)

    sf = CpuSharded()
    S = torch.from_numpy(np.random.default_rng(1).normal(size=(N, 6)))  # 6 columns with p_max 4: two chunks
    coeffs = mf.cheby_coefficients(mf.filter_kernel("heat", 60), lmax, 16)
    R = sf.apply(lmax, coeffs, S)
    if rank == 0:
        np.save(os.path.join(out_dir, "R_sharded.npy"), R.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_filter_host_logic_gloo(tmp_path, world):
    from oracle import cheby
    from meld_b200 import filter as mf, distributed as mdist

    mp.spawn(_worker_sharded_filter, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    g = load_golden("blobs2k5_wagner")
    S = np.random.default_rng(1).normal(size=(g["L"].shape[0], 6))
    coeffs = mf.cheby_coefficients(mf.filter_kernel("heat", 60), g["lmax"], 16)
    ref = cheby.cheby_op(g["L"], g["lmax"], coeffs, S)
    out = np.load(os.path.join(str(tmp_path), "R_sharded.npy"))
    assert np.abs(out - ref).max() <= 1e-12 * np.abs(ref).max()
    assert mdist.chunk_partition(10, 3) == (4, [0, 4, 8, 10])
    assert mdist.chunk_partition(7, 8)[1] == [0, 1, 2, 3, 4, 5, 6, 7, 7]
