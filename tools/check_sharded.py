"""torchrun check of the sharded build (run under gpurun --gpus N):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py
Every rank builds the graph with the candidate search sharded over ranks and compares it, bit for bit, with
the graph rank-locally built by meld_b200_knn_graph_build; prints timings of both."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meld_b200  # noqa: E402
from meld_b200 import synthetic  # noqa: E402
from meld_b200.graph import DeviceGraph  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    rank, world = dist.get_rank(), dist.get_world_size()
    Xh, labels, kw = synthetic.make_config("c4", N=n)
    X = torch.from_numpy(Xh).cuda()
    gk = dict(knn=kw["knn"], decay=40.0, thresh=1e-4, anisotropy=1.0)
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gs = DeviceGraph.from_data_sharded(X, **gk)
        torch.cuda.synchronize()
        dist.barrier()
        t_sh = time.perf_counter() - t0
        t0 = time.perf_counter()
        g1 = DeviceGraph.from_data(X, **gk)
        torch.cuda.synchronize()
        t_1 = time.perf_counter() - t0
        if rank == 0:
            print("iter {}: sharded x{} {:.3f} s   single {:.3f} s".format(it, world, t_sh, t_1), flush=True)
    Ls, L1 = gs.to_scipy_L(), g1.to_scipy_L()
    same = (np.array_equal(Ls.indptr, L1.indptr) and np.array_equal(Ls.indices, L1.indices)
            and np.array_equal(Ls.data, L1.data))
    print("rank {}: nnz {} identical to single-GPU build: {}".format(rank, Ls.nnz, same), flush=True)
    # end to end through the public API
    op = meld_b200.MELD(verbose=0, distributed=True, **kw)
    dens = op.fit_transform(Xh, labels)
    op1 = meld_b200.MELD(verbose=0, **kw)
    dens1 = op1.fit_transform(Xh, labels)
    print("rank {}: densities max rel diff {:.2e}".format(rank, float(np.abs(dens.values - dens1.values).max() / np.abs(dens1.values).max())), flush=True)
    dist.destroy_process_group()
    assert same


if __name__ == "__main__":
    main()
