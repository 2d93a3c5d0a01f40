"""Cross-check the tcgen05 candidate search against the SIMT one (development probe, run under gpurun)."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from meld_b200 import _native as nv, synthetic  # noqa: E402


def search(X, knn, simt):
    n, d = X.shape
    key2 = torch.empty(n, dtype=torch.float32, device="cuda")
    cnt = torch.empty(n, dtype=torch.int32, device="cuda")
    cap = C.c_int64()
    torch.cuda.synchronize()
    t = time.perf_counter()
    nv.check(nv.lib().meld_b200_debug_candidate_search(nv.ptr(X), n, d, knn, 40.0, 1e-4, 1.0,
                                                       nv.FLAG_SIMT_SEARCH if simt else 0, nv.current_stream_ptr(),
                                                       nv.ptr(key2), nv.ptr(cnt), C.byref(cap)), "debug_search")
    torch.cuda.synchronize()
    return key2.cpu().numpy(), cnt.cpu().numpy(), cap.value, time.perf_counter() - t


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    knn = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    only_tc = len(sys.argv) > 4 and sys.argv[4] == "tc"
    if len(sys.argv) > 5:
        nv.set_tuning(tc_multicast=int(sys.argv[5]))
    Xh, _ = synthetic.make_blobs(n, d, 8, 3, 10.0, seed=3)
    X = torch.from_numpy(Xh).cuda()
    k_tc, c_tc, cap_tc, t_tc = search(X, knn, False)
    print("tc   : key2[:5]", k_tc[:5], "cnt[:8]", c_tc[:8], "cap", cap_tc, "mean cnt %.1f max %d" % (c_tc.mean(), c_tc.max()),
          "time %.3fs" % t_tc, flush=True)
    _, _, _, t_tc2 = search(X, knn, False)
    print("tc second call time %.4fs  (2 N^2 K' flops x2 passes -> %.1f TFLOP/s)" % (
        t_tc2, 2 * 2.0 * n * n * (3 * d + 3) / t_tc2 / 1e12))
    if only_tc:
        return
    k_s, c_s, cap_s, t_s = search(X, knn, True)
    print("simt : key2[:5]", k_s[:5], "cnt[:8]", c_s[:8], "cap", cap_s, "mean cnt %.1f max %d" % (c_s.mean(), c_s.max()),
          "time %.3fs" % t_s)
    print("max |key2 diff| rel", np.abs(k_tc - k_s).max() / np.abs(k_s).max(), " cnt diff", np.abs(c_tc - c_s).max())


if __name__ == "__main__":
    main()
