"""Where does the end-to-end (host in, DataFrame out) step spend its time?  Development probe (gpurun)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meld_b200  # noqa: E402
from meld_b200 import synthetic  # noqa: E402
from meld_b200.graph import _as_device_f64  # noqa: E402


def main():
    Xh, labels, kw = synthetic.make_config("c4")
    X_pin = torch.from_numpy(Xh).pin_memory().numpy()
    sync = torch.cuda.synchronize

    def lap(t0, what, out):
        sync()
        t1 = time.perf_counter()
        out.append("{} {:.1f}".format(what, 1e3 * (t1 - t0)))
        return t1

    for rep in range(5):
        if rep == 3:
            os.environ["MELD_B200_NO_LABEL_THREAD"] = "1"
            print("-- label thread off")
        out = []
        t0 = time.perf_counter()
        tA = t0
        Xd = _as_device_f64(torch, X_pin)
        t0 = lap(t0, "h2d", out)
        op = meld_b200.MELD(verbose=0, **kw)
        op.fit(Xd)
        t0 = lap(t0, "fit", out)
        samples, codes = op._label_codes(labels)
        t0 = lap(t0, "label_codes", out)
        dens = op.transform(labels)
        t0 = lap(t0, "transform", out)
        out.append("total {:.1f}".format(1e3 * (t0 - tA)))
        sync()
        t0 = time.perf_counter()
        op2 = meld_b200.MELD(verbose=0, **kw)
        d2 = op2.fit_transform(X_pin, labels)
        sync()
        out.append("| fit_transform(host) {:.1f}".format(1e3 * (time.perf_counter() - t0)))
        out.append("timings_ {}".format({k: round(1e3 * v, 1) for k, v in op2.timings_.items()}))
        print("rep", rep, "; ".join(out), flush=True)


if __name__ == "__main__":
    main()
