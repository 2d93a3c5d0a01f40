"""Where do the slow end-to-end steps come from?  20 fit_transform calls per setting, every step's wall time."""
import gc, json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import meld_b200
from meld_b200 import synthetic

X, labels, kw = synthetic.make_config("c4")
Xp = torch.from_numpy(X).pin_memory().numpy()
codes = np.unique(labels, return_inverse=True)[1].astype(np.int64)


def run(tag, lab, env=None, n=20):
    for k, v in (env or {}).items():
        os.environ[k] = v
    for _ in range(3):
        meld_b200.MELD(verbose=0, **kw).fit_transform(Xp, lab)
    gc.collect(); gc.disable()
    ts, slow = [], None
    for _ in range(n):
        t = time.perf_counter()
        op = meld_b200.MELD(verbose=0, **kw)
        op.fit_transform(Xp, lab)
        ts.append(1e3 * (time.perf_counter() - t))
        if ts[-1] == max(ts):
            slow = {k: round(1e3 * v, 1) for k, v in op.timings_.items()}
    gc.enable()
    for k in (env or {}):
        os.environ.pop(k, None)
    print(json.dumps(dict(setting=tag, median=round(float(np.median(ts)), 2), mean=round(float(np.mean(ts)), 2),
                          max=round(max(ts), 2), all=[round(v, 1) for v in ts], slowest=slow)), flush=True)


run("default (string labels, label thread)", labels)
run("no label thread", labels, {"MELD_B200_NO_LABEL_THREAD": "1"})
run("integer labels (cheap factorisation)", codes)
run("MELD_B200_NO_ARENA", labels, {"MELD_B200_NO_ARENA": "1"})
