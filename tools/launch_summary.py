"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python tools/launch_summary.py gpurun_out/launches.csv "<the command that was profiled>" > profiles/…_summary.txt

The per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's, not absolutes.
"""

import csv
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    what = sys.argv[2] if len(sys.argv) > 2 else ""
    lines = open(path, newline="").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    tot, cnt = defaultdict(float), defaultdict(int)
    for row in csv.DictReader(lines[start:]):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        tot[row["Kernel Name"]] += ms
        cnt[row["Kernel Name"]] += 1
    total = sum(tot.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none launch list of `%s`" % what)
    print("# total %.1f ms over %d launches (cold-cache, serialised: compare SHARES)" % (total, sum(cnt.values())))
    for k in sorted(tot, key=tot.get, reverse=True):
        print("%10.2f ms %5.1f%% x %4d  %s" % (tot[k], 100 * tot[k] / total, cnt[k], k[:130]))


if __name__ == "__main__":
    main()
