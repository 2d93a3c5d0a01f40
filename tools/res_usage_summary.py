"""Registers / stack / shared memory of the library's kernels from `cuobjdump -res-usage` (no GPU needed).

    cuobjdump -res-usage meld_b200/libmeld_b200.so > /tmp/res.txt; python tools/res_usage_summary.py /tmp/res.txt
"""

import re
import subprocess
import sys

HOT = ("cheby_flat2_kernel<4, 1024, 1, 1, 0, 0>", "cheby_flat2_kernel<1, 1024, 1, 1, 1, 0>",
       "cheby_flat2_kernel<4, 1024, 1, 1, 0, 1>", "cheby_flat2_kernel<8, 1024, 2, 0, 0, 0>",
       "cheby_flat2_kernel<6, 1024, 2, 0, 0, 0>", "tc_search_kernel", "refine_dist_kernel", "kernel_values_kernel",
       "combine_basis_kernel", "lanczos_axpy_kernel(", "merge_rows_kernel", "fill_sym_kernel", "kmeans_assign_kernel<2>",
       "tile_proj_kernel<2>", "cheby_flat_kernel<4, 8, 1024>", "cheby_flat_pipe_kernel<1, 8, 1024>", "merge_lists_kernel",
       "tc_prep_kernel", "mirror_records_kernel", "lanczos_axpy_peer_kernel")


def main():
    rows, name = [], None
    for line in open(sys.argv[1]):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            rows.append((name,) + tuple(int(x) for x in m.groups()))
            name = None
    dem = subprocess.run(["cu++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
    dem = [re.sub(r"\((int|bool)\)", "", d).replace("<0>", "<0>") for d in dem]  # cu++filt prints (int)4, (bool)0
    out = [(d,) + r[1:] for r, d in zip(rows, dem)]
    spill = [o for o in out if o[2] > 0 or o[4] > 0]
    print("# cuobjdump -res-usage meld_b200/libmeld_b200.so (sm_100a): registers / stack bytes / static shared bytes / local bytes")
    print("# %d kernels in the library; %d have a stack frame or local memory (listed at the end)" % (len(out), len(spill)))
    fmt = "REG %3d  STACK %4d  SHARED %6d  LOCAL %3d  %s"
    for o in sorted((o for o in out if any(k in o[0] for k in HOT)), key=lambda o: o[0]):
        print(fmt % (o[1], o[2], o[3], o[4], o[0][:160]))
    print("# kernels with a stack frame or local memory:")
    for o in sorted(spill, key=lambda o: o[0]):
        print(fmt % (o[1], o[2], o[3], o[4], o[0][:160]))


if __name__ == "__main__":
    main()
