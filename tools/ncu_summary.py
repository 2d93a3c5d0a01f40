"""Condense an .ncu-rep into the handful of numbers DESIGN.md / the roofline refer to (run here, no GPU)."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
    "launch__cluster_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.sum",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sector_op_read_hit_rate.pct", "sm__cycles_active.avg",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor_subpipe_hmma.sum",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu summary of", rep)
    for r in rows[2:]:
        print("\n## kernel:", r[hdr.index("Kernel Name")][:100])
        for k in KEYS:
            if k in hdr and r[hdr.index(k)] not in ("", "nan", "-nan"):
                print("  {:95s} {} {}".format(k, r[hdr.index(k)], units[hdr.index(k)]))
        st = []
        for k in hdr:
            if k.startswith("smsp__pcsamp_warps_issue_stalled") and not k.endswith("not_issued"):
                try:
                    st.append((float(r[hdr.index(k)].replace(",", "")), k))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1.0
        print("  top warp-stall samples:", ", ".join("{} {:.0f}%".format(k[33:], 100 * v / tot) for v, k in sorted(st, reverse=True)[:5]))


if __name__ == "__main__":
    main()
