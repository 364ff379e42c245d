"""Summarise an .ncu-rep (top kernel) into a small text file for profiles/ (run where ncu is installed)."""
import csv, subprocess, sys, collections, re

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__icc_request_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg"]

def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    for r in rows:
        d = dict(zip(hdr, r))
        print("kernel:", d.get("Kernel Name"), "| id", d.get("ID"))
        for k in KEYS:
            if k in d:
                print("  %-70s %s %s" % (k, d[k], units[hdr.index(k)]))
        st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(d[h]) for h in hdr
              if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and d[h] not in ("", "n/a")}
        tot = sum(st.values()) or 1.0
        print("  warp stall samples (%):", ", ".join("%s %.1f" % (k, 100 * v / tot) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:7]))

if __name__ == "__main__":
    main()
