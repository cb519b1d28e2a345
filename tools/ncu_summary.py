"""Key counters of every launch in an `ncu --set full` report, as a table: python tools/ncu_summary.py <report.ncu-rep | raw.csv> [> profiles/...]
(reads the raw page: `ncu -i report --page raw --csv`; the .ncu-rep files themselves stay in gpurun_out/, which is scratch)."""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid (CTAs)"),
    ("launch__block_size", "block (threads)"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / CTA"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of peak)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy (%)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe (%)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe (%)"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe (%)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (%)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe (%)"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "L1 data pipe: all LSU wavefronts (% of peak)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts (% of peak)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global-load requests"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate (%)"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> L1 bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "L2 -> L1 throughput"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall / issue: wait (fixed-latency dependency)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall / issue: short scoreboard (smem, MUFU)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall / issue: long scoreboard (global, TMEM)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall / issue: barrier"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall / issue: branch resolving"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall / issue: not selected"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall / issue: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall / issue: dispatch"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall / issue: sleeping"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall / issue: no instruction"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall / issue: MIO throttle"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall / issue: LG throttle"),
]


def main():
    src = sys.argv[1]
    if src.endswith(".csv"):
        txt = open(src).read()
    else:
        txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")] for r in data]
    print("source: %s (%d launches; ncu --set full --clock-control none: per-launch times are cold-cache and serialised)" % (src, len(data)))
    w = max(len(m[1]) for m in METRICS) + 2
    print("%-*s %s" % (w, "", " | ".join("%-34s" % n[:34] for n in names)))
    for key, label in METRICS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for r in data:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = ("%.4g" % f) if abs(f) < 1e6 else ("%.4e" % f)
            except ValueError:
                pass
            vals.append("%-34s" % ("%s %s" % (v, units[i])))
        print("%-*s %s" % (w, label, " | ".join(vals)))


if __name__ == "__main__":
    main()
