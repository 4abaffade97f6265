"""Markdown summary of an `ncu --set full` report: one column per captured launch, the metrics the roofline discussion
in DESIGN.md uses.  Usage: python scripts/ncu_summary.py report.ncu-rep [title] > profiles/rNN_ncu_<what>.md"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_registers", "CTAs / SM (register limit)"),
    ("launch__occupancy_limit_shared_mem", "CTAs / SM (shared-memory limit)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 / shared-memory pipe"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor pipe"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active (of active cycles)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe cycles active (of elapsed cycles)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global loads)"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall: MIO throttle"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (shared memory)"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction (I-cache)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    name = hdr.index("Kernel Name")
    print("# %s\n" % title)
    print("`ncu --set full --clock-control none --import-source on`, one capture per kernel (cold cache, serialised).\n")
    print("| metric | " + " | ".join("`%s`" % r[name].split("(")[0].replace("void ", "") for r in data) + " |")
    print("|---|" + "---:|" * len(data))
    for key, label in METRICS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for r in data:
            v = r[i]
            try:
                f = float(v.replace(",", ""))
                v = ("%.0f" % f) if abs(f) >= 1000 or f == int(f) else ("%.2f" % f)
            except ValueError:
                pass
            vals.append(v + (" " + units[i] if units[i] else ""))
        print("| %s | " % label + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
