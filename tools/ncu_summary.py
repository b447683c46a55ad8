"""Summarise .ncu-rep captures (ncu --set full) into a markdown table for profiles/.
python tools/ncu_summary.py title1=path1.ncu-rep [title2=path2.ncu-rep ...] > profiles/rNN_ncu_summary.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
]


def main():
    for arg in sys.argv[1:]:
        title, path = arg.rsplit("=", 1)  # the title may hold "="
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(vals, units)))
            print("## %s\n" % title)
            print("| metric | value | unit |\n|---|---|---|")
            for k in KEYS:
                if k in d:
                    print("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
            print("\nWarp stall reasons (per issue-active cycle, > 0.03):\n")
            for h in hdr:
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        v = float(d[h][0])
                    except ValueError:
                        continue
                    if v > 0.03:
                        print("- %s: %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
            print()


if __name__ == "__main__":
    main()
