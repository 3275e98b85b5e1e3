#!/usr/bin/env python
"""tools/ncu_summary.py REPORT.ncu-rep [title] -> markdown summary of the metrics the roofline argument rests on."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__shared_mem_per_block_dynamic", "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
]


def main():
    rep = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else rep
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print(f"source: `{rep}` (ncu --set full --clock-control none), one column per captured launch\n")
    launches = rows[2:]
    print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(launches))) + " |")
    print("|---|---|" + "---|" * len(launches))
    print("| kernel | | " + " | ".join("`" + r[idx["Kernel Name"]].split("(")[0].replace("void <unnamed>::", "") + "`" for r in launches) + " |")
    for w in WANT:
        if w in idx:
            print(f"| {w} | {units[idx[w]]} | " + " | ".join(r[idx[w]] for r in launches) + " |")
    for r in launches:
        rd, wr, t = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]]), float(r[idx["gpu__time_duration.sum"]])
        ru, tu = units[idx["dram__bytes_read.sum"]], units[idx["gpu__time_duration.sum"]]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[ru]
        ts = {"us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "ns": 1e-9, "nsecond": 1e-9, "s": 1.0, "second": 1.0}[tu]
        print(f"\nDRAM traffic {((rd + wr) * scale) / 1e9:.3f} GB in {t * ts * 1e3:.3f} ms = {((rd + wr) * scale) / (t * ts) / 1e9:.0f} GB/s")


if __name__ == "__main__":
    main()
