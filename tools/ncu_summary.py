#!/usr/bin/env python
"""Turn an .ncu-rep (brought back from a gpurun call) into the small text summary committed under profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_name.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = ["source: %s (ncu --set full --clock-control none --import-source on)" % rep, ""]
    for r in rows[2:]:
        lines.append("kernel: " + r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                lines.append("  %-82s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        if "dram__bytes_read.sum" in hdr:
            def gb(name):
                v, u = float(r[hdr.index(name)]), units[hdr.index(name)]
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
            lines.append("  traffic (dram read + write) per launch: %.0f bytes" %
                         (gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")))
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", "0", "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    hidx = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
    if hidx:
        h = hidx[0]
        sh = srows[h]
        i_src, i_s = sh.index("Source"), sh.index("Warp Stall Sampling (All Samples)")
        data = []
        for r in srows[h + 1:(hidx[1] if len(hidx) > 1 else len(srows))]:
            if len(r) > i_s and r[0].startswith("0x"):
                data.append((int(r[i_s]) if r[i_s].isdigit() else 0, r[i_src].strip()))
        tot = sum(d[0] for d in data) or 1
        lines.append("top warp-stall sampling sites of the first launch (SASS, >= 2 %% of %d samples):" % tot)
        for k, (s, txt) in enumerate(data):
            if s >= 0.02 * tot:
                lines.append("  #%-4d %5.1f%%  %s" % (k, 100.0 * s / tot, txt[:100]))
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
