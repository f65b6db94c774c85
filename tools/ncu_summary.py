#!/usr/bin/env python3
"""Turn `ncu -i REPORT --page raw --csv` output into the short per-kernel text summary kept under profiles/.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/raw.csv; python tools/ncu_summary.py /tmp/raw.csv "title" > profiles/...
"""
import collections
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    groups = collections.OrderedDict()
    for r in rows[2:]:
        groups.setdefault(r[ik].split("(")[0], []).append(r)
    print("# " + (sys.argv[2] if len(sys.argv) > 2 else "ncu --set full --clock-control none"))
    for name, rs in groups.items():
        r = rs[-1]
        print(f"\n## {name}   ({len(rs)} captured launches; last one shown)")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"    {w:<84}{r[i]:>12} {units[i]}")


if __name__ == "__main__":
    main()
