#!/usr/bin/env python
"""Extract the metrics the roofline discussion uses from an .ncu-rep (`ncu -i rep --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__waves_per_multiprocessor", "waves / SM"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
]
STALLS = "smsp__average_warps_issue_stalled_"


def main(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("### `%s`  grid %s block %s\n" % (r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
        print("| metric | value | unit |\n|---|---:|---|")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print("| %s (`%s`) | %s | %s |" % (label, key, r[i], units[i]))
        st = [(h[len(STALLS):-len("_per_issue_active.ratio")], float(r[i].replace(",", ""))) for i, h in enumerate(hdr)
              if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")]
        top = sorted(st, key=lambda x: -x[1])[:7]
        print("| top stall reasons (warps per issue-active cycle) | %s | |" % ", ".join("%s %.2f" % t for t in top))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
