#!/usr/bin/env python
"""Summarise .ncu-rep captures on the GPU box (the reports themselves are too large to bring back):
   ncu_summary.py out.txt a.ncu-rep b.ncu-rep ...
For every captured launch: duration, DRAM bytes read / written, DRAM throughput %, registers, occupancy, the top warp
stall reasons (per issued instruction) and pipe utilisations -- from `ncu -i ... --page raw --csv`."""
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "L1 LSU wavefronts"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 wavefronts % of peak"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__block_size", "block size"),
    ("launch__grid_size", "grid size"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (shared mem), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe cycles active %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("nvltx__bytes.sum", "NVLink TX bytes"),
    ("nvlrx__bytes.sum", "NVLink RX bytes"),
]
STALL = re.compile(r"smsp__average_warps?_issue_stalled_(\w+)_per_issue_active\.ratio|smsp__average_warp_latency_issue_stalled_(\w+)\.ratio")


def main():
    out = open(sys.argv[1], "w")
    for rep in sys.argv[2:]:
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(r.stdout)))
        if len(rows) < 3:
            out.write("== %s: no data (%s)\n" % (rep, r.stderr[-200:]))
            continue
        hdr, units = rows[0], rows[1]
        col = {h: i for i, h in enumerate(hdr)}
        for row in rows[2:]:
            name = row[col.get("Kernel Name", 4)] if "Kernel Name" in col else "?"
            out.write("== %s\n   kernel: %s\n" % (rep, name[:160]))
            for key, label in WANT:
                if key in col:
                    out.write("   %-46s %s %s\n" % (label, row[col[key]], units[col[key]]))
            stalls = []
            for h, i in col.items():
                m = STALL.match(h)
                if m:
                    try:
                        stalls.append((float(row[i].replace(",", "")), m.group(1) or m.group(2)))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            out.write("   top stalls (warps per issue-active cycle): " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]) + "\n")
    out.close()


if __name__ == "__main__":
    main()
