#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, no GPU needed) into a markdown table and a
JSON record of per-launch DRAM traffic that bench.py quotes as roofline.traffic.

  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_x [--cells N]
writes profiles/r01_x.md and profiles/r01_x.json.
"""
import csv
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
         "s": 1.0, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    cells = None
    if "--cells" in sys.argv:
        cells = float(sys.argv[sys.argv.index("--cells") + 1])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    record, md = [], ["# %s\n\nsource: `%s` (ncu --set full --clock-control none)\n" % (out, rep)]
    for r in rows[2:]:
        md.append("## %s\n\n| metric | value | unit |\n|---|---|---|" % r[name_col])
        vals = {}
        for m in METRICS:
            if m not in hdr:
                continue
            c = hdr.index(m)
            md.append("| %s | %s | %s |" % (m, r[c], units[c]))
            try:
                vals[m] = float(r[c].replace(",", "")) * SCALE.get(units[c], 1.0)
            except ValueError:
                pass
        traffic = vals.get("dram__bytes_read.sum", 0.0) + vals.get("dram__bytes_write.sum", 0.0)
        rec = {"kernel": r[name_col], "time_s": vals.get("gpu__time_duration.sum"),
               "dram_bytes_read": vals.get("dram__bytes_read.sum"),
               "dram_bytes_write": vals.get("dram__bytes_write.sum"), "traffic_bytes": traffic,
               "registers": vals.get("launch__registers_per_thread"),
               "grid": vals.get("launch__grid_size"), "block": vals.get("launch__block_size")}
        if cells:
            rec["cells"] = cells
            rec["traffic_bytes_per_cell"] = traffic / cells
            md.append("| DRAM traffic per cell | %.2f | byte |" % (traffic / cells))
        if rec["time_s"]:
            md.append("| DRAM traffic / duration | %.1f | GB/s |" % (traffic / rec["time_s"] / 1e9))
        md.append("")
        record.append(rec)
    open(out + ".md", "w").write("\n".join(md) + "\n")
    json.dump({"source": rep, "launches": record}, open(out + ".json", "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
