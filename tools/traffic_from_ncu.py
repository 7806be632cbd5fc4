#!/usr/bin/env python
"""Turn `ncu --page raw --csv` exports of the search launches of one bench step into profiles/traffic_r2.json (what bench.py's
roofline.traffic reads) and a markdown table.  Usage: python tools/traffic_from_ncu.py <workload key> <T>=<csv> [<T>=<csv> ...]"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "time us"), ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "hmma inst %"),
        ("sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", "imma inst %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"), ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("launch__registers_per_thread", "regs")]


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    key = sys.argv[1]
    out = {"dram_bytes_by_T": {}, "launches": {}, "source": "ncu --set full --clock-control none, per-launch dram__bytes_read.sum + dram__bytes_write.sum summed over the "
                                                            "level's search launches of one bench step (profiles/search_kernels_r2.md)"}
    md = ["| level | launch | kernel | " + " | ".join(n for _, n in COLS) + " |", "|---|---|---|" + "---:|" * len(COLS)]
    for spec in sys.argv[2:]:
        T, path = spec.split("=")
        rows = list(csv.reader(open(path)))
        hdr, units, data = rows[0], rows[1], rows[2:]
        tot = 0.0
        for i, r in enumerate(data):
            cells = []
            for name, _ in COLS:
                if name not in hdr:
                    cells.append("n/a")
                    continue
                j = hdr.index(name)
                if name.startswith("dram__bytes"):
                    b = to_bytes(r[j], units[j])
                    tot += b
                    cells.append("%.1f MB" % (b / 1e6))
                elif name.startswith("gpu__time"):
                    cells.append("%s %s" % (r[j], units[j]))
                else:
                    cells.append(r[j])
            md.append("| T=%s | %d | `%s` | %s |" % (T, i + 1, r[hdr.index("Kernel Name")].split("(")[0][:40], " | ".join(cells)))
        out["dram_bytes_by_T"][T] = tot
        out["launches"][T] = len(data)
    path = os.path.join(ROOT, "profiles", "traffic_r2.json")
    allv = {}
    if os.path.exists(path):
        allv = json.load(open(path))
    allv[key] = out
    json.dump(allv, open(path, "w"), indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
