#!/usr/bin/env python
"""Summarise `.ncu-rep` files (one profiled launch each) into a text table + the DRAM-traffic JSON bench.py reads.

    python scripts/ncu_summary.py OUT.txt TRAFFIC.json name1=rep1.ncu-rep name2=rep2.ncu-rep ...
"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "lsu wavefronts %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2 throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "cycles active"),
    ("sm__cycles_elapsed.avg.per_second", "sm clock"),
]


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return None
    head, units, vals = rows[0], rows[1], rows[-1]
    d = {h: (v, u) for h, u, v in zip(head, units, vals)}
    d["_kernel"] = (d.get("Kernel Name", ("?", ""))[0], "")
    return d


def main():
    out_txt, out_json = sys.argv[1], sys.argv[2]
    traffic = {"source": "ncu --set full --clock-control none, one launch each (scripts/final_evidence.sh)",
               "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum)"}
    lines = []
    for item in sys.argv[3:]:
        name, rep = item.split("=", 1)
        d = read(rep)
        if d is None:
            lines.append(f"== {name}: no data in {rep}")
            continue
        lines.append(f"== {name}: {d['_kernel'][0][:110]}")
        for key, label in METRICS:
            if key in d:
                v, u = d[key]
                lines.append(f"   {label:24s} {v} {u}")
        try:
            traffic[name] = float(d["dram__bytes_read.sum"][0].replace(",", "")) + \
                float(d["dram__bytes_write.sum"][0].replace(",", ""))
        except Exception:
            pass
    with open(out_txt, "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(out_json, "w") as f:
        json.dump(traffic, f, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
