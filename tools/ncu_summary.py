#!/usr/bin/env python
"""Summaries of the ncu outputs that profiles/ keeps:
    python tools/ncu_summary.py launches <launches.csv>           per-kernel totals / shares of a launch list
    python tools/ncu_summary.py raw <report.ncu-rep>              key metrics of every kernel in a --set full capture
"""
import collections
import csv
import json
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"].startswith("n"):
            v /= 1000.0
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {t / c:.1f} | {100 * t / tot:.1f}% |")
    print(f"| total | {sum(a[0] for a in agg.values())} | {tot:.1f} | | |")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": re.sub(r"\(.*", "", vals[hdr.index("Kernel Name")]).replace("<unnamed>::", "").replace("void ", "")}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                d[w] = f"{vals[i]} {units[i]}"
        res.append(d)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
