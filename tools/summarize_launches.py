#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (one row per launch):
launches, mean / min duration and the kernel's share of the summed kernel time, as a markdown table.
Durations under ncu are serialised and cold-cache: the SHARE is what compares with bench.py's stage_ms_serial.

    python tools/summarize_launches.py profiles/r2n_launches.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[1:]:
        if r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^(void )?(fm::|\(anonymous namespace\)::|<unnamed>::)?", "", r[col["Kernel Name"]].split("(")[0])
        v = float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[col["Metric Unit"]], 1.0)
        a = agg.setdefault(name, {"n": 0, "sum": 0.0, "min": 1e30, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
        a["n"] += 1; a["sum"] += v; a["min"] = min(a["min"], v)
    total = sum(a["sum"] for a in agg.values())
    print("| kernel | grid | block | launches | mean us | min us | share of kernel time |")
    print("|---|---|---|---|---|---|---|")
    for name, a in agg.items():
        print(f"| `{name}` | {a['grid']} | {a['block']} | {a['n']} | {a['sum'] / a['n']:.1f} | {a['min']:.1f} | {100 * a['sum'] / total:.1f} % |")
    print(f"\nsum of kernel time: {total:.1f} us over {sum(a['n'] for a in agg.values())} launches")


if __name__ == "__main__":
    main(sys.argv[1])
