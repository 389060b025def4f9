#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export (one row per captured launch) as a markdown table.

    ncu -i gpurun_out/prof_TAG.ncu-rep --page raw --csv > gpurun_out/prof_TAG_raw.csv
    python tools/summarize_ncu.py gpurun_out/prof_TAG_raw.csv > profiles/TAG_summary.md
"""
import csv
import sys

KEYS = [
    ("duration us", "gpu__time_duration.sum"),
    ("regs/thread", "launch__registers_per_thread"),
    ("CTAs/SM (regs | smem | warps)", None),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue slots busy %", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("FMA pipe busy % (cycles)", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("FMA pipe instr % of peak", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
    ("ALU pipe instr % of peak", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("LSU pipe instr % of peak", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    ("tensor pipe busy % (cycles)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor pipe (i8/subpipe) busy %", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("warp instructions", "smsp__inst_executed.sum"),
    ("DRAM read MB", "dram__bytes_read.sum"),
    ("DRAM write MB", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("smem bank conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    names = [r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("fm::", "") for r in data]
    print("| metric | " + " | ".join(names) + " |")
    print("|---|" + "---|" * len(names))
    for label, key in KEYS:
        cells = []
        for r in data:
            if key is None:
                cells.append(" \\| ".join(str(int(float(r[col[k]]))) for k in
                                          ("launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
                                           "launch__occupancy_limit_warps")))
            elif key in col:
                v = r[col[key]]
                try:
                    f = float(v)
                    u = units[col[key]]
                    if label.endswith(" MB"):
                        f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    elif label.endswith(" us"):
                        f *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                    cells.append(f"{f:.1f}" if abs(f) < 1e4 else f"{f:.3e}")
                except ValueError:
                    cells.append(v)
            else:
                cells.append("n/a")
        print(f"| {label} | " + " | ".join(cells) + " |")
    print()
    print("Top stall reasons (warps stalled per issue-active cycle):")
    print()
    for name, r in zip(names, data):
        st = []
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print(f"* `{name}`: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:6]))


def traffic_json(path, out_path):
    """Per-kernel DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of an `ncu --set full` capture,
    for bench.py's roofline.traffic: {kernel: {dram_bytes, duration_us}, "_source": csv path}."""
    import json
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        out = json.load(open(out_path))                  # several captures (chain, channelizer) share one file
    except Exception:
        out = {}
    out["_source"] = ", ".join(sorted(set(filter(None, [out.get("_source"), path]))))
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("fm::", "").replace("<unnamed>::", "").split("<")[0]
        b = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[col[key]]) * scale.get(units[col[key]], 1.0)
        dur = float(r[col["gpu__time_duration.sum"]]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1.0)
        out[name] = {"dram_bytes": b, "duration_us": dur}
    json.dump(out, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[2] == "--traffic-json":
        traffic_json(sys.argv[1], sys.argv[3])
    else:
        main(sys.argv[1])
