#!/usr/bin/env python
"""Gathers the multi-GPU legs (tools/gpu_multi.sh N TAG, one `gpurun --gpus N` call each) into one JSON for profiles/:

    python tools/collect_multi_gpu.py TAG [N ...]   ->  profiles/TAG_multi_gpu.json

Per N: the bench line of `--workload wideband` (config 5: channels sharded, NCCL broadcast of the wideband block), the
bench line of the default workload (config 3 x N, streams sharded, no collective) and tools/h2d_ceiling.py's line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP_WIDE = ("n_gpus", "ms_per_step", "value", "unit", "scaling", "x_realtime_wideband", "wideband_MSps", "e2e", "gpu_launches",
             "rds_check", "roofline", "clocks", "config")
KEEP_STREAMS = ("n_gpus", "ms_per_step", "value", "unit", "scaling", "e2e", "gpu_launches", "rds_check", "clocks", "stage_ms_serial")


def last_json(path):
    try:
        lines = [ln for ln in open(path) if ln.startswith("{")]
        return json.loads(lines[-1]) if lines else None
    except OSError:
        return None


def main():
    tag = sys.argv[1]
    ns = [int(a) for a in sys.argv[2:]] or [1, 2, 4, 8]
    out = {"gpu_box": "N x NVIDIA B200 on one host, one process per GPU (torchrun)", "tag": tag,
           "config5_wideband_broadcast": [], "config3_streams_sharded": [], "host_ingest_ceiling": []}
    for n in ns:
        w = last_json(os.path.join(ROOT, "gpurun_out", f"{tag}_wideband_n{n}.log"))
        if w:
            out["config5_wideband_broadcast"].append({k: w[k] for k in KEEP_WIDE if k in w})
        s = last_json(os.path.join(ROOT, "gpurun_out", f"{tag}_streams_n{n}.log"))
        if s:
            out["config3_streams_sharded"].append({k: s[k] for k in KEEP_STREAMS if k in s})
        c = last_json(os.path.join(ROOT, "gpurun_out", f"{tag}_h2d_ceiling_n{n}.log"))
        if c:
            out["host_ingest_ceiling"].append(c)
            if s:
                out["config3_streams_sharded"][-1]["e2e_over_ceiling"] = s["e2e"]["value"] / c["h2d_plus_d2h"]["iq_MSps_ceiling"]
    path = os.path.join(ROOT, "profiles", f"{tag}_multi_gpu.json")
    json.dump(out, open(path, "w"), indent=1)
    for row in out["config5_wideband_broadcast"]:
        print("wideband N=%d: %.4f ms/step, e2e %.0f MS/s, PI %s" % (row["n_gpus"], row["ms_per_step"], row["e2e"]["value"], row["rds_check"]))
    for row in out["config3_streams_sharded"]:
        print("streams  N=%d: %.0f GS/s device, %.1f GS/s e2e, e2e/ceiling %s" % (row["n_gpus"], row["value"] / 1e3, row["e2e"]["value"] / 1e3,
                                                                               ("%.2f" % row["e2e_over_ceiling"]) if "e2e_over_ceiling" in row else "n/a"))
    print("wrote", path)


if __name__ == "__main__":
    main()
