"""Measurements of BASELINE.json configs 2, 4 and 5 (GPU box; `python tools/config_sweeps.py [--out F]`,
or under torchrun for the multi-rank wideband leg).  bench.py measures config 3 (the headline).

  config 2  one stream, block sizes 4096 .. 1048576, through the synchronous fmgpu_process_u8 call
            with HOST buffers (what Broadcast_FM_Demod::Process costs a caller): latency per block
            and x realtime.
  config 4  wideband 20.48 MS/s u8 capture, 100 stations at 200 kHz, channelizer (tensor and FP32
            kernels) -> 100 demodulators -> device RDS: wideband MS/s and x realtime, device-resident
            input, timed with CUDA events on the channelizer's stream.
  config 5  (WORLD_SIZE > 1) the same capture broadcast with NCCL from rank 0 every step, channels
            sharded c mod world_size; max over ranks.
Every number is timed after warm-up; nothing here runs under a profiler."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fm_radio_b200 as fm
from fm_radio_b200 import ChanMode, synth
from fm_radio_b200.batch import WidebandReceiver, gather_results

FS = 1_024_000.0


def config2(blocks=(4096, 16384, 65536, 262144, 1048576), seconds=2.0):
    rows = []
    for B in blocks:
        n_blk = max(8, int(seconds * FS / B))
        cap = synth.synth_u8_numpy(B * min(n_blk, 16), synth.StreamParams.for_stream(0))
        pin = torch.from_numpy(cap).pin_memory()
        d = fm.FMDemod(B, 1)
        views = [pin[2 * B * k:2 * B * (k + 1)] for k in range(min(n_blk, 16))]
        for k in range(4):
            d.process_u8(views[k % len(views)])
        lat = []
        for k in range(n_blk):
            t0 = time.perf_counter()
            d.process_u8(views[k % len(views)])
            lat.append(time.perf_counter() - t0)
        d.close()
        lat = np.array(lat) * 1e3
        rows.append({"block_size": B, "blocks": n_blk, "latency_ms_median": float(np.median(lat)),
                     "latency_ms_p99": float(np.percentile(lat, 99)),
                     "x_realtime": float(B / FS / (np.median(lat) * 1e-3))})
        print("config2", rows[-1], flush=True)
    return rows


def wideband(rank, world, mode, n_st=100, B=65536, steps=24, warmup=4, broadcast=False):
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device())
    D = 20
    cent = synth.wideband_centres(n_st)
    ps = [synth.StreamParams.for_stream(2000 + s) for s in range(n_st)]
    n_in = B * D
    n_cap = warmup + steps                                       # one continuous capture, 2.6 MB per block
    if rank == 0 or not broadcast:
        cap = synth.synth_wideband_u8(n_in * n_cap, cent, ps, device=dev)
    else:
        cap = torch.zeros(2 * n_in * n_cap, dtype=torch.uint8, device=dev)
    rx = WidebandReceiver(synth.FS_WIDEBAND, cent, rank, world, B, D, 192, mode=mode)
    blocks = [cap[2 * n_in * k:2 * n_in * (k + 1)] for k in range(n_cap)]
    stage = torch.empty(2 * n_in, dtype=torch.uint8, device=dev)

    def step(k):
        if broadcast and world > 1:
            if rank == 0:
                stage.copy_(blocks[k % n_cap])
            rx.broadcast_and_feed(stage)
            # the staging buffer may be overwritten only after the channelizer has read it
            torch.cuda.current_stream().wait_stream(torch.cuda.ExternalStream(rx.chan.stream))
        else:
            rx.feed(blocks[k % n_cap], torch.cuda.current_stream().cuda_stream)

    side = torch.cuda.Stream()            # a real (non-legacy) stream, so its handle is non-zero and the channelizer waits on it
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for k in range(warmup):
            step(k)
        rx.demod.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            step(warmup + k)
        rx.demod.sync()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    res = gather_results(rx.results())
    own_pi = sum(1 for (c, pi, _ps, _rt, _n) in res if pi == ps[c].pi_code)
    row = {"mode": ChanMode(rx.chan.mode).name, "world_size": world, "stations": n_st, "block_out": B, "steps": steps,
           "ms_per_step": dt / steps * 1e3, "wideband_MSps": n_in * steps / dt / 1e6,
           "x_realtime_wideband": n_in * steps / dt / synth.FS_WIDEBAND,
           "station_MSps_aggregate": n_st * B * steps / dt / 1e6,
           "signal_s": (warmup + steps) * B / FS, "stations_with_own_pi": own_pi,
           "broadcast_bytes_per_step": 2 * n_in if (broadcast and world > 1) else 0,
           "chan_launches": rx.chan.launch_count}
    rx.close()
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/config_sweeps.json")
    ap.add_argument("--skip-config2", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    out = {}
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        out["config5_wideband_broadcast"] = [wideband(rank, world, ChanMode.AUTO, broadcast=True)]
        dist.destroy_process_group()
    else:
        if not args.skip_config2:
            out["config2_single_stream_block_sweep"] = config2()
        out["config4_wideband_100_stations"] = [wideband(0, 1, m) for m in (ChanMode.TENSOR, ChanMode.FP32)]
    if rank == 0:
        out["gpu"] = torch.cuda.get_device_name()
        print(json.dumps(out, indent=1))
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
