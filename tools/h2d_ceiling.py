"""What can this host feed its GPUs?  N ranks (one per GPU, `torchrun --nproc-per-node N tools/h2d_ceiling.py`, or plain
`python` for N = 1) each stream pinned host -> device copies of the size bench.py's end-to-end path moves per step
(128 MiB of u8 IQ in, ~13 MB of PCM + symbols out, on two streams so the directions overlap as in the product) with
NO kernels running; prints aggregate GB/s -- the ceiling of any end-to-end number on this box -- next to the IQ MS/s it
corresponds to (2 bytes per IQ sample).  Timed after warm-up with a barrier on both sides, max over ranks."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import pynvml
    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception:
    pass
H2D, D2H = 128 << 20, 13_500_000
n_buf = 2
hin = [torch.empty(H2D, dtype=torch.uint8).pin_memory() for _ in range(n_buf)]
hout = [torch.empty(D2H, dtype=torch.uint8).pin_memory() for _ in range(n_buf)]
din = [torch.empty(H2D, dtype=torch.uint8, device="cuda") for _ in range(n_buf)]
dout = [torch.zeros(D2H, dtype=torch.uint8, device="cuda") for _ in range(n_buf)]
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def run(steps, both):
    for k in range(steps):
        with torch.cuda.stream(s_in):
            din[k % n_buf].copy_(hin[k % n_buf], non_blocking=True)
        if both:
            with torch.cuda.stream(s_out):
                hout[k % n_buf].copy_(dout[k % n_buf], non_blocking=True)
    torch.cuda.synchronize()


out = {}
for both in (False, True):
    run(8, both)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    steps = 64
    t0 = time.perf_counter()
    run(steps, both)
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    key = "h2d_plus_d2h" if both else "h2d_only"
    out[key] = {"h2d_GBps_aggregate": world * steps * H2D / dt / 1e9,
                "d2h_GBps_aggregate": (world * steps * D2H / dt / 1e9) if both else 0.0,
                "iq_MSps_ceiling": world * steps * (H2D // 2) / dt / 1e6, "ms_per_step": dt / steps * 1e3}
if rank == 0:
    print(json.dumps({"tool": "h2d_ceiling", "world_size": world, "cpus": os.cpu_count(), "h2d_bytes_per_step": H2D,
                      "d2h_bytes_per_step": D2H, **out}), flush=True)
if world > 1:
    dist.destroy_process_group()
