"""A few wideband steps (config 4: 100 stations from a 20.48 MS/s u8 capture) for profiling: run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel breakdown of the step, or under `ncu --set full -k regex:chan_`.
usage: python tools/chan_profile.py [tensor|fp32] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fm_radio_b200 import ChanMode, synth
from fm_radio_b200.batch import WidebandReceiver

mode = ChanMode.FP32 if (len(sys.argv) > 1 and sys.argv[1] == "fp32") else ChanMode.TENSOR
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n_st, B, D = 100, 65536, 20
dev = torch.device("cuda", 0)
cent = synth.wideband_centres(n_st)
ps = [synth.StreamParams.for_stream(2000 + s) for s in range(n_st)]
n_in = B * D
cap = synth.synth_wideband_u8(n_in * 4, cent, ps, device=dev)
rx = WidebandReceiver(synth.FS_WIDEBAND, cent, 0, 1, B, D, 192, mode=mode)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for k in range(2):
        rx.feed(cap[2 * n_in * (k % 4):2 * n_in * (k % 4 + 1)], side.cuda_stream)
    rx.demod.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        rx.feed(cap[2 * n_in * (k % 4):2 * n_in * (k % 4 + 1)], side.cuda_stream)
    rx.demod.sync()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"chan_profile mode {mode.name}: {dt / steps * 1e3:.4f} ms/step over {steps} steps", flush=True)
rx.close()
