"""Why does bench.py time 0.30 ms per step where tools/ab_lib.py times 0.28?  Variants of the timed loop."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
variant = sys.argv[1]
import torch
if "dist" in variant:
    import torch.distributed as dist  # noqa: F401
if "setdev" in variant:
    torch.cuda.set_device(0)
import fm_radio_b200 as fm
from fm_radio_b200 import synth

S, B, n_in, steps = 1024, 65536, 4, 240
dev = torch.device("cuda", 0)
params = [synth.StreamParams.for_stream(s) for s in range(S)]
cap = torch.empty((n_in, S, 2 * B), dtype=torch.uint8, device=dev)
for s0 in range(0, S, 64):
    piece = synth.synth_u8_torch(B * n_in, params[s0:s0 + 64], dev)
    cap[:, s0:s0 + 64] = piece.view(64, n_in, 2 * B).transpose(0, 1)
torch.cuda.synchronize()
if "throwaway" in variant:
    for _ in range(int(os.environ.get('N_THROW', '1'))):
        t = fm.FMDemod(B, 2 if 'small' in variant else S, device=0, pipeline_depth=4)
        if 'norun' not in variant: t.enqueue_u8_device(cap[0][:2].contiguous() if 'small' in variant else cap[0]); t.sync()
        t.close()
g = fm.FMDemod(B, S, device=0, pipeline_depth=4)
if "part" in variant: print('partition', g.partition())
g.set_control(fm.Control.AUDIO_PCM_RATE_HZ, 48000)
ext = torch.cuda.current_stream().cuda_stream
g.wait_external_stream(ext)
if "ramp" in variant:
    t_w = time.perf_counter(); n = 0
    while (time.perf_counter() - t_w) < 0.5:
        for k in range(8):
            g.enqueue_u8_device(cap[(n + k) % n_in])
        n += 8
        g.sync()
for k in range(6):
    g.enqueue_u8_device(cap[k % n_in])
g.sync(); torch.cuda.synchronize()
if "launchcount" in variant:
    _ = g.launch_count
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.wait_external_stream(ext)
for k in range(steps):
    g.enqueue_u8_device(cap[(6 + k) % n_in])
g.signal_external_stream(ext)
e1.record()
g.sync(); torch.cuda.synchronize()
print(f"{variant:28s} {e0.elapsed_time(e1) / steps:.4f} ms/step", flush=True)
g.close()
