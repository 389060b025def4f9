"""Host-side cost of enqueueing one block (kernel-parameter preparation + CUDA API calls), measured with the GPU
kept behind the host: wall time of N enqueue calls issued back to back right after a sync, for N small enough not to
fill the launch queue.  usage: python tools/host_overhead.py [n_streams]"""
import ctypes as C, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fm_radio_b200 as fm

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = 65536
cap = torch.randint(0, 255, (2, S, 2 * B), dtype=torch.uint8, device="cuda")
g = fm.FMDemod(B, S, pipeline_depth=4)
g.set_control(fm.Control.AUDIO_PCM_RATE_HZ, 48000)
for k in range(8):
    g.enqueue_u8_device(cap[k % 2])
g.sync()
L, h = g.L, g.h
p0, p1 = cap[0].data_ptr(), cap[1].data_ptr()
for n in (4, 8, 16):
    g.sync()
    t0 = time.perf_counter()
    for k in range(n):
        L.fmgpu_enqueue_u8_device(h, p0 if k & 1 else p1)
    t1 = time.perf_counter()
    g.sync()
    t2 = time.perf_counter()
    print(f"S={S} n={n}: host enqueue {1e6 * (t1 - t0) / n:.1f} us/block; wall incl. GPU {1e3 * (t2 - t0) / n:.4f} ms/block", flush=True)
# python-level call (api.py wrapper + tensor indexing), as bench.py issues it
g.sync()
t0 = time.perf_counter()
for k in range(8):
    g.enqueue_u8_device(cap[k % 2])
t1 = time.perf_counter()
g.sync()
print(f"S={S} via api.py + tensor indexing: host enqueue {1e6 * (t1 - t0) / 8:.1f} us/block", flush=True)
g.close()
