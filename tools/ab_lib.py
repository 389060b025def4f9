"""A/B timing of two builds of libfmgpu.so on the same box: device-resident, pipelined, 1024 streams.
usage: python tools/ab_lib.py libA.so libB.so [steps]"""
import ctypes as C, sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fm_radio_b200 import synth
from fm_radio_b200.api import _Config

S, B, n_in = 1024, 65536, 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 48
dev = torch.device("cuda", 0)
params = [synth.StreamParams.for_stream(s) for s in range(S)]
cap = torch.empty((n_in, S, 2 * B), dtype=torch.uint8, device=dev)
for s0 in range(0, S, 64):
    piece = synth.synth_u8_torch(B * n_in, params[s0:s0 + 64], dev)
    cap[:, s0:s0 + 64] = piece.view(64, n_in, 2 * B).transpose(0, 1)
torch.cuda.synchronize()
ext = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    for path in sys.argv[1:3]:
        L = C.CDLL(os.path.abspath(path))
        vp = C.c_void_p
        L.fmgpu_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
        for f in ("fmgpu_destroy", "fmgpu_sync"): getattr(L, f).argtypes = [vp]
        for f in ("fmgpu_enqueue_u8_device", "fmgpu_wait_external_stream", "fmgpu_signal_external_stream"): getattr(L, f).argtypes = [vp, vp]
        h = vp()
        assert L.fmgpu_create(C.byref(_Config(B, S, 0, 0, 4)), C.byref(h)) == 0
        L.fmgpu_set_control.argtypes = [vp, C.c_int, C.c_double]
        if not os.environ.get('AB_NO_PCM'): L.fmgpu_set_control(h, 6, 48000.0)     # audio output stage on, as bench.py
        L.fmgpu_wait_external_stream(h, ext)
        for k in range(int(os.environ.get('AB_WARM', '6'))): L.fmgpu_enqueue_u8_device(h, cap[k % n_in].data_ptr())
        L.fmgpu_sync(h); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); L.fmgpu_wait_external_stream(h, ext)
        for k in range(steps): L.fmgpu_enqueue_u8_device(h, cap[(6 + k) % n_in].data_ptr())
        L.fmgpu_signal_external_stream(h, ext); e1.record()
        L.fmgpu_sync(h); torch.cuda.synchronize()
        ms = (C.c_float * 7)()
        L.fmgpu_profile_stages7.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_float * 7)]
        L.fmgpu_profile_stages7(h, cap[0].data_ptr(), 8, C.byref(ms))
        print(f"{os.path.basename(path):24s} {e0.elapsed_time(e1) / steps:.4f} ms/step   serial k1..k7 ms: "
              + " ".join(f"{x:.4f}" for x in ms), flush=True)
        L.fmgpu_destroy(h)
