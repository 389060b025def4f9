"""Development check (GPU box): channelizer, both kernels, against the float64 checker; then the
wideband chain (channelizer -> demodulators -> device RDS)."""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fm_radio_b200 as fm
from fm_radio_b200 import synth, ChanMode
from oracle import bind

L = bind.lib("port")._cdll
L.fmo_channelize_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
D, NN = 20, 192
n_st = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
nblk = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cent = synth.wideband_centres(n_st)
ps = [synth.StreamParams.for_stream(2000 + s) for s in range(n_st)]
dev = torch.device("cuda", 0)
t0 = time.time()
iq_d = synth.synth_wideband_u8(B * D * nblk, cent, ps, device=dev)
torch.cuda.synchronize()
iq = iq_d.cpu().numpy()
print("synth %.1fs" % (time.time() - t0), iq[:8], flush=True)

outs = {}
for mode in (ChanMode.FP32, ChanMode.TENSOR):
    ch = fm.Channelizer(synth.FS_WIDEBAND, cent, D, NN, B, mode=mode)
    hz, inc = ch.freqs()
    ys = [ch.process_u8(iq[2 * B * D * k:2 * B * D * (k + 1)]) for k in range(nblk)]
    outs[mode] = np.concatenate(ys, axis=1)
    print(mode.name, "launches", ch.launch_count, "rms", np.sqrt(np.mean(np.abs(outs[mode]) ** 2)), flush=True)
    b = ch.get_b().copy()
    ch.close()
sel = sorted(set([0, 1, n_st // 2 - 1, n_st // 2, n_st - 1, n_st // 3]))
ref = np.zeros((len(sel), B * nblk, 2), np.float64)
L.fmo_channelize_f64(iq.ctypes.data, B * D * nblk, None, 0, D, NN, b.ctypes.data, inc[sel].ctypes.data, len(sel), ref.ctypes.data)
refc = ref[..., 0] + 1j * ref[..., 1]
for mode, y in outs.items():
    for j, c in enumerate(sel):
        e = np.abs(y[c] - refc[j]).max()
        rms = np.sqrt(np.mean(np.abs(refc[j]) ** 2))
        print("%-6s ch %3d max|d| %.3e rms %.3f rel %.2e" % (mode.name, c, e, rms, e / rms), flush=True)
print("tensor vs fp32 max|d| %.3e" % np.abs(outs[ChanMode.TENSOR] - outs[ChanMode.FP32]).max())
