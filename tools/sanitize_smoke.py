"""Small shapes through EVERY production kernel, for `compute-sanitizer` (memcheck / synccheck) on the GPU box:

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py

Covers: the tcgen05 K1 and the FP32 K1 (u8), the cf32 K1, K2..K7 with keep_intermediates on and off, a ragged
stream count (3: an odd stream pair in K2 / K4), the stream-pipelined path and the CUDA-graph path (small blocks), the
device RDS rings, the audio PCM stage, both channelizer kernels (tensor pipeline with and without the general tile
factor, FP32) with a partial channel group, and the stand-alone polyphase / resampler entry points.
Sizes are tiny on purpose: the sanitizer runs kernels 10-100 x slower."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import fm_radio_b200 as fm
from fm_radio_b200 import Buf, ChanMode, Control, synth


def chain(B, S, n_blocks, keep, options=(), cf32=False):
    caps = np.stack([synth.synth_u8_numpy(B * n_blocks, synth.StreamParams.for_stream(s)) for s in range(S)])
    d = fm.FMDemod(B, S, device=0, keep_intermediates=keep, pipeline_depth=2)
    for name in options:
        d.set_option(name, 1)
    d.set_control(Control.AUDIO_PCM_RATE_HZ, 48000)
    for k in range(n_blocks):
        blk = caps[:, 2 * B * k:2 * B * (k + 1)]
        if cf32:
            x = (blk.astype(np.float32) - 127.0).reshape(S, B, 2)
            d.process_cf32(np.ascontiguousarray(x).view(np.complex64).reshape(S, B))
        else:
            d.process_u8(np.ascontiguousarray(blk))
        for s in range(S):
            assert np.isfinite(d.get(Buf.AUDIO_OUT, s)).all()
    d.rds_fetch()
    d.rds_db(0), d.rds_db_ext(S - 1), d.rds_counts(0)
    d.get(Buf.AUDIO_PCM_S16, 0)
    d.close()


def channelizer(mode, centres, B=2048, D=20, NN=192, n_blocks=3):
    rng = np.random.default_rng(1)
    ch = fm.Channelizer(synth.FS_WIDEBAND, centres, D, NN, B, mode=mode)
    for _ in range(n_blocks):
        ch.process_u8(rng.integers(0, 256, 2 * B * D, dtype=np.uint8))
    ch.close()


def main():
    chain(4096, 3, 6, keep=True)                                   # graph path, ragged pair, every debug buffer
    chain(4096, 2, 4, keep=False, options=("k1_fp32", "k3_exact", "k4_v1", "k5_literal"))
    chain(65536, 3, 3, keep=False)                                 # stream pipeline, full-size tiles
    chain(8192, 2, 3, keep=False, cf32=True)
    raster = synth.wideband_centres(100)[[0, 17, 50, 99]]
    channelizer(ChanMode.TENSOR, raster)                           # sign-only tile factor, partial group
    channelizer(ChanMode.TENSOR, np.array([-3.3333e6, 1.2345e6, 5.0e6 + 1.0]))    # general tile factor
    channelizer(ChanMode.TENSOR, synth.wideband_centres(40))       # two groups
    channelizer(ChanMode.FP32, raster)
    ds = fm.PolyphaseDownsampler(4, 16, True)
    ds.process((np.ones(4 * 256) + 0j).astype(np.complex64), 256)
    print("sanitize_smoke: done")


if __name__ == "__main__":
    main()
