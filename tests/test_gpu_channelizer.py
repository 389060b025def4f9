"""GPU suite (-m gpu) for the wideband front end (BASELINE config 4): the CUDA channelizer (tensor-core
and FP32 kernels), called through the C-ABI, against the float64 checker (oracle/fm_oracle.c
fmo_channelize_f64), and the wideband chain channelizer -> demodulators -> device RDS against the
checker chain (float64 channelizer -> CPU demodulator).

Tolerances: channel samples max-abs <= 1e-4 x RMS (feed-forward; measured 1e-6 tensor, 3e-6 FP32);
audio behind the loops after lock (blocks >= 48): max-abs <= 1e-4 or SNR >= 60 dB; RDS bit-exact."""
import ctypes as C

import numpy as np
import pytest

import fm_radio_b200 as fm
from fm_radio_b200 import Buf, ChanMode, synth
from fm_radio_b200.batch import WidebandReceiver
from oracle import bind
from tests import helpers as H

pytestmark = pytest.mark.gpu


def _checker(iq, n0, D, NN, b, inc, hist=None):
    L = bind.lib("port")._cdll
    L.fmo_channelize_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]
    n_in = iq.size // 2
    inc = np.ascontiguousarray(inc, np.uint32)
    out = np.zeros((len(inc), n_in // D, 2), np.float64)
    L.fmo_channelize_f64(iq.ctypes.data, n_in, None if hist is None else hist.ctypes.data, n0, D, NN,
                         np.ascontiguousarray(b, np.float32).ctypes.data, inc.ctypes.data, len(inc), out.ctypes.data)
    return out[..., 0] + 1j * out[..., 1]


def _wideband(n_samples, idx, seed0=2000):
    import torch
    cent = synth.wideband_centres(100)[idx]
    ps = [synth.StreamParams.for_stream(seed0 + s) for s in idx]
    iq = synth.synth_wideband_u8(n_samples, cent, ps, device=torch.device("cuda", 0))
    return cent, ps, iq


@pytest.mark.parametrize("mode", [ChanMode.TENSOR, ChanMode.FP32])
@pytest.mark.parametrize("shape", [(20, 192, 100), (16, 128, 37), (8, 64, 5)])
def test_channelizer_matches_float64_checker(mode, shape):
    D, NN, n_ch = shape
    B, nblk = 2048, 3
    rng = np.random.default_rng(D)
    # random bytes are the hardest input (full-scale, white): every tap and every digit plane matters
    iq = rng.integers(0, 256, 2 * B * D * nblk, dtype=np.uint8)
    cent = rng.uniform(-0.5, 0.5, n_ch) * synth.FS_WIDEBAND
    ch = fm.Channelizer(synth.FS_WIDEBAND, cent, D, NN, B, mode=mode)
    assert ch.mode == mode
    _, inc = ch.freqs()
    got = np.concatenate([ch.process_u8(iq[2 * B * D * k:2 * B * D * (k + 1)]) for k in range(nblk)], axis=1)
    sel = sorted(set([0, n_ch // 2, n_ch - 1, n_ch // 3]))
    ref = _checker(iq, 0, D, NN, ch.get_b().copy(), inc[sel])
    for j, c in enumerate(sel):
        rms = np.sqrt(np.mean(np.abs(ref[j]) ** 2))
        assert np.abs(got[c] - ref[j]).max() <= 1e-4 * rms, (mode, c, np.abs(got[c] - ref[j]).max(), rms)
    ch.close()


def test_tensor_and_fp32_kernels_agree_and_device_path_equals_host_path():
    import torch
    D, NN, B, nblk = 20, 192, 4096, 4
    idx = list(range(0, 100, 3))
    cent, _, iq_d = _wideband(B * D * nblk, idx)
    iq = iq_d.cpu().numpy()
    a = fm.Channelizer(synth.FS_WIDEBAND, cent, D, NN, B, mode=ChanMode.TENSOR)
    b = fm.Channelizer(synth.FS_WIDEBAND, cent, D, NN, B, mode=ChanMode.FP32)
    c = fm.Channelizer(synth.FS_WIDEBAND, cent, D, NN, B, mode=ChanMode.TENSOR)
    c.wait_external_stream(torch.cuda.current_stream().cuda_stream)
    for k in range(nblk):
        blk = iq[2 * B * D * k:2 * B * D * (k + 1)]
        ya, yb = a.process_u8(blk), b.process_u8(blk)
        assert np.abs(ya - yb).max() <= 1e-4 * np.sqrt(np.mean(np.abs(ya) ** 2))
        ptr = c.enqueue_u8_device(iq_d[2 * B * D * k:2 * B * D * (k + 1)])
        c.sync()
        yc = torch.empty((len(idx), B), dtype=torch.complex64, device="cuda")
        C.cdll.LoadLibrary("libcudart.so").cudaMemcpy(C.c_void_p(yc.data_ptr()), C.c_void_p(ptr), C.c_size_t(yc.numel() * 8), 3)
        assert np.array_equal(yc.cpu().numpy(), ya)             # same kernel, same bytes: identical bits
    for o in (a, b, c):
        o.close()


def test_channelizer_rejects_bad_arguments():
    with pytest.raises(fm.FMGPUError):
        fm.Channelizer(synth.FS_WIDEBAND, [0.0], 20, 192, 1000)                  # block_out not a multiple of 128
    with pytest.raises(fm.FMGPUError):
        fm.Channelizer(synth.FS_WIDEBAND, [0.0], 20, 100, 1024, mode=ChanMode.TENSOR)   # taps not a multiple of 64
    ch = fm.Channelizer(synth.FS_WIDEBAND, [0.0], 20, 100, 1024)                 # auto falls back to the FP32 kernel
    assert ch.mode == ChanMode.FP32
    with pytest.raises(fm.FMGPUError):
        ch.process_u8(np.zeros(100, np.uint8))                                   # wrong block length
    ch.close()


@pytest.mark.parametrize("kind", H.cpu_checker_kinds())
def test_wideband_chain_matches_checker_chain(kind):
    """8 stations in one capture, 3.4 s: channelizer -> 8 demodulators -> K6, against the float64 channelizer
    -> CPU demodulator for two of them (an edge channel and one next to 0 Hz)."""
    D, NN, B, nblk = 20, 192, H.B, 52
    idx = [0, 13, 37, 49, 50, 64, 88, 99]
    cent, ps, iq_d = _wideband(B * D * nblk, idx, seed0=3000)
    rx = WidebandReceiver(synth.FS_WIDEBAND, cent, block_out=B, decimation=D, n_taps=NN, pipeline_depth=1)
    _, inc = rx.chan.freqs()
    proto = rx.chan.get_b().copy()
    sel = [0, 4]
    chk = [bind.CpuDemod(B, kind) for _ in sel]
    hist = None
    worst = 0.0
    for k in range(nblk):
        blk_d = iq_d[2 * B * D * k:2 * B * D * (k + 1)]
        slot = rx.feed(blk_d, None if k else __import__("torch").cuda.current_stream().cuda_stream)
        rx.demod.fetch_outputs(slot)
        rx.demod.sync()
        blk = blk_d.cpu().numpy()
        ref = _checker(blk, k * B * D, D, NN, proto, inc[sel], hist)
        hist = blk[-2 * NN:].copy()
        for j, c in enumerate(sel):
            chk[j].process_cf32(ref[j].astype(np.complex64))
            if k >= H.LOCK_BLOCK:
                got, want = rx.demod.get(Buf.AUDIO_OUT, c), chk[j].get("audio_out")
                d = np.abs(got - want).max()
                worst = max(worst, d)
                assert d <= 1e-4 or H.snr_db(got, want) >= 60.0, (k, c, d, H.snr_db(got, want))
    res = rx.results()
    for (cid, pi, psn, rt, n_groups), p in zip(res, ps):
        assert pi == p.pi_code and psn == p.ps.encode(), (cid, hex(pi), psn)
        assert n_groups >= 30
    for j, c in enumerate(sel):
        for a, b in zip(rx.demod.rds_groups(c), chk[j].groups()):
            assert np.array_equal(a, b)                    # bit-exact group sequence, validity flags, block types
        assert rx.demod.rds_bytes(c) == chk[j].rds_bytes()
        assert rx.demod.rds_db(c) == chk[j].db()
    rx.close()


def test_config4_hundred_stations_each_decode_their_own_pi():
    """BASELINE config 4 at full size: 100 stations on the 200 kHz raster in one 20.48 MS/s capture, 1.5 s,
    blocks pipelined through channelizer and demodulators; every station's PI / PS must come out."""
    import torch
    D, NN, B, nblk = 20, 192, H.B, 24
    idx = list(range(100))
    cent, ps, iq_d = _wideband(B * D * nblk, idx, seed0=4000)
    rx = WidebandReceiver(synth.FS_WIDEBAND, cent, block_out=B, decimation=D, n_taps=NN)
    assert rx.chan.mode == ChanMode.TENSOR
    ext = torch.cuda.current_stream().cuda_stream
    for k in range(nblk):
        rx.feed(iq_d[2 * B * D * k:2 * B * D * (k + 1)], ext if k == 0 else None)
    res = rx.results()
    bad = [(cid, hex(pi)) for (cid, pi, psn, rt, n), p in zip(res, ps) if pi != p.pi_code or psn != p.ps.encode()]
    assert not bad, bad
    assert rx.chan.launch_count == nblk and rx.demod.launch_count >= 7 * nblk
    rx.close()
