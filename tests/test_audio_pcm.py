"""Audio output stage: Resample() 32 kHz -> device rate
(audio/resampled_pcm_player.cpp:37-54), the scraper's int16 conversion (fm_scraper.cpp:74-78) and
PolyphaseUpsampler<T> (dsp/polyphase_filter.h:90-185).

CPU part: the C restatement against the golden fixture written by the unmodified reference
(tests/golden/make_golden_audio.py) and against the live reference where oracle/_ref exists.
GPU part (-m gpu): kernel K7 and the stand-alone entry points through the C-ABI against the restatement
(bit-exact: same operations, one rounding each) and the fixture.

Tolerances: int16 PCM bit-exact everywhere.  Resampled floats: the CUDA path and the restatement are
bit-identical; against the reference build (-ffast-math contracts f0*(1-k) + f1*k into one FMA) <= 1e-6 of
full scale.  Upsampler: <= 1e-5 of the RMS (the reference's AVX partial sums order the taps differently).
"""
import os

import numpy as np
import pytest

from oracle import bind
from tests import helpers as H

GOLD = np.load(os.path.join(H.GOLDEN, "golden_audio_pcm.npz"))


def _port_resample(x, n_out):
    y = np.zeros((n_out, 2), np.float32)
    bind.lib("port").resample_linear(np.ascontiguousarray(x).ctypes.data, x.shape[0], y.ctypes.data, n_out)
    return y


def _port_s16(x):
    s = np.zeros(x.shape, np.int16)
    bind.lib("port").frames_to_s16(np.ascontiguousarray(x).ctypes.data, x.shape[0], s.ctypes.data)
    return s


def test_restatement_resample_matches_reference_fixture():
    for i, (n_in, n_out) in enumerate(GOLD["cases"]):
        y = _port_resample(GOLD[f"rs{i}_in"], int(n_out))
        assert np.abs(y - GOLD[f"rs{i}_out"]).max() <= 1e-6, (i, n_in, n_out)
        # held last frame: the final outputs never read past the block (resampled_pcm_player.cpp:45-46)
        assert np.isfinite(y).all()


def test_restatement_s16_matches_reference_fixture_bit_exact():
    assert np.array_equal(_port_s16(GOLD["s16_in"]), GOLD["s16_out"])
    # the documented corner cases: truncation toward zero, wrap outside int16, 0 for NaN / inf / beyond int32
    s = _port_s16(np.array([[1.0, -1.0], [1.06, -1.06], [np.nan, np.inf], [1e12, 3.2e-5]], np.float32))
    assert s.tolist() == [[31128, -31128], [-32540, 32540], [0, 0], [0, 0]]


def test_restatement_upsampler_matches_reference_fixture():
    Lf, K, calls = int(GOLD["us_L"]), int(GOLD["us_K"]), int(GOLD["us_calls"])
    x = GOLD["us_in"]
    y = np.zeros(x.size * Lf, np.float32)
    b = np.ascontiguousarray(GOLD["us_b"])
    bind.lib("port").polyphase_us_f32(Lf, K, b.ctypes.data, x.ctypes.data, y.ctypes.data, x.size // calls, calls)
    assert np.abs(y - GOLD["us_out"]).max() <= 1e-5 * np.sqrt(np.mean(GOLD["us_out"] ** 2))


@pytest.mark.skipif(not bind.available("ref"), reason="oracle/_ref (the compiled reference) is not present")
def test_restatement_matches_live_reference():
    rng = np.random.default_rng(5)
    Lr = bind.lib("ref")
    for n_in, n_out in ((2048, 3072), (1000, 1378), (64, 96), (2048, 2048 * 6)):
        x = rng.uniform(-1.2, 1.2, (n_in, 2)).astype(np.float32)
        y = np.zeros((n_out, 2), np.float32)
        Lr.resample_linear(x.ctypes.data, n_in, y.ctypes.data, n_out)
        assert np.abs(_port_resample(x, n_out) - y).max() <= 1e-6
    x = (rng.standard_normal((5000, 2)) * 0.8).astype(np.float32)
    s = np.zeros(x.shape, np.int16)
    Lr.frames_to_s16(x.ctypes.data, x.shape[0], s.ctypes.data)
    assert np.array_equal(_port_s16(x), s)


def test_wav_container_matches_the_reference_layout(tmp_path):
    """fm_scraper.cpp:107-160: 44-byte PCM header (stereo, 16 bit), size fields rewritten after every write; readable
    by the standard library; reference_sizes=True reproduces the reference's frame-count-in-a-byte-field quirk."""
    import wave
    from fm_radio_b200.wav import WavWriter, wav_header
    assert len(wav_header(32000, 0)) == 44
    pcm = _port_s16(GOLD["s16_in"])
    path = tmp_path / "a.wav"
    with WavWriter(str(path), 48000) as w:
        w.write(pcm[:1000]); w.write(pcm[1000:])
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"
    assert int.from_bytes(raw[40:44], "little") == pcm.size * 2 and int.from_bytes(raw[4:8], "little") == 36 + pcm.size * 2
    assert int.from_bytes(raw[24:28], "little") == 48000 and int.from_bytes(raw[28:32], "little") == 48000 * 4
    assert raw[44:] == pcm.astype("<i2").tobytes()
    with wave.open(str(path)) as r:
        assert (r.getnchannels(), r.getsampwidth(), r.getframerate(), r.getnframes()) == (2, 2, 48000, pcm.shape[0])
    ref = tmp_path / "b.wav"
    with WavWriter(str(ref), 32000, reference_sizes=True) as w:
        w.write(pcm)
    raw = open(ref, "rb").read()
    assert int.from_bytes(raw[40:44], "little") == pcm.shape[0] and raw[44:] == pcm.astype("<i2").tobytes()


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_resample_linear_bit_exact_vs_restatement_and_fixture():
    import fm_radio_b200 as fm
    for i, (n_in, n_out) in enumerate(GOLD["cases"]):
        x = GOLD[f"rs{i}_in"]
        y = fm.resample_linear(x, int(n_out))
        assert np.array_equal(y, _port_resample(x, int(n_out))), i
        assert np.abs(y - GOLD[f"rs{i}_out"]).max() <= 1e-6
    rng = np.random.default_rng(11)
    for n_in, n_out in ((3000, 4500), (2048, 2048 * 6), (17, 5), (1, 1)):
        x = rng.uniform(-1.5, 1.5, (n_in, 2)).astype(np.float32)
        assert np.array_equal(fm.resample_linear(x, n_out), _port_resample(x, n_out)), (n_in, n_out)


@pytest.mark.gpu
def test_gpu_frames_to_s16_bit_exact():
    import fm_radio_b200 as fm
    assert np.array_equal(fm.frames_to_s16(GOLD["s16_in"]), GOLD["s16_out"])
    rng = np.random.default_rng(12)
    x = (rng.standard_normal((100000, 2)) * 0.9).astype(np.float32)
    assert np.array_equal(fm.frames_to_s16(x), _port_s16(x))


@pytest.mark.gpu
def test_gpu_polyphase_upsampler_matches_fixture_and_restatement():
    import fm_radio_b200 as fm
    Lf, K, calls = int(GOLD["us_L"]), int(GOLD["us_K"]), int(GOLD["us_calls"])
    x = GOLD["us_in"]
    n = x.size // calls
    up = fm.PolyphaseUpsampler(GOLD["us_b"], Lf, K, False)
    y = np.concatenate([up.process(x[c * n:(c + 1) * n]) for c in range(calls)])
    rms = np.sqrt(np.mean(GOLD["us_out"] ** 2))
    assert np.abs(y - GOLD["us_out"]).max() <= 1e-5 * rms
    # ragged calls (shorter than K, then longer) carry the K-sample history like push_value / push_values
    up2 = fm.PolyphaseUpsampler(GOLD["us_b"], Lf, K, False)
    cuts = [0, 5, 6, 40, 41, 700, x.size]
    y2 = np.concatenate([up2.process(x[a:b]) for a, b in zip(cuts[:-1], cuts[1:])])
    assert np.abs(y2 - y).max() <= 1e-6 * rms
    # complex input = the same filter on both components
    upc = fm.PolyphaseUpsampler(GOLD["us_b"], Lf, K, True)
    xc = (x[:600] + 1j * x[600:1200]).astype(np.complex64)
    yc = upc.process(xc)
    upr, upi = fm.PolyphaseUpsampler(GOLD["us_b"], Lf, K, False), fm.PolyphaseUpsampler(GOLD["us_b"], Lf, K, False)
    assert np.array_equal(yc.real, upr.process(x[:600])) and np.array_equal(yc.imag, upi.process(x[600:1200]))


@pytest.mark.gpu
@pytest.mark.parametrize("rate", [48000, 44100])
def test_gpu_chain_audio_output_stage(rate):
    """K7 inside the chain: every block's GetAudioOut resampled + converted exactly as the restatement does it
    on the same frames; the stage does not perturb the 32 kHz audio or the RDS symbols."""
    import fm_radio_b200 as fm
    from fm_radio_b200 import Buf, Control
    iq = H.capture("seed0")
    S = 3
    a, b = fm.FMDemod(H.B, S), fm.FMDemod(H.B, S)
    a.set_control(Control.AUDIO_PCM_RATE_HZ, rate)
    with pytest.raises(fm.FMGPUError):
        b.get(Buf.AUDIO_PCM_S16)                      # nothing processed / stage off
    M = int(np.float32(rate) / np.float32(32000.0) * np.float32(H.B // 32))
    for k in range(4):
        blk = np.stack([iq[2 * H.B * (k + s):2 * H.B * (k + s + 1)] for s in range(S)])
        a.process_u8(blk); b.process_u8(blk)
        for s in range(S):
            audio = a.get(Buf.AUDIO_OUT, s).reshape(-1, 2)
            assert np.array_equal(audio, b.get(Buf.AUDIO_OUT, s).reshape(-1, 2))
            assert np.array_equal(a.get(Buf.RDS_PRED_SYM, s), b.get(Buf.RDS_PRED_SYM, s))
            pcm = a.get(Buf.AUDIO_PCM_F32, s).reshape(-1, 2)
            assert pcm.shape == (M, 2)
            want = _port_resample(audio, M)
            assert np.array_equal(pcm, want)
            assert np.array_equal(a.get(Buf.AUDIO_PCM_S16, s).reshape(-1, 2), _port_s16(want))
    with pytest.raises(fm.FMGPUError):
        b.get(Buf.AUDIO_PCM_F32)                      # stage off on this handle
    # switching the stage off again stops producing (launch count per block back to the base chain)
    n0 = a.launch_count
    a.set_control(Control.AUDIO_PCM_RATE_HZ, 0)
    a.process_u8(blk)
    n1 = a.launch_count
    b0 = b.launch_count
    b.process_u8(blk)
    assert n1 - n0 == b.launch_count - b0
    with pytest.raises(fm.FMGPUError):
        a.set_control(Control.AUDIO_PCM_RATE_HZ, 1000)
    a.close(); b.close()


@pytest.mark.gpu
def test_gpu_chain_audio_stage_pipelined_equals_synchronous():
    """The asynchronous path (device input, ring of slots, pinned int16 mirror) gives the same PCM bytes."""
    import torch
    import fm_radio_b200 as fm
    from fm_radio_b200 import Buf, Control
    iq = H.capture("seed0")
    S, nblk = 4, 6
    blocks = [np.stack([iq[2 * H.B * (k + s):2 * H.B * (k + s + 1)] for s in range(S)]) for k in range(nblk)]
    sync = fm.FMDemod(H.B, S)
    sync.set_control(Control.AUDIO_PCM_RATE_HZ, 48000)
    want = []
    for blk in blocks:
        sync.process_u8(blk)
        want.append([sync.get(Buf.AUDIO_PCM_S16, s) for s in range(S)])
    pipe = fm.FMDemod(H.B, S, pipeline_depth=3)
    pipe.set_control(Control.AUDIO_PCM_RATE_HZ, 48000)
    dev = [torch.from_numpy(blk).cuda() for blk in blocks]
    torch.cuda.synchronize()
    got = []
    for k in range(nblk):
        slot = pipe.enqueue_u8_device(dev[k])
        pipe.fetch_outputs(slot)
        pipe.sync()
        got.append([pipe.get(Buf.AUDIO_PCM_S16, s) for s in range(S)])
    for k in range(nblk):
        for s in range(S):
            assert np.array_equal(got[k][s], want[k][s]), (k, s)
    sync.close(); pipe.close()
