"""Generates tests/golden/golden_audio_pcm.npz from the UNMODIFIED reference (oracle/_ref/libfmref.so):
Resample() of audio/resampled_pcm_player.cpp:37-54, the int16 conversion of fm_scraper.cpp:74-78 and
PolyphaseUpsampler<float> (dsp/polyphase_filter.h:90-185) on seeded inputs.

    python tests/golden/make_golden_audio.py      (needs `make -C oracle ref`, i.e. /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bind  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def audio_like(n, seed):
    """two-tone stereo frames in [-1.2, 1.2] plus a little noise: the shape of GetAudioOut"""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / 32000.0
    l = 0.6 * np.sin(2 * np.pi * 1000 * t) + 0.3 * np.sin(2 * np.pi * 3500 * t + 0.3)
    r = 0.6 * np.sin(2 * np.pi * 1700 * t + 1.0) + 0.3 * np.sin(2 * np.pi * 5200 * t)
    return (np.stack([l, r], 1) + 0.01 * rng.standard_normal((n, 2))).astype(np.float32)


def main():
    L = bind.lib("ref")
    out = {}
    # (n_in, n_out): the chain's 2048 -> 3072 (32 k -> 48 k at B = 65536), 32 k -> 44.1 k, a down-sampling case, tiny blocks
    cases = [(2048, 3072), (2048, int((44100.0 / 32000.0) * 2048.0)), (2048, 1024), (32, 48), (128, 176)]
    out["cases"] = np.array(cases, np.int32)
    for i, (n_in, n_out) in enumerate(cases):
        x = audio_like(n_in, 100 + i)
        y = np.zeros((n_out, 2), np.float32)
        L.resample_linear(x.ctypes.data, n_in, y.ctypes.data, n_out)
        out[f"rs{i}_in"] = x
        out[f"rs{i}_out"] = y
    # int16 conversion: in-range audio, clipping-range values (wrap), non-finite
    x = audio_like(4096, 7) * 1.5
    x[:8] = np.array([[0.0, -0.0], [1.0, -1.0], [1.0526316, -1.0526316], [1.06, -1.06], [40.0, -40.0],
                      [1e12, -1e12], [np.nan, np.inf], [3.2e-5, -3.2e-5]], np.float32)
    s = np.zeros((4096, 2), np.int16)
    L.frames_to_s16(x.ctypes.data, 4096, s.ctypes.data)
    out["s16_in"], out["s16_out"] = x, s
    # PolyphaseUpsampler<float>: L = 3, K = 16 with a windowed-sinc prototype, three consecutive calls
    Lf, K, n_in, calls = 3, 16, 500, 3
    b = np.zeros(Lf * K, np.float32)
    L.create_fir_lpf(b.ctypes.data, Lf * K, 1.0 / Lf)
    x = audio_like(n_in * calls, 9)[:, 0].copy()
    y = np.zeros(n_in * calls * Lf, np.float32)
    L.polyphase_us_f32(Lf, K, b.ctypes.data, x.ctypes.data, y.ctypes.data, n_in, calls)
    out["us_L"], out["us_K"], out["us_calls"] = np.int32(Lf), np.int32(K), np.int32(calls)
    out["us_b"], out["us_in"], out["us_out"] = b, x, y
    np.savez_compressed(os.path.join(HERE, "golden_audio_pcm.npz"), **out)
    print("wrote golden_audio_pcm.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
