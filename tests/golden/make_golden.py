"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libfmref.so).

Run in the build container (needs /root/reference to have been compiled by `make -C oracle ref`):
    python tests/golden/make_golden.py
The fixtures pin (a) the synthetic-capture generator, (b) the C restatement oracle/fm_oracle.c and
(c) the CUDA path to outputs of the real reference on boxes where the reference library is absent.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fm_radio_b200 import synth  # noqa: E402
from oracle import bind  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
B = 65536
N_BLOCKS = 70
SAMPLE_BLOCKS = (0, 1, 24, 50, 69)
BUFS = ("fm_demod", "fm_out_iq", "pilot", "pll_dt", "pll", "audio_lpr", "audio_lmr", "rds", "audio_out")


def pick(x):
    """first 32 samples + every 61st sample: a few hundred values per buffer"""
    return np.concatenate([x[:32], x[32::61]])


def make(seed_params, tag, block_size=B, n_blocks=N_BLOCKS):
    iq = synth.synth_u8_numpy(block_size * n_blocks, seed_params)
    ref = bind.CpuDemod(block_size, "ref")
    out = {"block_size": block_size, "n_blocks": n_blocks,
           "input_sha256": np.frombuffer(hashlib.sha256(iq.tobytes()).digest(), np.uint8),
           "input_head": iq[:256].copy()}
    counts, lmr_phase, agc_pilot, agc_rds, syms = [], [], [], [], []
    for k in range(n_blocks):
        ref.process_u8(iq[2 * block_size * k:2 * block_size * (k + 1)])
        s = ref.get("rds_pred_sym")
        counts.append(len(s)); syms.append(s)
        lmr_phase.append(ref.scalar("audio_lmr_phase_error"))
        agc_pilot.append(ref.scalar("agc_pilot_gain")); agc_rds.append(ref.scalar("agc_rds_gain"))
        if k in SAMPLE_BLOCKS:
            for name in BUFS:
                out[f"blk{k}_{name}"] = pick(ref.get(name))
    out["sym_counts"] = np.array(counts, np.int32)
    out["symbols"] = np.concatenate(syms).astype(np.float32)
    out["lmr_phase"] = np.array(lmr_phase, np.float32)
    out["agc_pilot"] = np.array(agc_pilot, np.float32)
    out["agc_rds"] = np.array(agc_rds, np.float32)
    d, v, t = ref.groups()
    out["groups_data"], out["groups_valid"], out["groups_type"] = d, v, t
    out["rds_bytes"] = np.frombuffer(ref.rds_bytes(), np.uint8)
    db = ref.db()
    out["db_pi"] = np.array([db["pi"]], np.uint16)
    out["db_pty"] = np.array([db["pty"]], np.uint8)
    out["db_ps"] = np.frombuffer(db["ps"], np.uint8)
    out["db_rt"] = np.frombuffer(db["rt"], np.uint8)
    for name in ("fm_in", "fm_out", "hilbert", "audio_lpr", "audio_lmr", "rds", "deemphasis", "peak_pilot", "pll_lpf",
                 "bpsk_ted_lpf", "bpsk_pll_lpf"):
        b, a = ref.taps(name)
        out[f"taps_{name}_b"] = b
        out[f"taps_{name}_a"] = a
    path = os.path.join(HERE, f"golden_{tag}.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", len(d), "groups; PI %04X" % db["pi"], db["ps"])


if __name__ == "__main__":
    make(synth.StreamParams(), "seed0")
    make(synth.StreamParams.for_stream(7), "stream7")
    make(synth.StreamParams(), "seed0_b4096", block_size=4096, n_blocks=16 * 40)
