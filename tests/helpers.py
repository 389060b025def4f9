"""Shared helpers of the test-suite (test infrastructure: may import oracle/)."""
from __future__ import annotations

import functools
import hashlib
import os

import numpy as np

from fm_radio_b200 import synth
from oracle import bind

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
B = 65536
LOCK_BLOCK = 48          # "after lock": blocks >= 48 at B = 65536, i.e. t >= 3.07 s

REF_TAP_NAMES = ("fm_in", "fm_out", "hilbert", "audio_lpr", "audio_lmr", "rds", "deemphasis", "peak_pilot",
                 "pll_lpf", "bpsk_ted_lpf", "bpsk_pll_lpf")
IIR_NAMES = ("deemphasis", "peak_pilot", "pll_lpf", "bpsk_ted_lpf", "bpsk_pll_lpf")


@functools.lru_cache(maxsize=4)
def capture(tag: str = "seed0", n_blocks: int = 70, block_size: int = B) -> np.ndarray:
    p = {"seed0": synth.StreamParams(), "stream7": synth.StreamParams.for_stream(7)}[tag]
    return synth.synth_u8_numpy(block_size * n_blocks, p)


def golden(tag: str):
    return np.load(os.path.join(GOLDEN, f"golden_{tag}.npz"))


def sha256(a: np.ndarray) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), np.uint8)


def pick(x):
    return np.concatenate([x[:32], x[32::61]])


def cpu_checker_kinds():
    kinds = ["port"]
    if bind.available("ref"):
        kinds.append("ref")
    return kinds


def rel_rms(err: np.ndarray, ref: np.ndarray) -> float:
    return float(np.sqrt(np.mean(np.abs(err) ** 2)) / max(np.sqrt(np.mean(np.abs(ref) ** 2)), 1e-30))


def snr_db(test: np.ndarray, ref: np.ndarray) -> float:
    return -20.0 * np.log10(max(rel_rms(test - ref, ref), 1e-30))


def wrap_turn_diff(a, b):
    d = np.abs(a - b)
    return np.minimum(d, np.abs(1.0 - d))


def copy_taps(src: "bind.CpuDemod", dst_set):
    """Copy every filter of a CPU checker (after its first Process) through dst_set(name, b, a)."""
    for name in REF_TAP_NAMES:
        b, a = src.taps(name)
        dst_set(name, b, a[:len(b)] if name in IIR_NAMES else None)


def oracle_stream_job(args):
    """Process-pool worker (spawned, no CUDA): synthesise stream `s` of config 3 (for_stream(s)), run a CPU checker
    over it and return (capture bytes, groups, rds bytes, db).  kind = "port" | "ref"."""
    s, n_blocks, block_size, kind = args
    cap = synth.synth_u8_numpy(block_size * n_blocks, synth.StreamParams.for_stream(s))
    chk = bind.CpuDemod(block_size, kind)
    for k in range(n_blocks):
        chk.process_u8(cap[2 * block_size * k:2 * block_size * (k + 1)])
    return cap, chk.groups(), chk.rds_bytes(), chk.db()
