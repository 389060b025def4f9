"""CPU suite for the host side of the product: the C-ABI library loads and exports every symbol
include/fmgpu.h declares, the host filter designers and the host RDS bit path agree with the
checkers, and -- without a GPU -- compute entry points fail loudly instead of falling back."""
import ctypes as C
import dataclasses
import os
import re

import numpy as np
import pytest

import fm_radio_b200 as fm
from fm_radio_b200 import api, synth
from oracle import bind
from tests import helpers as H


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(H.ROOT, "include", "fmgpu.h")).read()
    declared = set(re.findall(r"\b(fmgpu_[A-Za-z0-9_]+)\s*\(", hdr))
    assert declared == set(api.EXPORTED_SYMBOLS), declared ^ set(api.EXPORTED_SYMBOLS)
    L = C.CDLL(api.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert b"sm_100a" in fm.lib().fmgpu_version()


def test_no_cpu_fallback(have_gpu):
    if have_gpu:
        pytest.skip("GPU present")
    with pytest.raises(fm.FMGPUError) as e:
        fm.FMDemod(65536, 1)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(fm.FMGPUError):
        fm.PolyphaseDownsampler(4, 16, True)


def test_product_does_not_import_oracle():
    pkg = os.path.join(H.ROOT, "fm_radio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "fm_oracle" not in src and "libfmref" not in src, f


@pytest.mark.parametrize("kind", H.cpu_checker_kinds())
def test_host_designers_match_checker(kind):
    Lc = bind.lib(kind)
    rng = np.random.default_rng(3)
    for N in (16, 64, 65, 128):
        for k in rng.uniform(0.02, 0.95, 4):
            ref = np.zeros(N, np.float32)
            Lc.create_fir_lpf(ref.ctypes.data, N, float(k))
            assert np.abs(fm.create_fir_lpf(N, float(k)) - ref).max() < 2e-7
            Lc.create_fir_hpf(ref.ctypes.data, N, float(k))
            assert np.abs(fm.create_fir_hpf(N, float(k)) - ref).max() < 2e-7
            Lc.create_fir_bpf(ref.ctypes.data, N, float(k) * 0.5, float(k))
            assert np.abs(fm.create_fir_bpf(N, float(k) * 0.5, float(k)) - ref).max() < 2e-7
        Lc.create_fir_hilbert(ref.ctypes.data, N)
        assert np.abs(fm.create_fir_hilbert(N) - ref).max() < 1e-7
    for k in (0.003, 0.1, 0.5, 0.9):
        rb, ra = np.zeros(2, np.float32), np.zeros(2, np.float32)
        Lc.create_iir_single_pole_lpf(rb.ctypes.data, ra.ctypes.data, k)
        b, a = fm.create_iir_single_pole_lpf(k)
        assert np.allclose(b, rb, rtol=2e-6) and np.allclose(a, ra, rtol=2e-6)
    for wid, wname in enumerate(fm.WINDOWS):               # the designers' window argument (filter_designer.h:9-11)
        for N, k in ((64, 0.2375), (128, 0.03125), (65, 0.6)):
            ref = np.zeros(N, np.float32)
            Lc.create_fir_lpf_window(ref.ctypes.data, N, k, wid)
            assert np.abs(fm.create_fir_lpf_window(N, k, wname) - ref).max() < 2e-7, wname
    assert np.array_equal(fm.create_fir_lpf_window(64, 0.3), fm.create_fir_lpf(64, 0.3))
    # method-2 peak filter: ONE parameter set per process on the reference side (its normalisation memoises the first
    # call's parameters in a static lambda, filter_designer.cpp:347)
    rb, ra = np.zeros(3, np.float32), np.zeros(3, np.float32)
    Lc.create_iir_peak_2_filter(rb.ctypes.data, ra.ctypes.data, 0.296875, 0.999, 20.0)
    b, a = fm.create_iir_peak_2_filter(0.296875, 0.999, 20.0)
    assert np.allclose(b, rb, rtol=1e-3) and np.allclose(a, ra, rtol=2e-6)
    assert (fm.api.TOTAL_TAPS_IIR_SINGLE_POLE_LPF, fm.api.TOTAL_TAPS_IIR_SECOND_ORDER_NOTCH_FILTER,
            fm.api.TOTAL_TAPS_IIR_SECOND_ORDER_PEAK_FILTER) == (2, 3, 3)
    if kind == "port":      # the reference's peak/notch designers memoise their first (k, r), see test_oracle.py
        for k, r in ((0.296875, 0.9999), (0.1, 0.99)):
            rb, ra = np.zeros(3, np.float32), np.zeros(3, np.float32)
            Lc.create_iir_peak_1_filter(rb.ctypes.data, ra.ctypes.data, k, r)
            b, a = fm.create_iir_peak_1_filter(k, r)
            assert np.allclose(b, rb, rtol=1e-3) and np.allclose(a, ra, rtol=2e-6)
            Lc.create_iir_notch_filter(rb.ctypes.data, ra.ctypes.data, k, r)
            b, a = fm.create_iir_notch_filter(k, r)
            assert np.allclose(b, rb, rtol=1e-3) and np.allclose(a, ra, rtol=2e-6)


def _symbols_from_bits(bits, rng, noise=0.2, flip=0.0):
    """Chips as the BPSK synchroniser emits them: two per differential bit, soft values."""
    d = np.bitwise_xor.accumulate(bits)
    lvl = d.astype(np.float32) * 2 - 1
    chips = np.empty(2 * len(bits), np.float32)
    chips[0::2] = lvl
    chips[1::2] = -lvl
    chips += rng.standard_normal(len(chips)).astype(np.float32) * noise
    if flip > 0:
        m = rng.random(len(chips)) < flip
        chips[m] *= -1
    return chips


@pytest.mark.parametrize("kind", H.cpu_checker_kinds())
@pytest.mark.parametrize("flip", [0.0, 0.004, 0.02])
def test_host_rds_decoder_matches_checker(kind, flip):
    rng = np.random.default_rng(11)
    p = synth.StreamParams.for_stream(3)
    bits = synth.rds_bits(p, 104 * 60)
    # junk before the first group exercises the block-A search; chip-phase offset exercises PushBit's skipping
    pre = rng.integers(0, 2, 37).astype(np.uint8)
    sym = _symbols_from_bits(np.concatenate([pre, bits]), rng, flip=flip)[1:]
    mine, ref = fm.RDSDecoder(), bind.CpuRds(kind)
    for chunk in np.array_split(sym, 23):          # ragged pushes, including across byte/packet boundaries
        mine.push_symbols(chunk)
        ref.push_symbols(chunk)
    mine.push_symbols(np.zeros(0, np.float32))     # empty input is a no-op
    for a, b in zip(mine.groups(), ref.groups()):
        assert np.array_equal(a, b)
    assert mine.rds_bytes() == ref.rds_bytes()
    assert mine.db() == ref.db()
    if flip == 0.0:
        d, v, _ = mine.groups()
        assert len(d) >= 57 and v.all()
        assert mine.db()["pi"] == p.pi_code and mine.db()["ps"] == p.ps.encode()


def _ext_params(seed: int) -> synth.StreamParams:
    return dataclasses.replace(synth.StreamParams.for_stream(seed), extended_groups=True, tp=1, ta=seed & 1, ms=(seed >> 1) & 1,
                               di=0b1011 if seed % 3 else 0b0100, ptyn=("Jazz\rXY " if seed % 2 else "NEWSTALK"),
                               mjd=61331 + 400 * seed, hour=7 + seed, minute=55, lto=(-7 if seed % 2 else 11))


@pytest.mark.parametrize("kind", H.cpu_checker_kinds())
@pytest.mark.parametrize("flip", [0.0, 0.004, 0.02])
def test_host_rds_database_groups_0a_4a_10a_match_checker(kind, flip):
    """TA/TP, M/S, DI, clock (4A) and programme type name (10A) of RDS_Database (rds_database.h:26-53):
    product host decoder == plain-C restatement == the compiled reference, after every ragged push."""
    rng = np.random.default_rng(5)
    for seed in (1, 2, 6):
        p = _ext_params(seed)
        bits = synth.rds_bits(p, 104 * 96)
        sym = _symbols_from_bits(np.concatenate([rng.integers(0, 2, 11).astype(np.uint8), bits]), rng, flip=flip)
        mine, ref = fm.RDSDecoder(), bind.CpuRds(kind)
        for chunk in np.array_split(sym, 17):
            mine.push_symbols(chunk)
            ref.push_symbols(chunk)
            assert mine.db_ext() == ref.db_ext()
            assert mine.db() == ref.db()
        if flip == 0.0:
            e = mine.db_ext()
            assert e["programme_type_name"] == p.ptyn.replace("\r", "\0").encode()
            assert (e["year"], e["month"], e["day"]) == _ymd(p.mjd) and e["hour"] == p.hour and e["local_time_offset"] == p.lto
            assert e["traffic_announcement"] == 2 * p.tp + p.ta and e["is_music"] == p.ms
            assert [e["is_dynamic_program_type"], e["is_compressed"], e["is_artificial_head"], e["is_stereo"]] == [(p.di >> k) & 1 for k in (3, 2, 1, 0)]


def _ymd(mjd: int):
    import datetime
    d = datetime.date(1858, 11, 17) + datetime.timedelta(days=mjd)
    return d.year, d.month, d.day


def test_rds_clock_date_conversion_over_the_17_bit_mjd_range():
    """modified_julian_date.h:8-24 as restated in rds_core.h (32-bit arithmetic): every 97th day of the 17-bit MJD
    range through a hand-built 4A group, against the calendar."""
    p = synth.StreamParams(extended_groups=True)
    rng = np.random.default_rng(0)
    for mjd in list(range(0, 1 << 17, 97 * 64)) + [0, 15078, 51544, 61331, (1 << 17) - 1]:
        q = dataclasses.replace(p, mjd=mjd, hour=23, minute=3, lto=0)
        words = synth.rds_group_words(q, 7)
        bits = np.concatenate([[(synth.rds_block(w, n) >> b) & 1 for b in range(25, -1, -1)] for w, n in zip(words, "ABCD")] * 4).astype(np.uint8)
        mine = fm.RDSDecoder()
        mine.push_symbols(_symbols_from_bits(bits, rng, flip=0.0))
        e = mine.db_ext()
        assert (e["year"], e["month"], e["day"]) == _ymd(mjd), mjd
        assert (e["hour"], e["minute"]) == (23, 3)


def test_rds_encoder_crc_known_answer():
    # EN 50067 annex B worked example style check: syndrome of every valid block is zero and a
    # single flipped bit is corrected back by the host decoder
    p = synth.StreamParams()
    for w, name in zip(synth.rds_group_words(p, 0), "ABCD"):
        blk = synth.rds_block(w, name)
        x = blk ^ synth.RDS_OFFSET[name]
        reg = 0
        for bit in range(25, -1, -1):
            reg = (reg << 1) | ((x >> bit) & 1)
            if reg & 0x400:
                reg ^= 0b0110111001
        assert reg & 0x3FF == 0


def test_stream_partition_helpers():
    from fm_radio_b200 import batch
    for n, w in ((8192, 8), (1024, 3), (5, 8), (1, 1)):
        parts = [batch.shard_streams(n, r, w) for r in range(w)]
        flat = [s for p in parts for s in p]
        assert sorted(flat) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
