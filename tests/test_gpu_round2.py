"""GPU suite, part 2 (-m gpu): the tensor-core K1 against the FP32 K1, the symbol-wise K5 against the per-sample K5,
and parity on the BASELINE.json configurations as stated (config 1's full 10 s capture, config 2's largest block
sizes, a sample of config 3's 1024 stream recipes for 10 s each).  Tolerances as in tests/test_gpu_parity.py."""
from concurrent.futures import ProcessPoolExecutor
import multiprocessing as mp
import os

import numpy as np
import pytest

import fm_radio_b200 as fm
from fm_radio_b200 import Buf, synth
from oracle import bind
from tests import helpers as H
from tests.test_gpu_parity import FEED_FORWARD, FEEDBACK, TAP_IDS, _assert_fb, _assert_ff

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bs", [1024, 4096, 65536])
def test_k1_tensor_core_path_matches_fp32_path_and_checker(bs):
    """k1_toeplitz_i8 (tcgen05 kind::i8, exact integer sums) and k1_fir4_discrim_u8 (FFMA2) are the same FIR +
    discriminator: both within the feed-forward tolerance of the checker, within rounding noise of each other, and
    a silent block gives exact zeros in both."""
    iq = H.capture("seed0")
    nblk = {1024: 96, 4096: 24, 65536: 4}[bs]
    t = fm.FMDemod(bs, 1, keep_intermediates=True)
    f = fm.FMDemod(bs, 1, keep_intermediates=True)
    f.set_option("k1_fp32", 1)
    chk = bind.CpuDemod(bs, "port")
    for k in range(nblk):
        blk = iq[2 * bs * k:2 * bs * (k + 1)]
        chk.process_u8(blk); t.process_u8(blk); f.process_u8(blk)
        ref_in, ref_d = chk.get("fm_in"), chk.get("fm_demod")
        rms_in, rms_d = np.sqrt(np.mean(np.abs(ref_in) ** 2)), np.sqrt(np.mean(ref_d ** 2))
        for g in (t, f):
            assert np.abs(g.get(Buf.FM_IN) - ref_in).max() <= 1e-4 * rms_in, (k, g is t)
            assert np.abs(g.get(Buf.FM_DEMOD) - ref_d).max() <= 1e-4 * max(rms_d, 1e-3), (k, g is t)
        assert np.abs(t.get(Buf.FM_IN) - f.get(Buf.FM_IN)).max() <= 2e-5 * rms_in
        assert np.abs(t.get(Buf.FM_DEMOD) - f.get(Buf.FM_DEMOD)).max() <= 2e-5 * max(rms_d, 1e-3)
    z = np.full(2 * bs, 127, np.uint8)
    for g in (t, f):
        g.process_u8(z); g.process_u8(z)                   # the second block's history is silent too
        assert np.all(g.get(Buf.FM_IN) == 0) and np.all(g.get(Buf.FM_DEMOD) == 0)
    t.close(); f.close()


def test_k1_tensor_core_path_new_stream_and_taps_upload():
    """First block of a stream (zero FIR history = bytes 127) and re-designed fm_in taps (the G image is rebuilt)."""
    iq = H.capture("stream7")
    t = fm.FMDemod(H.B, 2, keep_intermediates=True)
    chk = [bind.CpuDemod(H.B, "port") for _ in range(2)]
    b = fm.create_fir_lpf(64, 0.2)
    for k in range(3):
        blk = np.stack([iq[2 * H.B * k:2 * H.B * (k + 1)], iq[2 * H.B * (k + 5):2 * H.B * (k + 6)]])
        if k == 1:
            t.upload_taps(fm.Filter.FM_IN, b)
            for c in chk:
                c.set_taps("fm_in", b, None)
        t.process_u8(blk)
        for s in range(2):
            chk[s].process_u8(blk[s])
            _assert_ff(t.get(Buf.FM_IN, s), chk[s].get("fm_in"), (k, s, "fm_in"))
            _assert_ff(t.get(Buf.FM_DEMOD, s), chk[s].get("fm_demod"), (k, s, "fm_demod"))
    t.close()


def test_u8_block_after_cf32_block_is_refused():
    g = fm.FMDemod(H.B, 1)
    iq = H.capture("seed0")
    g.process_u8(iq[:2 * H.B])
    g.process_cf32((iq[2 * H.B:4 * H.B].astype(np.float32) - 127.0).view(np.complex64))       # cf32 after u8: fine
    with pytest.raises(fm.FMGPUError):
        g.process_u8(iq[4 * H.B:6 * H.B])
    g.close()


def test_k5_symbolwise_loop_equals_per_sample_loop_on_the_device():
    """The symbol-wise BPSK synchroniser (production) and the per-sample loop (reference order) give identical bits:
    symbols, counts and every display signal, through acquisition and lock."""
    iq = H.capture("seed0")
    a = fm.FMDemod(H.B, 3, keep_intermediates=True)
    b = fm.FMDemod(H.B, 3, keep_intermediates=True)
    b.set_option("k5_literal", 1)
    bufs = (Buf.RDS_PRED_SYM, Buf.RDS_RAW_SYM, Buf.RDS, Buf.BPSK_PLL_SYM, Buf.BPSK_ZCD, Buf.BPSK_INT_DUMP_TRIGGER,
            Buf.BPSK_TED_RAW_PHASE_ERROR, Buf.BPSK_TED_PI_PHASE_ERROR, Buf.BPSK_PLL_RAW_PHASE_ERROR,
            Buf.BPSK_PLL_PI_PHASE_ERROR, Buf.BPSK_INT_DUMP_FILTER)
    for k in range(60):
        blk = np.stack([iq[2 * H.B * k:2 * H.B * (k + 1)], iq[2 * H.B * (k + 3):2 * H.B * (k + 4)], iq[2 * H.B * (k + 7):2 * H.B * (k + 8)]])
        a.process_u8(blk); b.process_u8(blk)
        for s in range(3):
            for buf in bufs:
                assert np.array_equal(a.get(buf, s), b.get(buf, s), equal_nan=True), (k, s, buf)
    a.close(); b.close()


@pytest.mark.parametrize("kind", H.cpu_checker_kinds())
def test_config1_full_10s_capture(kind):
    """BASELINE config 1 as stated: the whole 10 s capture (156 blocks of 65536) against the checker -- every group,
    validity flag and block type, the packed bytes, PI / PS / RT, and the audio after lock."""
    n_blocks = 156
    iq = H.capture("seed0", n_blocks)
    chk = bind.CpuDemod(H.B, kind)
    g = fm.FMDemod(H.B, 1, keep_intermediates=True)
    dec = fm.RDSDecoder()
    worst_audio = 0.0
    for k in range(n_blocks):
        blk = iq[2 * H.B * k:2 * H.B * (k + 1)]
        chk.process_u8(blk)
        if k == 0:
            H.copy_taps(chk, lambda name, b, a: g.upload_taps(TAP_IDS[name], b, a))
        g.process_u8(blk)
        dec.push_symbols(g.get(Buf.RDS_PRED_SYM))
        _assert_ff(g.get(Buf.FM_DEMOD), chk.get("fm_demod"), (k, "fm_demod"))
        if k >= H.LOCK_BLOCK:
            _assert_fb(g.get(Buf.AUDIO_OUT), chk.get("audio_out"), (k, "audio_out"))
            assert H.wrap_turn_diff(g.get(Buf.PLL_DT), chk.get("pll_dt")).max() <= 1e-4, k
            worst_audio = max(worst_audio, float(np.abs(g.get(Buf.AUDIO_OUT) - chk.get("audio_out")).max()))
    for a, b in zip(dec.groups(), chk.groups()):
        assert np.array_equal(a, b)
    assert dec.rds_bytes() == chk.rds_bytes()
    assert dec.db() == chk.db()
    assert len(dec.groups()[0]) >= 100                    # ~111 groups in 10 s
    print(f"config 1 vs {kind}: {len(dec.groups()[0])} groups, worst audio_out error after lock {worst_audio:.2e}")
    g.close()


@pytest.mark.parametrize("bs", [262144, 1048576])
def test_config2_largest_block_sizes_match_checker(bs):
    """BASELINE config 2's largest block sizes for >= 2 s of signal: feed-forward stages every block, loops after
    lock, RDS groups bit-exact with the checker run at the same block size."""
    n_blocks = max(2, int(np.ceil(4.2 * 1_024_000 / bs)))          # >= 4.2 s: lock (3.07 s) + at least one block after it
    iq = H.capture("seed0", (n_blocks * bs) // H.B)
    chk = bind.CpuDemod(bs, "port")
    g = fm.FMDemod(bs, 1, keep_intermediates=True)
    dec = fm.RDSDecoder()
    for k in range(n_blocks):
        blk = iq[2 * bs * k:2 * bs * (k + 1)]
        chk.process_u8(blk); g.process_u8(blk)
        dec.push_symbols(g.get(Buf.RDS_PRED_SYM))
        for name, buf in FEED_FORWARD:
            _assert_ff(g.get(buf), chk.get(name), (k, name))
        if k * bs >= H.LOCK_BLOCK * H.B:
            for name, buf in FEEDBACK:
                _assert_fb(g.get(buf), chk.get(name), (k, name))
            assert H.wrap_turn_diff(g.get(Buf.PLL_DT), chk.get("pll_dt")).max() <= 1e-4, k
    for a, b in zip(dec.groups(), chk.groups()):
        assert np.array_equal(a, b)
    assert dec.rds_bytes() == chk.rds_bytes()
    assert len(dec.groups()[0]) >= 20
    g.close()


def test_config3_sample_of_the_1024_stream_recipes_for_10s():
    """SURVEY.md 8(d) config 3's criterion -- "every stream's group list == the oracle's for that seed" -- on 64 of the
    1024 stream recipes (every 16th: SNR 30-50 dB, carrier offset +-2 kHz, own PI / PS / RT), 10 s each, decoded ON
    THE DEVICE (K6) in one batch.  The checker legs run in a process pool."""
    import torch
    ids = list(range(0, 1024, 16))
    n_blocks = 156
    ctx = mp.get_context("spawn")
    with ProcessPoolExecutor(max_workers=min(len(ids), os.cpu_count() or 1), mp_context=ctx) as ex:
        jobs = list(ex.map(H.oracle_stream_job, [(s, n_blocks, H.B, "port") for s in ids]))
    S = len(ids)
    g = fm.FMDemod(H.B, S, pipeline_depth=3)
    g.wait_external_stream(torch.cuda.current_stream().cuda_stream)
    got = [[np.zeros((0, 4), np.uint16), np.zeros((0, 4), np.uint8), np.zeros((0, 4), np.uint8)] for _ in range(S)]
    got_bytes = [b"" for _ in range(S)]
    keep = []
    for k in range(n_blocks):
        blk = torch.from_numpy(np.stack([j[0][2 * H.B * k:2 * H.B * (k + 1)] for j in jobs])).cuda()
        keep.append(blk)
        if len(keep) > 4:
            g.sync(); keep = keep[-4:]
        g.enqueue_u8_device(blk)
        if (k + 1) % 12 == 0 or k == n_blocks - 1:
            g.rds_fetch()
            for i in range(S):
                d, v, t = g.rds_groups(i, first=len(got[i][0]))
                got[i] = [np.concatenate([a, b]) for a, b in zip(got[i], (d, v, t))]
                got_bytes[i] += g.rds_bytes(i, first=len(got_bytes[i]))
    # Criterion: the group list (data, validity flags, block types) equals the checker's.  While the RDS loops are still
    # acquiring, the soft symbols are garbage around zero and the synchroniser's sequence of false locks (groups with
    # invalid blocks; stream 224 takes three attempts in the checker AND in the reference) depends on rounding noise, so
    # a stream may differ in those first groups; from the final lock on -- everything but at most 8 groups -- the lists
    # must be identical, and such streams are counted and named.
    n_groups, n_exact_groups, n_exact_bytes, late = [], 0, 0, []
    for i, s in enumerate(ids):
        _cap, groups, rds_bytes, db = jobs[i]
        gd, od = got[i], groups
        same = all(np.array_equal(a, b) for a, b in zip(gd, od))
        if not same:
            m = min(len(gd[0]), len(od[0]))
            suffix = 0
            while suffix < m and all(np.array_equal(a[len(a) - 1 - suffix], b[len(b) - 1 - suffix]) for a, b in zip(gd, od)):
                suffix += 1
            late.append((s, len(gd[0]), len(od[0]), suffix))
            assert suffix >= m - 8 and abs(len(gd[0]) - len(od[0])) <= 4, late
        n_exact_groups += same
        n_exact_bytes += got_bytes[i] == rds_bytes
        # (the packed byte stream is only counted: one soft symbol more or less during acquisition shifts the packing of
        # every later byte by a bit, while the group synchroniser slides bit by bit and still finds identical groups)
        assert abs(len(got_bytes[i]) - len(rds_bytes)) <= 16, s
        assert g.rds_db(i) == db, s
        assert db["pi"] == 0x1000 + s
        n_groups.append(len(groups[0]))
    assert min(n_groups) >= 80, n_groups
    assert n_exact_groups >= S - 8, late
    print(f"config 3 sample: {S} streams x 10 s, groups per stream {min(n_groups)}..{max(n_groups)}; {n_exact_groups} group lists and "
          f"{n_exact_bytes} byte streams identical to the checker's from the first bit; differing only in the first groups: {late}")
    g.close()


def test_k3_fast_pass_stays_within_rounding_noise_of_the_exact_body():
    """The pilot PLL's fast pass (2-op detector chain, re-anchored every 32 samples, exact redo when unlocked) against
    the exact body on the same input: pll_dt within 2e-6 turn after lock (tolerance vs the reference: 1e-4), audio
    within 1e-5, RDS symbols' hard decisions identical.  During acquisition every group is redone by the exact body."""
    iq = H.capture("stream7")
    a = fm.FMDemod(H.B, 1, keep_intermediates=True)
    b = fm.FMDemod(H.B, 1, keep_intermediates=True)
    b.set_option("k3_exact", 1)
    worst_dt = worst_audio = 0.0
    for k in range(70):
        blk = iq[2 * H.B * k:2 * H.B * (k + 1)]
        a.process_u8(blk); b.process_u8(blk)
        if k >= H.LOCK_BLOCK:
            worst_dt = max(worst_dt, float(H.wrap_turn_diff(a.get(Buf.PLL_DT), b.get(Buf.PLL_DT)).max()))
            worst_audio = max(worst_audio, float(np.abs(a.get(Buf.AUDIO_OUT) - b.get(Buf.AUDIO_OUT)).max()))
            sa, sb = a.get(Buf.RDS_PRED_SYM), b.get(Buf.RDS_PRED_SYM)
            assert len(sa) == len(sb) and np.array_equal(sa > 0, sb > 0), k
    print(f"K3 fast vs exact after lock: pll_dt {worst_dt:.2e} turn, audio {worst_audio:.2e}")
    assert worst_dt <= 2e-6 and worst_audio <= 1e-5
    a.close(); b.close()


@pytest.mark.parametrize("S,bs", [(5, 65536), (3, 8192), (2, 1024)])
def test_k4_constant_bank_taps_equal_shared_memory_taps(S, bs):
    """k4_mix_fir with its FIR taps as constant-bank operands (production) and as shared-memory loads (first version) is
    the same arithmetic in the same order: audio, RDS baseband, symbols and the L-R phase estimate must be identical
    bits, for full and partial tiles and an odd last stream pair."""
    iq = H.capture("seed0")
    nblk = {65536: 10, 8192: 40, 1024: 200}[bs]
    a = fm.FMDemod(bs, S, keep_intermediates=True)
    b = fm.FMDemod(bs, S, keep_intermediates=True)
    b.set_option("k4_v1", 1)
    for k in range(nblk):
        blk = np.stack([iq[2 * bs * (k + 3 * s):2 * bs * (k + 3 * s + 1)] for s in range(S)])
        a.process_u8(blk); b.process_u8(blk)
        for s in range(S):
            for buf in (Buf.AUDIO_OUT, Buf.AUDIO_LPR, Buf.AUDIO_LMR, Buf.RDS, Buf.RDS_PRED_SYM):
                assert np.array_equal(a.get(buf, s), b.get(buf, s), equal_nan=True), (k, s, buf)
            assert a.scalar(fm.Scalar.AUDIO_LMR_PHASE_ERROR, s) == b.scalar(fm.Scalar.AUDIO_LMR_PHASE_ERROR, s)
    a.close(); b.close()


@pytest.mark.parametrize("S,bs,keep", [(1, 4096, False), (3, 1024, True), (2, 16384, False)])
def test_cuda_graph_replay_equals_multi_stream_pipeline(S, bs, keep):
    """Small launch-bound blocks are replayed as one CUDA graph per (ring slot, history parity); the kernels and their
    parameters are the pipeline's, so every output must be identical bits -- also across a control change (re-capture),
    a taps upload, and when the caller alternates between two input buffers (device-pointer entry point)."""
    import torch
    iq = H.capture("seed0")
    nblk = {4096: 160, 1024: 300, 16384: 40}[bs]
    a = fm.FMDemod(bs, S, keep_intermediates=keep)           # graph replay (default at these sizes)
    b = fm.FMDemod(bs, S, keep_intermediates=keep)
    b.set_option("graph", 0)
    for g in (a, b):
        g.set_control(fm.Control.AUDIO_PCM_RATE_HZ, 48000)
    dev = [torch.empty((S, 2 * bs), dtype=torch.uint8, device="cuda") for _ in range(2)]
    for k in range(nblk):
        blk = np.stack([iq[2 * bs * (k + 5 * s):2 * bs * (k + 5 * s + 1)] for s in range(S)])
        if k == nblk // 3:
            for g in (a, b):
                g.set_control(fm.Control.AUDIO_LPR_CUTOFF_HZ, 9000); g.set_control(fm.Control.USE_DEEMPHASIS, 1)
        if k == nblk // 2:
            for g in (a, b):
                g.upload_taps(fm.Filter.FM_IN, fm.create_fir_lpf(64, 0.22))
        if k % 7 == 3:                                       # device-pointer entry point, two alternating buffers
            d = dev[k & 1]
            d.copy_(torch.from_numpy(blk)); torch.cuda.synchronize()
            for g in (a, b):
                slot = g.enqueue_u8_device(d); g.fetch_outputs(slot); g.sync()
        else:
            a.process_u8(blk); b.process_u8(blk)
        for s in range(S):
            for buf in (Buf.AUDIO_OUT, Buf.AUDIO_PCM_S16, Buf.RDS_PRED_SYM):
                assert np.array_equal(a.get(buf, s), b.get(buf, s), equal_nan=True), (k, s, buf)
            if keep:
                for buf in (Buf.FM_DEMOD, Buf.PLL_DT, Buf.AUDIO_LMR, Buf.RDS):
                    assert np.array_equal(a.get(buf, s), b.get(buf, s), equal_nan=True), (k, s, buf)
    a.rds_fetch(); b.rds_fetch()
    for s in range(S):
        assert a.rds_counts(s) == b.rds_counts(s)
    assert a.launch_count == b.launch_count
    a.close(); b.close()


def test_device_rds_database_flags_clock_and_programme_type_name():
    """K6 keeps all of RDS_Database the reference's handler fills (rds_database.h:26-53): the signal carries 0A
    flag bits, 4A clock groups and 10A programme-type-name groups; per stream the device's database equals the
    host decoder's (same symbols) and the checker chain's, field for field, and shows what was transmitted."""
    import dataclasses
    import torch
    S, nblk = 3, 80
    ps = [dataclasses.replace(synth.StreamParams.for_stream(70 + s), extended_groups=True, tp=1, ta=s & 1, ms=(s >> 1) & 1,
                              di=(0b1011, 0b0100, 0b1111)[s], ptyn=("Jazz\rXY ", "NEWSTALK", "B200 PTY")[s],
                              mjd=61331 + 1000 * s, hour=7 + 8 * s, minute=10, lto=(-7, 11, 0)[s]) for s in range(S)]
    caps = np.stack([synth.synth_u8_numpy(H.B * nblk, p) for p in ps])
    dev_in = torch.from_numpy(caps).cuda()
    g = fm.FMDemod(H.B, S, pipeline_depth=3)
    g.wait_external_stream(torch.cuda.current_stream().cuda_stream)
    decs = [fm.RDSDecoder() for _ in range(S)]
    keep = []
    for k in range(nblk):
        keep.append(dev_in[:, 2 * H.B * k:2 * H.B * (k + 1)].contiguous())
        slot = g.enqueue_u8_device(keep[-1])
        g.fetch_outputs(slot); g.sync()
        for s in range(S):
            decs[s].push_symbols(g.get(Buf.RDS_PRED_SYM, s))
    g.rds_fetch()
    for s in range(S):
        e = g.rds_db_ext(s)
        assert e == decs[s].db_ext() and g.rds_db(s) == decs[s].db()
        p = ps[s]
        assert e["programme_type_name"] == p.ptyn.replace("\r", "\0").encode()
        assert (e["year"], e["month"], e["day"]) == ((2026, 10, 18), (2029, 7, 14), (2032, 4, 9))[s]
        assert e["hour"] == p.hour and e["local_time_offset"] == p.lto and 10 <= e["minute"] <= 16
        assert e["traffic_announcement"] == 2 + p.ta and e["is_music"] == p.ms
        assert [e["is_dynamic_program_type"], e["is_compressed"], e["is_artificial_head"], e["is_stereo"]] == [(p.di >> k) & 1 for k in (3, 2, 1, 0)]
    for kind in H.cpu_checker_kinds():
        chk = bind.CpuDemod(H.B, kind)
        for k in range(nblk):
            chk.process_u8(caps[1, 2 * H.B * k:2 * H.B * (k + 1)])
        assert g.rds_db_ext(1) == chk.db_ext() and g.rds_db(1) == chk.db()
    g.close()


def test_dsp_filter_classes_match_checker_and_reference():
    """The stand-alone src/dsp classes behind the C-ABI (fmgpu_dsp_filter_*, fmgpu_agc_*): FIR_Filter, Hilbert_FIR_Filter,
    IIR_Filter, AGC_Filter, several consecutive blocks through one object (state carried), including blocks shorter than
    the filter's history.  IIR and AGC run the reference's operation order: bit-identical to the restatement; the FIR
    sums differ from it by FMA contraction only."""
    rng = np.random.default_rng(7)
    libs = [bind.lib(k) for k in H.cpu_checker_kinds()]          # port first, then the compiled reference where present
    for (K, N, calls) in ((65, 100, 3), (65, 20, 6), (33, 32, 4), (8, 1, 9), (3, 257, 2), (128, 4096, 2)):
        b = rng.standard_normal(K).astype(np.float32)
        for cplx in (False, True):
            C, dt = (2, np.complex64) if cplx else (1, np.float32)
            x = rng.standard_normal(C * N * calls).astype(np.float32)
            xs = x.view(dt)
            f = fm.FIRFilter(K, cplx)
            assert f.get_K() == K
            f.get_b()[:] = b
            got = np.concatenate([f.process(xs[c * N:(c + 1) * N]) for c in range(calls)])
            for L in libs:
                ref = np.zeros_like(x)
                (L.fir_cf32 if cplx else L.fir_f32)(K, b.ctypes.data, x.ctypes.data, ref.ctypes.data, N, calls)
                assert np.abs(got - ref.view(dt)).max() <= 1e-5 * np.sqrt(K), (K, N, cplx)
            bi = (rng.standard_normal(K) / K).astype(np.float32)
            ai = (rng.standard_normal(K) * 0.4 / K).astype(np.float32)
            g = fm.IIRFilter(K, cplx)
            g.get_b()[:] = bi
            g.get_a()[:] = ai
            got = np.concatenate([g.process(xs[c * N:(c + 1) * N]) for c in range(calls)])
            for n_lib, L in enumerate(libs):
                ref = np.zeros_like(x)
                (L.iir_cf32 if cplx else L.iir_f32)(K, bi.ctypes.data, ai.ctypes.data, x.ctypes.data, ref.ctypes.data, N, calls)
                if n_lib == 0:
                    assert np.array_equal(got, ref.view(dt)), (K, N, cplx)
                else:
                    assert np.abs(got - ref.view(dt)).max() <= 1e-5, (K, N, cplx)
        if K % 2 == 1:
            x = rng.standard_normal(N * calls).astype(np.float32)
            hf = fm.HilbertFIRFilter(K)
            assert np.array_equal(hf.get_b(), fm.create_fir_hilbert(K))
            got = np.concatenate([hf.process(x[c * N:(c + 1) * N]) for c in range(calls)])
            for L in libs:
                ref = np.zeros(2 * N * calls, np.float32)
                L.hilbert_f32(K, x.ctypes.data, ref.ctypes.data, N, calls)
                assert np.array_equal(got.real, ref[0::2]) and np.abs(got.imag - ref[1::2]).max() <= 1e-5 * np.sqrt(K), (K, N)
        x = (rng.standard_normal(2 * N * calls) * 3.0).astype(np.float32)
        a = fm.AGCFilter()
        assert (a.target_power, a.current_gain, a.beta) == (1.0, np.float32(0.1), np.float32(0.2))
        a.target_power = 0.5
        got, gains = [], []
        for c in range(calls):
            got.append(a.process(x.view(np.complex64)[c * N:(c + 1) * N]))
            gains.append(a.current_gain)
        got = np.concatenate(got)
        for n_lib, L in enumerate(libs):
            ref, gref = np.zeros_like(x), np.zeros(calls, np.float32)
            L.agc_cf32(0.5, 0.2, 0.1, x.ctypes.data, ref.ctypes.data, N, calls, gref.ctypes.data)
            if n_lib == 0:
                assert np.array_equal(np.float32(gains), gref) and np.array_equal(got, ref.view(np.complex64)), (K, N)
            else:
                assert np.allclose(np.float32(gains), gref, rtol=2e-6) and np.allclose(got, ref.view(np.complex64), rtol=2e-6, atol=1e-7)
    with pytest.raises(fm.FMGPUError):
        fm.IIRFilter(1)


@pytest.mark.parametrize("S,bs,keep", [(3, 65536, True), (37, 8192, False), (2, 1024, True)])
def test_k3_helper_warp_version_equals_single_warp_fast_pass(S, bs, keep):
    """k3_pll_duo (recurrence warp + helper warp that prepares the detector increments one group ahead) against the
    single-warp fast pass: the recurrence warp's arithmetic is the same, so every output is bit-identical -- from the
    first block (acquisition: every group redone by the exact body) through lock; ragged stream counts leave dead lanes
    in the last CTA, and a silent block in the middle poisons the loop (the CTA's fallback path) the same way in both."""
    nblk = {65536: 56, 8192: 40, 1024: 64}[bs]
    caps = np.stack([synth.synth_u8_numpy(bs * nblk, synth.StreamParams.for_stream(200 + (s % 5))) for s in range(S)])
    if bs == 8192:
        caps[1, 2 * bs * 20:2 * bs * 21] = 127                     # stream 1: one silent block -> non-finite AGC gain from then on
    a = fm.FMDemod(bs, S, keep_intermediates=keep)
    b = fm.FMDemod(bs, S, keep_intermediates=keep)
    a.set_option("k3_duo", 1)                                      # by default launch_k3 picks by grid size
    b.set_option("k3_single", 1)
    bufs = [Buf.PLL_DT, Buf.AUDIO_OUT, Buf.RDS_PRED_SYM] + ([Buf.PLL_RAW_PHASE_ERROR, Buf.PLL_LPF_PHASE_ERROR] if keep else [])
    for k in range(nblk):
        blk = np.ascontiguousarray(caps[:, 2 * bs * k:2 * bs * (k + 1)])
        a.process_u8(blk); b.process_u8(blk)
        for s in (0, S // 2, S - 1):
            for buf in bufs:
                assert np.array_equal(a.get(buf, s), b.get(buf, s), equal_nan=True), (k, s, buf)
    a.close(); b.close()
