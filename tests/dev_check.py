"""Test infrastructure (imports oracle/).  Development check run on the GPU box (python tests/dev_check.py): S=1 keep_intermediates vs the reference, stage by stage."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fm_radio_b200 as fm
from fm_radio_b200 import Buf, Filter, Scalar
from oracle import bind

B = 65536
nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 64
t0 = time.time()
iq = fm.synth.synth_u8_numpy(B * nblk)
print("synth %.1fs" % (time.time() - t0), flush=True)
ref = bind.CpuDemod(B, "ref")
g = fm.FMDemod(B, 1, keep_intermediates=True)
ref.process_u8(iq[:2 * B])   # settle UpdateFilters so taps exist
ref = bind.CpuDemod(B, "ref")
tmp = bind.CpuDemod(B, "ref"); tmp.process_u8(iq[:2 * B])
names = {Filter.FM_IN: "fm_in", Filter.FM_OUT: "fm_out", Filter.HILBERT: "hilbert", Filter.AUDIO_LPR: "audio_lpr",
         Filter.AUDIO_LMR: "audio_lmr", Filter.RDS: "rds", Filter.DEEMPHASIS: "deemphasis", Filter.PEAK_PILOT: "peak_pilot",
         Filter.PLL_LPF: "pll_lpf", Filter.BPSK_TED_LPF: "bpsk_ted_lpf", Filter.BPSK_PLL_LPF: "bpsk_pll_lpf"}
for f, n in names.items():
    br, ar = tmp.taps(n)
    bg, ag = g.download_taps(f)
    d = np.abs(br - bg).max()
    da = 0 if ag is None else np.abs(ar[:len(ag)] - ag).max()
    print("taps %-12s max diff b %.2e a %.2e" % (n, d, da))
    g.upload_taps(f, br, ar[:len(br)] if ag is not None else None)

pairs = [("fm_demod", Buf.FM_DEMOD), ("fm_out_iq", Buf.FM_OUT_IQ), ("pilot", Buf.PILOT), ("pll_dt", Buf.PLL_DT),
         ("pll", Buf.PLL), ("pll_raw_phase_error", Buf.PLL_RAW_PHASE_ERROR), ("pll_lpf_phase_error", Buf.PLL_LPF_PHASE_ERROR),
         ("audio_lpr", Buf.AUDIO_LPR), ("audio_lmr", Buf.AUDIO_LMR), ("rds", Buf.RDS), ("audio_out", Buf.AUDIO_OUT),
         ("bpsk_pll_sym", Buf.BPSK_PLL_SYM)]
worst = {n: [0.0, 0.0, 0.0] for n, _ in pairs}
rdsdec = fm.RDSDecoder()
sym_g, sym_r = [], []
for k in range(nblk):
    blk = iq[2 * B * k:2 * B * (k + 1)]
    ref.process_u8(blk)
    g.process_u8(blk)
    s = g.get(Buf.RDS_PRED_SYM)
    rdsdec.push_symbols(s)
    sym_g.append(s); sym_r.append(ref.get("rds_pred_sym"))
    for n, b in pairs:
        a = ref.get(n); c = g.get(b)
        d = np.abs(a - c)
        if n == "pll_dt": d = np.minimum(d, np.abs(1 - d))
        m = float(d.max())
        rms = float(np.sqrt(np.mean(np.abs(a) ** 2)))
        erms = float(np.sqrt(np.mean(d ** 2)))
        w = worst[n]
        w[0] = max(w[0], m)
        if k >= 48: w[1] = max(w[1], m); w[2] = max(w[2], erms / max(rms, 1e-30))
    if k in (0, 1, 8, 24, 47, nblk - 1):
        print("blk %3d nsym %d/%d lmr_phase %.5f/%.5f agc %.4f/%.4f rdsagc %.3f/%.3f" % (
            k, len(s), len(ref.get("rds_pred_sym")), g.scalar(Scalar.AUDIO_LMR_PHASE_ERROR), ref.scalar("audio_lmr_phase_error"),
            g.scalar(Scalar.AGC_PILOT_GAIN), ref.scalar("agc_pilot_gain"), g.scalar(Scalar.AGC_RDS_GAIN), ref.scalar("agc_rds_gain")), flush=True)
for n, _ in pairs:
    w = worst[n]
    print("%-22s max|d| all %.3e  after-lock %.3e  rel-rms after-lock %.3e (%.1f dB)" % (n, w[0], w[1], w[2], -20 * np.log10(max(w[2], 1e-30))))
sg, sr = np.concatenate(sym_g), np.concatenate(sym_r)
print("symbols", len(sg), len(sr))
n = min(len(sg), len(sr))
print("symbol sign agreement (aligned by index): %.4f" % np.mean(np.sign(sg[:n]) == np.sign(sr[:n])))
dr, vr, tr = ref.groups(); dg, vg, tg = rdsdec.groups()
print("groups ref %d gpu %d equal=%s" % (len(dr), len(dg), np.array_equal(dr, dg) and np.array_equal(vr, vg) and np.array_equal(tr, tg)))
print("rds bytes equal:", ref.rds_bytes() == rdsdec.rds_bytes(), len(ref.rds_bytes()))
print(ref.db()); print(rdsdec.db())
print("launches", g.launch_count)
