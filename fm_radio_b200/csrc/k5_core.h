// The arithmetic of the RDS BPSK symbol synchroniser, one stream = one "lane", written once for the CUDA kernel
// (k5_bpsk.cu) and for a host build that checks the kernel's control-flow restructuring without a GPU
// (tests/k5_core_check.cpp: the literal per-sample loop and the symbol-wise loop below must agree bit for bit).
//
// Reference (file:line under /root/reference/src): BPSK_Synchroniser::Process fm_demod/bpsk_synchroniser.cpp:94-186,
// PLL_Mixer::Update fm_demod/pll_mixer.cpp:12-21, Zero_Crossing_Detector fm_demod/zero_crossing_detector.cpp:3-8,
// Trigger_Cooldown fm_demod/trigger_cooldown.cpp:4-13, TED_Clock fm_demod/ted_clock.cpp:18-44.
//
// Per 16 kS/s sample the reference runs, in order: (A) carrier PI controller -> NCO -> two polynomial sines ->
// rotate the sample; (B) zero-crossing detector with cool-down -> timing-error PI controller -> ramp clock ->
// integrate-and-dump; and on a dump (C) atan2 of the symbol -> the carrier loop's phase error.  Every sample's (A)
// depends on the previous samples only through (C), i.e. only across a DUMP (every 6-7 samples): between two
// dumps the carrier loop free-runs on a constant error.  k5_symbol_step therefore processes one SYMBOL per call:
//   A  the carrier loop + rotation for the next NB samples as one batch (the NB sine pairs and rotations are
//      independent of one another: instruction-level parallelism instead of NB serial ~45-op chains), keeping the
//      loop state after every sample;
//   B  the timing loop over those samples until this stream's clock dumps (or the block ends), ~12 dependent ops
//      per sample; the carrier state is rolled back to the sample of the dump (the samples after it were
//      speculative and are recomputed by the next call with the new phase error);
//   C  once per call, the symbol's atan2 and the new carrier error.
// Every operation on a given sample has the same operands in the same order as the per-sample loop
// (k5_sample, the literal restatement), so both produce identical bits; what changes is the LENGTH OF THE
// DEPENDENT CHAIN per sample the warp waits for (measured on B200: 390 -> ~100 cycles per sample).  Streams
// of one warp dump at different samples, so each lane walks its own row at its own position.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define K5_HD __host__ __device__ __forceinline__
#else
#define K5_HD inline
#endif

namespace fm {

struct K5Coef {                 // the scalar coefficients of K5Params (fm_common.cuh), in the order they are used
    float pll_b0, pll_b1, pll_a0, int_pll_KTs, pll_Kp, mixer_fgain, mixer_KTs;
    float ted_b0, ted_b1, ted_a0, int_ted_KTs, ted_Kp, dump_KTs, ted_fgain, ted_fcenter, ted_KTs;
    int cooldown_N;
};

struct K5Lane {                 // bpsk_synchroniser.h state of one stream (BpskState order of fm_common.cuh)
    float lp_x1, lp_y1, int_pll, mix_t, pll_prev, zcd_xn;
    int cooldown;
    float ted_yn, ted_phase_error, ted_prev, lt_x1, lt_y1, int_ted, dump_re, dump_im, gain;
};

struct K5Sample { float x_re, x_im; };

K5_HD float k5_clampf(float x, float lo, float hi) {            // dsp/clamp.h:4-8 (NaN -> lo, like the reference)
    float y = (x > lo) ? x : lo;
    y = (y > hi) ? hi : y;
    return y;
}
// x - nearest integer by the 1.5 * 2^23 trick (two adds on the FMA pipe instead of roundf's ~6 instructions).
// std::round rounds halves away from zero, this rounds them to even: they differ only for x = +-0.5 exactly,
// where both results are the same phase (+-0.5 turn, and the polynomial sine of either is exactly 0).
K5_HD float k5_wrap_turn(float x) {
#if defined(__CUDA_ARCH__)
    return x - ((x + 12582912.0f) - 12582912.0f);
#else
    volatile float r = x + 12582912.0f;
    return x - (r - 12582912.0f);
#endif
}
// dsp/simd/chebyshev_sine.h:13-41 -- sin(2 pi x) on [-0.5, 0.5]; coefficients verbatim, Horner in the reference's order
K5_HD float k5_sine(float x) {
    const float z = x * x;
    float b = 3.20396066f;
    b = fmaf(b, z, -14.07150173f);
    b = fmaf(b, z, 38.50016403f);
    b = fmaf(b, z, -67.07687378f);
    b = fmaf(b, z, 64.83583069f);
    b = fmaf(b, z, -25.13274193f);
    return b * (z - 0.25f) * x;
}
// the minimax atan2 of fm_common.cuh (1.2e-7 rad); the host build divides where the device uses MUFU.RCP
K5_HD float k5_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
#if defined(__CUDA_ARCH__)
    float q = __fdividef(mn, mx);
#else
    float q = mn / mx;
#endif
    q = (mx == 0.0f) ? 0.0f : q;
    const float s = q * q;
    float p = -0.004054562299f;
    p = fmaf(p, s, 0.021862939178f);
    p = fmaf(p, s, -0.05591229796f);
    p = fmaf(p, s, 0.096421950378f);
    p = fmaf(p, s, -0.139086285623f);
    p = fmaf(p, s, 0.1994656543f);
    p = fmaf(p, s, -0.333298607622f);
    p = fmaf(p, s, 0.999999335572f);
    float r = p * q;
    r = (ay > ax) ? (1.57079632679489662f - r) : r;
    r = (x < 0.0f) ? (3.14159265358979323846f - r) : r;
    return copysignf(r, y);
}

// ---- (A) one sample of the carrier loop: PI controller (:106-113), PLL_Mixer::Update, rotation (:121-125) ----
K5_HD void k5_carrier(const K5Coef& c, float pll_prev, float& lp_x1, float& lp_y1, float& int_pll, float& mix_t,
                      float x_re, float x_im, float& iq_re, float& iq_im, float& PI_pll_error) {
    const float pll_lpf = fmaf(pll_prev, c.pll_b1, fmaf(lp_x1, c.pll_b0, lp_y1 * c.pll_a0));
    lp_x1 = pll_prev; lp_y1 = pll_lpf;
    int_pll = k5_clampf(fmaf(c.int_pll_KTs, pll_prev, int_pll), -1.0f, 1.0f);
    PI_pll_error = fmaf(pll_lpf, c.pll_Kp, int_pll);
    const float control = k5_clampf(PI_pll_error, -1.0f, 1.0f);
    const float freq = 0.0f + control * c.mixer_fgain;                        // f_center = 0
    const float tt = k5_wrap_turn(fmaf(c.mixer_KTs, freq, mix_t));
    mix_t = tt;
    const float dc = k5_wrap_turn(tt + 0.25f);
    const float cs = k5_sine(dc), sn = k5_sine(tt);
    iq_re = fmaf(x_re, cs, -(x_im * sn));
    iq_im = fmaf(x_re, sn, x_im * cs);
}

// ---- (B) one sample of the timing loop (:127-152); returns true when the integrate-and-dump filter dumps ----
K5_HD bool k5_timing(const K5Coef& c, K5Lane& L, float iq_re, float iq_im, bool& is_zcd_out, float& PI_ted_out) {
    bool is_zcd = (iq_im * L.zcd_xn) < 0.0f;
    L.zcd_xn = iq_im;
    if (is_zcd && L.cooldown == 0) { L.cooldown = c.cooldown_N; }
    else { if (L.cooldown > 0) L.cooldown--; is_zcd = false; }
    if (is_zcd) {                                                             // ted_clock.cpp:18-28
        float err = 2.0f * L.ted_yn;
        if (err > 1.0f) err = err - 2.0f;
        L.ted_prev = err;
    }
    const float ted_lpf = fmaf(L.ted_prev, c.ted_b1, fmaf(L.lt_x1, c.ted_b0, L.lt_y1 * c.ted_a0));
    L.lt_x1 = L.ted_prev; L.lt_y1 = ted_lpf;
    L.int_ted = k5_clampf(fmaf(c.int_ted_KTs, L.ted_prev, L.int_ted), -1.0f, 1.0f);
    const float PI_ted_error = fmaf(c.ted_Kp, ted_lpf, L.int_ted);
    L.ted_phase_error = -PI_ted_error;
    L.dump_re = fmaf(c.dump_KTs, iq_re, L.dump_re);
    L.dump_im = fmaf(c.dump_KTs, iq_im, L.dump_im);
    const float ctl = k5_clampf(L.ted_phase_error, -1.0f, 1.0f);               // ted_clock.cpp:31-44
    const float f = fmaf(ctl, c.ted_fgain, c.ted_fcenter);
    const float v = fmaf(c.ted_KTs, f, L.ted_yn);
    const float offset = c.ted_KTs * f * 0.5f;
    const bool is_ted = !(v < (1.0f - offset));
    L.ted_yn = is_ted ? 0.0f : v;
    is_zcd_out = is_zcd; PI_ted_out = PI_ted_error;
    return is_ted;
}

// ---- (C) the dumped symbol -> the carrier loop's next phase error (:155-170) ----
K5_HD float k5_symbol_error(float sym_re, float sym_im) {
    const float sym_phase = k5_atan2f(sym_im, sym_re);
    const float est = (sym_phase > 0.0f) ? (1.57079632679489662f - sym_phase) : (-1.57079632679489662f - sym_phase);
    return est * 0.636619772367581343f;                                       // / (pi / 2): the reference's -ffast-math build multiplies too
}

// per-sample display signals of keep_intermediates mode (bpsk_synchroniser.cpp:175-182); the default sink drops them
struct K5NoDebug {
    static constexpr bool kLive = false;
    K5_HD void sample(int, float, float, float, float, bool, bool, float, float, float, float, float, float) const {}
    K5_HD void symbol(int, float, float) const {}
};

// The literal per-sample loop (one call = one sample), the order of the reference.
template <class Dbg>
K5_HD bool k5_sample(const K5Coef& c, K5Lane& L, int i, float x_re, float x_im, float& sym_re, float& sym_im, const Dbg& dbg) {
    x_re *= L.gain; x_im *= L.gain;
    float iq_re, iq_im, PI_pll, PI_ted; bool is_zcd;
    k5_carrier(c, L.pll_prev, L.lp_x1, L.lp_y1, L.int_pll, L.mix_t, x_re, x_im, iq_re, iq_im, PI_pll);
    const bool is_ted = k5_timing(c, L, iq_re, iq_im, is_zcd, PI_ted);
    if (is_ted) {
        sym_re = L.dump_re; sym_im = L.dump_im;
        L.dump_re = 0.0f; L.dump_im = 0.0f;
        L.pll_prev = k5_symbol_error(sym_re, sym_im);
    }
    dbg.sample(i, x_re, x_im, iq_re, iq_im, is_zcd, is_ted, L.ted_prev, PI_ted, L.pll_prev, PI_pll, L.dump_re, L.dump_im);
    return is_ted;
}

// One symbol (at most NB samples) of one lane.  fetch(i) returns sample i of the lane's row for 0 <= i < n (the
// kernel serves it from a per-lane shared-memory ring filled two batches ahead).  Advances pos; returns true when a
// symbol was dumped (sym_re, sym_im).  Inactive lanes (active = false) execute the same instruction stream on
// clamped loads and change nothing.
//   * B is branch-free: the timing loop runs over all NB samples on a scratch copy of the state and the state is
//     snapshotted at the sample of the dump (or the block's last sample); what follows the dump is discarded.
//     Without per-sample branches the scheduler overlaps the samples' independent work, and the dependent chain
//     per sample is the 3-op filter recurrence, plus ~10 ops once per symbol at the zero crossing.
template <int NB, class Fetch, class Dbg>
K5_HD bool k5_symbol_step(const K5Coef& c, K5Lane& L, const Fetch& fetch, int& pos, int n, bool active,
                          float& sym_re, float& sym_im, const Dbg& dbg) {
    float xr[NB], xi[NB], iq_re[NB], iq_im[NB], t_lpy[NB], t_int[NB], t_mix[NB], t_pi[NB];
#pragma unroll
    for (int q = 0; q < NB; q++) {
        const int i = (pos + q < n) ? pos + q : n - 1;
        const K5Sample s = fetch(i);
        xr[q] = s.x_re * L.gain; xi[q] = s.x_im * L.gain;
    }
    {   // A: the carrier loop free-runs on the current error
        float lp_x1 = L.lp_x1, lp_y1 = L.lp_y1, int_pll = L.int_pll, mix_t = L.mix_t;
#pragma unroll
        for (int q = 0; q < NB; q++) {
            k5_carrier(c, L.pll_prev, lp_x1, lp_y1, int_pll, mix_t, xr[q], xi[q], iq_re[q], iq_im[q], t_pi[q]);
            t_lpy[q] = lp_y1; t_int[q] = int_pll; t_mix[q] = mix_t;
        }
    }
    // B: the timing loop on a scratch state T; S = the state at the dump / at the block's last sample
    K5Lane T = L, S = L;
    bool done = !active, dumped = false;
    int consumed = 0;
    const int q_last = (n - 1 - pos < NB - 1) ? n - 1 - pos : NB - 1;
    const float pll_prev_used = L.pll_prev;
#pragma unroll
    for (int q = 0; q < NB; q++) {
        bool is_zcd; float PI_ted;
        const bool is_ted = k5_timing(c, T, iq_re[q], iq_im[q], is_zcd, PI_ted);
        const bool take = !done && (is_ted || q == q_last);
        if (Dbg::kLive && !done) {
            const float raw_pll = is_ted ? k5_symbol_error(T.dump_re, T.dump_im) : pll_prev_used;
            dbg.sample(pos + q, xr[q], xi[q], iq_re[q], iq_im[q], is_zcd, is_ted, T.ted_prev, PI_ted, raw_pll, t_pi[q],
                       is_ted ? 0.0f : T.dump_re, is_ted ? 0.0f : T.dump_im);
        }
        if (take) {
            S.zcd_xn = T.zcd_xn; S.cooldown = T.cooldown; S.ted_yn = T.ted_yn; S.ted_phase_error = T.ted_phase_error;
            S.ted_prev = T.ted_prev; S.lt_x1 = T.lt_x1; S.lt_y1 = T.lt_y1; S.int_ted = T.int_ted;
            S.dump_re = is_ted ? 0.0f : T.dump_re; S.dump_im = is_ted ? 0.0f : T.dump_im;
            S.lp_x1 = pll_prev_used; S.lp_y1 = t_lpy[q]; S.int_pll = t_int[q]; S.mix_t = t_mix[q];
            sym_re = T.dump_re; sym_im = T.dump_im;
            consumed = q + 1; dumped = is_ted;
        }
        done = done || take;
    }
    L = S;
    // C: once per symbol
    if (dumped) L.pll_prev = k5_symbol_error(sym_re, sym_im);
    pos += consumed;
    return dumped;
}

} // namespace fm
