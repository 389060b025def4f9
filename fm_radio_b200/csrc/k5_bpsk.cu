// K5: RDS AGC + BPSK symbol synchroniser -- the second non-linear sequential recurrence (16 kS/s).
//
// Replaces Broadcast_FM_Demod::SynchroniseRDS (broadcast_fm_demod.cpp:538-547):
//   AGC_Filter<cf32>::process, target 0.5                     dsp/agc.h:12-19
//   BPSK_Synchroniser::Process                                fm_demod/bpsk_synchroniser.cpp:94-186
//     PLL_Mixer::Update                                       fm_demod/pll_mixer.cpp:12-21
//     Zero_Crossing_Detector::process                         fm_demod/zero_crossing_detector.cpp:3-8
//     Trigger_Cooldown::on_trigger                            fm_demod/trigger_cooldown.cpp:4-13
//     TED_Clock::get_timing_error / update                    fm_demod/ted_clock.cpp:18-44
//   imag extraction of the dumped symbols                     broadcast_fm_demod.cpp:542-546
// One thread per stream (same reasoning as K3; this loop is 8x shorter and data dependent, so it is
// kept in the reference's exact operation order).  The block's RDS power arrives as per-tile
// partial sums from K4 and is reduced here in a fixed order.
#include "fm_common.cuh"

namespace fm {

template <bool KEEP>
__global__ void __launch_bounds__(32)
k5_bpsk(const float2* __restrict__ rds_in, const float* __restrict__ rds_power_partial,
        float* __restrict__ state, float* __restrict__ pred_sym, int* __restrict__ sym_count,
        float2* __restrict__ dbg_rds, float2* __restrict__ dbg_raw_sym, float2* __restrict__ dbg_pll_sym,
        uint8_t* __restrict__ dbg_zcd, uint8_t* __restrict__ dbg_dump_trig,
        float* __restrict__ dbg_ted_raw, float* __restrict__ dbg_ted_pi,
        float* __restrict__ dbg_pll_raw, float* __restrict__ dbg_pll_pi, float2* __restrict__ dbg_dump_filter,
        const __grid_constant__ K5Params p)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n_streams) return;
    const int S = p.n_streams;
#define ST(f) state[(size_t)(f) * S + s]
    float lp_x1 = ST(BP_LPF_PLL_X1), lp_y1 = ST(BP_LPF_PLL_Y1), int_pll = ST(BP_INT_PLL);
    float mix_t = ST(BP_MIX_T), pll_prev = ST(BP_PLL_PREV_ERR);
    float zcd_xn = ST(BP_ZCD_XN);
    int cooldown = (int)ST(BP_COOLDOWN);
    float ted_yn = ST(BP_TED_YN), ted_phase_error = ST(BP_TED_PHASE_ERR), ted_prev = ST(BP_TED_PREV_ERR);
    float lt_x1 = ST(BP_LPF_TED_X1), lt_y1 = ST(BP_LPF_TED_Y1), int_ted = ST(BP_INT_TED);
    float dump_re = ST(BP_DUMP_RE), dump_im = ST(BP_DUMP_IM);
    float gain = ST(BP_AGC_GAIN);

    // agc.h:12-19
    float pw = 0.0f;
    for (int i = 0; i < p.n_tiles_k4; i++) pw += rds_power_partial[(size_t)s * p.n_tiles_k4 + i];
    const float avg_power = pw / (float)p.n;
    const float target_gain = sqrtf(p.agc_target / avg_power);
    gain = gain + p.agc_beta * (target_gain - gain);

    const float4* x4 = (const float4*)(rds_in + (size_t)s * p.n);
    const size_t o = (size_t)s * p.n;
    int total = 0;
    // 8 samples (two 32-byte sectors) per lane are loaded one group ahead of their use, so the
    // global-load latency never sits on the recurrence
    float4 nx[4] = { x4[0], x4[1], x4[2], x4[3] };
    for (int i0 = 0; i0 < p.n; i0 += 8) {
    const float4 cur[4] = { nx[0], nx[1], nx[2], nx[3] };
    if (i0 + 8 < p.n) {
#pragma unroll
        for (int q = 0; q < 4; q++) nx[q] = x4[(i0 >> 1) + 4 + q];
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int i = i0 + j;
        float2 xi = (j & 1) ? make_float2(cur[j >> 1].z, cur[j >> 1].w) : make_float2(cur[j >> 1].x, cur[j >> 1].y);
        xi.x *= gain; xi.y *= gain;
        // PI controller of the carrier PLL (:106-113)
        const float pll_lpf = fmaf(pll_prev, p.pll_b[1], fmaf(lp_x1, p.pll_b[0], lp_y1 * p.pll_a[0]));
        lp_x1 = pll_prev; lp_y1 = pll_lpf;
        int_pll = clampf(fmaf(p.int_pll_KTs, pll_prev, int_pll), -1.0f, 1.0f);
        const float PI_pll_error = fmaf(pll_lpf, p.pll_Kp, int_pll);
        // PLL_Mixer::Update, f_center = 0
        const float control = clampf(PI_pll_error, -1.0f, 1.0f);
        const float freq = 0.0f + control * p.mixer_fgain;
        float tt = fmaf(p.mixer_KTs, freq, mix_t);
        tt = tt - roundf(tt);
        mix_t = tt;
        float dc = tt + 0.25f;
        dc = dc - roundf(dc);
        const float c = chebyshev_sine(dc), sn = chebyshev_sine(tt);
        const float iq_re = xi.x * c - xi.y * sn;
        const float iq_im = xi.x * sn + xi.y * c;
        // zero crossing on Q with cooldown (:127-132)
        bool is_zcd = (iq_im * zcd_xn) < 0.0f;
        zcd_xn = iq_im;
        if (is_zcd && cooldown == 0) { cooldown = p.cooldown_N; }
        else { if (cooldown > 0) cooldown--; is_zcd = false; }
        if (is_zcd) {
            float err = 2.0f * ted_yn;
            if (err > 1.0f) err = err - 2.0f;
            ted_prev = err;
        }
        // TED PI controller (:134-143)
        const float ted_lpf = fmaf(ted_prev, p.ted_b[1], fmaf(lt_x1, p.ted_b[0], lt_y1 * p.ted_a[0]));
        lt_x1 = ted_prev; lt_y1 = ted_lpf;
        int_ted = clampf(fmaf(p.int_ted_KTs, ted_prev, int_ted), -1.0f, 1.0f);
        const float PI_ted_error = fmaf(p.ted_Kp, ted_lpf, int_ted);
        ted_phase_error = -PI_ted_error;
        // integrate and dump (:146)
        dump_re = fmaf(p.dump_KTs, iq_re, dump_re);
        dump_im = fmaf(p.dump_KTs, iq_im, dump_im);
        // TED_Clock::update (ted_clock.cpp:31-44)
        bool is_ted;
        {
            const float ctl = clampf(ted_phase_error, -1.0f, 1.0f);
            const float f = fmaf(ctl, p.ted_fgain, p.ted_fcenter);
            const float v = fmaf(p.ted_KTs, f, ted_yn);
            ted_yn = v;
            const float offset = p.ted_KTs * f / 2.0f;
            is_ted = !(v < (1.0f - offset));
            if (is_ted) ted_yn = 0.0f;
        }
        if (is_ted) {
            const float sym_re = dump_re, sym_im = dump_im;
            dump_re = 0.0f; dump_im = 0.0f;
            const float sym_phase = atan2f(sym_im, sym_re);
            const float est = (sym_phase > 0.0f) ? (PI_F / 2.0f - sym_phase) : (-PI_F / 2.0f - sym_phase);
            pll_prev = est / (PI_F / 2.0f);
            pred_sym[o + total] = sym_im;
            if (KEEP) dbg_raw_sym[o + total] = make_float2(sym_re, sym_im);
            total++;
        }
        if (KEEP) {
            dbg_rds[o + i] = xi;
            dbg_pll_sym[o + i] = make_float2(iq_re, iq_im);
            dbg_zcd[o + i] = is_zcd ? 1 : 0;
            dbg_dump_trig[o + i] = is_ted ? 1 : 0;
            dbg_ted_raw[o + i] = ted_prev;
            dbg_ted_pi[o + i] = PI_ted_error;
            dbg_pll_raw[o + i] = pll_prev;
            dbg_pll_pi[o + i] = PI_pll_error;
            dbg_dump_filter[o + i] = make_float2(dump_re, dump_im);
        }
    }
    }
    sym_count[s] = total;
    ST(BP_LPF_PLL_X1) = lp_x1; ST(BP_LPF_PLL_Y1) = lp_y1; ST(BP_INT_PLL) = int_pll;
    ST(BP_MIX_T) = mix_t; ST(BP_PLL_PREV_ERR) = pll_prev; ST(BP_ZCD_XN) = zcd_xn;
    ST(BP_COOLDOWN) = (float)cooldown;
    ST(BP_TED_YN) = ted_yn; ST(BP_TED_PHASE_ERR) = ted_phase_error; ST(BP_TED_PREV_ERR) = ted_prev;
    ST(BP_LPF_TED_X1) = lt_x1; ST(BP_LPF_TED_Y1) = lt_y1; ST(BP_INT_TED) = int_ted;
    ST(BP_DUMP_RE) = dump_re; ST(BP_DUMP_IM) = dump_im; ST(BP_AGC_GAIN) = gain;
#undef ST
}

cudaError_t launch_k5(const float2* rds_in, const float* rds_power_partial, float* state, float* pred_sym,
                      int* sym_count, const K5Debug& d, const K5Params& p, cudaStream_t st)
{
    const int grid = (p.n_streams + 31) / 32;
    if (p.keep)
        k5_bpsk<true><<<grid, 32, 0, st>>>(rds_in, rds_power_partial, state, pred_sym, sym_count, d.rds, d.raw_sym,
            d.pll_sym, d.zcd, d.dump_trig, d.ted_raw, d.ted_pi, d.pll_raw, d.pll_pi, d.dump_filter, p);
    else
        k5_bpsk<false><<<grid, 32, 0, st>>>(rds_in, rds_power_partial, state, pred_sym, sym_count, nullptr, nullptr,
            nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, p);
    return cudaGetLastError();
}

} // namespace fm
