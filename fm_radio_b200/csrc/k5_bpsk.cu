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
//
// One thread per stream, one warp per CTA (same reasoning as K3: the loop is non-linear, so the parallelism is
// across streams and the time of a batch is n_samples x the latency of the dependent chain).  The arithmetic lives
// in k5_core.h; this file is the warp loop around it.  First version: the reference's per-sample order, ~390
// cycles per sample (libm atan2f, two roundf and two serial polynomial sines on the chain of EVERY sample,
// because in a warp some lane dumps a symbol at almost every sample).  This version walks SYMBOL by symbol
// (k5_symbol_step): every lane advances to its own next dump, so the dump work (atan2 -> carrier error) is done
// once per symbol for all 32 lanes together, the carrier loop + rotation of the symbol's samples runs as one batch
// with instruction-level parallelism, and the timing loop runs branch-free over the batch (state snapshotted at the
// dump), so only its short filter recurrence remains per sample.  Lanes therefore sit at different samples of their
// rows, each streaming it through a private shared-memory ring filled by cp.async two batches ahead (with plain
// loads at the head of the step, ncu showed 30 % of the kernel's time waiting for them: the address depends on
// where the previous symbol ended).
// Bit-identical to the per-sample loop (k5_sample; checked on the host by tests/test_k5_core_cpu.py and on the
// device by FMGPU_K5_LITERAL=1 in tests/test_gpu_parity.py).
#include "fm_common.cuh"
#include "k5_core.h"
#include <cstdlib>

namespace fm {

constexpr int K5_NB = 7;        // samples per batch: a locked clock dumps every 6 or 7 samples (16000 / 2375 = 6.74)

// Each lane streams ITS row through a private ring of 64 samples in shared memory (512 B per lane), filled by the
// lane's own cp.async 8 samples (64 B) at a time, at least two batches ahead of use.  16-byte unit u of lane l sits
// at unit u ^ l of the lane's region, which spreads the lanes' reads over the banks.  A lane reads only what it
// copied itself (cp.async.wait_group orders that), so no warp-level synchronisation is involved.
constexpr int K5_RING = 64;     // samples per lane

struct K5RingFetch {
    const float2* row;          // the lane's row in global memory
    const char* ring;           // the lane's 512-byte region
    int lane;
    __device__ __forceinline__ K5Sample operator()(int i) const {
        const int u = ((i >> 1) ^ lane) & (K5_RING / 2 - 1);
        const float2 v = *(const float2*)(ring + (u << 4) + ((i & 1) << 3));
        return K5Sample{ v.x, v.y };
    }
    // copy samples [first, first + 8) of the row into the ring (first is a multiple of 8)
    __device__ __forceinline__ void issue(int first) const {
        const unsigned base = (unsigned)__cvta_generic_to_shared(ring);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int u = (((first >> 1) + k) ^ lane) & (K5_RING / 2 - 1);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(base + (unsigned)(u << 4)), "l"(row + first + 2 * k));
        }
    }
};

struct K5DirectFetch {          // the per-sample loop's plain loads
    const float2* row;
    __device__ __forceinline__ K5Sample operator()(int i) const { const float2 v = __ldg(row + i); return K5Sample{ v.x, v.y }; }
};

struct K5LiveDebug {            // keep_intermediates: the display arrays of bpsk_synchroniser.cpp:175-182
    static constexpr bool kLive = true;
    K5Debug d; size_t o;
    __device__ __forceinline__ void sample(int i, float x_re, float x_im, float iq_re, float iq_im, bool is_zcd, bool is_ted,
                                           float ted_raw, float ted_pi, float pll_raw, float pll_pi, float dump_re, float dump_im) const {
        d.rds[o + i] = make_float2(x_re, x_im);
        d.pll_sym[o + i] = make_float2(iq_re, iq_im);
        d.zcd[o + i] = is_zcd ? 1 : 0;
        d.dump_trig[o + i] = is_ted ? 1 : 0;
        d.ted_raw[o + i] = ted_raw;
        d.ted_pi[o + i] = ted_pi;
        d.pll_raw[o + i] = pll_raw;
        d.pll_pi[o + i] = pll_pi;
        d.dump_filter[o + i] = make_float2(dump_re, dump_im);
    }
};

template <bool KEEP, bool LITERAL>
__global__ void __launch_bounds__(32)
k5_bpsk(const float2* __restrict__ rds_in, const float* __restrict__ rds_power_partial,
        float* __restrict__ state, float* __restrict__ pred_sym, int* __restrict__ sym_count,
        const K5Debug dbgp, const __grid_constant__ K5Params p)
{
    const int s_raw = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = s_raw < p.n_streams;
    const int s = valid ? s_raw : p.n_streams - 1;
    const int S = p.n_streams;
#define ST(f) state[(size_t)(f) * S + s]
    K5Lane L;
    L.lp_x1 = ST(BP_LPF_PLL_X1); L.lp_y1 = ST(BP_LPF_PLL_Y1); L.int_pll = ST(BP_INT_PLL);
    L.mix_t = ST(BP_MIX_T); L.pll_prev = ST(BP_PLL_PREV_ERR);
    L.zcd_xn = ST(BP_ZCD_XN);
    L.cooldown = (int)ST(BP_COOLDOWN);
    L.ted_yn = ST(BP_TED_YN); L.ted_phase_error = ST(BP_TED_PHASE_ERR); L.ted_prev = ST(BP_TED_PREV_ERR);
    L.lt_x1 = ST(BP_LPF_TED_X1); L.lt_y1 = ST(BP_LPF_TED_Y1); L.int_ted = ST(BP_INT_TED);
    L.dump_re = ST(BP_DUMP_RE); L.dump_im = ST(BP_DUMP_IM);
    float gain = ST(BP_AGC_GAIN);

    // agc.h:12-19
    float pw = 0.0f;
    for (int i = 0; i < p.n_tiles_k4; i++) pw += rds_power_partial[(size_t)s * p.n_tiles_k4 + i];
    const float avg_power = pw / (float)p.n;
    const float target_gain = sqrtf(p.agc_target / avg_power);
    gain = gain + p.agc_beta * (target_gain - gain);
    L.gain = gain;

    K5Coef c;
    c.pll_b0 = p.pll_b[0]; c.pll_b1 = p.pll_b[1]; c.pll_a0 = p.pll_a[0]; c.int_pll_KTs = p.int_pll_KTs; c.pll_Kp = p.pll_Kp;
    c.mixer_fgain = p.mixer_fgain; c.mixer_KTs = p.mixer_KTs;
    c.ted_b0 = p.ted_b[0]; c.ted_b1 = p.ted_b[1]; c.ted_a0 = p.ted_a[0]; c.int_ted_KTs = p.int_ted_KTs; c.ted_Kp = p.ted_Kp;
    c.dump_KTs = p.dump_KTs; c.ted_fgain = p.ted_fgain; c.ted_fcenter = p.ted_fcenter; c.ted_KTs = p.ted_KTs;
    c.cooldown_N = p.cooldown_N;

    const size_t o = (size_t)s * p.n;
    const int n = p.n;
    __shared__ __align__(16) char s_ring[32 * K5_RING * 8];
    const K5DirectFetch fetch{ rds_in + o };
    int total = 0;
    K5LiveDebug live{ dbgp, o };
    K5NoDebug none;
    if (LITERAL) {
        if (valid)
            for (int i = 0; i < n; i++) {
                const K5Sample x = fetch(i);
                float sr = 0.0f, si = 0.0f;
                const bool d = KEEP ? k5_sample(c, L, i, x.x_re, x.x_im, sr, si, live) : k5_sample(c, L, i, x.x_re, x.x_im, sr, si, none);
                if (d) {
                    pred_sym[o + total] = si;
                    if (KEEP) dbgp.raw_sym[o + total] = make_float2(sr, si);
                    total++;
                }
            }
    } else {
        // Symbol by symbol.  Ring invariant: after a step's issue, filled >= pos + 16 (pos advances by at most K5_NB = 7
        // per step and one 8-sample chunk is issued per step while filled < pos + 23), so the samples of a step
        // (< pos + 7) were issued by the PREVIOUS step's group at the latest: wait_group 1 after this step's commit.
        const K5RingFetch rf{ rds_in + o, s_ring + threadIdx.x * (K5_RING * 8), (int)threadIdx.x };
        int pos = valid ? 0 : n, filled = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (filled < n) { rf.issue(filled); filled += 8; }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        while (__any_sync(0xffffffffu, pos < n)) {
            if (filled < n && filled < pos + K5_NB + 16) { rf.issue(filled); filled += 8; }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            float sr = 0.0f, si = 0.0f;
            const bool d = KEEP ? k5_symbol_step<K5_NB>(c, L, rf, pos, n, pos < n, sr, si, live)
                                : k5_symbol_step<K5_NB>(c, L, rf, pos, n, pos < n, sr, si, none);
            if (d) {
                pred_sym[o + total] = si;
                if (KEEP) dbgp.raw_sym[o + total] = make_float2(sr, si);
                total++;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    if (!valid) return;
    sym_count[s] = total;
    ST(BP_LPF_PLL_X1) = L.lp_x1; ST(BP_LPF_PLL_Y1) = L.lp_y1; ST(BP_INT_PLL) = L.int_pll;
    ST(BP_MIX_T) = L.mix_t; ST(BP_PLL_PREV_ERR) = L.pll_prev; ST(BP_ZCD_XN) = L.zcd_xn;
    ST(BP_COOLDOWN) = (float)L.cooldown;
    ST(BP_TED_YN) = L.ted_yn; ST(BP_TED_PHASE_ERR) = L.ted_phase_error; ST(BP_TED_PREV_ERR) = L.ted_prev;
    ST(BP_LPF_TED_X1) = L.lt_x1; ST(BP_LPF_TED_Y1) = L.lt_y1; ST(BP_INT_TED) = L.int_ted;
    ST(BP_DUMP_RE) = L.dump_re; ST(BP_DUMP_IM) = L.dump_im; ST(BP_AGC_GAIN) = gain;
#undef ST
}

cudaError_t launch_k5(const float2* rds_in, const float* rds_power_partial, float* state, float* pred_sym,
                      int* sym_count, const K5Debug& d, const K5Params& p, cudaStream_t st)
{
    const bool literal = p.literal != 0;                 // A/B aid (fmgpu_set_option "k5_literal"): the per-sample loop
    const int grid = (p.n_streams + 31) / 32;
    const K5Debug nd{};
    if (p.keep) {
        if (literal) k5_bpsk<true, true><<<grid, 32, 0, st>>>(rds_in, rds_power_partial, state, pred_sym, sym_count, d, p);
        else         k5_bpsk<true, false><<<grid, 32, 0, st>>>(rds_in, rds_power_partial, state, pred_sym, sym_count, d, p);
    } else {
        if (literal) k5_bpsk<false, true><<<grid, 32, 0, st>>>(rds_in, rds_power_partial, state, pred_sym, sym_count, nd, p);
        else         k5_bpsk<false, false><<<grid, 32, 0, st>>>(rds_in, rds_power_partial, state, pred_sym, sym_count, nd, p);
    }
    return cudaGetLastError();
}

} // namespace fm
