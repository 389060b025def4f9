// K3: the 19 kHz pilot PLL -- a non-linear, strictly sequential per-sample recurrence at 128 kS/s.
//
// Replaces the loop of Broadcast_FM_Demod::LockOntoPilot (broadcast_fm_demod.cpp:426-456) with its
// helpers IIR_Filter<float> K=2 one sample at a time (dsp/iir_filter.h:40-69), Integrator_Block
// (dsp/integrator.h:9-13), clamp (dsp/clamp.h:4-8), PLL_Mixer::Update (fm_demod/pll_mixer.cpp:12-21)
// and the block-wise AGC gain update of AGC_Filter (dsp/agc.h:12-19).
//
// Parallelisation: the loop cannot be scanned (clamps, phase wrap), so parallelism is across
// streams only: ONE THREAD PER STREAM, one warp (32 streams) per CTA so that the warps spread over
// the SMs; every other stage of the chain overlaps with it on other CUDA streams.  What is
// optimised here is the LENGTH OF THE DEPENDENT CHAIN per sample, because aggregate throughput
// of a 1024-stream batch is bounded by n_samples * chain latency:
//   reference chain  e -> PI -> NCO -> 2 x polynomial sine -> complex multiply -> atan2f -> e
//   this kernel      e -> PI -> NCO -> (theta[n] + t) wrapped to one turn        -> e
// using theta[n] = arg(pilot[n])/2pi precomputed in parallel by K2 (arg(a*b) = arg(a)+arg(b);
// arg of the reference's oscillator sample is 2 pi t up to its 4e-8 polynomial error).  The AGC
// gain is a positive scale and does not enter the angle; if it is not finite (all-zero block:
// sqrt(1/0)) the reference's pilot becomes NaN and poisons the loop for good -- reproduced through
// `poison`.  Per-thread I/O is whole 32-byte sectors (LDG.128 / STG.128 on the lane's own row),
// loaded a full 32-sample group ahead, so memory latency never sits on the chain.
#include "fm_common.cuh"

namespace fm {

template <bool KEEP>
__global__ void __launch_bounds__(32)
k3_pll(const float* __restrict__ theta, const float* __restrict__ power, float* __restrict__ state,
       float* __restrict__ pll_dt, float* __restrict__ dbg_raw, float* __restrict__ dbg_pi,
       const __grid_constant__ K3Params p)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n_streams) return;
    const int S = p.n_streams;
    float x1 = state[PLL_LPF_X1 * S + s];
    float y1 = state[PLL_LPF_Y1 * S + s];
    float integ = state[PLL_INT * S + s];
    float t = state[PLL_T * S + s];
    float e = state[PLL_E_PREV * S + s];
    float gain = state[PLL_AGC_GAIN * S + s];

    // dsp/agc.h:12-19 (block-wise): P = mean |x|^2, g += beta*(sqrt(target/P) - g)
    const float avg_power = power[s] / (float)p.n;
    const float target_gain = sqrtf(p.agc_target / avg_power);
    gain = gain + p.agc_beta * (target_gain - gain);
    const float poison = 0.0f * gain;          // 0 for a finite gain, NaN for inf/NaN

    const float4* th4 = (const float4*)(theta + (size_t)s * p.n);
    float4* dt4 = (float4*)(pll_dt + (size_t)s * p.n);
    float4* raw4 = KEEP ? (float4*)(dbg_raw + (size_t)s * p.n) : nullptr;
    float4* pi4 = KEEP ? (float4*)(dbg_pi + (size_t)s * p.n) : nullptr;

    const float b0 = p.lpf_b[0], b1 = p.lpf_b[1], a0 = p.lpf_a[0];
    const float int_KTs = p.int_KTs, Kp = p.Kp, f_gain = p.f_gain, f_center = p.f_center, mixer_KTs = p.mixer_KTs;
    // Groups of 32 samples (four 32-byte sectors per lane), loaded one whole group (~2000 cycles of
    // recurrence) ahead: the ncu source view of the 8-sample version showed 46 % of the stall samples
    // on the first use of the prefetched registers (DRAM latency > 8 iterations).
    constexpr int G = 32;
    float4 nx[G / 4];
#pragma unroll
    for (int q = 0; q < G / 4; q++) nx[q] = th4[q];
    for (int i = 0; i < p.n; i += G) {
        float4 cur[G / 4];
#pragma unroll
        for (int q = 0; q < G / 4; q++) cur[q] = nx[q];
        if (i + G < p.n) {
#pragma unroll
            for (int q = 0; q < G / 4; q++) nx[q] = th4[((i + G) >> 2) + q];
        }
#pragma unroll
        for (int q = 0; q < G / 4; q++) {
            const float th[4] = { cur[q].x, cur[q].y, cur[q].z, cur[q].w };
            float dt[4], raw[4], pie[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // IIR1: y = xn[0]*b[0] + yn[0]*a[0] + xn[1]*b[1]; the part that does not depend on the
                // newest error is formed first so only one FFMA sits on the e -> lpf path.
                const float m = fmaf(x1, b0, y1 * a0);
                const float lpf = fmaf(e, b1, m);
                x1 = e; y1 = lpf;
                integ = clampf(fmaf(int_KTs, e, integ), -1.0f, 1.0f);
                const float PI_error = fmaf(lpf, Kp, integ);
                // PLL_Mixer::Update
                const float control = clampf(PI_error, -1.0f, 1.0f);
                const float freq = fmaf(control, f_gain, f_center);
                const float tu = fmaf(mixer_KTs, freq, t);                 // |tu| < 0.66
                // t - round(t): the nearest integer comes from two FADDs with 1.5 * 2^23 (all on the
                // FMA pipe, 4-cycle latency each, no FRND / compare-select chain).  Ties (|tu| == 0.5
                // exactly) round to even instead of away from zero; both results are the same phase.
                t = tu - ((tu + 12582912.0f) - 12582912.0f);
                // phase detector: arg(pilot * pll) = 2 pi * wrap(theta + t)
                const float u = th[j] + tu;                                 // |u| < 1.16
                const float uw = u - ((u + 12582912.0f) - 12582912.0f);
                e = fmaf(uw, TWO_PI_F, poison);
                dt[j] = t;
                if (KEEP) { raw[j] = e; pie[j] = PI_error; }
            }
            dt4[(i >> 2) + q] = make_float4(dt[0], dt[1], dt[2], dt[3]);
            if (KEEP) {
                raw4[(i >> 2) + q] = make_float4(raw[0], raw[1], raw[2], raw[3]);
                pi4[(i >> 2) + q] = make_float4(pie[0], pie[1], pie[2], pie[3]);
            }
        }
    }
    state[PLL_LPF_X1 * S + s] = x1;
    state[PLL_LPF_Y1 * S + s] = y1;
    state[PLL_INT * S + s] = integ;
    state[PLL_T * S + s] = t;
    state[PLL_E_PREV * S + s] = e;
    state[PLL_AGC_GAIN * S + s] = gain;
}

cudaError_t launch_k3(const float* theta, const float* power, float* state, float* pll_dt,
                      float* dbg_raw, float* dbg_pi, const K3Params& p, cudaStream_t st)
{
    const int grid = (p.n_streams + 31) / 32;
    if (p.keep) k3_pll<true><<<grid, 32, 0, st>>>(theta, power, state, pll_dt, dbg_raw, dbg_pi, p);
    else        k3_pll<false><<<grid, 32, 0, st>>>(theta, power, state, pll_dt, dbg_raw, dbg_pi, p);
    return cudaGetLastError();
}

} // namespace fm
