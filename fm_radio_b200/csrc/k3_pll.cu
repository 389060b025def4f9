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
// optimised here is the LENGTH OF THE DEPENDENT CHAIN per sample, because the time of a
// 1024-stream batch is n_samples x chain latency (one warp per SM sub-partition, nothing to hide
// latency behind):
//   reference   e -> IIR, integrator+clamp -> PI -> clamp -> freq -> t+=, wrap -> 2 x polynomial sine
//                 -> complex multiply -> atan2f -> e                         (~50 dependent ops)
//   first cut   e -> fma, min, max -> fma, min, max -> fma -> fma -> add -> 3-op wrap -> fma -> e
//               13 dependent ops of 4-5 cycles; measured 101 cycles per sample on B200
//   exact body  w -> {fma.sat, fma.sat} -> sub -> {fma.sat, fma.sat} -> sub -> fma -> fma -> 3-op wrap -> w
//               9 dependent ops, all on the FMA pipe (no FMA<->ALU pipe crossings); 57 cycles per sample
//   this kernel w -> {fma, fma} -> fma -> fma -> fma -> 3-op wrap -> w
//               7 dependent ops: groups of 32 samples run with the clamps taken as inactive (exact whenever
//               they are: clamp(x) returns x bit for bit for |x| <= 1) and are redone by the exact body from
//               the saved state if a clamp would have acted -- never, for a locking or locked loop
// using
//   * theta[n] = arg(pilot[n])/2pi precomputed in parallel by K2 (arg(a*b) = arg(a)+arg(b); arg of
//     the reference's oscillator sample is 2 pi t up to its 4e-8 polynomial error), so the phase
//     detector is wrap(theta + t);
//   * the error kept in TURNS (w = e / 2pi) with 2pi folded into the coefficients that consume it;
//   * clamp(x, -1, 1) == sat(x) - sat(-x) exactly (one of the two terms is zero), and sat() is a free
//     modifier of the FFMA that produces x, so a clamp costs one dependent FADD instead of two
//     ALU-pipe FMNMX;
//   * the phase detector input formed straight from freq, fma(KTs, freq, theta + t), next to the
//     NCO update fma(KTs, freq, t) instead of after it.  (Merging freq into the increment,
//     (t + KTs*fc) + (KTs*fg)*c, saves one more op but was rejected: it moves the NCO's rounding
//     points and the loop's static phase error with them -- 3.7e-5 turns away from the reference
//     instead of 9e-7, measured.)
//   fast pass   w -> fma -> fma -> w: the wraps leave the chain as well.  From the two NCO / detector updates above,
//               w[n+1] = wrap(w[n] + (theta[n+1] - theta[n] + KTs*fc) + KTs*fg*c[n+1]); the middle term e[n+1] depends on
//               the input only, and for a locked loop nothing wraps (|w| ~ 1e-3 turn), so inside a group of 32 samples
//               w[n+1] = fma(c[n+1], KTs*fg, w[n] + e[n+1]),  c[n+1] = fma(w[n+1], Kp*b1 + ci, Kp*m + integ)
//               (the control from the merged coefficient; the filter, integrator and NCO STATES keep their own exact
//               recurrences beside the chain).  The chain per sample is 2 ops + the IIR's own 3-op recurrence instead
//               of 7; the loop becomes issue-bound at ~20 instructions per sample.  The detector value of the group's
//               LAST sample is re-anchored on the exact expression wrap(theta + t), so rounding cannot accumulate in w
//               (inside a group it stays below 2e-7 turn), and a group in which |w| >= 1/4 turn, a clamp could have
//               acted or a NaN appeared is redone from the saved state by the exact body.  Not bit-identical to the
//               exact body: the control differs in its last bit (below the NCO's 2^-9 Hz frequency quantum) and w
//               carries less rounding noise; pll_dt stays within 2e-6 turn of the reference (tolerance 1e-4).
// The AGC gain is a positive scale and does not enter the angle; if it is not finite (all-zero
// block: sqrt(1/0)) the reference's pilot becomes NaN and poisons the loop for good -- such a lane
// takes the slow loop below, which keeps the reference's clamp-of-NaN behaviour.  Per-thread I/O is
// whole 32-byte sectors on the lane's own row; the input streams through a shared-memory ring filled by
// cp.async seven 32-sample groups ahead, so memory latency never sits on the chain.
#include "fm_common.cuh"
#include <algorithm>
#include <atomic>
#include <cmath>

namespace fm {

__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float d;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// x - nearest integer.  WRAP 0: two FADDs with 1.5 * 2^23 (FMA pipe, ties to even); WRAP 1: FRND.
template <int WRAP>
__device__ __forceinline__ float wrap_turn(float x) {
    if (WRAP == 0) return x - ((x + 12582912.0f) - 12582912.0f);
    return x - rintf(x);
}

constexpr int K3_RING = 8;      // groups of theta in flight per warp (8 x 4 KB of shared memory)

template <bool KEEP, int WRAP, bool FAST>
__global__ void __launch_bounds__(32)
k3_pll(const float* __restrict__ theta, const float* __restrict__ power, float* __restrict__ state,
       float* __restrict__ pll_dt, float* __restrict__ dbg_raw, float* __restrict__ dbg_pi,
       const __grid_constant__ K3Params p)
{
    __shared__ __align__(128) float4 s_ring[K3_RING * 32 * 8];     // [slot][lane][8 chunks of 16 bytes]
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= p.n_streams) return;
    const int S = p.n_streams;
    float x1 = state[PLL_LPF_X1 * S + s];       // previous phase error, in turns
    float y1 = state[PLL_LPF_Y1 * S + s];
    float integ = state[PLL_INT * S + s];
    float t = state[PLL_T * S + s];
    float w = state[PLL_E_PREV * S + s];        // phase error in turns
    float gain = state[PLL_AGC_GAIN * S + s];
    float th_last = state[PLL_TH_PREV * S + s];  // theta of the previous block's last sample (fast pass only)

    // dsp/agc.h:12-19 (block-wise): P = mean |x|^2, g += beta*(sqrt(target/P) - g)
    const float avg_power = power[s] / (float)p.n;
    const float target_gain = sqrtf(p.agc_target / avg_power);
    gain = gain + p.agc_beta * (target_gain - gain);
    const bool poisoned = !(fabsf(gain) <= 3.0e38f);            // inf or NaN

    const float4* th4 = (const float4*)(theta + (size_t)s * p.n);
    float4* dt4 = (float4*)(pll_dt + (size_t)s * p.n);
    float4* raw4 = KEEP ? (float4*)(dbg_raw + (size_t)s * p.n) : nullptr;
    float4* pi4 = KEEP ? (float4*)(dbg_pi + (size_t)s * p.n) : nullptr;

    const float a0 = p.lpf_a[0], Kp = p.Kp;
    const float b0t = p.lpf_b[0] * TWO_PI_F, b1t = p.lpf_b[1] * TWO_PI_F, ci = p.int_KTs * TWO_PI_F;
    const float f_gain = p.f_gain, f_center = p.f_center, mixer_KTs = p.mixer_KTs;

    if (!poisoned) {
        // Groups of 32 samples = one 128-byte line per lane.  theta was written by K2 together with 2x as much
        // fm_out_iq, so by now part of it has left the L2, and under the FIR kernels' DRAM traffic a miss costs
        // far more than the ~1500 cycles one group of register look-ahead buys (ncu: long_scoreboard was this
        // kernel's top stall, and it ran 1.35x slower inside the pipeline than alone).  So each lane streams
        // its row through a ring of K3_RING groups in shared memory with cp.async, K3_RING - 1 groups
        // (~10 000 cycles) ahead of use; chunk c of lane l sits at chunk c ^ (l & 7) of the lane's 128-byte
        // row, so the lanes' LDS.128 are bank-conflict free.  A lane reads only what it copied itself
        // (cp.async.wait_group orders that), so no warp-level synchronisation is involved.
        constexpr int G = 32;
        const int lane = threadIdx.x;
        const int n_groups = p.n / G;
        char* ring = (char*)s_ring + lane * 128;
        const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring);
        auto issue = [&](int g) {                        // copy group g into ring slot g % K3_RING (or nothing past the end)
            if (g < n_groups) {
                const char* src = (const char*)(th4 + (size_t)g * (G / 4));
                const unsigned dst = ring_sa + (unsigned)(g % K3_RING) * (32 * 128);
#pragma unroll
                for (int c = 0; c < 8; c++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + (unsigned)((c ^ (lane & 7)) << 4)), "l"(src + c * 16));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto fetch = [&](int g, float4 (&dst)[G / 4]) {  // ring slot -> registers
            const char* row = ring + (g % K3_RING) * (32 * 128);
#pragma unroll
            for (int c = 0; c < 8; c++) dst[c] = *(const float4*)(row + ((c ^ (lane & 7)) << 4));
        };
#pragma unroll
        for (int g = 0; g < K3_RING - 1; g++) issue(g);
        asm volatile("cp.async.wait_group %0;" :: "n"(K3_RING - 2) : "memory");      // group 0 has landed
        float4 nx[G / 4];
        fetch(0, nx);
        for (int i = 0; i < p.n; i += G) {
            const int g = i / G;
            float4 cur[G / 4];
#pragma unroll
            for (int q = 0; q < G / 4; q++) cur[q] = nx[q];
            issue(g + K3_RING - 1);                      // into the slot of group g - 1, read into registers one iteration ago
            asm volatile("cp.async.wait_group %0;" :: "n"(K3_RING - 2) : "memory");  // group g + 1 has landed
            if (g + 1 < n_groups) fetch(g + 1, nx);
            // Speculative pass: both clamps taken as inactive, which shortens the chain from 9 to 7 dependent
            // ops (w -> {fma, fma} -> fma -> fma -> fma -> 3-op wrap).  clamp(x) = sat(x) - sat(-x) returns x
            // itself, bit for bit, whenever |x| <= 1, so the pass is exact unless a clamp would have acted
            // (or a NaN met fma.sat); the group is then redone from the saved state by the exact body.  A locked
            // or locking loop never clamps: |PI| stays below 0.05.
            const float sx1 = x1, sy1 = y1, sinteg = integ, st = t, sw = w;
            bool redo;
            if (FAST) {
                // fast pass (header): 2-op detector chain, exact state recurrences beside it
                const float cw = fmaf(Kp, b1t, ci), kfg = mixer_KTs * f_gain, d_nom = mixer_KTs * f_center;
                float wmax = 0.0f, thp = th_last, t_before_last = t, freq_last = f_center;
#pragma unroll
                for (int q = 0; q < G / 4; q++) {
                    const float th[4] = { cur[q].x, cur[q].y, cur[q].z, cur[q].w };
                    float dt[4], raw[4], pie[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        // w = the detector value of THIS sample's input (state on entry: the previous sample's)
                        const float m = fmaf(x1, b0t, y1 * a0);
                        const float g = fmaf(m, Kp, integ);
                        const float control = fmaf(w, cw, g);
                        const float lpf = fmaf(w, b1t, m);
                        integ = fmaf(ci, w, integ);
                        x1 = w; y1 = lpf;
                        const float freq = fmaf(control, f_gain, f_center);
                        t_before_last = t; freq_last = freq;
                        t = wrap_turn<WRAP>(fmaf(mixer_KTs, freq, t));
                        const float e = wrap_turn<WRAP>((th[j] - thp) + d_nom);
                        thp = th[j];
                        w = fmaf(control, kfg, w + e);
                        wmax = fmaxf(wmax, fabsf(w));
                        dt[j] = t;
                        if (KEEP) { raw[j] = w * TWO_PI_F; pie[j] = control; }
                    }
                    dt4[(i >> 2) + q] = make_float4(dt[0], dt[1], dt[2], dt[3]);
                    if (KEEP) {
                        raw4[(i >> 2) + q] = make_float4(raw[0], raw[1], raw[2], raw[3]);
                        pi4[(i >> 2) + q] = make_float4(pie[0], pie[1], pie[2], pie[3]);
                    }
                }
                // re-anchor the detector on the exact expression of the group's last sample
                w = wrap_turn<WRAP>(fmaf(mixer_KTs, freq_last, thp + t_before_last));
                redo = !(wmax < 0.25f) || !(fabsf(integ) <= p.integ_safe);
                if (!redo) th_last = thp;
            } else {
#pragma unroll
            for (int q = 0; q < G / 4; q++) {
                const float th[4] = { cur[q].x, cur[q].y, cur[q].z, cur[q].w };
                float dt[4], raw[4], pie[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float m = fmaf(x1, b0t, y1 * a0);
                    const float a = th[j] + t;
                    const float lpf = fmaf(w, b1t, m);
                    integ = fmaf(ci, w, integ);
                    x1 = w; y1 = lpf;
                    const float control = fmaf(lpf, Kp, integ);
                    const float freq = fmaf(control, f_gain, f_center);
                    t = wrap_turn<WRAP>(fmaf(mixer_KTs, freq, t));
                    w = wrap_turn<WRAP>(fmaf(mixer_KTs, freq, a));
                    dt[j] = t;
                    if (KEEP) { raw[j] = w * TWO_PI_F; pie[j] = control; }
                }
                dt4[(i >> 2) + q] = make_float4(dt[0], dt[1], dt[2], dt[3]);
                if (KEEP) {
                    raw4[(i >> 2) + q] = make_float4(raw[0], raw[1], raw[2], raw[3]);
                    pi4[(i >> 2) + q] = make_float4(pie[0], pie[1], pie[2], pie[3]);
                }
            }
                redo = !(fabsf(integ) <= p.integ_safe);
            }
            // |w| <= 1/2 turn bounds the low-pass output and the integrator's step, so ONE test of the integrator at
            // the end of the group proves that neither clamp could have acted anywhere inside it (integ_safe is
            // derived on the host from the loop's coefficients, launch_k3; NaN fails the test; <= 0 = always redo)
            if (redo) {
                x1 = sx1; y1 = sy1; integ = sinteg; t = st; w = sw;
                th_last = cur[G / 4 - 1].w;
#pragma unroll 1
                for (int q = 0; q < G / 4; q++) {
                    const float4 c4 = th4[(i >> 2) + q];
                    const float th[4] = { c4.x, c4.y, c4.z, c4.w };
                    float dt[4], raw[4], pie[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        // off the chain: everything that depends only on the previous iteration's t, x1, y1
                        const float m = fmaf(x1, b0t, y1 * a0);
                        const float a = th[j] + t;                              // |a| <= 1
                        // IIR1 y = xn[0]*b[0] + yn[0]*a[0] + xn[1]*b[1] on the newest error
                        const float lpf = fmaf(w, b1t, m);
                        // integrator with clamp(-1, 1) = sat(x) - sat(-x)
                        integ = fma_sat(ci, w, integ) - fma_sat(-ci, w, -integ);
                        x1 = w; y1 = lpf;
                        // PI error and PLL_Mixer::Update's clamp, the same way
                        const float control = fma_sat(lpf, Kp, integ) - fma_sat(-lpf, Kp, -integ);
                        // NCO (pll_mixer.cpp:12-21), in the reference's rounding structure -- freq is
                        // quantised to ulp(19000) = 2^-9 Hz and the loop's static phase error (1 rad per
                        // Hz of NCO bias) follows that quantisation: t += KTs*freq; t -= round(t)
                        const float freq = fmaf(control, f_gain, f_center);
                        t = wrap_turn<WRAP>(fmaf(mixer_KTs, freq, t));
                        // phase detector: arg(pilot * pll) / 2pi = wrap(theta + t), straight from freq
                        w = wrap_turn<WRAP>(fmaf(mixer_KTs, freq, a));
                        dt[j] = t;
                        if (KEEP) { raw[j] = w * TWO_PI_F; pie[j] = fmaf(lpf, Kp, integ); }
                    }
                    dt4[(i >> 2) + q] = make_float4(dt[0], dt[1], dt[2], dt[3]);
                    if (KEEP) {
                        raw4[(i >> 2) + q] = make_float4(raw[0], raw[1], raw[2], raw[3]);
                        pi4[(i >> 2) + q] = make_float4(pie[0], pie[1], pie[2], pie[3]);
                    }
                }
            }
        }
    } else {
        // Poisoned lane: e is NaN from here on; clamp(NaN) = -1 as dsp/clamp.h:4-8 evaluates it, so
        // the NCO runs at f_center - f_gain.  Reference operation order, no latency tricks.
        const float nan = 0.0f * gain;
        float x1r = x1 * TWO_PI_F;
        for (int i = 0; i < p.n; i++) {
            const float e = fmaf(w, TWO_PI_F, nan);
            const float lpf = fmaf(e, p.lpf_b[1], fmaf(x1r, p.lpf_b[0], y1 * a0));
            x1r = e; y1 = lpf;
            integ = clampf(fmaf(p.int_KTs, e, integ), -1.0f, 1.0f);
            const float PI_error = fmaf(lpf, Kp, integ);
            const float control = clampf(PI_error, -1.0f, 1.0f);
            const float tu = fmaf(p.mixer_KTs, fmaf(control, p.f_gain, p.f_center), t);
            t = tu - roundf(tu);
            pll_dt[(size_t)s * p.n + i] = t;
            if (KEEP) { dbg_raw[(size_t)s * p.n + i] = e; dbg_pi[(size_t)s * p.n + i] = PI_error; }
        }
        x1 = nan; w = nan;
        th_last = theta[(size_t)s * p.n + p.n - 1];
    }
    state[PLL_LPF_X1 * S + s] = x1;
    state[PLL_LPF_Y1 * S + s] = y1;
    state[PLL_INT * S + s] = integ;
    state[PLL_T * S + s] = t;
    state[PLL_E_PREV * S + s] = w;
    state[PLL_AGC_GAIN * S + s] = gain;
    state[PLL_TH_PREV * S + s] = th_last;
}

// One 32-sample group by the exact body (the redo path of the fast passes), theta read from global memory.
template <bool KEEP, int WRAP>
__device__ __forceinline__ void k3_exact_group(const float4* __restrict__ th4, float4* __restrict__ dt4, float4* __restrict__ raw4, float4* __restrict__ pi4,
                                               int i, bool live, float& x1, float& y1, float& integ, float& t, float& w,
                                               float a0, float Kp, float b0t, float b1t, float ci, float f_gain, float f_center, float mixer_KTs)
{
#pragma unroll 1
    for (int q = 0; q < 8; q++) {
        const float4 c4 = th4[(i >> 2) + q];
        const float th[4] = { c4.x, c4.y, c4.z, c4.w };
        float dt[4], raw[4], pie[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float m = fmaf(x1, b0t, y1 * a0);
            const float a = th[j] + t;
            const float lpf = fmaf(w, b1t, m);
            integ = fma_sat(ci, w, integ) - fma_sat(-ci, w, -integ);
            x1 = w; y1 = lpf;
            const float control = fma_sat(lpf, Kp, integ) - fma_sat(-lpf, Kp, -integ);
            const float freq = fmaf(control, f_gain, f_center);
            t = wrap_turn<WRAP>(fmaf(mixer_KTs, freq, t));
            w = wrap_turn<WRAP>(fmaf(mixer_KTs, freq, a));
            dt[j] = t;
            if (KEEP) { raw[j] = w * TWO_PI_F; pie[j] = fmaf(lpf, Kp, integ); }
        }
        if (live) {
            dt4[(i >> 2) + q] = make_float4(dt[0], dt[1], dt[2], dt[3]);
            if (KEEP) {
                raw4[(i >> 2) + q] = make_float4(raw[0], raw[1], raw[2], raw[3]);
                pi4[(i >> 2) + q] = make_float4(pie[0], pie[1], pie[2], pie[3]);
            }
        }
    }
}

// ---- fast pass with a helper warp ----------------------------------------------------------------------------------
// ncu on the single-warp fast pass: the warp issues 22 instructions per sample at 1.8 cycles each (selected 1.00 +
// dispatch 0.29 + wait 0.24 + long_scoreboard 0.16) = 39.7 cycles per sample, while its dependent chain is 2 FMAs: the
// loop is bound by how many instructions ONE warp has to issue, and 6 of them per sample (the detector increment
// e[n] = wrap((theta[n] - theta[n-1]) + KTs fc), a function of the input alone, and the ring traffic behind it) do not
// belong to the recurrence at all.  Here a second warp of the CTA does that part, one 32-sample group ahead: it streams
// theta through the cp.async ring, computes e with exactly the same operations, and hands it (and the group's last theta,
// for the re-anchor) to the recurrence warp through two shared-memory buffers guarded by named barriers (full / empty per
// buffer).  The recurrence warp's arithmetic is unchanged, so the results are bit-identical to the single-warp fast pass
// (tests/test_gpu_round2.py).  Measured: it pays when the CTAs have an SM each -- 100 stations (4 CTAs): config-4 step
// 0.181 -> 0.144 ms -- and does not when two or more pairs share an SM: 1024 streams (32 CTAs on the 16-SM recurrence
// partition) 0.176 against 0.171 ms alone and 0.28 against 0.205 ms next to K5 / K6; two pairs in one 128-thread CTA
// likewise.  launch_k3 therefore uses it only while the grid fits the partition one CTA per SM (p.exact = 0: automatic).
// If any lane of the CTA is poisoned (non-finite AGC gain) the CTA falls back to the exact body for every group (rare:
// an all-zero input block), which keeps the barriers warp-uniform.
constexpr int K3D_PAIRS = 1;                                        // (recurrence, helper) warp pairs per CTA (2 pairs per CTA measured slower)
constexpr int K3D_THREADS = 64 * K3D_PAIRS;
constexpr int K3D_SMEM = K3D_PAIRS * (K3_RING * 32 * 128 + 2 * 32 * 128 + 2 * 32 * 4);
template <bool KEEP>
__global__ void __launch_bounds__(K3D_THREADS)
k3_pll_duo(const float* __restrict__ theta, const float* __restrict__ power, float* __restrict__ state,
           float* __restrict__ pll_dt, float* __restrict__ dbg_raw, float* __restrict__ dbg_pi,
           const __grid_constant__ K3Params p)
{
    constexpr int WRAP = 0, G = 32;
    extern __shared__ __align__(128) unsigned char k3d_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);             // warp-uniform
    const bool helper = warp >= K3D_PAIRS;
    const int pair = helper ? warp - K3D_PAIRS : warp;
    unsigned char* my = k3d_smem + pair * (K3D_SMEM / K3D_PAIRS);
    float4* s_ring = (float4*)my;                                   // theta: [slot][lane][8 chunks of 16 bytes] (helper warp)
    float4* s_e = (float4*)(my + K3_RING * 32 * 128);               // e:     [buffer][lane][8 chunks], same swizzle
    float (*s_thl)[32] = (float (*)[32])(my + K3_RING * 32 * 128 + 2 * 32 * 128);      // theta of the group's last sample
    const int S = p.n_streams;
    const int s_raw = (blockIdx.x * K3D_PAIRS + pair) * 32 + lane;
    const bool live = s_raw < S;
    const int s = live ? s_raw : S - 1;                             // dead lanes shadow the last stream and store nothing
    float x1 = state[PLL_LPF_X1 * S + s], y1 = state[PLL_LPF_Y1 * S + s], integ = state[PLL_INT * S + s];
    float t = state[PLL_T * S + s], w = state[PLL_E_PREV * S + s], gain = state[PLL_AGC_GAIN * S + s];
    float th_last = state[PLL_TH_PREV * S + s];
    const float avg_power = power[s] / (float)p.n;
    const float target_gain = sqrtf(p.agc_target / avg_power);
    gain = gain + p.agc_beta * (target_gain - gain);
    const bool poisoned = !(fabsf(gain) <= 3.0e38f);
    const bool any_poisoned = __syncthreads_or(poisoned ? 1 : 0) != 0;

    const float4* th4 = (const float4*)(theta + (size_t)s * p.n);
    float4* dt4 = (float4*)(pll_dt + (size_t)s * p.n);
    float4* raw4 = KEEP ? (float4*)(dbg_raw + (size_t)s * p.n) : nullptr;
    float4* pi4 = KEEP ? (float4*)(dbg_pi + (size_t)s * p.n) : nullptr;
    const float a0 = p.lpf_a[0], Kp = p.Kp;
    const float b0t = p.lpf_b[0] * TWO_PI_F, b1t = p.lpf_b[1] * TWO_PI_F, ci = p.int_KTs * TWO_PI_F;
    const float f_gain = p.f_gain, f_center = p.f_center, mixer_KTs = p.mixer_KTs;
    const float d_nom = mixer_KTs * f_center;
    const int n_groups = p.n / G;
    const unsigned e_sa = (unsigned)__cvta_generic_to_shared((char*)s_e + lane * 128);
    auto bar_sync = [](int id) { asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory"); };
    auto bar_arrive = [](int id) { asm volatile("bar.arrive %0, 64;" :: "r"(id) : "memory"); };
    const int FULL = 1 + 4 * pair, EMPTY = 3 + 4 * pair;            // barrier ids FULL + b, EMPTY + b, per pair

    if (helper) {
        if (any_poisoned) return;
        char* ring = (char*)s_ring + lane * 128;
        const unsigned ring_sa = (unsigned)__cvta_generic_to_shared(ring);
        auto issue = [&](int g) {
            if (g < n_groups) {
                const char* src = (const char*)(th4 + (size_t)g * (G / 4));
                const unsigned dst = ring_sa + (unsigned)(g % K3_RING) * (32 * 128);
#pragma unroll
                for (int c = 0; c < 8; c++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst + (unsigned)((c ^ (lane & 7)) << 4)), "l"(src + c * 16));
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int g = 0; g < K3_RING - 1; g++) issue(g);
        float thp = th_last;
        for (int g = 0; g < n_groups; g++) {
            const int b = g & 1;
            asm volatile("cp.async.wait_group %0;" :: "n"(K3_RING - 2) : "memory");      // group g has landed
            float4 cur[G / 4];
            const char* row = ring + (g % K3_RING) * (32 * 128);
#pragma unroll
            for (int c = 0; c < 8; c++) cur[c] = *(const float4*)(row + ((c ^ (lane & 7)) << 4));
            issue(g + K3_RING - 1);                                  // into the slot just read
            float4 ev[G / 4];
#pragma unroll
            for (int q = 0; q < G / 4; q++) {
                const float th[4] = { cur[q].x, cur[q].y, cur[q].z, cur[q].w };
                float e[4];
#pragma unroll
                for (int j = 0; j < 4; j++) { e[j] = wrap_turn<WRAP>((th[j] - thp) + d_nom); thp = th[j]; }
                ev[q] = make_float4(e[0], e[1], e[2], e[3]);
            }
            if (g >= 2) bar_sync(EMPTY + b);                          // the recurrence warp has read group g - 2 out of buffer b
#pragma unroll
            for (int c = 0; c < 8; c++)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(e_sa + (unsigned)(b * 32 * 128 + ((c ^ (lane & 7)) << 4))),
                             "f"(ev[c].x), "f"(ev[c].y), "f"(ev[c].z), "f"(ev[c].w) : "memory");
            s_thl[b][lane] = thp;
            __threadfence_block();
            bar_arrive(FULL + b);
        }
        return;
    }

    // ---------------- recurrence warp ----------------
    if (!any_poisoned) {
        const float cw = fmaf(Kp, b1t, ci), kfg = mixer_KTs * f_gain;
        for (int g = 0; g < n_groups; g++) {
            const int b = g & 1, i = g * G;
            bar_sync(FULL + b);
            float4 cur[G / 4];
#pragma unroll
            for (int c = 0; c < 8; c++)
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cur[c].x), "=f"(cur[c].y), "=f"(cur[c].z), "=f"(cur[c].w)
                             : "r"(e_sa + (unsigned)(b * 32 * 128 + ((c ^ (lane & 7)) << 4))) : "memory");
            const float thl = s_thl[b][lane];
            if (g + 2 < n_groups) bar_arrive(EMPTY + b);              // buffer b may take group g + 2
            const float sx1 = x1, sy1 = y1, sinteg = integ, st = t, sw = w;
            float wmax = 0.0f, t_before_last = t, freq_last = f_center;
#pragma unroll
            for (int q = 0; q < G / 4; q++) {
                const float e[4] = { cur[q].x, cur[q].y, cur[q].z, cur[q].w };
                float dt[4], raw[4], pie[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float m = fmaf(x1, b0t, y1 * a0);
                    const float gg = fmaf(m, Kp, integ);
                    const float control = fmaf(w, cw, gg);
                    const float lpf = fmaf(w, b1t, m);
                    integ = fmaf(ci, w, integ);
                    x1 = w; y1 = lpf;
                    const float freq = fmaf(control, f_gain, f_center);
                    t_before_last = t; freq_last = freq;
                    t = wrap_turn<WRAP>(fmaf(mixer_KTs, freq, t));
                    w = fmaf(control, kfg, w + e[j]);
                    wmax = fmaxf(wmax, fabsf(w));
                    dt[j] = t;
                    if (KEEP) { raw[j] = w * TWO_PI_F; pie[j] = control; }
                }
                if (live) {
                    dt4[(i >> 2) + q] = make_float4(dt[0], dt[1], dt[2], dt[3]);
                    if (KEEP) {
                        raw4[(i >> 2) + q] = make_float4(raw[0], raw[1], raw[2], raw[3]);
                        pi4[(i >> 2) + q] = make_float4(pie[0], pie[1], pie[2], pie[3]);
                    }
                }
            }
            // re-anchor the detector on the exact expression of the group's last sample
            w = wrap_turn<WRAP>(fmaf(mixer_KTs, freq_last, thl + t_before_last));
            const bool redo = !(wmax < 0.25f) || !(fabsf(integ) <= p.integ_safe);
            th_last = thl;
            if (redo) {
                x1 = sx1; y1 = sy1; integ = sinteg; t = st; w = sw;
                k3_exact_group<KEEP, WRAP>(th4, dt4, raw4, pi4, i, live, x1, y1, integ, t, w, a0, Kp, b0t, b1t, ci, f_gain, f_center, mixer_KTs);
            }
        }
    } else if (!poisoned) {
        for (int i = 0; i < p.n; i += G)
            k3_exact_group<KEEP, WRAP>(th4, dt4, raw4, pi4, i, live, x1, y1, integ, t, w, a0, Kp, b0t, b1t, ci, f_gain, f_center, mixer_KTs);
        th_last = theta[(size_t)s * p.n + p.n - 1];
    } else {
        // Poisoned lane: as in k3_pll (reference operation order, clamp(NaN) = -1)
        const float nan = 0.0f * gain;
        float x1r = x1 * TWO_PI_F;
        for (int i = 0; i < p.n; i++) {
            const float e = fmaf(w, TWO_PI_F, nan);
            const float lpf = fmaf(e, p.lpf_b[1], fmaf(x1r, p.lpf_b[0], y1 * a0));
            x1r = e; y1 = lpf;
            integ = clampf(fmaf(p.int_KTs, e, integ), -1.0f, 1.0f);
            const float PI_error = fmaf(lpf, Kp, integ);
            const float control = clampf(PI_error, -1.0f, 1.0f);
            const float tu = fmaf(p.mixer_KTs, fmaf(control, p.f_gain, p.f_center), t);
            t = tu - roundf(tu);
            if (live) {
                pll_dt[(size_t)s * p.n + i] = t;
                if (KEEP) { dbg_raw[(size_t)s * p.n + i] = e; dbg_pi[(size_t)s * p.n + i] = PI_error; }
            }
        }
        x1 = nan; w = nan;
        th_last = theta[(size_t)s * p.n + p.n - 1];
    }
    if (live) {
        state[PLL_LPF_X1 * S + s] = x1;
        state[PLL_LPF_Y1 * S + s] = y1;
        state[PLL_INT * S + s] = integ;
        state[PLL_T * S + s] = t;
        state[PLL_E_PREV * S + s] = w;
        state[PLL_AGC_GAIN * S + s] = gain;
        state[PLL_TH_PREV * S + s] = th_last;
    }
}

cudaError_t launch_k3(const float* theta, const float* power, float* state, float* pll_dt,
                      float* dbg_raw, float* dbg_pi, const K3Params& p_in, cudaStream_t st)
{
    K3Params p = p_in;
    const int grid = (p.n_streams + 31) / 32;
    {
        // Bounds for the speculative pass.  The phase error is wrapped to |w| <= 1/2 turn, so in the kernel's
        // units |lpf| <= (|b0| + |b1|) * pi / (1 - |a0|) and the integrator moves by at most int_KTs * pi per
        // sample.  If |integ| <= integ_safe after a 32-sample group, then |integ| <= 1 and |Kp*lpf + integ| <= 1
        // held at every sample of it (1e-3 of margin covers the roundings).
        const double pi = 3.14159265358979323846;
        const double a0 = std::fabs((double)p.lpf_a[0]);
        const double lpf_max = a0 < 1.0 ? (std::fabs((double)p.lpf_b[0]) + std::fabs((double)p.lpf_b[1])) * pi / (1.0 - a0) : 1e30;
        const double safe = 1.0 - std::fabs((double)p.Kp) * lpf_max - 32.0 * std::fabs((double)p.int_KTs) * pi - 1e-3;
        p.integ_safe = safe > 0.0 ? (float)safe : 0.0f;
    }
    if (p.exact == 1) {
        if (p.keep) k3_pll<true, 0, false><<<grid, 32, 0, st>>>(theta, power, state, pll_dt, dbg_raw, dbg_pi, p);
        else        k3_pll<false, 0, false><<<grid, 32, 0, st>>>(theta, power, state, pll_dt, dbg_raw, dbg_pi, p);
    } else {
        const int grid2 = (p.n_streams + 32 * K3D_PAIRS - 1) / (32 * K3D_PAIRS);
        const bool duo = p.exact == 3 || (p.exact == 0 && grid2 <= std::max(1, p.rec_sms));
        if (!duo) {                                  // the fast pass in one warp
            if (p.keep) k3_pll<true, 0, true><<<grid, 32, 0, st>>>(theta, power, state, pll_dt, dbg_raw, dbg_pi, p);
            else        k3_pll<false, 0, true><<<grid, 32, 0, st>>>(theta, power, state, pll_dt, dbg_raw, dbg_pi, p);
        } else {
            static std::atomic<bool> configured[64];     // the attribute is per device
            int dev = 0;
            cudaGetDevice(&dev);
            if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
                cudaFuncSetAttribute(k3_pll_duo<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K3D_SMEM);
                cudaFuncSetAttribute(k3_pll_duo<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K3D_SMEM);
                if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
            }
            if (p.keep) k3_pll_duo<true><<<grid2, K3D_THREADS, K3D_SMEM, st>>>(theta, power, state, pll_dt, dbg_raw, dbg_pi, p);
            else        k3_pll_duo<false><<<grid2, K3D_THREADS, K3D_SMEM, st>>>(theta, power, state, pll_dt, dbg_raw, dbg_pi, p);
        }
    }
    return cudaGetLastError();
}

} // namespace fm
