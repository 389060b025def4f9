// K2: fm_demod (256 kS/s) -> MPX at 128 kS/s -> analytic MPX -> 19 kHz pilot resonator -> pilot angle.
//
// Replaces (reference file:line under /root/reference/src):
//   PolyphaseDownsampler<float>::process, M=2 K=32 NN=64     dsp/polyphase_filter.h:41-64,190-195
//   optional de-emphasis IIR_Filter<float> K=2               dsp/iir_filter.h:40-69, broadcast_fm_demod.cpp:404-406
//   Hilbert_FIR_Filter<float>::process, 65 taps              dsp/hilbert_fir_filter.h:26-46
//   IIR_Filter<complex<float>> K=3 (pilot peak filter)       dsp/iir_filter.h:40-69, broadcast_fm_demod.cpp:421
//   AGC_Filter::calculate_average_power (sum only)           dsp/agc.h:21-30
// and adds the B200-first step that takes atan2 OFF the PLL's serial critical path:
//   theta[n] = arg(pilot[n]) / 2 pi  (turns)
// The reference PLL computes arg(pilot[n] * pll[n]) inside its per-sample feedback loop
// (broadcast_fm_demod.cpp:449-450); since arg(a*b) = arg(a) + arg(b) and arg(pll[n]) = 2 pi t[n]
// up to the polynomial sine's 4e-8 error, K3 only needs theta[n] + t[n] wrapped to one turn.
// AGC scaling (a positive real gain) does not change the angle, so the pilot is not scaled here.
//
// PACKED FP32.  The signals here are REAL, so unlike K1 (complex sample x real tap) there is no
// natural pair for sm_100's two-wide FP32 instructions -- and the first version of this kernel,
// scalar FFMA with every LDS / shuffle / address instruction stealing an FMA issue slot, reached
// 23 % of the FP32 peak (ncu: issue slots 53 % busy, FMA pipe 37 %).  This version makes the pair
// out of TWO STREAMS: one CTA walks streams 2p and 2p+1 together and every value is a float2
// (stream 2p, stream 2p+1).  Each tap is one FFMA2 whose tap operand is a uniform-register scalar
// broadcast to both halves; the shared-memory windows are LDS.128 of two pairs; the linear-recurrence
// scans, the atan2 polynomial and the power sum are packed the same way.  Per stream the arithmetic
// (operation order, roundings) is exactly the scalar program's.  An odd last stream is paired with
// itself and its second half discarded.
//
// One CTA per stream pair walks the block in chunks of 1024 MPX samples, carrying all filter state
// in shared memory / registers, so the IIR state carry across chunks is exact.  FIRs are register
// tiled (8 consecutive outputs per thread, sliding LDS.128 windows over bank-skewed arrays).  The
// Hilbert transformer's even taps are exactly zero as designed (filter_designer.cpp:369-383): the
// host checks that and selects the variant that skips them (33 instead of 65 products).
//
// The two LINEAR recurrences (pilot resonator, pole radius 0.9999; optional de-emphasis pole) run
// block-parallel as a linear-recurrence scan: each thread runs its 8 samples from a zero state,
// the 128 end states are combined with a Kogge-Stone scan over warp shuffles using host-computed
// powers of the state matrix (A^8, A^16, ... A^256), and each thread then re-runs its 8 samples
// from its true start state, so inside a thread the arithmetic is the sequential recurrence.
#include "fm_common.cuh"

namespace fm {

// bank-skewed shared layouts in float2 (pair) units: s_in pads 2 pairs per 16, s_out 2 pairs per 8,
// so the per-thread LDS.128 windows (thread stride 16 resp. 8 pairs) are conflict free.
__device__ __forceinline__ int a_in(int i) { return i + 2 * (i >> 4); }
__device__ __forceinline__ int a_out(int i) { return i + 2 * (i >> 3); }

constexpr int K2_S_IN = (K2_NN + 2 * K2_CH) + 2 * ((K2_NN + 2 * K2_CH) >> 4) + 8;
constexpr int K2_S_OUT = (64 + K2_CH) + 2 * ((64 + K2_CH) >> 3) + 8;
constexpr int K2_WARPS = K2_THREADS / 32;

__device__ __forceinline__ float2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 shfl_up2(float2 v, int d) {
    return make_float2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}

template <bool SPARSE_HILBERT>
__global__ void __launch_bounds__(K2_THREADS, 4)
k2_mpx(const float* __restrict__ fm_demod, float* __restrict__ hist_demod, float* __restrict__ hist_out,
       float* __restrict__ scal, float2* __restrict__ fm_out_iq, float* __restrict__ theta,
       float* __restrict__ power, float2* __restrict__ pilot_dbg, const __grid_constant__ K2Params p, int n_streams)
{
    __shared__ __align__(16) float2 s_in[K2_S_IN];        // [64 hist + 2048] fm_demod pairs, skewed
    __shared__ __align__(16) float2 s_out[K2_S_OUT];      // [64 hist + 1024] fm_out pairs, skewed
    __shared__ float2 s_red[K2_WARPS];                    // warp partial powers
    __shared__ float2 s_wt[K2_WARPS][4];                  // warp totals of the scans
    __shared__ float2 s_edge[K2_WARPS][4];                // last two Hilbert outputs of each warp: re, im of x[-2], x[-1]
    __shared__ float2 s_st[12];                           // carried recurrence states
    __shared__ float4 s_Q[32];                            // A^(8 lane): per-lane operand, so not a constant-bank read
    __shared__ float s_dQ[32];                            // alpha^(8 lane)
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const int sA = 2 * blockIdx.x;
    const bool hasB = sA + 1 < n_streams;
    const int sB = hasB ? sA + 1 : sA;

    if (t < 32) { s_Q[t] = make_float4(p.pk_Q[t][0], p.pk_Q[t][1], p.pk_Q[t][2], p.pk_Q[t][3]); s_dQ[t] = p.de_Q[t]; }
    if (t < K2_NN) s_in[a_in(t)] = make_float2(hist_demod[(size_t)sA * K2_NN + t], hist_demod[(size_t)sB * K2_NN + t]);
    if (t < 64) s_out[a_out(t)] = make_float2(hist_out[(size_t)sA * 64 + t], hist_out[(size_t)sB * 64 + t]);
    float* scA = scal + (size_t)sA * K2_SCAL_N;
    float* scB = scal + (size_t)sB * K2_SCAL_N;
    // carried state: s_st[0..3] = pilot y1r, y2r, y1i, y2i; [4], [5] = de-emphasis x1, y1;
    //                s_st[8..11] = pilot filter inputs x[n-2] re, im, x[n-1] re, im
    if (t < 12) {
        const int idx[12] = { K2_PK_Y1R, K2_PK_Y2R, K2_PK_Y1I, K2_PK_Y2I, K2_DEEMPH_X1, K2_DEEMPH_Y1, 0, 0,
                              K2_PK_X2R, K2_PK_X2I, K2_PK_X1R, K2_PK_X1I };
        s_st[t] = (t == 6 || t == 7) ? make_float2(0.f, 0.f) : make_float2(scA[idx[t]], scB[idx[t]]);
    }
    float2 power_total = make_float2(0.0f, 0.0f);               // thread 0 only

    for (int c0 = 0; c0 < p.n_out; c0 += K2_CH) {
        const int nch = min(K2_CH, p.n_out - c0);               // multiple of 128
        const bool active = (t * K2_R) < nch;
        __syncthreads();
        // ---- A: stage 2*nch fm_demod samples of both streams behind the 64-sample history ----
        {
            const float4* srcA = (const float4*)(fm_demod + (size_t)sA * 2 * p.n_out + 2 * (size_t)c0);
            const float4* srcB = (const float4*)(fm_demod + (size_t)sB * 2 * p.n_out + 2 * (size_t)c0);
            for (int q = t; q < (2 * nch) / 4; q += K2_THREADS) {
                const float4 a = __ldg(srcA + q), b = __ldg(srcB + q);
                float2* d = s_in + a_in(K2_NN + 4 * q);
                *(float4*)(d) = make_float4(a.x, b.x, a.y, b.y);
                *(float4*)(d + 2) = make_float4(a.z, b.z, a.w, b.w);
            }
        }
        __syncthreads();
        // ---- B: /2 FIR, output o = 8t+r: sum_k b[k] * s_in[16t + 2r + 2 + k] ----
        if (active) {
            float2 acc[K2_R];
#pragma unroll
            for (int r = 0; r < K2_R; r++) acc[r] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int g = 0; g < 5; g++) {
                const float2* wp = s_in + a_in(16 * t + 16 * g);
                float2 w[16];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const float4 v = *(const float4*)(wp + 2 * q);
                    w[2 * q] = make_float2(v.x, v.y); w[2 * q + 1] = make_float2(v.z, v.w);
                }
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    const int n = 16 * g + m;
#pragma unroll
                    for (int r = 0; r < K2_R; r++) {
                        const int k = n - 2 * r - 2;
                        if (k >= 0 && k < K2_NN) acc[r] = __ffma2_rn(w[m], bc(p.taps_fm_out[k]), acc[r]);
                    }
                }
            }
            float2* d = s_out + a_out(64 + K2_R * t);
#pragma unroll
            for (int q = 0; q < K2_R / 2; q++) *(float4*)(d + 2 * q) = make_float4(acc[2 * q].x, acc[2 * q].y, acc[2 * q + 1].x, acc[2 * q + 1].y);
        }
        __syncthreads();
        // carry the fm_demod history: last 64 staged inputs -> front (read now, write after sync)
        float2 carry_in = make_float2(0.0f, 0.0f);
        if (t < K2_NN) carry_in = s_in[a_in(2 * nch + t)];
        // ---- B2: optional de-emphasis, y[n] = b[0]x[n-1] + a[0]y[n-1] + b[1]x[n] (iir_filter.h:62-68),
        //      as a first-order linear-recurrence scan over the chunk ----
        if (p.use_deemph) {
            float2 xv[K2_R + 1];
            float2 cz = make_float2(0.0f, 0.0f);
            const float2 y_carried = s_st[5];       // read before the barrier: rewritten after it
            const float2 b0 = bc(p.deemph_b[0]), b1 = bc(p.deemph_b[1]), a0 = bc(p.deemph_a[0]);
            if (active) {
                const float2* d = s_out + a_out(64 + K2_R * t);
                xv[0] = (t == 0) ? s_st[4] : s_out[a_out(64 + K2_R * t - 1)];
#pragma unroll
                for (int q = 0; q < K2_R / 2; q++) {
                    const float4 v = *(const float4*)(d + 2 * q);
                    xv[1 + 2 * q] = make_float2(v.x, v.y); xv[2 + 2 * q] = make_float2(v.z, v.w);
                }
                float2 y1 = (t == 0) ? y_carried : make_float2(0.0f, 0.0f);
#pragma unroll
                for (int j = 0; j < K2_R; j++) y1 = __ffma2_rn(y1, a0, __ffma2_rn(xv[j], b0, __fmul2_rn(xv[j + 1], b1)));
                cz = y1;
            }
#pragma unroll
            for (int l = 0; l < 5; l++) {
                const float2 o = shfl_up2(cz, 1 << l);
                if (lane >= (1 << l)) cz = __ffma2_rn(bc(p.de_P[l]), o, cz);
            }
            if (lane == 31) s_wt[warp][0] = cz;
            __syncthreads();                        // also orders every read of pre-filter x before the writes below
            float2 g = make_float2(0.0f, 0.0f);
            for (int w = 0; w < warp; w++) g = __ffma2_rn(bc(p.de_P[5]), g, s_wt[w][0]);
            float2 y1 = shfl_up2(cz, 1);
            if (lane == 0) y1 = make_float2(0.0f, 0.0f);
            y1 = __ffma2_rn(bc(s_dQ[lane]), g, y1);          // true y[8t-1]
            if (t == 0) y1 = y_carried;
            if (active) {
                float2 yv[K2_R];
#pragma unroll
                for (int j = 0; j < K2_R; j++) { y1 = __ffma2_rn(y1, a0, __ffma2_rn(xv[j], b0, __fmul2_rn(xv[j + 1], b1))); yv[j] = y1; }
                float2* d = s_out + a_out(64 + K2_R * t);
#pragma unroll
                for (int q = 0; q < K2_R / 2; q++) *(float4*)(d + 2 * q) = make_float4(yv[2 * q].x, yv[2 * q].y, yv[2 * q + 1].x, yv[2 * q + 1].y);
                if (K2_R * (t + 1) == nch) { s_st[4] = xv[K2_R]; s_st[5] = y1; }
            }
            __syncthreads();
        }
        // ---- C: Hilbert, imag[o] = sum_k b[k]*s_out[o+k], real[o] = s_out[o+32] ----
        float2 xre[K2_R], xim[K2_R];
#pragma unroll
        for (int r = 0; r < K2_R; r++) { xre[r] = make_float2(0.0f, 0.0f); xim[r] = make_float2(0.0f, 0.0f); }
        if (active) {
#pragma unroll
            for (int g = 0; g < 9; g++) {
                const float2* wp = s_out + a_out(8 * t + 8 * g);
                float2 w[8];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 v = *(const float4*)(wp + 2 * q);
                    w[2 * q] = make_float2(v.x, v.y); w[2 * q + 1] = make_float2(v.z, v.w);
                }
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    const int n = 8 * g + m;
#pragma unroll
                    for (int r = 0; r < K2_R; r++) {
                        const int k = n - r;
                        if (k >= 0 && k < K2_HILB && (!SPARSE_HILBERT || (k & 1)))
                            xim[r] = __ffma2_rn(w[m], bc(p.taps_hilbert[k]), xim[r]);
                        if (k == 32) xre[r] = w[m];
                    }
                }
            }
            float4* gA = (float4*)(fm_out_iq + (size_t)sA * p.n_out + c0 + K2_R * t);
            float4* gB = (float4*)(fm_out_iq + (size_t)sB * p.n_out + c0 + K2_R * t);
#pragma unroll
            for (int q = 0; q < K2_R / 2; q++) {
                gA[q] = make_float4(xre[2 * q].x, xim[2 * q].x, xre[2 * q + 1].x, xim[2 * q + 1].x);
                if (hasB) gB[q] = make_float4(xre[2 * q].y, xim[2 * q].y, xre[2 * q + 1].y, xim[2 * q + 1].y);
            }
        }
        // the two samples before this thread's first one: previous lane / previous warp / carried
        float2 pm2r = shfl_up2(xre[K2_R - 2], 1), pm2i = shfl_up2(xim[K2_R - 2], 1);
        float2 pm1r = shfl_up2(xre[K2_R - 1], 1), pm1i = shfl_up2(xim[K2_R - 1], 1);
        if (lane == 31) { s_edge[warp][0] = xre[K2_R - 2]; s_edge[warp][1] = xim[K2_R - 2]; s_edge[warp][2] = xre[K2_R - 1]; s_edge[warp][3] = xim[K2_R - 1]; }
        __syncthreads();
        // carry histories for the next chunk / block
        if (t < K2_NN) s_in[a_in(t)] = carry_in;
        float2 carry_out = make_float2(0.0f, 0.0f);
        if (t < 64) carry_out = s_out[a_out(nch + t)];
        if (lane == 0) {
            if (warp == 0) { pm2r = s_st[8]; pm2i = s_st[9]; pm1r = s_st[10]; pm1i = s_st[11]; }
            else { pm2r = s_edge[warp - 1][0]; pm2i = s_edge[warp - 1][1]; pm1r = s_edge[warp - 1][2]; pm1i = s_edge[warp - 1][3]; }
        }
        // ---- D: pilot peak filter (iir_filter.h:62-68, K = 3; a[2] multiplies the always-zero yn[K-1]):
        //   y[n] = u[n] + a[1] y[n-1] + a[0] y[n-2],  u[n] = b[0]x[n-2] + b[1]x[n-1] + b[2]x[n]
        // as a second-order linear-recurrence scan, real and imaginary parts (and both streams) side by side ----
        float2 yr[K2_R], yi[K2_R];
        {
            const float2 b0 = bc(p.peak_b[0]), b1 = bc(p.peak_b[1]), b2 = bc(p.peak_b[2]);
            const float2 a0 = bc(p.peak_a[0]), a1 = bc(p.peak_a[1]);
            float2 ur[K2_R], ui[K2_R];
            float2 c1r = bc(0.f), c2r = bc(0.f), c1i = bc(0.f), c2i = bc(0.f);     // (y[n-1], y[n-2]) after this thread's samples
            if (active) {
#pragma unroll
                for (int j = 0; j < K2_R; j++) {
                    const float2 x0r = (j >= 2) ? xre[j - 2] : (j == 1 ? pm1r : pm2r), x0i = (j >= 2) ? xim[j - 2] : (j == 1 ? pm1i : pm2i);
                    const float2 x1r = (j >= 1) ? xre[j - 1] : pm1r, x1i = (j >= 1) ? xim[j - 1] : pm1i;
                    ur[j] = __ffma2_rn(x0r, b0, __ffma2_rn(x1r, b1, __fmul2_rn(xre[j], b2)));
                    ui[j] = __ffma2_rn(x0i, b0, __ffma2_rn(x1i, b1, __fmul2_rn(xim[j], b2)));
                }
                if (t == 0) { c1r = s_st[0]; c2r = s_st[1]; c1i = s_st[2]; c2i = s_st[3]; }
#pragma unroll
                for (int j = 0; j < K2_R; j++) {
                    const float2 nr = __ffma2_rn(c1r, a1, __ffma2_rn(c2r, a0, ur[j]));
                    const float2 ni = __ffma2_rn(c1i, a1, __ffma2_rn(c2i, a0, ui[j]));
                    c2r = c1r; c1r = nr; c2i = c1i; c1i = ni;
                }
            }
#pragma unroll
            for (int l = 0; l < 5; l++) {
                const float2 o1r = shfl_up2(c1r, 1 << l), o2r = shfl_up2(c2r, 1 << l);
                const float2 o1i = shfl_up2(c1i, 1 << l), o2i = shfl_up2(c2i, 1 << l);
                if (lane >= (1 << l)) {
                    const float2 P0 = bc(p.pk_P[l][0]), P1 = bc(p.pk_P[l][1]), P2 = bc(p.pk_P[l][2]), P3 = bc(p.pk_P[l][3]);
                    c1r = __fadd2_rn(c1r, __ffma2_rn(P0, o1r, __fmul2_rn(P1, o2r))); c2r = __fadd2_rn(c2r, __ffma2_rn(P2, o1r, __fmul2_rn(P3, o2r)));
                    c1i = __fadd2_rn(c1i, __ffma2_rn(P0, o1i, __fmul2_rn(P1, o2i))); c2i = __fadd2_rn(c2i, __ffma2_rn(P2, o1i, __fmul2_rn(P3, o2i)));
                }
            }
            if (lane == 31) { s_wt[warp][0] = c1r; s_wt[warp][1] = c2r; s_wt[warp][2] = c1i; s_wt[warp][3] = c2i; }
            __syncthreads();
            float2 g1r = bc(0.f), g2r = bc(0.f), g1i = bc(0.f), g2i = bc(0.f);      // state at the start of this warp
            {
                const float2 P0 = bc(p.pk_P[5][0]), P1 = bc(p.pk_P[5][1]), P2 = bc(p.pk_P[5][2]), P3 = bc(p.pk_P[5][3]);
                for (int w = 0; w < warp; w++) {
                    const float2 n1r = __fadd2_rn(__ffma2_rn(P0, g1r, __fmul2_rn(P1, g2r)), s_wt[w][0]), n2r = __fadd2_rn(__ffma2_rn(P2, g1r, __fmul2_rn(P3, g2r)), s_wt[w][1]);
                    const float2 n1i = __fadd2_rn(__ffma2_rn(P0, g1i, __fmul2_rn(P1, g2i)), s_wt[w][2]), n2i = __fadd2_rn(__ffma2_rn(P2, g1i, __fmul2_rn(P3, g2i)), s_wt[w][3]);
                    g1r = n1r; g2r = n2r; g1i = n1i; g2i = n2i;
                }
            }
            // true start state of this thread = (previous lane's inclusive prefix) + A^(8 lane) * warp start
            float2 e1r = shfl_up2(c1r, 1), e2r = shfl_up2(c2r, 1), e1i = shfl_up2(c1i, 1), e2i = shfl_up2(c2i, 1);
            if (lane == 0) { e1r = bc(0.f); e2r = bc(0.f); e1i = bc(0.f); e2i = bc(0.f); }
            const float4 Q = s_Q[lane];
            e1r = __fadd2_rn(e1r, __ffma2_rn(bc(Q.x), g1r, __fmul2_rn(bc(Q.y), g2r))); e2r = __fadd2_rn(e2r, __ffma2_rn(bc(Q.z), g1r, __fmul2_rn(bc(Q.w), g2r)));
            e1i = __fadd2_rn(e1i, __ffma2_rn(bc(Q.x), g1i, __fmul2_rn(bc(Q.y), g2i))); e2i = __fadd2_rn(e2i, __ffma2_rn(bc(Q.z), g1i, __fmul2_rn(bc(Q.w), g2i)));
            if (t == 0) { e1r = s_st[0]; e2r = s_st[1]; e1i = s_st[2]; e2i = s_st[3]; }
            __syncthreads();                                         // everyone has read s_st / s_wt / s_edge
            if (active) {
#pragma unroll
                for (int j = 0; j < K2_R; j++) {
                    yr[j] = __ffma2_rn(e1r, a1, __ffma2_rn(e2r, a0, ur[j]));
                    yi[j] = __ffma2_rn(e1i, a1, __ffma2_rn(e2i, a0, ui[j]));
                    e2r = e1r; e1r = yr[j]; e2i = e1i; e1i = yi[j];
                }
                if (K2_R * (t + 1) == nch) {                         // last sample of the chunk: carry
                    s_st[0] = e1r; s_st[1] = e2r; s_st[2] = e1i; s_st[3] = e2i;
                    s_st[8] = xre[K2_R - 2]; s_st[9] = xim[K2_R - 2]; s_st[10] = xre[K2_R - 1]; s_st[11] = xim[K2_R - 1];
                }
            }
        }
        if (t < 64) s_out[a_out(t)] = carry_out;
        // ---- E: pilot angle in turns + |y|^2 partial sums ----
        float2 pw = make_float2(0.0f, 0.0f);
        if (active) {
            float2 th[K2_R];
#pragma unroll
            for (int j = 0; j < K2_R; j++) {
                th[j] = __fmul2_rn(fm_atan2f_x2(yi[j].x, yr[j].x, yi[j].y, yr[j].y), bc(INV_TWO_PI_F));
                pw = __fadd2_rn(pw, __ffma2_rn(yr[j], yr[j], __fmul2_rn(yi[j], yi[j])));
            }
            float4* dA = (float4*)(theta + (size_t)sA * p.n_out + c0 + K2_R * t);
            dA[0] = make_float4(th[0].x, th[1].x, th[2].x, th[3].x);
            dA[1] = make_float4(th[4].x, th[5].x, th[6].x, th[7].x);
            if (hasB) {
                float4* dB = (float4*)(theta + (size_t)sB * p.n_out + c0 + K2_R * t);
                dB[0] = make_float4(th[0].y, th[1].y, th[2].y, th[3].y);
                dB[1] = make_float4(th[4].y, th[5].y, th[6].y, th[7].y);
            }
            if (p.keep) {
                float4* kA = (float4*)(pilot_dbg + (size_t)sA * p.n_out + c0 + K2_R * t);
                float4* kB = (float4*)(pilot_dbg + (size_t)sB * p.n_out + c0 + K2_R * t);
#pragma unroll
                for (int q = 0; q < K2_R / 2; q++) {
                    kA[q] = make_float4(yr[2 * q].x, yi[2 * q].x, yr[2 * q + 1].x, yi[2 * q + 1].x);
                    if (hasB) kB[q] = make_float4(yr[2 * q].y, yi[2 * q].y, yr[2 * q + 1].y, yi[2 * q + 1].y);
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            pw.x += __shfl_xor_sync(0xffffffffu, pw.x, off);
            pw.y += __shfl_xor_sync(0xffffffffu, pw.y, off);
        }
        if (lane == 0) s_red[warp] = pw;
        __syncthreads();
        if (t == 0) {
            float2 sum = make_float2(0.0f, 0.0f);
            for (int w = 0; w < K2_WARPS; w++) { sum.x += s_red[w].x; sum.y += s_red[w].y; }
            power_total.x += sum.x; power_total.y += sum.y;
        }
    }
    __syncthreads();
    if (t < K2_NN) {
        const float2 v = s_in[a_in(t)];
        hist_demod[(size_t)sA * K2_NN + t] = v.x;
        if (hasB) hist_demod[(size_t)sB * K2_NN + t] = v.y;
    }
    if (t < 64) {
        const float2 v = s_out[a_out(t)];
        hist_out[(size_t)sA * 64 + t] = v.x;
        if (hasB) hist_out[(size_t)sB * 64 + t] = v.y;
    }
    if (t < 12 && t != 6 && t != 7) {
        const int idx[12] = { K2_PK_Y1R, K2_PK_Y2R, K2_PK_Y1I, K2_PK_Y2I, K2_DEEMPH_X1, K2_DEEMPH_Y1, 0, 0,
                              K2_PK_X2R, K2_PK_X2I, K2_PK_X1R, K2_PK_X1I };
        scA[idx[t]] = s_st[t].x;
        if (hasB) scB[idx[t]] = s_st[t].y;
    }
    if (t == 0) {
        power[sA] = power_total.x;
        if (hasB) power[sB] = power_total.y;
    }
}

cudaError_t launch_k2(const float* fm_demod, float* hist_demod, float* hist_out, float* scal,
                      float2* fm_out_iq, float* theta, float* power, float2* pilot_dbg,
                      const K2Params& p, int n_streams, cudaStream_t st)
{
    // filter_designer.cpp:369-383 leaves every even tap exactly zero; uploaded taps may not
    bool sparse = true;
    for (int k = 0; k < K2_HILB; k += 2) sparse = sparse && (p.taps_hilbert[k] == 0.0f);
    const int grid = (n_streams + 1) / 2;
    if (sparse) k2_mpx<true><<<grid, K2_THREADS, 0, st>>>(fm_demod, hist_demod, hist_out, scal, fm_out_iq, theta, power, pilot_dbg, p, n_streams);
    else        k2_mpx<false><<<grid, K2_THREADS, 0, st>>>(fm_demod, hist_demod, hist_out, scal, fm_out_iq, theta, power, pilot_dbg, p, n_streams);
    return cudaGetLastError();
}

} // namespace fm
