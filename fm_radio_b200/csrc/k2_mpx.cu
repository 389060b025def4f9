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
// One CTA per stream walks the block in chunks of 2048 MPX samples, carrying all filter state in
// shared memory / registers, so the IIR state carry across chunks is exact.  FIRs are register
// tiled (8 consecutive outputs per thread, sliding LDS.128 windows over bank-skewed shared
// arrays, taps as constant-bank operands of fully unrolled FFMAs).
//
// The two LINEAR recurrences (pilot resonator, pole radius 0.9999; optional de-emphasis pole) run
// block-parallel as a linear-recurrence scan: each thread runs its 8 samples from a zero state,
// the 256 end states are combined with a Kogge-Stone scan over warp shuffles using host-computed
// powers of the state matrix (A^8, A^16, ... A^256), and each thread then re-runs its 8 samples
// from its true start state, so inside a thread the arithmetic is the sequential recurrence.
// (Measured on B200, 1024 streams x 8192 samples: two sequential lanes per stream 445 us per
// launch; scan version: see profiles/.)
#include "fm_common.cuh"

namespace fm {

// bank-skewed shared layouts: s_in pads 4 floats per 16, s_out pads 4 floats per 8, so the
// per-thread LDS.128 windows (thread stride 16 resp. 8 floats) are conflict free.
__device__ __forceinline__ int a_in(int i) { return i + 4 * (i >> 4); }
__device__ __forceinline__ int a_out(int i) { return i + 4 * (i >> 3); }

constexpr int K2_S_IN = (K2_NN + 2 * K2_CH) + 4 * ((K2_NN + 2 * K2_CH) >> 4) + 16;
constexpr int K2_S_OUT = (64 + K2_CH) + 4 * ((64 + K2_CH) >> 3) + 16;
constexpr int K2_SMEM_BYTES = (K2_S_IN + K2_S_OUT) * 4 + 2 * K2_CH * 8 + 16 + 256 + 32 * 5 * 4;

__global__ void __launch_bounds__(K2_THREADS, 2)
k2_mpx(const float* __restrict__ fm_demod, float* __restrict__ hist_demod, float* __restrict__ hist_out,
       float* __restrict__ scal, float2* __restrict__ fm_out_iq, float* __restrict__ theta,
       float* __restrict__ power, float2* __restrict__ pilot_dbg, const __grid_constant__ K2Params p)
{
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;                                   // [64 hist + 4096] fm_demod, skewed
    float* s_out = s_in + K2_S_IN;                        // [64 hist + 2048] fm_out, skewed
    float* s_iq = s_out + K2_S_OUT + 4;                   // [-2..2047] x (re,im): two samples of history in front
    float* s_y = s_iq + 2 * K2_CH;                        // [2048] y (re,im)
    float* s_red = s_y + 2 * K2_CH;                       // [8] warp partials
    float* s_wt = s_red + 8;                              // [8][4] warp totals of the scans
    float* s_st = s_wt + 32;                              // [16] carried recurrence states
    float* s_Q = s_st + 16;                               // [32][4] A^(8 lane): per-lane operand, so not a constant-bank read
    float* s_dQ = s_Q + 128;                              // [32] alpha^(8 lane)
    const int t = threadIdx.x;
    const int s = blockIdx.x;
    const int lane = t & 31, warp = t >> 5;

    if (t < 128) s_Q[t] = p.pk_Q[t >> 2][t & 3];
    if (t < 32) s_dQ[t] = p.de_Q[t];
    if (t < K2_NN) s_in[a_in(t)] = hist_demod[(size_t)s * K2_NN + t];
    if (t < 64) s_out[a_out(t)] = hist_out[(size_t)s * 64 + t];
    float* sc = scal + (size_t)s * K2_SCAL_N;
    // carried recurrence state: pilot filter x[n-2], x[n-1] sit in front of s_iq, y[n-1], y[n-2] in s_st
    if (t == 0) {
        s_iq[-4] = sc[K2_PK_X2R]; s_iq[-3] = sc[K2_PK_X2I]; s_iq[-2] = sc[K2_PK_X1R]; s_iq[-1] = sc[K2_PK_X1I];
        s_st[0] = sc[K2_PK_Y1R]; s_st[1] = sc[K2_PK_Y2R]; s_st[2] = sc[K2_PK_Y1I]; s_st[3] = sc[K2_PK_Y2I];
        s_st[4] = sc[K2_DEEMPH_X1]; s_st[5] = sc[K2_DEEMPH_Y1];
    }
    float power_total = 0.0f;                                   // thread 0 only

    for (int c0 = 0; c0 < p.n_out; c0 += K2_CH) {
        const int nch = min(K2_CH, p.n_out - c0);               // multiple of 128
        const bool active = (t * K2_R) < nch;
        __syncthreads();
        // ---- A: stage 2*nch fm_demod samples behind the 64-sample history ----
        {
            const float4* src = (const float4*)(fm_demod + (size_t)s * 2 * p.n_out + 2 * (size_t)c0);
            for (int q = t; q < (2 * nch) / 4; q += K2_THREADS) {
                const float4 v = __ldg(src + q);
                *(float4*)(s_in + a_in(K2_NN + 4 * q)) = v;
            }
        }
        __syncthreads();
        // ---- B: /2 FIR, output o = 8t+r: sum_k b[k] * s_in[16t + 2r + 2 + k] ----
        if (active) {
            float acc[K2_R];
#pragma unroll
            for (int r = 0; r < K2_R; r++) acc[r] = 0.0f;
#pragma unroll
            for (int g = 0; g < 5; g++) {
                const float* wp = s_in + a_in(16 * t + 16 * g);
                float w[16];
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 v = *(const float4*)(wp + 4 * q);
                    w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
                }
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    const int n = 16 * g + m;
#pragma unroll
                    for (int r = 0; r < K2_R; r++) {
                        const int k = n - 2 * r - 2;
                        if (k >= 0 && k < K2_NN) acc[r] = fmaf(w[m], p.taps_fm_out[k], acc[r]);
                    }
                }
            }
            float* d = s_out + a_out(64 + K2_R * t);
            *(float4*)(d) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *(float4*)(d + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        __syncthreads();
        // carry the fm_demod history: last 64 staged inputs -> front (read now, write after sync)
        float carry_in = 0.0f;
        if (t < K2_NN) carry_in = s_in[a_in(2 * nch + t)];
        // ---- B2: optional de-emphasis, y[n] = b[0]x[n-1] + a[0]y[n-1] + b[1]x[n] (iir_filter.h:62-68),
        //      as a first-order linear-recurrence scan over the chunk ----
        if (p.use_deemph) {
            float xv[K2_R + 1];
            float cz = 0.0f;
            const float y_carried = s_st[5];        // read before the barrier: rewritten after it
            if (active) {
                const float* d = s_out + a_out(64 + K2_R * t);
                const float4 v0 = *(const float4*)d, v1 = *(const float4*)(d + 4);
                xv[0] = (t == 0) ? s_st[4] : s_out[a_out(64 + K2_R * t - 1)];
                xv[1] = v0.x; xv[2] = v0.y; xv[3] = v0.z; xv[4] = v0.w; xv[5] = v1.x; xv[6] = v1.y; xv[7] = v1.z; xv[8] = v1.w;
                float y1 = (t == 0) ? y_carried : 0.0f;
#pragma unroll
                for (int j = 0; j < K2_R; j++) y1 = fmaf(y1, p.deemph_a[0], fmaf(xv[j], p.deemph_b[0], xv[j + 1] * p.deemph_b[1]));
                cz = y1;
            }
#pragma unroll
            for (int l = 0; l < 5; l++) {
                const float o = __shfl_up_sync(0xffffffffu, cz, 1 << l);
                if (lane >= (1 << l)) cz = fmaf(p.de_P[l], o, cz);
            }
            if (lane == 31) s_wt[warp] = cz;
            __syncthreads();                        // also orders every read of pre-filter x before the writes below
            float g = 0.0f;
            for (int w = 0; w < warp; w++) g = fmaf(p.de_P[5], g, s_wt[w]);
            float y1 = __shfl_up_sync(0xffffffffu, cz, 1);
            if (lane == 0) y1 = 0.0f;
            y1 = fmaf(s_dQ[lane], g, y1);         // true y[8t-1]
            if (t == 0) y1 = y_carried;
            if (active) {
                float yv[K2_R];
#pragma unroll
                for (int j = 0; j < K2_R; j++) { y1 = fmaf(y1, p.deemph_a[0], fmaf(xv[j], p.deemph_b[0], xv[j + 1] * p.deemph_b[1])); yv[j] = y1; }
                float* d = s_out + a_out(64 + K2_R * t);
                *(float4*)d = make_float4(yv[0], yv[1], yv[2], yv[3]);
                *(float4*)(d + 4) = make_float4(yv[4], yv[5], yv[6], yv[7]);
                if (K2_R * (t + 1) == nch) { s_st[4] = xv[K2_R]; s_st[5] = y1; }
            }
            __syncthreads();
        }
        // ---- C: Hilbert, imag[o] = sum_k b[k]*s_out[o+k], real[o] = s_out[o+32] ----
        if (active) {
            float acc[K2_R], re[K2_R];
#pragma unroll
            for (int r = 0; r < K2_R; r++) acc[r] = 0.0f;
#pragma unroll
            for (int g = 0; g < 9; g++) {
                const float* wp = s_out + a_out(8 * t + 8 * g);
                const float4 v0 = *(const float4*)(wp);
                const float4 v1 = *(const float4*)(wp + 4);
                const float w[8] = { v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w };
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    const int n = 8 * g + m;
#pragma unroll
                    for (int r = 0; r < K2_R; r++) {
                        const int k = n - r;
                        if (k >= 0 && k < K2_HILB) acc[r] = fmaf(w[m], p.taps_hilbert[k], acc[r]);
                        if (k == 32) re[r] = w[m];
                    }
                }
            }
            float4* d = (float4*)(s_iq + 2 * K2_R * t);
            float4* gdst = (float4*)(fm_out_iq + (size_t)s * p.n_out + c0 + K2_R * t);
#pragma unroll
            for (int q = 0; q < K2_R / 2; q++) {
                const float4 v = make_float4(re[2 * q], acc[2 * q], re[2 * q + 1], acc[2 * q + 1]);
                d[q] = v;
                gdst[q] = v;
            }
        }
        __syncthreads();
        // carry histories for the next chunk / block
        if (t < K2_NN) s_in[a_in(t)] = carry_in;
        float carry_out = 0.0f;
        if (t < 64) carry_out = s_out[a_out(nch + t)];
        // ---- D: pilot peak filter (iir_filter.h:62-68, K = 3; a[2] multiplies the always-zero yn[K-1]):
        //   y[n] = u[n] + a[1] y[n-1] + a[0] y[n-2],  u[n] = b[0]x[n-2] + b[1]x[n-1] + b[2]x[n]
        // as a second-order linear-recurrence scan, real and imaginary parts side by side ----
        {
            const float b0 = p.peak_b[0], b1 = p.peak_b[1], b2 = p.peak_b[2];
            const float a0 = p.peak_a[0], a1 = p.peak_a[1];
            float ur[K2_R], ui[K2_R];
            float c1r = 0.f, c2r = 0.f, c1i = 0.f, c2i = 0.f;      // (y[n-1], y[n-2]) after this thread's samples
            if (active) {
                const float4* xs = (const float4*)(s_iq + 2 * K2_R * t - 4);     // x[8t-2 .. 8t+7]
                float xr[K2_R + 2], xi[K2_R + 2];
#pragma unroll
                for (int q = 0; q < (K2_R + 2) / 2; q++) { const float4 v = xs[q]; xr[2 * q] = v.x; xi[2 * q] = v.y; xr[2 * q + 1] = v.z; xi[2 * q + 1] = v.w; }
#pragma unroll
                for (int j = 0; j < K2_R; j++) {
                    ur[j] = fmaf(xr[j], b0, fmaf(xr[j + 1], b1, xr[j + 2] * b2));
                    ui[j] = fmaf(xi[j], b0, fmaf(xi[j + 1], b1, xi[j + 2] * b2));
                }
                if (t == 0) { c1r = s_st[0]; c2r = s_st[1]; c1i = s_st[2]; c2i = s_st[3]; }
#pragma unroll
                for (int j = 0; j < K2_R; j++) {
                    const float yr = fmaf(c1r, a1, fmaf(c2r, a0, ur[j]));
                    const float yi = fmaf(c1i, a1, fmaf(c2i, a0, ui[j]));
                    c2r = c1r; c1r = yr; c2i = c1i; c1i = yi;
                }
            }
#pragma unroll
            for (int l = 0; l < 5; l++) {
                const float o1r = __shfl_up_sync(0xffffffffu, c1r, 1 << l), o2r = __shfl_up_sync(0xffffffffu, c2r, 1 << l);
                const float o1i = __shfl_up_sync(0xffffffffu, c1i, 1 << l), o2i = __shfl_up_sync(0xffffffffu, c2i, 1 << l);
                if (lane >= (1 << l)) {
                    c1r += fmaf(p.pk_P[l][0], o1r, p.pk_P[l][1] * o2r); c2r += fmaf(p.pk_P[l][2], o1r, p.pk_P[l][3] * o2r);
                    c1i += fmaf(p.pk_P[l][0], o1i, p.pk_P[l][1] * o2i); c2i += fmaf(p.pk_P[l][2], o1i, p.pk_P[l][3] * o2i);
                }
            }
            if (lane == 31) { s_wt[4 * warp] = c1r; s_wt[4 * warp + 1] = c2r; s_wt[4 * warp + 2] = c1i; s_wt[4 * warp + 3] = c2i; }
            __syncthreads();
            float g1r = 0.f, g2r = 0.f, g1i = 0.f, g2i = 0.f;        // state at the start of this warp
            for (int w = 0; w < warp; w++) {
                const float n1r = fmaf(p.pk_P[5][0], g1r, p.pk_P[5][1] * g2r) + s_wt[4 * w], n2r = fmaf(p.pk_P[5][2], g1r, p.pk_P[5][3] * g2r) + s_wt[4 * w + 1];
                const float n1i = fmaf(p.pk_P[5][0], g1i, p.pk_P[5][1] * g2i) + s_wt[4 * w + 2], n2i = fmaf(p.pk_P[5][2], g1i, p.pk_P[5][3] * g2i) + s_wt[4 * w + 3];
                g1r = n1r; g2r = n2r; g1i = n1i; g2i = n2i;
            }
            // true start state of this thread = (previous lane's inclusive prefix) + A^(8 lane) * warp start
            float e1r = __shfl_up_sync(0xffffffffu, c1r, 1), e2r = __shfl_up_sync(0xffffffffu, c2r, 1);
            float e1i = __shfl_up_sync(0xffffffffu, c1i, 1), e2i = __shfl_up_sync(0xffffffffu, c2i, 1);
            if (lane == 0) { e1r = 0.f; e2r = 0.f; e1i = 0.f; e2i = 0.f; }
            const float4 Q = *(const float4*)(s_Q + 4 * lane);
            e1r += fmaf(Q.x, g1r, Q.y * g2r); e2r += fmaf(Q.z, g1r, Q.w * g2r);
            e1i += fmaf(Q.x, g1i, Q.y * g2i); e2i += fmaf(Q.z, g1i, Q.w * g2i);
            if (t == 0) { e1r = s_st[0]; e2r = s_st[1]; e1i = s_st[2]; e2i = s_st[3]; }
            __syncthreads();                                         // everyone has read s_st / s_wt
            if (active) {
                float4* yd = (float4*)(s_y + 2 * K2_R * t);
#pragma unroll
                for (int q = 0; q < K2_R / 2; q++) {
                    const float y0r = fmaf(e1r, a1, fmaf(e2r, a0, ur[2 * q])), y0i = fmaf(e1i, a1, fmaf(e2i, a0, ui[2 * q]));
                    const float y1r = fmaf(y0r, a1, fmaf(e1r, a0, ur[2 * q + 1])), y1i = fmaf(y0i, a1, fmaf(e1i, a0, ui[2 * q + 1]));
                    e2r = y0r; e1r = y1r; e2i = y0i; e1i = y1i;
                    yd[q] = make_float4(y0r, y0i, y1r, y1i);
                }
                if (K2_R * (t + 1) == nch) {                         // last sample of the chunk: carry
                    s_st[0] = e1r; s_st[1] = e2r; s_st[2] = e1i; s_st[3] = e2i;
                    const float4 xl = *(const float4*)(s_iq + 2 * nch - 4);
                    s_st[8] = xl.x; s_st[9] = xl.y; s_st[10] = xl.z; s_st[11] = xl.w;
                }
            }
        }
        __syncthreads();
        if (t < 64) s_out[a_out(t)] = carry_out;
        if (t == 0) { s_iq[-4] = s_st[8]; s_iq[-3] = s_st[9]; s_iq[-2] = s_st[10]; s_iq[-1] = s_st[11]; }
        // ---- E: pilot angle in turns + |y|^2 partial sums ----
        float pw = 0.0f;
        if (active) {
            float th[K2_R];
            const float4* ys = (const float4*)(s_y + 2 * K2_R * t);
#pragma unroll
            for (int q = 0; q < K2_R / 2; q++) {
                const float4 v = ys[q];
                th[2 * q] = atan2f(v.y, v.x) * INV_TWO_PI_F;
                th[2 * q + 1] = atan2f(v.w, v.z) * INV_TWO_PI_F;
                pw += v.x * v.x + v.y * v.y;
                pw += v.z * v.z + v.w * v.w;
                if (p.keep) ((float4*)(pilot_dbg + (size_t)s * p.n_out + c0 + K2_R * t))[q] = v;
            }
            float4* d = (float4*)(theta + (size_t)s * p.n_out + c0 + K2_R * t);
            d[0] = make_float4(th[0], th[1], th[2], th[3]);
            d[1] = make_float4(th[4], th[5], th[6], th[7]);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) pw += __shfl_xor_sync(0xffffffffu, pw, off);
        if (lane == 0) s_red[warp] = pw;
        __syncthreads();
        if (t == 0) {
            float sum = 0.0f;
            for (int w = 0; w < K2_THREADS / 32; w++) sum += s_red[w];
            power_total += sum;
        }
    }
    __syncthreads();
    if (t < K2_NN) hist_demod[(size_t)s * K2_NN + t] = s_in[a_in(t)];
    if (t < 64) hist_out[(size_t)s * 64 + t] = s_out[a_out(t)];
    if (t == 0) {
        sc[K2_PK_X2R] = s_iq[-4]; sc[K2_PK_X2I] = s_iq[-3]; sc[K2_PK_X1R] = s_iq[-2]; sc[K2_PK_X1I] = s_iq[-1];
        sc[K2_PK_Y1R] = s_st[0]; sc[K2_PK_Y2R] = s_st[1]; sc[K2_PK_Y1I] = s_st[2]; sc[K2_PK_Y2I] = s_st[3];
        sc[K2_DEEMPH_X1] = s_st[4]; sc[K2_DEEMPH_Y1] = s_st[5];
        power[s] = power_total;
    }
}

cudaError_t launch_k2(const float* fm_demod, float* hist_demod, float* hist_out, float* scal,
                      float2* fm_out_iq, float* theta, float* power, float2* pilot_dbg,
                      const K2Params& p, int n_streams, cudaStream_t st)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k2_mpx, cudaFuncAttributeMaxDynamicSharedMemorySize, K2_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    k2_mpx<<<n_streams, K2_THREADS, K2_SMEM_BYTES, st>>>(fm_demod, hist_demod, hist_out, scal, fm_out_iq, theta, power, pilot_dbg, p);
    return cudaGetLastError();
}

} // namespace fm
