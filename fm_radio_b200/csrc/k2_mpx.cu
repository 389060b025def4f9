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
#include "fm_common.cuh"

namespace fm {

// bank-skewed shared layouts: s_in pads 4 floats per 16, s_out pads 4 floats per 8, so the
// per-thread LDS.128 windows (thread stride 16 resp. 8 floats) are conflict free.
__device__ __forceinline__ int a_in(int i) { return i + 4 * (i >> 4); }
__device__ __forceinline__ int a_out(int i) { return i + 4 * (i >> 3); }

constexpr int K2_S_IN = (K2_NN + 2 * K2_CH) + 4 * ((K2_NN + 2 * K2_CH) >> 4) + 16;
constexpr int K2_S_OUT = (64 + K2_CH) + 4 * ((64 + K2_CH) >> 3) + 16;
constexpr int K2_SMEM_BYTES = (K2_S_IN + K2_S_OUT) * 4 + 2 * K2_CH * 8 + 64;

__global__ void __launch_bounds__(K2_THREADS, 2)
k2_mpx(const float* __restrict__ fm_demod, float* __restrict__ hist_demod, float* __restrict__ hist_out,
       float* __restrict__ scal, float2* __restrict__ fm_out_iq, float* __restrict__ theta,
       float* __restrict__ power, float2* __restrict__ pilot_dbg, const __grid_constant__ K2Params p)
{
    extern __shared__ __align__(16) float smem[];
    float* s_in = smem;                                   // [64 hist + 4096] fm_demod, skewed
    float* s_out = s_in + K2_S_IN;                        // [64 hist + 2048] fm_out, skewed
    float* s_iq = s_out + K2_S_OUT;                       // [2048] x (re,im)
    float* s_y = s_iq + 2 * K2_CH;                        // [2048] y (re,im)
    float* s_red = s_y + 2 * K2_CH;                       // [8] warp partials
    const int t = threadIdx.x;
    const int s = blockIdx.x;
    const int lane = t & 31, warp = t >> 5;

    if (t < K2_NN) s_in[a_in(t)] = hist_demod[(size_t)s * K2_NN + t];
    if (t < 64) s_out[a_out(t)] = hist_out[(size_t)s * 64 + t];
    float* sc = scal + (size_t)s * K2_SCAL_N;
    // pilot IIR state of component `lane` (0 = re, 1 = im), held by lanes 0 and 1 of warp 0
    float pk_x1 = 0.f, pk_x2 = 0.f, pk_y1 = 0.f, pk_y2 = 0.f;
    if (warp == 0 && lane < 2) {
        pk_x1 = sc[K2_PK_X1R + lane]; pk_x2 = sc[K2_PK_X2R + lane];
        pk_y1 = sc[K2_PK_Y1R + lane]; pk_y2 = sc[K2_PK_Y2R + lane];
    }
    float de_x1 = sc[K2_DEEMPH_X1], de_y1 = sc[K2_DEEMPH_Y1];   // used by thread 0 only
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
        // ---- B2: optional de-emphasis, y[n] = b[0]x[n-1] + a[0]y[n-1] + b[1]x[n] (iir_filter.h:62-68) ----
        if (p.use_deemph) {
            if (t == 0) {
                for (int i = 0; i < nch; i++) {
                    const float x = s_out[a_out(64 + i)];
                    const float y = fmaf(de_y1, p.deemph_a[0], fmaf(de_x1, p.deemph_b[0], x * p.deemph_b[1]));
                    de_x1 = x; de_y1 = y;
                    s_out[a_out(64 + i)] = y;
                }
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
        // ---- D: pilot peak filter, sequential, lane 0 = real part, lane 1 = imaginary part ----
        //   y[n] = b[0]x[n-2] + b[1]x[n-1] + b[2]x[n] + a[0]y[n-2] + a[1]y[n-1]   (iir_filter.h:62-68;
        //   a[2] multiplies the always-zero yn[K-1]).  Only one FFMA is on the y[n-1] -> y[n] chain.
        if (warp == 0 && lane < 2) {
            const float b0 = p.peak_b[0], b1 = p.peak_b[1], b2 = p.peak_b[2];
            const float a0 = p.peak_a[0], a1 = p.peak_a[1];
#pragma unroll 4
            for (int i = 0; i < nch; i++) {
                const float x0 = s_iq[2 * i + lane];
                const float u = fmaf(pk_y2, a0, fmaf(pk_x2, b0, fmaf(pk_x1, b1, x0 * b2)));
                const float y = fmaf(pk_y1, a1, u);
                pk_x2 = pk_x1; pk_x1 = x0;
                pk_y2 = pk_y1; pk_y1 = y;
                s_y[2 * i + lane] = y;
            }
        }
        __syncthreads();
        if (t < 64) s_out[a_out(t)] = carry_out;
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
    if (warp == 0 && lane < 2) {
        sc[K2_PK_X1R + lane] = pk_x1; sc[K2_PK_X2R + lane] = pk_x2;
        sc[K2_PK_Y1R + lane] = pk_y1; sc[K2_PK_Y2R + lane] = pk_y2;
    }
    if (t == 0) {
        sc[K2_DEEMPH_X1] = de_x1; sc[K2_DEEMPH_Y1] = de_y1;
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
