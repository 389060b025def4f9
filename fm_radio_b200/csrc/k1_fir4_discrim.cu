// K1: u8 (or cf32) baseband IQ -> 64-tap /4 polyphase FIR -> polar discriminator -> fm_demod.
//
// Replaces, fused in one pass (reference file:line under /root/reference/src):
//   App::Run u8 -> cf32 unpack, (float)u8 - 127.0f           app.cpp:56-65
//   PolyphaseDownsampler<cf32>::process, M=4 K=16 NN=64      dsp/polyphase_filter.h:41-64,197-201
//     y[i] = sum_k b[k] * x[(i+1)*4 - 64 + k]                (c32_f32_cum_mul.cpp:70-111 on AVX)
//   FM_Demod::Process, atan2 + wrapped first difference      fm_demod/fm_demod.cpp:30-45
//
// Design (FP32-FMA-pipe bound, 64 FLOP per input IQ sample):
//   * one CTA = 2048 consecutive outputs of one stream = 8192 IQ samples (+64 samples of halo);
//     the u8 tile is read once with 128-bit coalesced loads, unpacked with PRMT + FADD (no I2F)
//     and staged in shared memory as fp32, 16-frame segments skewed by 4 banks so that the
//     per-thread LDS.128 window reads are conflict free;
//   * each thread owns 16 consecutive outputs (+ the one before them, recomputed bit-identically
//     so the discriminator's first difference needs no cross-thread or cross-CTA exchange):
//     34 register accumulators, the window slides one 4-sample frame (2 x LDS.128) per 128 FFMA;
//   * the 64 taps ride in the kernel parameter block (__grid_constant__), i.e. constant bank 0,
//     and the tap loop is fully unrolled so every FFMA takes its tap as a c[0][imm] operand:
//     no register, no load, no issue slot spent on coefficients;
//   * epilogue: atan2f, wrapped difference, gain; 4 x STG.128 per thread.
// The previous block's last 64 samples are the only state (ping-pong buffers, because the CTA
// that reads the history is not the CTA that writes it); the discriminator's prev_theta is
// recomputed from that history instead of being stored.
#include "fm_common.cuh"

namespace fm {

__device__ __forceinline__ float u8_to_f32_m127(uint32_t w, int byte) {
    // (2^23 + u8) built by PRMT into the mantissa of 0x4B000000, then one exact FADD.
    const uint32_t sel = 0x7440u | (uint32_t)byte;
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - 8388735.0f;
}

__device__ __forceinline__ float wrap_phase(float x) {          // fm_demod.cpp:6-10
    if (x >= PI_F) return x - 2.0f * PI_F;
    else if (x <= -PI_F) return x + 2.0f * PI_F;
    return x;
}

template <bool U8>
__global__ void __launch_bounds__(K1_THREADS, 3)
k1_fir4_discrim(const void* __restrict__ iq, const float2* __restrict__ hist_in,
                float2* __restrict__ hist_out, float* __restrict__ fm_demod,
                const __grid_constant__ K1Params p)
{
    extern __shared__ __align__(16) float smem[];
    const int t = threadIdx.x;
    const int tile = blockIdx.x;
    const int s = blockIdx.y;
    const int o0 = tile * K1_TILE;
    const int n_valid = min(K1_TILE, p.n_out - o0);      // outputs of this tile (multiple of 16)
    const int n_frames = n_valid + 16;                   // frames to stage, frame f <-> input frame o0-16+f
    const size_t n_in = (size_t)p.n_out * K1_M;          // IQ samples per stream per block

    // ---- stage the tile: frame f lives at smem[(f>>4)*K1_SEG + (f&15)*8 .. +8) ----
    const int f_first = (tile == 0) ? 16 : 0;            // tile 0 takes its halo from the history
    if (tile == 0 && t < K1_HIST) {
        const float2 v = hist_in[(size_t)s * K1_HIST + t];
        const int f = t >> 2;
        float* d = smem + (f >> 4) * K1_SEG + (f & 15) * 8 + (t & 3) * 2;
        d[0] = v.x; d[1] = v.y;
    }
    if (U8) {
        // 16 bytes = 8 IQ samples = 2 frames per thread per iteration
        const uint8_t* src = (const uint8_t*)iq + ((size_t)s * n_in + (size_t)(o0 - 16) * K1_M) * 2;
        for (int pp = (f_first >> 1) + t; pp < (n_frames >> 1); pp += K1_THREADS) {
            const uint4 w = __ldg((const uint4*)(src + (size_t)pp * 16));
            const int f = pp * 2;
            float* d = smem + (f >> 4) * K1_SEG + (f & 15) * 8;     // f even: both frames in one segment
            float4 v;
            v.x = u8_to_f32_m127(w.x, 0); v.y = u8_to_f32_m127(w.x, 1); v.z = u8_to_f32_m127(w.x, 2); v.w = u8_to_f32_m127(w.x, 3);
            *(float4*)(d + 0) = v;
            v.x = u8_to_f32_m127(w.y, 0); v.y = u8_to_f32_m127(w.y, 1); v.z = u8_to_f32_m127(w.y, 2); v.w = u8_to_f32_m127(w.y, 3);
            *(float4*)(d + 4) = v;
            v.x = u8_to_f32_m127(w.z, 0); v.y = u8_to_f32_m127(w.z, 1); v.z = u8_to_f32_m127(w.z, 2); v.w = u8_to_f32_m127(w.z, 3);
            *(float4*)(d + 8) = v;
            v.x = u8_to_f32_m127(w.w, 0); v.y = u8_to_f32_m127(w.w, 1); v.z = u8_to_f32_m127(w.w, 2); v.w = u8_to_f32_m127(w.w, 3);
            *(float4*)(d + 12) = v;
        }
    } else {
        // 16 bytes = 2 IQ samples = half a frame
        const float4* src = (const float4*)((const float2*)iq + (size_t)s * n_in + (ptrdiff_t)(o0 - 16) * K1_M);
        for (int hf = f_first * 2 + t; hf < n_frames * 2; hf += K1_THREADS) {
            const float4 v = __ldg(src + hf);
            const int f = hf >> 1;
            *(float4*)(smem + (f >> 4) * K1_SEG + (f & 15) * 8 + (hf & 1) * 4) = v;
        }
    }
    __syncthreads();

    // ---- history for the next block: the last 64 IQ samples of this stream's block ----
    if (o0 + n_valid == p.n_out && t < K1_HIST) {
        const int f = n_valid + (t >> 2);
        const float* d = smem + (f >> 4) * K1_SEG + (f & 15) * 8 + (t & 3) * 2;
        hist_out[(size_t)s * K1_HIST + t] = make_float2(d[0], d[1]);
    }

    if (t * K1_R >= n_valid) return;

    // ---- 17 outputs per thread: r = -1 (previous thread's last output, for the difference) .. 15 ----
    // output o = 16t + r uses frames f = o+1 .. o+16 of the staged tile with tap group f-o-1.
    float ar[K1_R + 1], ai[K1_R + 1];
#pragma unroll
    for (int r = 0; r <= K1_R; r++) { ar[r] = 0.0f; ai[r] = 0.0f; }
    const float* base = smem + t * K1_SEG;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const float* fp = base + ((j < 16) ? j * 8 : K1_SEG + (j - 16) * 8);
        const float4 a = *(const float4*)fp;
        const float4 b = *(const float4*)(fp + 4);
#pragma unroll
        for (int r = -1; r < K1_R; r++) {
            const int g = j - r - 1;                    // tap group, static after unrolling
            if (g >= 0 && g < 16) {
                ar[r + 1] = fmaf(a.x, p.taps[4 * g + 0], ar[r + 1]); ai[r + 1] = fmaf(a.y, p.taps[4 * g + 0], ai[r + 1]);
                ar[r + 1] = fmaf(a.z, p.taps[4 * g + 1], ar[r + 1]); ai[r + 1] = fmaf(a.w, p.taps[4 * g + 1], ai[r + 1]);
                ar[r + 1] = fmaf(b.x, p.taps[4 * g + 2], ar[r + 1]); ai[r + 1] = fmaf(b.y, p.taps[4 * g + 2], ai[r + 1]);
                ar[r + 1] = fmaf(b.z, p.taps[4 * g + 3], ar[r + 1]); ai[r + 1] = fmaf(b.w, p.taps[4 * g + 3], ai[r + 1]);
            }
        }
    }

    // ---- discriminator epilogue (fm_demod.cpp:36-44) ----
    float prev = atan2f(ai[0], ar[0]);
    float out[K1_R];
#pragma unroll
    for (int r = 0; r < K1_R; r++) {
        const float th = atan2f(ai[r + 1], ar[r + 1]);
        out[r] = wrap_phase(th - prev) * p.discrim_gain;
        prev = th;
    }
    float4* dst = (float4*)(fm_demod + (size_t)s * p.n_out + o0 + t * K1_R);
#pragma unroll
    for (int q = 0; q < K1_R / 4; q++) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
}

constexpr int K1_SMEM_BYTES = (K1_THREADS + 1) * K1_SEG * (int)sizeof(float);

cudaError_t launch_k1(bool u8, const void* iq, const float2* hist_in, float2* hist_out, float* fm_demod,
                      const K1Params& p, cudaStream_t st)
{
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k1_fir4_discrim<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k1_fir4_discrim<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const dim3 grid((p.n_out + K1_TILE - 1) / K1_TILE, p.n_streams);
    if (u8) k1_fir4_discrim<true><<<grid, K1_THREADS, K1_SMEM_BYTES, st>>>(iq, hist_in, hist_out, fm_demod, p);
    else    k1_fir4_discrim<false><<<grid, K1_THREADS, K1_SMEM_BYTES, st>>>(iq, hist_in, hist_out, fm_demod, p);
    return cudaGetLastError();
}

} // namespace fm
