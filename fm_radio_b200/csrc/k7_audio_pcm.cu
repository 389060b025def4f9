// K7: the audio output stage -- what the reference's consumers do with every OnAudioOut block.
//
// Replaces (reference file:line under /root/reference/src):
//   Resample(buf_in, buf_out), linear interpolation 32 kHz -> device rate     audio/resampled_pcm_player.cpp:37-54
//     called per block by Resampled_PCM_Player::ConsumeBuffer (:15-28): M = (int)(L * N) output frames,
//     the read position restarts at 0 every block and the last frame is held (no state across blocks)
//   Audio_Scraper::on_audio_data float -> int16 conversion                     fm_scraper.cpp:74-78
//     Frame<int16_t>(frame * (32767 * 0.95f)), a static_cast per channel (audio/frame.h:67-74)
//   PolyphaseUpsampler<T>::process (the in-tree higher-quality interpolator)   dsp/polyphase_filter.h:90-185
//
// HBM-bound and tiny next to the FIR stages: per IQ sample 0.25 B of audio in, 0.375 B (f32) + 0.19 B
// (s16) out at 48 kHz.  One thread per output frame, every access coalesced.
//
// Resample() walks the read position with a float accumulator (j += step), so where output i reads
// depends on 3071 rounded additions before it.  That walk depends only on (N, M): the host runs it once,
// in the same float arithmetic, and the kernel reads (j0, k) per output from a table shared by all streams.
#include "fm_common.cuh"

namespace fm {

// static_cast<int16_t>(float) as the reference's x86 build performs it: cvttss2si (truncate toward zero,
// 0x80000000 when out of int32 range or NaN), then the low 16 bits.
__device__ __forceinline__ short f32_to_s16_x86(float v) {
    const int i = (fabsf(v) < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;
    return (short)(i & 0xFFFF);
}

constexpr int K7_U = 4;         // output frames per thread, loads of all four issued before the first use

__global__ void __launch_bounds__(256)
k7_audio_pcm(const float2* __restrict__ audio, const K7Entry* __restrict__ table,
             float2* __restrict__ pcm_f32, short2* __restrict__ pcm_s16, int n_in, int n_out, float s16_scale)
{
    const int s = blockIdx.y;
    const float2* in = audio + (size_t)s * n_in;
    for (int i0 = (blockIdx.x * blockDim.x) * K7_U + threadIdx.x; i0 < n_out; i0 += gridDim.x * blockDim.x * K7_U) {
        K7Entry e[K7_U];
        float2 f0[K7_U], f1[K7_U];
#pragma unroll
        for (int u = 0; u < K7_U; u++) {
            const int i = i0 + u * 256;
            e[u] = (i < n_out) ? table[i] : K7Entry{ 0, 0.0f };
        }
#pragma unroll
        for (int u = 0; u < K7_U; u++) {
            f0[u] = __ldg(in + e[u].j0);
            f1[u] = (e[u].j0 + 1 < n_in) ? __ldg(in + e[u].j0 + 1) : f0[u];
        }
#pragma unroll
        for (int u = 0; u < K7_U; u++) {
            const int i = i0 + u * 256;
            if (i >= n_out) continue;
            // buf_out[i] = f0*(1.0f-k) + f1*k, each Frame operator rounding once (frame.h:10-16, 42-49)
            const float a = __fsub_rn(1.0f, e[u].k);
            float2 o;
            o.x = __fadd_rn(__fmul_rn(f0[u].x, a), __fmul_rn(f1[u].x, e[u].k));
            o.y = __fadd_rn(__fmul_rn(f0[u].y, a), __fmul_rn(f1[u].y, e[u].k));
            const size_t k = (size_t)s * n_out + i;
            if (pcm_f32) pcm_f32[k] = o;
            if (pcm_s16) pcm_s16[k] = make_short2(f32_to_s16_x86(__fmul_rn(o.x, s16_scale)), f32_to_s16_x86(__fmul_rn(o.y, s16_scale)));
        }
    }
}

// n_out == n_in and table == nullptr: conversion only (the scraper's 32 kHz WAV path)
__global__ void __launch_bounds__(256)
k7_frames_to_s16(const float2* __restrict__ frames, short2* __restrict__ out, size_t n, float s16_scale)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float2 f = frames[i];
        out[i] = make_short2(f32_to_s16_x86(__fmul_rn(f.x, s16_scale)), f32_to_s16_x86(__fmul_rn(f.y, s16_scale)));
    }
}

cudaError_t launch_k7(const float2* audio, const K7Entry* table, float2* pcm_f32, short2* pcm_s16,
                      int n_in, int n_out, int n_streams, cudaStream_t st)
{
    const float scale = 32767.0f * 0.95f;                // CONVERT_RESCALE, fm_scraper.cpp:74
    const int bx = (n_out + 256 * K7_U - 1) / (256 * K7_U);
    k7_audio_pcm<<<dim3(bx > 8 ? 8 : bx, n_streams), 256, 0, st>>>(audio, table, pcm_f32, pcm_s16, n_in, n_out, scale);
    return cudaGetLastError();
}

cudaError_t launch_frames_to_s16(const float2* frames, short2* out, size_t n, cudaStream_t st)
{
    const float scale = 32767.0f * 0.95f;
    const size_t blocks = (n + 255) / 256;
    k7_frames_to_s16<<<(unsigned)(blocks > 1184 ? 1184 : (blocks ? blocks : 1)), 256, 0, st>>>(frames, out, n, scale);
    return cudaGetLastError();
}

// The read positions of Resample(): j starts at 0 and advances by step = (float)n_in / (float)n_out in
// float (resampled_pcm_player.cpp:41-53).  Host code, plain IEEE float (this file is not built with fast math).
void k7_build_table(int n_in, int n_out, K7Entry* out)
{
    const float step = (float)n_in / (float)n_out;
    volatile float j = 0.0f;                             // volatile: every addition rounds to float, no reassociation
    for (int i = 0; i < n_out; i++) {
        const float jj = j;
        const int j0 = (int)jj;
        out[i].j0 = j0;
        out[i].k = jj - (float)j0;
        j = jj + step;
    }
}

// PolyphaseUpsampler<T>::process (dsp/polyphase_filter.h:131-160): y[i*L + phase] = sum_{t<K} bp[phase*K + t] * ext[i + 1 + t],
// ext = (K history samples) ++ (N new samples), bp = the constructor's repacked taps (:106-116).
// One output per thread, taps ascending as apply_filter (:176-184).
template <bool CPLX>
__global__ void polyphase_us_kernel(const float* __restrict__ ext, const float* __restrict__ bp,
                                    float* __restrict__ y, int L, int K, int n_in)
{
    extern __shared__ float s_bp[];
    for (int k = threadIdx.x; k < L * K; k += blockDim.x) s_bp[k] = bp[k];
    __syncthreads();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_in * L) return;
    const int i = o / L, phase = o - i * L;
    const float* b0 = s_bp + phase * K;
    if (CPLX) {
        const float2* w = (const float2*)ext + i + 1;
        float ar = 0.0f, ai = 0.0f;
        for (int t = 0; t < K; t++) { const float2 v = w[t]; ar = fmaf(v.x, b0[t], ar); ai = fmaf(v.y, b0[t], ai); }
        ((float2*)y)[o] = make_float2(ar, ai);
    } else {
        const float* w = ext + i + 1;
        float a = 0.0f;
        for (int t = 0; t < K; t++) a = fmaf(w[t], b0[t], a);
        y[o] = a;
    }
}

cudaError_t launch_polyphase_us(const float* ext, const float* bp, float* y, int L, int K, int n_in,
                                int is_complex, cudaStream_t st)
{
    const int threads = 128;
    const int grid = (n_in * L + threads - 1) / threads;
    const size_t sm = (size_t)L * K * sizeof(float);
    if (is_complex) polyphase_us_kernel<true><<<grid, threads, sm, st>>>(ext, bp, y, L, K, n_in);
    else            polyphase_us_kernel<false><<<grid, threads, sm, st>>>(ext, bp, y, L, K, n_in);
    return cudaGetLastError();
}

} // namespace fm
