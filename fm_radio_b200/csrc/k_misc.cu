// Small kernels off the hot path: GUI-buffer finalisation (keep_intermediates) and the stand-alone
// polyphase decimator behind the src/dsp PolyphaseDownsampler<T> API surface.
#include "fm_common.cuh"

namespace fm {

// pilot_buf as the reference exposes it is AFTER the AGC (broadcast_fm_demod.cpp:423); pll_buf is
// the oscillator sample (S(t+1/4), S(t)) (:443-445, 453).  Both are display-only here because K3
// works on angles, so they are produced in parallel after the fact.
__global__ void kdbg_finalize(float2* __restrict__ pilot, const float* __restrict__ pll_state,
                              const float* __restrict__ pll_dt, float2* __restrict__ pll_out, int n, int n_streams)
{
    const int s = blockIdx.y;
    const float gain = pll_state[PLL_AGC_GAIN * n_streams + s];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t k = (size_t)s * n + i;
        float2 v = pilot[k];
        v.x *= gain; v.y *= gain;
        pilot[k] = v;
        const float t = pll_dt[k];
        float dc = t + 0.25f;
        dc = dc - roundf(dc);
        pll_out[k] = make_float2(chebyshev_sine(dc), chebyshev_sine(t));
    }
}

cudaError_t launch_kdbg(float2* pilot, const float* pll_state, const float* pll_dt, float2* pll_out,
                        int n, int n_streams, cudaStream_t st)
{
    const dim3 grid((n + 255) / 256 > 64 ? 64 : (n + 255) / 256, n_streams);
    kdbg_finalize<<<grid, 256, 0, st>>>(pilot, pll_state, pll_dt, pll_out, n, n_streams);
    return cudaGetLastError();
}

// y[i] = sum_k b[k] * ext[(i+1)*M + k], ext = (NN history samples) ++ (n_out*M new samples).
// Generic M / NN, one output per thread, taps broadcast from shared memory.
template <bool CPLX>
__global__ void polyphase_ds_kernel(const float* __restrict__ ext, const float* __restrict__ taps,
                                    float* __restrict__ y, int M, int NN, int n_out)
{
    extern __shared__ float s_b[];
    for (int k = threadIdx.x; k < NN; k += blockDim.x) s_b[k] = taps[k];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    if (CPLX) {
        const float2* w = (const float2*)ext + (size_t)(i + 1) * M;
        float ar = 0.0f, ai = 0.0f;
        for (int k = 0; k < NN; k++) { const float2 v = w[k]; ar = fmaf(v.x, s_b[k], ar); ai = fmaf(v.y, s_b[k], ai); }
        ((float2*)y)[i] = make_float2(ar, ai);
    } else {
        const float* w = ext + (size_t)(i + 1) * M;
        float a = 0.0f;
        for (int k = 0; k < NN; k++) a = fmaf(w[k], s_b[k], a);
        y[i] = a;
    }
}

cudaError_t launch_polyphase_ds(const float* ext, const float* taps, float* y, int M, int NN, int n_out,
                                int is_complex, cudaStream_t st)
{
    const int threads = 128;
    const int grid = (n_out + threads - 1) / threads;
    if (is_complex) polyphase_ds_kernel<true><<<grid, threads, NN * sizeof(float), st>>>(ext, taps, y, M, NN, n_out);
    else            polyphase_ds_kernel<false><<<grid, threads, NN * sizeof(float), st>>>(ext, taps, y, M, NN, n_out);
    return cudaGetLastError();
}

} // namespace fm
