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

// GUI mode only: the COMPLEX outputs of the two audio decimators, temp_audio_buf as the reference has it after
// broadcast_fm_demod.cpp:475 (L+R: filt_poly_ds_lpf_audio_lpr on fm_out_iq) and :490 (L-R: filt_poly_ds_lpf_audio_lmr
// on the 38 kHz mixdown), the sources of its two audio spectra (:481, :523).  The fused K4 keeps only what the audio
// needs of them (the real part of the first, the imaginary part of the second), so for the display they are
// recomputed here, one output per thread, straight from the definitions: y[o] = sum_k b[k] v(4 (o + 1) - 128 + k),
// v(n < 0) from the carried histories; the mixdown is the reference's formula (apply_harmonic_pll.cpp:16-23).
__global__ void kdbg_audio_iq(const float2* __restrict__ fm_out_iq, const float* __restrict__ pll_dt,
                              const float2* __restrict__ hist_iq, const float2* __restrict__ hist_m2,
                              const float* __restrict__ lmr_phase_used, float2* __restrict__ lpr_iq,
                              float2* __restrict__ lmr_iq, const __grid_constant__ K4Params p)
{
    const int s = blockIdx.y;
    const int n_out = p.n >> 2;
    const float off = lmr_phase_used[s];
    const float2* x = fm_out_iq + (size_t)s * p.n;
    const float* dt = pll_dt + (size_t)s * p.n;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += gridDim.x * blockDim.x) {
        float ar = 0.0f, ai = 0.0f, br = 0.0f, bi = 0.0f;
        for (int k = 0; k < K4_NN; k++) {
            const int n = 4 * (o + 1) - K4_NN + k;
            float2 v, m;
            if (n < 0) {
                v = hist_iq[(size_t)s * K4_NN + K4_NN + n];
                m = hist_m2[(size_t)s * K4_NN + K4_NN + n];
            } else {
                v = x[n];
                float ph = fmaf(dt[n], p.harmonic_lmr, off);
                float pc = ph + 0.25f;
                pc -= rintf(pc); ph -= rintf(ph);
                const float c = chebyshev_sine(pc), sn = chebyshev_sine(ph);
                m = make_float2(v.x * c - v.y * sn, v.x * sn + v.y * c);
            }
            ar = fmaf(v.x, p.taps_lpr[k], ar); ai = fmaf(v.y, p.taps_lpr[k], ai);
            br = fmaf(m.x, p.taps_lmr[k], br); bi = fmaf(m.y, p.taps_lmr[k], bi);
        }
        lpr_iq[(size_t)s * n_out + o] = make_float2(ar, ai);
        lmr_iq[(size_t)s * n_out + o] = make_float2(br, bi);
    }
}

cudaError_t launch_kdbg_audio_iq(const float2* fm_out_iq, const float* pll_dt, const float2* hist_iq, const float2* hist_m2,
                                 const float* lmr_phase_used, float2* lpr_iq, float2* lmr_iq, const K4Params& p, cudaStream_t st)
{
    const int n_out = p.n >> 2;
    const dim3 grid((n_out + 127) / 128 > 32 ? 32 : (n_out + 127) / 128, p.n_streams);
    kdbg_audio_iq<<<grid, 128, 0, st>>>(fm_out_iq, pll_dt, hist_iq, hist_m2, lmr_phase_used, lpr_iq, lmr_iq, p);
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

// ---- stand-alone src/dsp filter classes (FIR_Filter is polyphase_ds_kernel with M = 1) ----
// Hilbert_FIR_Filter<float>::process (dsp/hilbert_fir_filter.h:26-46): y[i] = { x delayed by (K-1)/2, FIR(x) };
// ext = (K history samples) ++ x, so the delayed sample of output i is ext[i + 1 + (K-1)/2].
__global__ void hilbert_pack_kernel(const float* __restrict__ ext, const float* __restrict__ fir, float2* __restrict__ y, int n, int mid)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = make_float2(ext[i + 1 + mid], fir[i]);
}

// IIR_Filter<T>::process (dsp/iir_filter.h:40-69), direct form I, strictly sequential: one thread walks the block in the
// reference's operation order -- y = 0; y += (xn[i] b[i] + yn[i] a[i]) for i = 0 .. K-1 -- with explicit roundings (no
// contraction), so the result is bit-identical to the scalar program.  state = xn[K] ++ yn[K] (C floats per entry).
template <int C>
__global__ void iir_seq_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int K,
                               const float* __restrict__ b, const float* __restrict__ a, float* __restrict__ state)
{
    extern __shared__ float s_iir[];                 // b[K], a[K], xn[K*C], yn[K*C]
    float* sb = s_iir; float* sa = sb + K; float* xn = sa + K; float* yn = xn + K * C;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    for (int i = 0; i < K; i++) { sb[i] = b[i]; sa[i] = a[i]; }
    for (int i = 0; i < K * C; i++) { xn[i] = state[i]; yn[i] = state[K * C + i]; }
    for (int t = 0; t < n; t++) {
        for (int i = 0; i < (K - 1) * C; i++) xn[i] = xn[i + C];
        for (int c = 0; c < C; c++) xn[(K - 1) * C + c] = x[(size_t)t * C + c];
        float acc[C];
        for (int c = 0; c < C; c++) acc[c] = 0.0f;
        for (int i = 0; i < K; i++)
            for (int c = 0; c < C; c++)
                acc[c] = __fadd_rn(acc[c], __fadd_rn(__fmul_rn(xn[i * C + c], sb[i]), __fmul_rn(yn[i * C + c], sa[i])));
        for (int c = 0; c < C; c++) y[(size_t)t * C + c] = acc[c];
        for (int i = 0; i < (K - 2) * C; i++) yn[i] = yn[i + C];
        if (K >= 2) for (int c = 0; c < C; c++) yn[(K - 2) * C + c] = acc[c];
    }
    for (int i = 0; i < K * C; i++) { state[i] = xn[i]; state[K * C + i] = yn[i]; }
}

cudaError_t launch_iir_seq(const float* x, float* y, int n, int K, const float* b, const float* a, float* state, int is_complex, cudaStream_t st)
{
    const size_t smem = (size_t)(2 * K + 2 * K * (is_complex ? 2 : 1)) * sizeof(float);
    if (is_complex) iir_seq_kernel<2><<<1, 32, smem, st>>>(x, y, n, K, b, a, state);
    else            iir_seq_kernel<1><<<1, 32, smem, st>>>(x, y, n, K, b, a, state);
    return cudaGetLastError();
}

cudaError_t launch_hilbert_pack(const float* ext, const float* fir, float2* y, int n, int mid, cudaStream_t st)
{
    hilbert_pack_kernel<<<(n + 127) / 128, 128, 0, st>>>(ext, fir, y, n, mid);
    return cudaGetLastError();
}

// AGC_Filter<cf32>::calculate_average_power (dsp/agc.h:21-30): the reference adds I*I + Q*Q sample by sample into one
// float; one thread does exactly that (explicit roundings), so the gain matches the scalar program bit for bit.
__global__ void agc_power_seq_kernel(const float2* __restrict__ x, int n, float* __restrict__ sum)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float acc = 0.0f;
    for (int i = 0; i < n; i++) { const float2 v = x[i]; acc = __fadd_rn(acc, __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y))); }
    *sum = acc;
}
__global__ void agc_scale_kernel(const float2* __restrict__ x, float2* __restrict__ y, int n, float g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float2 v = x[i]; y[i] = make_float2(__fmul_rn(g, v.x), __fmul_rn(g, v.y)); }
}
cudaError_t launch_agc_power_seq(const float2* x, int n, float* sum, cudaStream_t st) { agc_power_seq_kernel<<<1, 32, 0, st>>>(x, n, sum); return cudaGetLastError(); }
cudaError_t launch_agc_scale(const float2* x, float2* y, int n, float g, cudaStream_t st) { agc_scale_kernel<<<(n + 127) / 128, 128, 0, st>>>(x, y, n, g); return cudaGetLastError(); }

} // namespace fm
