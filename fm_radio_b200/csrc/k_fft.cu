// Display spectra: the forward DFT + FFT shift behind the reference's UpdateFFTCalc
// (fm_demod/broadcast_fm_demod.cpp:27-40): CalculateFFT (dsp/calculate_fft.cpp:43-50, FFTW3f there -- an external
// library that is not vendored; what it computes is the plain forward DFT X[k] = sum_n x[n] exp(-2 pi i n k / N))
// followed by InplaceFFTShift (dsp/fftshift.h:21-33).  The dB magnitude / averaging step
// (Calculate_FFT_Mag::Process, dsp/calculate_fft_mag.cpp:11-45) stays where the reference runs it, on the host in
// the shim, with the reference's own class.
//
// Off the hot path: it runs only when a GUI window raises a trigger (never in fm_demod_benchmark), on one
// stream's buffer of 1 024 .. 65 536 points.  Radix-2 Stockham autosort, one launch per stage, twiddles from
// sincospif of an exactly representable fraction; the FFT shift is folded into the last stage's store.
#include "fm_common.cuh"

namespace fm {

// stage with butterflies of span ns (1, 2, 4 .. n/2): out[(j / ns) * 2 ns + k (+ ns)] = in[j] +- w^k in[j + n/2], k = j % ns
template <bool REAL_IN>
__global__ void fft_stage(const float* __restrict__ in, float2* __restrict__ out, int n, int ns, int shift)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = n >> 1;
    if (j >= half) return;
    float2 a, b;
    if (REAL_IN) { a = make_float2(in[j], 0.0f); b = make_float2(in[j + half], 0.0f); }
    else { a = ((const float2*)in)[j]; b = ((const float2*)in)[j + half]; }
    const int k = j & (ns - 1);
    float sn, cs;
    sincospif(-(float)k / (float)ns, &sn, &cs);          // w = exp(-i pi k / ns)
    const float2 wb = make_float2(b.x * cs - b.y * sn, b.x * sn + b.y * cs);
    int j0 = ((j - k) << 1) + k;
    int j1 = j0 + ns;
    if (shift) { j0 = (j0 + half) & (n - 1); j1 = (j1 + half) & (n - 1); }    // InplaceFFTShift on the final stage
    out[j0] = make_float2(a.x + wb.x, a.y + wb.y);
    out[j1] = make_float2(a.x - wb.x, a.y - wb.y);
}

// in: n complex (or real) samples; work0 / work1: n complex each; returns the buffer holding the result
cudaError_t launch_fft(const float* in, int in_is_real, float2* work0, float2* work1, int n, int fftshift,
                       float2** result, cudaStream_t st)
{
    if (n < 2 || (n & (n - 1)) != 0) return cudaErrorInvalidValue;
    const int threads = 256, grid = (n / 2 + threads - 1) / threads;
    const float* src = in;
    float2* dst = work0;
    bool first = true;
    for (int ns = 1; ns < n; ns <<= 1) {
        const int last = (ns << 1) == n;
        if (first && in_is_real) fft_stage<true><<<grid, threads, 0, st>>>(src, dst, n, ns, last && fftshift);
        else                     fft_stage<false><<<grid, threads, 0, st>>>(src, dst, n, ns, last && fftshift);
        first = false;
        src = (const float*)dst;
        dst = (dst == work0) ? work1 : work0;
    }
    *result = (float2*)src;
    return cudaGetLastError();
}

} // namespace fm
