// K4: 2nd / 3rd pilot-harmonic mixdown fused with the three 128-tap polyphase decimators and the
// stereo mix.  K4b: per-stream update of the L-R phase offset.
//
// Replaces Broadcast_FM_Demod::ExtractComponents + MixAudio (reference file:line under src/):
//   L+R: PolyphaseDownsampler<cf32> M=4 K=32 NN=128, real part     broadcast_fm_demod.cpp:475-479
//   apply_harmonic_pll_auto(h=2, off=audio_lmr_phase_error)        :485-488, dsp/simd/apply_harmonic_pll.cpp:88-139
//   L-R: same /4 FIR on the mixed signal, imag part                :490, 519-521
//   phase estimator over every 10th output                         :496-517
//   apply_harmonic_pll_auto(h=3, off=0) + /8 FIR M=8 K=16 NN=128   :527-533
//   MixAudio                                                       :549-585
//   AGC_Filter power sum for the RDS AGC (sum only)                dsp/agc.h:21-30
//
// One CTA = 1024 MPX-rate samples of one stream (+128 samples of FIR history).  The mixed signals
// are produced ONCE per sample into bank-skewed planar shared arrays (the oscillator comes from the
// reference's polynomial sine, coefficients verbatim) and the four warps then run the four FIR
// roles: L+R (real FIR only: the reference discards the imaginary part, :477-479), L-R real,
// L-R imag, RDS real+imag.  Each thread owns 8 (/4) or 4 (/8) consecutive outputs and slides a
// register window of input quads: one LDS.128 of samples + one broadcast LDS.128 of taps per 32
// (16) FFMAs.  History of the MIXED signals is carried (not recomputed) because the reference's
// FIR history holds samples mixed with the previous block's phase offset.
#include "fm_common.cuh"

namespace fm {

constexpr int K4_LEN = K4_NN + K4_TS;                         // 1152 staged samples
constexpr int K4_PLEN = K4_LEN + 4 * (K4_LEN >> 5);           // padded: +4 floats per 32
__device__ __forceinline__ int a4(int i) { return i + 4 * (i >> 5); }
__device__ __forceinline__ int a4q(int q) { return 4 * q + 4 * (q >> 3); }   // quad q = floats [4q,4q+4)

__device__ __forceinline__ float dot4(const float4 x, const float4 b, float acc) {
    acc = fmaf(x.x, b.x, acc); acc = fmaf(x.y, b.y, acc);
    acc = fmaf(x.z, b.z, acc); acc = fmaf(x.w, b.w, acc);
    return acc;
}

// R consecutive outputs of a decimate-by-M 128-tap FIR (M = 4: R = 8, M = 8: R = 4).
// Output r, tap quad pq reads array quad q0 + (M/4)*r + pq; the window of W = (M/4)*(R-1)+1 quads
// slides by one quad per pq.  Sum order: taps ascending, as the scalar reference.
template <int M, int R>
__device__ __forceinline__ void fir128(const float* __restrict__ sig, const float* __restrict__ taps, int q0, float (&acc)[R])
{
    constexpr int QS = M / 4;
    constexpr int W = QS * (R - 1) + 1;
    float4 win[W];
#pragma unroll
    for (int j = 0; j < W; j++) win[j] = *(const float4*)(sig + a4q(q0 + j));
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0.0f;
#pragma unroll 1
    for (int pb = 0; pb < 32; pb += 8) {
#pragma unroll
        for (int pp = 0; pp < 8; pp++) {
            const int pq = pb + pp;
            const float4 b = *(const float4*)(taps + 4 * pq);
#pragma unroll
            for (int r = 0; r < R; r++) acc[r] = dot4(win[QS * r], b, acc[r]);
#pragma unroll
            for (int j = 0; j < W - 1; j++) win[j] = win[j + 1];
            win[W - 1] = *(const float4*)(sig + a4q(q0 + W + pq));   // may read <= 1 quad past the data: padded
        }
    }
}

__global__ void __launch_bounds__(K4_THREADS)
k4_mix_fir(const float2* __restrict__ fm_out_iq, const float* __restrict__ pll_dt,
           const float* __restrict__ hist_x_in, const float2* __restrict__ hist_m2_in, const float2* __restrict__ hist_m3_in,
           float* __restrict__ hist_x_out, float2* __restrict__ hist_m2_out, float2* __restrict__ hist_m3_out,
           const float* __restrict__ lmr_phase, float2* __restrict__ audio_out, float2* __restrict__ rds_out,
           float* __restrict__ est_partial, float* __restrict__ rds_power_partial,
           float* __restrict__ dbg_lpr, float* __restrict__ dbg_lmr, const __grid_constant__ K4Params p)
{
    __shared__ __align__(16) float s_sig[5][K4_PLEN + 8];       // xr, m2r, m2i, m3r, m3i
    __shared__ __align__(16) float s_taps[3][K4_NN];
    __shared__ __align__(16) float s_res[3][K4_TS / 4];         // lpr, lmr_re, lmr_im
    __shared__ float s_est[K4_THREADS];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int tile = blockIdx.x, s = blockIdx.y;
    const int n0 = tile * K4_TS;
    const int nts = min(K4_TS, p.n - n0);                       // multiple of 128

    for (int k = t; k < K4_NN; k += K4_THREADS) {
        s_taps[0][k] = p.taps_lpr[k]; s_taps[1][k] = p.taps_lmr[k]; s_taps[2][k] = p.taps_rds[k];
    }
    if (t < 8) {
#pragma unroll
        for (int a = 0; a < 5; a++) s_sig[a][K4_PLEN + t] = 0.0f;   // the quad the window may over-read
    }
    // ---- stage + mix: sample idx of the staged array <-> MPX sample n0 - 128 + idx ----
    const float off2 = lmr_phase[s];
    for (int idx = t; idx < K4_NN + nts; idx += K4_THREADS) {
        float xr, m2r, m2i, m3r, m3i;
        const int n = n0 - K4_NN + idx;
        if (n < 0) {
            xr = hist_x_in[(size_t)s * K4_NN + idx];
            const float2 h2 = hist_m2_in[(size_t)s * K4_NN + idx];
            const float2 h3 = hist_m3_in[(size_t)s * K4_NN + idx];
            m2r = h2.x; m2i = h2.y; m3r = h3.x; m3i = h3.y;
        } else {
            const float2 x = fm_out_iq[(size_t)s * p.n + n];
            const float dt = pll_dt[(size_t)s * p.n + n];
            xr = x.x;
            {   // apply_harmonic_pll.cpp:16-23 with round-to-nearest-even as the AVX path (:129)
                float ds = fmaf(dt, p.harmonic_lmr, off2);
                float dc = ds + 0.25f;
                ds = ds - rintf(ds); dc = dc - rintf(dc);
                const float c = chebyshev_sine(dc), sn = chebyshev_sine(ds);
                m2r = x.x * c - x.y * sn; m2i = x.x * sn + x.y * c;
            }
            {
                float ds = dt * p.harmonic_rds;
                float dc = ds + 0.25f;
                ds = ds - rintf(ds); dc = dc - rintf(dc);
                const float c = chebyshev_sine(dc), sn = chebyshev_sine(ds);
                m3r = x.x * c - x.y * sn; m3i = x.x * sn + x.y * c;
            }
        }
        const int a = a4(idx);
        s_sig[0][a] = xr; s_sig[1][a] = m2r; s_sig[2][a] = m2i; s_sig[3][a] = m3r; s_sig[4][a] = m3i;
    }
    __syncthreads();

    // ---- history for the next block: the last 128 staged samples of the stream's last tile ----
    if (n0 + nts == p.n) {
        const int a = a4(nts + t);                               // K4_THREADS == K4_NN
        hist_x_out[(size_t)s * K4_NN + t] = s_sig[0][a];
        hist_m2_out[(size_t)s * K4_NN + t] = make_float2(s_sig[1][a], s_sig[2][a]);
        hist_m3_out[(size_t)s * K4_NN + t] = make_float2(s_sig[3][a], s_sig[4][a]);
    }

    const int n_audio = nts >> 2, n_rds = nts >> 3;
    float rds_pw = 0.0f;
    if (warp < 3) {
        // /4 FIR: output o = 8*lane + r reads staged samples 4o+4+k, i.e. quads o+1+pq
        if (8 * lane < n_audio) {
            float acc[8];
            fir128<4, 8>(s_sig[warp], s_taps[warp == 0 ? 0 : 1], 8 * lane + 1, acc);
            float4* d = (float4*)(&s_res[warp][8 * lane]);
            d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
    } else {
        // /8 FIR: output o = 4*lane + r reads staged samples 8o+8+k, i.e. quads 2o+2+pq
        if (4 * lane < n_rds) {
            float re[4], im[4];
            fir128<8, 4>(s_sig[3], s_taps[2], 8 * lane + 2, re);
            fir128<8, 4>(s_sig[4], s_taps[2], 8 * lane + 2, im);
            float4* d = (float4*)(rds_out + (size_t)s * (p.n >> 3) + (n0 >> 3) + 4 * lane);
            d[0] = make_float4(re[0], im[0], re[1], im[1]);
            d[1] = make_float4(re[2], im[2], re[3], im[3]);
#pragma unroll
            for (int r = 0; r < 4; r++) rds_pw += re[r] * re[r] + im[r] * im[r];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) rds_pw += __shfl_xor_sync(0xffffffffu, rds_pw, off);
        if (lane == 0) rds_power_partial[(size_t)s * p.n_tiles + tile] = rds_pw;
    }
    __syncthreads();

    // ---- MixAudio (:549-585) + phase-estimator partial sum (:496-511), 2 outputs per thread ----
    float est = 0.0f;
    if (2 * t < n_audio) {
        const int o = 2 * t;
        const size_t gi = (size_t)(n0 >> 2) + o;                // audio index within the block
        float4 fr;
        float* f = &fr.x;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const float lpr = s_res[0][o + j], lre = s_res[1][o + j], lmr = s_res[2][o + j];
            float L, R;
            if (p.audio_out_mode == 2) { L = fmaf(p.stereo_mix, lmr, lpr); R = fmaf(-p.stereo_mix, lmr, lpr); }
            else if (p.audio_out_mode == 1) { L = lmr; R = lmr; }
            else { L = lpr; R = lpr; }
            f[2 * j] = L * 2.0f; f[2 * j + 1] = R * 2.0f;
            if ((gi + j) % 10 == 0) {
                const float phase = atan2f(lmr, lre);
                est += (phase > 0.0f) ? (PI_F / 2.0f - phase) : (-PI_F / 2.0f - phase);
            }
        }
        *(float4*)(audio_out + (size_t)s * (p.n >> 2) + gi) = fr;
        if (p.keep) {
            *(float2*)(dbg_lpr + (size_t)s * (p.n >> 2) + gi) = make_float2(s_res[0][o], s_res[0][o + 1]);
            *(float2*)(dbg_lmr + (size_t)s * (p.n >> 2) + gi) = make_float2(s_res[2][o], s_res[2][o + 1]);
        }
    }
    s_est[t] = est;
    __syncthreads();
    if (warp == 0) {
        float v = (s_est[4 * lane] + s_est[4 * lane + 1]) + (s_est[4 * lane + 2] + s_est[4 * lane + 3]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) est_partial[(size_t)s * p.n_tiles + tile] = v;
    }
}

// broadcast_fm_demod.cpp:511-516: avg over ceil(N_audio/10) samples, err += 0.1*avg, fmod 2 pi.
__global__ void k4b_lmr_phase(const float* __restrict__ est_partial, float* __restrict__ lmr_phase,
                              int n_tiles, int n_audio, int n_streams)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    float sum = 0.0f;
    for (int i = 0; i < n_tiles; i++) sum += est_partial[(size_t)s * n_tiles + i];
    const int total_samples = (n_audio + 9) / 10;
    const float avg = sum / (float)total_samples;
    float e = lmr_phase[s];
    e += 0.1f * avg;
    e = fmodf(e, 2.0f * PI_F);
    lmr_phase[s] = e;
}

cudaError_t launch_k4(const float2* fm_out_iq, const float* pll_dt,
                      const float* hist_x_in, const float2* hist_m2_in, const float2* hist_m3_in,
                      float* hist_x_out, float2* hist_m2_out, float2* hist_m3_out,
                      float* lmr_phase, float2* audio_out, float2* rds_out, float* est_partial,
                      float* rds_power_partial, float* dbg_lpr, float* dbg_lmr, const K4Params& p, cudaStream_t st)
{
    const dim3 grid(p.n_tiles, p.n_streams);
    k4_mix_fir<<<grid, K4_THREADS, 0, st>>>(fm_out_iq, pll_dt, hist_x_in, hist_m2_in, hist_m3_in,
                                            hist_x_out, hist_m2_out, hist_m3_out, lmr_phase, audio_out, rds_out,
                                            est_partial, rds_power_partial, dbg_lpr, dbg_lmr, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k4b_lmr_phase<<<(p.n_streams + 127) / 128, 128, 0, st>>>(est_partial, lmr_phase, p.n_tiles, p.n >> 2, p.n_streams);
    return cudaGetLastError();
}

} // namespace fm
