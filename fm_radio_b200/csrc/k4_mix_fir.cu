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
// One CTA = 1024 MPX-rate samples (+128 samples of FIR history) of TWO streams, 2p and 2p+1, held
// as packed float2 (stream 2p, stream 2p+1) exactly like K2: the signals are real-valued planes, so
// the pair for sm_100's two-wide FP32 instructions is made of two streams, every tap is one FFMA2
// and every mixing / polynomial-sine step is a packed op.  (First version: one stream per CTA,
// scalar FFMA, 34 % of the FP32 peak with the issue slots as the limit.)  An odd last stream is
// paired with itself and its second half discarded.
//
// The mixed signals are produced ONCE per sample into bank-skewed planar shared arrays (the
// oscillator is the reference's polynomial sine, coefficients verbatim); the global loads of a
// thread's 9 samples are all issued before the first use.  Three of the four warps then run three FIR
// roles of equal weight (1024 / 1152 / 1024 FFMA2 per lane; the fourth only mixes and post-processes --
// measured faster than three-warp CTAs, 0.121 vs 0.133 ms, and than splitting RDS over two warps):
//   warp 0  L+R   : /4 FIR of the real plane (the reference discards the imaginary part, :477-479)
//   warp 1  L-R   : /4 FIR of the imaginary plane of the 38 kHz mixdown -- the audio -- and the real
//                   part ONLY at every 10th output, the ones the phase estimator reads (:496-511)
//   warp 2  RDS   : /8 FIR of the real, then the imaginary plane of the 57 kHz mixdown
// each thread owning 8 (/4) or 4 (/8) consecutive outputs and sliding a register window of sample
// quads (taps are stored duplicated in shared memory so that one broadcast LDS.128 yields two ready
// FFMA2 operands).  Results go through shared memory (aliased onto the dead
// sample planes) to the MixAudio / estimator epilogue, which writes coalesced.  History of the MIXED
// signals is carried (not recomputed) because the reference's FIR history holds samples mixed with
// the previous block's phase offset.
#include "fm_common.cuh"
#include <atomic>

namespace fm {

constexpr int K4_LEN = K4_NN + K4_TS;                         // 1152 staged samples
__device__ __host__ __forceinline__ constexpr int a4(int i) { return i + 2 * (i >> 5); }   // pair units: +2 pairs per 32
constexpr int K4_PLEN = a4(K4_LEN) + 8;                       // padded plane length (+ the quad a window may over-read)
constexpr int K4_SPARSE_MAX = 32;                             // >= ceil(256 / 10) estimator outputs per tile
constexpr int K4_SMEM_BYTES = (5 * K4_PLEN + 3 * K4_NN) * (int)sizeof(float2);
// epilogue arrays, aliased onto the sample planes once every FIR has finished
constexpr int K4_RES_LPR = 0, K4_RES_LMR = K4_TS / 4, K4_RES_SPARSE = 2 * (K4_TS / 4), K4_RES_EST = K4_RES_SPARSE + K4_SPARSE_MAX,
              K4_RES_END = K4_RES_EST + K4_THREADS;
static_assert(K4_RES_END <= 5 * K4_PLEN, "epilogue arrays must fit in the planes");

__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// x - nearest integer (ties to even, as the AVX path's _MM_FROUND_TO_NEAREST_INT, apply_harmonic_pll.cpp:129),
// two at a time: the 1.5 * 2^23 trick is exact for |x| < 2^22
__device__ __forceinline__ float2 wrap2(float2 x) {
    const float2 big = bc2(12582912.0f);
    return __fadd2_rn(x, neg2(__fadd2_rn(__fadd2_rn(x, big), neg2(big))));
}

// dsp/simd/chebyshev_sine.h:13-41, two at a time (same Horner order, every step one FFMA per half)
__device__ __forceinline__ float2 chebyshev_sine2(float2 x) {
    const float2 z = __fmul2_rn(x, x);
    float2 b = bc2(3.20396066f);
    b = __ffma2_rn(b, z, bc2(-14.07150173f));
    b = __ffma2_rn(b, z, bc2(38.50016403f));
    b = __ffma2_rn(b, z, bc2(-67.07687378f));
    b = __ffma2_rn(b, z, bc2(64.83583069f));
    b = __ffma2_rn(b, z, bc2(-25.13274193f));
    return __fmul2_rn(__fmul2_rn(b, __fadd2_rn(z, bc2(-0.25f))), x);
}

// R consecutive outputs of a decimate-by-M 128-tap FIR over a skewed plane of pairs (M = 4: R = 8,
// M = 8: R = 4).  Output r, tap quad pq reads sample quad (s0/4) + (M/4)*r + pq.  The register window is a
// ring of 8 quads indexed modulo 8; the tap loop is unrolled by 8, so every ring index is static (the
// window is renamed, never moved) while the code stays small enough for the instruction cache (the fully
// unrolled version stalled on instruction fetch, ncu no_instruction 1.5).  Step pq's new quad, pq + 8,
// goes into the slot of quad pq right after its last use.  taps2[k] = (b[k], b[k]).  Sum order: taps
// ascending, as the scalar reference.
template <int M, int R, bool CTAPS>
__device__ __forceinline__ void fir128(const float2* __restrict__ sig, const float2* __restrict__ taps2, const float (&ctaps)[K4_NN], int s0, float2 (&acc)[R])
{
    constexpr int QS = M / 4;
    static_assert(QS * (R - 1) + 1 <= 8, "window must fit the ring");
    float4 win[8][2];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        win[j][0] = *(const float4*)(sig + a4(s0 + 4 * j));
        win[j][1] = *(const float4*)(sig + a4(s0 + 4 * j + 2));
    }
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int pb = 0; pb < 32; pb += 8) {
#pragma unroll
        for (int pp = 0; pp < 8; pp++) {
            const int pq = pb + pp;
            // the four taps of this step: CTAPS = straight from the kernel parameters (constant bank, uniform-register
            // operands: no shared-memory traffic -- the broadcast tap loads were half of this kernel's LDS count and the
            // shared-memory pipe was as busy as the FMA pipe), else from the duplicated copy in shared memory
            float4 b01, b23;
            if (CTAPS) {
                b01 = make_float4(ctaps[4 * pq], ctaps[4 * pq], ctaps[4 * pq + 1], ctaps[4 * pq + 1]);
                b23 = make_float4(ctaps[4 * pq + 2], ctaps[4 * pq + 2], ctaps[4 * pq + 3], ctaps[4 * pq + 3]);
            } else {
                b01 = *(const float4*)(taps2 + 4 * pq);               // (b0, b0, b1, b1)
                b23 = *(const float4*)(taps2 + 4 * pq + 2);
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float4 x01 = win[(pp + QS * r) & 7][0], x23 = win[(pp + QS * r) & 7][1];
                acc[r] = __ffma2_rn(make_float2(x01.x, x01.y), make_float2(b01.x, b01.y), acc[r]);
                acc[r] = __ffma2_rn(make_float2(x01.z, x01.w), make_float2(b01.z, b01.w), acc[r]);
                acc[r] = __ffma2_rn(make_float2(x23.x, x23.y), make_float2(b23.x, b23.y), acc[r]);
                acc[r] = __ffma2_rn(make_float2(x23.z, x23.w), make_float2(b23.z, b23.w), acc[r]);
                if (r == 0) {   // quad pq is dead now: its slot takes quad pq + 8 (<= 1 quad past the data at the very end: padded)
                    win[pp][0] = *(const float4*)(sig + a4(s0 + 4 * (pq + 8)));
                    win[pp][1] = *(const float4*)(sig + a4(s0 + 4 * (pq + 8) + 2));
                }
            }
        }
    }
}

template <bool CTAPS>
__global__ void __launch_bounds__(K4_THREADS, 4)
k4_mix_fir(const float2* __restrict__ fm_out_iq, const float* __restrict__ pll_dt,
           const float* __restrict__ hist_x_in, const float2* __restrict__ hist_m2_in, const float2* __restrict__ hist_m3_in,
           float* __restrict__ hist_x_out, float2* __restrict__ hist_m2_out, float2* __restrict__ hist_m3_out,
           const float* __restrict__ lmr_phase, float2* __restrict__ audio_out, float2* __restrict__ rds_out,
           float* __restrict__ est_partial, float* __restrict__ rds_power_partial,
           float* __restrict__ dbg_lpr, float* __restrict__ dbg_lmr, const __grid_constant__ K4Params p)
{
    extern __shared__ __align__(16) float2 smem4[];
    float2* s_sig = smem4;                                      // [5][K4_PLEN]: xr, m2r, m2i, m3r, m3i
    float2* s_taps = smem4 + 5 * K4_PLEN;                       // [3][K4_NN] duplicated taps: lpr, lmr, rds
    float2* s_res = smem4;                                      // epilogue arrays (aliased, see K4_RES_*)
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int tile = blockIdx.x;
    const int sA = 2 * blockIdx.y;
    const bool hasB = sA + 1 < p.n_streams;
    const int sB = hasB ? sA + 1 : sA;
    const int n0 = tile * K4_TS;
    const int nts = min(K4_TS, p.n - n0);                       // multiple of 128

    for (int k = t; k < K4_NN; k += K4_THREADS) {
        s_taps[k] = bc2(p.taps_lpr[k]); s_taps[K4_NN + k] = bc2(p.taps_lmr[k]); s_taps[2 * K4_NN + k] = bc2(p.taps_rds[k]);
    }
    if (t < 8) {
#pragma unroll
        for (int a = 0; a < 5; a++) s_sig[a * K4_PLEN + K4_PLEN - 8 + t] = make_float2(0.0f, 0.0f);   // the quad a window may over-read
    }
    // ---- stage + mix: sample idx of the staged array <-> MPX sample n0 - 128 + idx ----
    // the L-R phase offset of the block (broadcast_fm_demod.cpp:485-488), as a rotation: (cos, sin)(2 pi off)
    const float2 off2 = wrap2(make_float2(lmr_phase[sA], lmr_phase[sB]));
    const float2 off_c = chebyshev_sine2(wrap2(__fadd2_rn(off2, bc2(0.25f)))), off_s = chebyshev_sine2(off2);
    const float2* xA = fm_out_iq + (size_t)sA * p.n, * xB = fm_out_iq + (size_t)sB * p.n;
    const float* dA = pll_dt + (size_t)sA * p.n, * dB = pll_dt + (size_t)sB * p.n;
    constexpr int NIT = (K4_LEN + K4_THREADS - 1) / K4_THREADS;   // 9
    float2 vxa[NIT], vxb[NIT], vdt[NIT];
#pragma unroll
    for (int it = 0; it < NIT; it++) {                          // every global load first
        const int idx = t + it * K4_THREADS;
        const int n = n0 - K4_NN + idx;
        if (idx < K4_NN + nts && n >= 0) {
            vxa[it] = __ldg(xA + n); vxb[it] = __ldg(xB + n);
            vdt[it] = make_float2(__ldg(dA + n), __ldg(dB + n));
        } else { vxa[it] = vxb[it] = vdt[it] = make_float2(0.0f, 0.0f); }
    }
#pragma unroll
    for (int it = 0; it < NIT; it++) {
        const int idx = t + it * K4_THREADS;
        const int n = n0 - K4_NN + idx;
        if (idx >= K4_NN + nts) continue;
        float2 xr, m2r, m2i, m3r, m3i;
        if (n < 0) {                                            // tile 0: the halo is the carried history
            xr = make_float2(hist_x_in[(size_t)sA * K4_NN + idx], hist_x_in[(size_t)sB * K4_NN + idx]);
            const float2 h2a = hist_m2_in[(size_t)sA * K4_NN + idx], h2b = hist_m2_in[(size_t)sB * K4_NN + idx];
            const float2 h3a = hist_m3_in[(size_t)sA * K4_NN + idx], h3b = hist_m3_in[(size_t)sB * K4_NN + idx];
            m2r = make_float2(h2a.x, h2b.x); m2i = make_float2(h2a.y, h2b.y);
            m3r = make_float2(h3a.x, h3b.x); m3i = make_float2(h3a.y, h3b.y);
        } else {
            xr = make_float2(vxa[it].x, vxb[it].x);
            const float2 xi = make_float2(vxa[it].y, vxb[it].y);
            // apply_harmonic_pll.cpp:16-23: y = x * (S(phi + 1/4 wrapped), S(phi wrapped)), phi = dt*h + off, for
            // h = 2 (+ the L-R phase offset) and h = 3.  The reference evaluates its polynomial sine four times per
            // sample; here the fundamental (cos, sin)(2 pi dt) is evaluated ONCE with that polynomial (dt is already
            // wrapped to [-1/2, 1/2]; cos(2 pi t) = S(1/4 - |t|) needs no wrap) and the harmonics follow by angle
            // addition -- 41 packed instructions per sample instead of 66 (the mixdown was 43 % of this kernel's FMA
            // work).  The products add ~1e-7 of rounding to a unit oscillator, the size of the polynomial's own error
            // (3.6e-8 mean, 1.5e-7 max) and of the reference's phase rounding (ulp(phi) = 6e-8 turns = 4e-7 rad).
            const float2 t1 = vdt[it];
            const float2 c1 = chebyshev_sine2(__fadd2_rn(bc2(0.25f), make_float2(-fabsf(t1.x), -fabsf(t1.y))));
            const float2 s1 = chebyshev_sine2(t1);
            const float2 c2 = __ffma2_rn(c1, c1, neg2(__fmul2_rn(s1, s1)));
            const float2 s2 = __fmul2_rn(__fadd2_rn(s1, s1), c1);
            const float2 c3 = __ffma2_rn(c2, c1, neg2(__fmul2_rn(s2, s1)));
            const float2 s3 = __ffma2_rn(s2, c1, __fmul2_rn(c2, s1));
            const float2 C2 = __ffma2_rn(c2, off_c, neg2(__fmul2_rn(s2, off_s)));     // e^{j 2 pi (2 dt + off)}
            const float2 S2 = __ffma2_rn(s2, off_c, __fmul2_rn(c2, off_s));
            m2r = __ffma2_rn(xr, C2, neg2(__fmul2_rn(xi, S2))); m2i = __ffma2_rn(xr, S2, __fmul2_rn(xi, C2));
            m3r = __ffma2_rn(xr, c3, neg2(__fmul2_rn(xi, s3))); m3i = __ffma2_rn(xr, s3, __fmul2_rn(xi, c3));
        }
        const int a = a4(idx);
        s_sig[a] = xr; s_sig[K4_PLEN + a] = m2r; s_sig[2 * K4_PLEN + a] = m2i; s_sig[3 * K4_PLEN + a] = m3r; s_sig[4 * K4_PLEN + a] = m3i;
    }
    __syncthreads();

    // ---- history for the next block: the last 128 staged samples of the streams' last tile ----
    if (n0 + nts == p.n) for (int i = t; i < K4_NN; i += K4_THREADS) {
        const int a = a4(nts + i);
        const float2 x = s_sig[a], r2 = s_sig[K4_PLEN + a], i2 = s_sig[2 * K4_PLEN + a], r3 = s_sig[3 * K4_PLEN + a], i3 = s_sig[4 * K4_PLEN + a];
        hist_x_out[(size_t)sA * K4_NN + i] = x.x;
        hist_m2_out[(size_t)sA * K4_NN + i] = make_float2(r2.x, i2.x);
        hist_m3_out[(size_t)sA * K4_NN + i] = make_float2(r3.x, i3.x);
        if (hasB) {
            hist_x_out[(size_t)sB * K4_NN + i] = x.y;
            hist_m2_out[(size_t)sB * K4_NN + i] = make_float2(r2.y, i2.y);
            hist_m3_out[(size_t)sB * K4_NN + i] = make_float2(r3.y, i3.y);
        }
    }

    // ---- the FIR roles; results stay in registers until every warp has finished with the planes ----
    const int n_audio = nts >> 2, n_rds = nts >> 3;
    const int gi0 = n0 >> 2;                                    // audio index of the tile's first output within the block
    const int o_first = (10 - gi0 % 10) % 10;                   // first output of the tile the estimator reads
    float2 acc8[8], sparse = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int r = 0; r < 8; r++) acc8[r] = make_float2(0.0f, 0.0f);
    if (warp < 2) {
        // /4 FIR: output o = 8*lane + r reads staged samples 4o + 4 + k
        if (8 * lane < n_audio) {
            if (warp == 0) fir128<4, 8, CTAPS>(s_sig, s_taps, p.taps_lpr, 32 * lane + 4, acc8);
            else           fir128<4, 8, CTAPS>(s_sig + 2 * K4_PLEN, s_taps + K4_NN, p.taps_lmr, 32 * lane + 4, acc8);
        }
    } else if (warp == 2) {
        // /8 FIR, real then imaginary plane: output o = 4*lane + r reads staged samples 8o + 8 + k.
        // This warp owns its outputs completely: RDS samples and their AGC power partial leave from registers.
        float2 pw = make_float2(0.0f, 0.0f);
        if (4 * lane < n_rds) {
            float2 re[4], im[4];
            fir128<8, 4, CTAPS>(s_sig + 3 * K4_PLEN, s_taps + 2 * K4_NN, p.taps_rds, 32 * lane + 8, re);
            fir128<8, 4, CTAPS>(s_sig + 4 * K4_PLEN, s_taps + 2 * K4_NN, p.taps_rds, 32 * lane + 8, im);
            float4* dA4 = (float4*)(rds_out + (size_t)sA * (p.n >> 3) + (n0 >> 3) + 4 * lane);
            dA4[0] = make_float4(re[0].x, im[0].x, re[1].x, im[1].x);
            dA4[1] = make_float4(re[2].x, im[2].x, re[3].x, im[3].x);
            if (hasB) {
                float4* dB4 = (float4*)(rds_out + (size_t)sB * (p.n >> 3) + (n0 >> 3) + 4 * lane);
                dB4[0] = make_float4(re[0].y, im[0].y, re[1].y, im[1].y);
                dB4[1] = make_float4(re[2].y, im[2].y, re[3].y, im[3].y);
            }
#pragma unroll
            for (int r = 0; r < 4; r++) pw = __fadd2_rn(pw, __ffma2_rn(re[r], re[r], __fmul2_rn(im[r], im[r])));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { pw.x += __shfl_xor_sync(0xffffffffu, pw.x, off); pw.y += __shfl_xor_sync(0xffffffffu, pw.y, off); }
        if (lane == 0) {
            rds_power_partial[(size_t)sA * p.n_tiles + tile] = pw.x;
            if (hasB) rds_power_partial[(size_t)sB * p.n_tiles + tile] = pw.y;
        }
    } else {
        // warp 3 (no FIR role of its own): the 26 sparse outputs, off warp 1's critical path
        // real part of L-R where the estimator looks: output o_first + 10*lane, taps ascending
        const int o = o_first + 10 * lane;
        if (o < n_audio) {
            const float2* sg = s_sig + K4_PLEN;
            const float2* tp = s_taps + K4_NN;
#pragma unroll 4
            for (int q = 0; q < 32; q++) {
                const float4 x01 = *(const float4*)(sg + a4(4 * o + 4 + 4 * q)), x23 = *(const float4*)(sg + a4(4 * o + 4 + 4 * q + 2));
                const float4 b01 = *(const float4*)(tp + 4 * q), b23 = *(const float4*)(tp + 4 * q + 2);
                sparse = __ffma2_rn(make_float2(x01.x, x01.y), make_float2(b01.x, b01.y), sparse);
                sparse = __ffma2_rn(make_float2(x01.z, x01.w), make_float2(b01.z, b01.w), sparse);
                sparse = __ffma2_rn(make_float2(x23.x, x23.y), make_float2(b23.x, b23.y), sparse);
                sparse = __ffma2_rn(make_float2(x23.z, x23.w), make_float2(b23.z, b23.w), sparse);
            }
        }
    }
    __syncthreads();                                            // planes and taps are dead from here on
    if (warp < 2) {
        float2* d = s_res + (warp == 0 ? K4_RES_LPR : K4_RES_LMR) + 8 * lane;
#pragma unroll
        for (int q = 0; q < 4; q++) *(float4*)(d + 2 * q) = make_float4(acc8[2 * q].x, acc8[2 * q].y, acc8[2 * q + 1].x, acc8[2 * q + 1].y);
    }
    if (warp == 3) s_res[K4_RES_SPARSE + lane] = sparse;
    __syncthreads();

    // ---- MixAudio (:549-585) + phase-estimator partial sum (:496-511), 2 outputs per thread ----
    float2 est = make_float2(0.0f, 0.0f);
    for (int it = t; 2 * it < n_audio; it += K4_THREADS) {
        const int o = 2 * it;
        const size_t gi = (size_t)gi0 + o;                      // audio index within the block
        float4 frA, frB;
        float* fA = &frA.x; float* fB = &frB.x;
        float2 lprs[2], lmrs[2];
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const float2 lpr = s_res[K4_RES_LPR + o + j], lmr = s_res[K4_RES_LMR + o + j];
            lprs[j] = lpr; lmrs[j] = lmr;
            float2 L, R;
            if (p.audio_out_mode == 2) { L = __ffma2_rn(bc2(p.stereo_mix), lmr, lpr); R = __ffma2_rn(bc2(-p.stereo_mix), lmr, lpr); }
            else if (p.audio_out_mode == 1) { L = lmr; R = lmr; }
            else { L = lpr; R = lpr; }
            fA[2 * j] = L.x * 2.0f; fA[2 * j + 1] = R.x * 2.0f;
            fB[2 * j] = L.y * 2.0f; fB[2 * j + 1] = R.y * 2.0f;
            if ((gi + j) % 10 == 0) {
                const float2 lre = s_res[K4_RES_SPARSE + (o + j - o_first) / 10];
                const float pa = atan2f(lmr.x, lre.x), pb = atan2f(lmr.y, lre.y);
                est.x += (pa > 0.0f) ? (PI_F / 2.0f - pa) : (-PI_F / 2.0f - pa);
                est.y += (pb > 0.0f) ? (PI_F / 2.0f - pb) : (-PI_F / 2.0f - pb);
            }
        }
        *(float4*)(audio_out + (size_t)sA * (p.n >> 2) + gi) = frA;
        if (hasB) *(float4*)(audio_out + (size_t)sB * (p.n >> 2) + gi) = frB;
        if (p.keep) {
            *(float2*)(dbg_lpr + (size_t)sA * (p.n >> 2) + gi) = make_float2(lprs[0].x, lprs[1].x);
            *(float2*)(dbg_lmr + (size_t)sA * (p.n >> 2) + gi) = make_float2(lmrs[0].x, lmrs[1].x);
            if (hasB) {
                *(float2*)(dbg_lpr + (size_t)sB * (p.n >> 2) + gi) = make_float2(lprs[0].y, lprs[1].y);
                *(float2*)(dbg_lmr + (size_t)sB * (p.n >> 2) + gi) = make_float2(lmrs[0].y, lmrs[1].y);
            }
        }
    }
    s_res[K4_RES_EST + t] = est;
    __syncthreads();
    if (warp == 0) {
        constexpr int EPL = K4_THREADS / 32;                    // partial sums per lane
        float2 v = s_res[K4_RES_EST + EPL * lane];
#pragma unroll
        for (int q = 1; q < EPL; q++) { const float2 e = s_res[K4_RES_EST + EPL * lane + q]; v.x += e.x; v.y += e.y; }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { v.x += __shfl_xor_sync(0xffffffffu, v.x, off); v.y += __shfl_xor_sync(0xffffffffu, v.y, off); }
        if (lane == 0) {
            est_partial[(size_t)sA * p.n_tiles + tile] = v.x;
            if (hasB) est_partial[(size_t)sB * p.n_tiles + tile] = v.y;
        }
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
    // the attribute is per device: one flag per ordinal (a handle may live on any device of the process)
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(k4_mix_fir<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K4_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k4_mix_fir<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K4_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    const dim3 grid(p.n_tiles, (p.n_streams + 1) / 2);
    if (p.balanced) k4_mix_fir<true><<<grid, K4_THREADS, K4_SMEM_BYTES, st>>>(fm_out_iq, pll_dt, hist_x_in, hist_m2_in, hist_m3_in,
                                            hist_x_out, hist_m2_out, hist_m3_out, lmr_phase, audio_out, rds_out,
                                            est_partial, rds_power_partial, dbg_lpr, dbg_lmr, p);
    else k4_mix_fir<false><<<grid, K4_THREADS, K4_SMEM_BYTES, st>>>(fm_out_iq, pll_dt, hist_x_in, hist_m2_in, hist_m3_in,
                                            hist_x_out, hist_m2_out, hist_m3_out, lmr_phase, audio_out, rds_out,
                                            est_partial, rds_power_partial, dbg_lpr, dbg_lmr, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k4b_lmr_phase<<<(p.n_streams + 127) / 128, 128, 0, st>>>(est_partial, lmr_phase, p.n_tiles, p.n >> 2, p.n_streams);
    return cudaGetLastError();
}

} // namespace fm
