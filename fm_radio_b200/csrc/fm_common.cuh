// Shared definitions of the sm_100a demodulation kernels.
//
// Stage map (one CUDA stream per stage, see fmgpu.cu):
//   K1  k1_fir4_discrim  u8/cf32 IQ -> 64-tap /4 polyphase FIR -> atan2 discriminator -> fm_demod
//   K2  k2_mpx           fm_demod -> 64-tap /2 FIR -> [deemphasis] -> 65-tap Hilbert -> fm_out_iq
//                        -> 19 kHz peak IIR -> pilot angle (turns) + block power
//   K3  k3_pll           per-stream pilot PLL recurrence -> pll_dt
//   K4  k4_mix_fir       harmonic mixdown x2/x3 fused with the three 128-tap decimators, stereo mix
//   K4b k4b_lmr_phase    per-stream L-R phase-offset update (applied to the next block)
//   K5  k5_bpsk          per-stream RDS AGC + BPSK symbol synchroniser -> soft symbols
//   K6  k6_rds           per-stream RDS bit path: symbols -> groups -> PI / PS / RadioText
//   K7  k7_audio_pcm     audio output stage: 32 kHz frames -> device-rate frames (48 kHz) + int16 PCM
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

namespace fm {

constexpr int K1_NN = 64;       // taps of filt_poly_ds_lpf_fm_in (broadcast_fm_demod.cpp:133-144)
constexpr int K1_M = 4;
constexpr int K1_R = 16;        // consecutive outputs per thread
constexpr int K1_THREADS = 128;
constexpr int K1_TILE = K1_R * K1_THREADS;          // 2048 outputs = 8192 IQ samples per CTA
constexpr int K1_SEG = 16 * 8 + 4;                  // floats per 16-frame segment (+4 pad: bank skew)
constexpr int K1_HIST = 64;                         // IQ samples of history
constexpr float K1_TAP_SCALE = 1.2676506002282294e30f;   // 2^100
constexpr float K1_UNSCALE = 562949953421312.0f;         // 2^49 = 2^149 / 2^100

constexpr int K2_THREADS = 128;
constexpr int K2_R = 8;
constexpr int K2_CH = K2_THREADS * K2_R;            // 1024 fm_out samples per chunk (of each of the two streams of a CTA)
constexpr int K2_NN = 64;                           // filt_poly_ds_lpf_fm_out (:146-157)
constexpr int K2_HILB = 65;                         // filt_hilbert_transform (:192-195)

constexpr int K4_THREADS = 128;
constexpr int K4_TS = 1024;                         // fm_out-rate samples per CTA
constexpr int K4_NN = 128;                          // audio / rds decimator taps (:240-274)

constexpr float TWO_PI_F = 6.283185307179586f;
constexpr float INV_TWO_PI_F = 0.15915494309189535f;
constexpr float PI_F = 3.14159265358979323846f;

// dsp/simd/chebyshev_sine.h:13-41 -- sin(2 pi x) on [-0.5, 0.5]; coefficients verbatim, Horner in
// the reference's order (the compiler contracts each step to one FFMA, as the AVX2+FMA build does).
__device__ __forceinline__ float chebyshev_sine(float x) {
    const float z = x * x;
    float b = 3.20396066f;
    b = fmaf(b, z, -14.07150173f);
    b = fmaf(b, z, 38.50016403f);
    b = fmaf(b, z, -67.07687378f);
    b = fmaf(b, z, 64.83583069f);
    b = fmaf(b, z, -25.13274193f);
    return b * (z - 0.25f) * x;
}

// atan2 of the discriminator (K1) and of the pilot angle (K2), shared by every kernel (so the cf32 and u8
// entry points give identical bits): minimax odd polynomial of degree 15 on [0, 1], 1.2e-7 rad max error in fp32, one MUFU.RCP.
__device__ __forceinline__ float fm_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float q = __fdividef(mn, fmaxf(mx, 1e-30f));    // atan2(0, 0) = 0 as the reference's libm (0 / 1e-30, no select)
    const float s = q * q;
    float p = -0.004054562299f;
    p = fmaf(p, s, 0.021862939178f);
    p = fmaf(p, s, -0.05591229796f);
    p = fmaf(p, s, 0.096421950378f);
    p = fmaf(p, s, -0.139086285623f);
    p = fmaf(p, s, 0.1994656543f);
    p = fmaf(p, s, -0.333298607622f);
    p = fmaf(p, s, 0.999999335572f);
    float r = p * q;
    r = (ay > ax) ? (0.5f * PI_F - r) : r;
    r = (x < 0.0f) ? (PI_F - r) : r;
    return copysignf(r, y);
}

// Two atan2 at once for the packed FP32 pipe (FFMA2 / FMUL2, sm_100): operation for operation the
// same IEEE sequence as fm_atan2f, so both give identical bits for identical inputs.
__device__ __forceinline__ float2 fm_atan2f_x2(float y0, float x0, float y1, float x1) {
    const float ax0 = fabsf(x0), ay0 = fabsf(y0), ax1 = fabsf(x1), ay1 = fabsf(y1);
    const float mx0 = fmaxf(ax0, ay0), mn0 = fminf(ax0, ay0), mx1 = fmaxf(ax1, ay1), mn1 = fminf(ax1, ay1);
    const float2 q = make_float2(__fdividef(mn0, fmaxf(mx0, 1e-30f)), __fdividef(mn1, fmaxf(mx1, 1e-30f)));
    const float2 s = __fmul2_rn(q, q);
    float2 p = make_float2(-0.004054562299f, -0.004054562299f);
    p = __ffma2_rn(p, s, make_float2(0.021862939178f, 0.021862939178f));
    p = __ffma2_rn(p, s, make_float2(-0.05591229796f, -0.05591229796f));
    p = __ffma2_rn(p, s, make_float2(0.096421950378f, 0.096421950378f));
    p = __ffma2_rn(p, s, make_float2(-0.139086285623f, -0.139086285623f));
    p = __ffma2_rn(p, s, make_float2(0.1994656543f, 0.1994656543f));
    p = __ffma2_rn(p, s, make_float2(-0.333298607622f, -0.333298607622f));
    p = __ffma2_rn(p, s, make_float2(0.999999335572f, 0.999999335572f));
    float2 r = __fmul2_rn(p, q);
    const float2 ra = __fadd2_rn(make_float2(0.5f * PI_F, 0.5f * PI_F), make_float2(-r.x, -r.y));     // pi/2 - r, both halves at once
    r.x = (ay0 > ax0) ? ra.x : r.x;
    r.y = (ay1 > ax1) ? ra.y : r.y;
    const float2 rb = __fadd2_rn(make_float2(PI_F, PI_F), make_float2(-r.x, -r.y));                  // pi - r
    r.x = (x0 < 0.0f) ? rb.x : r.x;
    r.y = (x1 < 0.0f) ? rb.y : r.y;
    return make_float2(copysignf(r.x, y0), copysignf(r.y, y1));
}

// dsp/clamp.h:4-8
__device__ __forceinline__ float clampf(float x, float lo, float hi) {
    float y = (x > lo) ? x : lo;
    y = (y > hi) ? hi : y;
    return y;
}

// round-half-away like std::round for |x| < 2^22 (apply_harmonic_pll.cpp:19, pll_mixer.cpp:18)
__device__ __forceinline__ float round_half_away(float x) { return roundf(x); }

struct K1Params {
    float taps[K1_NN];          // reference memory order: b[NN-1] multiplies the newest sample
    float taps_s[K1_NN];        // taps * 2^100: the u8 kernel multiplies them with the bytes taken as fp32 denormals (u * 2^-149)
    float neg_dc;               // -127 * sum(taps): the reference's "- 127.0f" (app.cpp:59-60), applied once per output
    float first_neg_dc[K1_R];   // stream start: output r of the first block met real samples on its last 4(r+1) taps only
    float discrim_gain;         // A = 0.5*Fs/(2 pi Fd) (fm_demod.cpp:36-39)
    int n_out;                  // B/4 outputs per stream
    int parity;                 // history ping-pong: read [parity], write [parity^1]
    int n_streams;
    int first_block;            // no block before this one: the discriminator's previous angle is 0
    float2* dbg_fm_in;          // keep_intermediates: [S][n_out] FIR outputs before the discriminator (GUI spectrum), else null
};

struct K2Params {
    float taps_fm_out[K2_NN];
    float taps_hilbert[K2_HILB];
    float deemph_b[2], deemph_a[2];
    float peak_b[3], peak_a[3];
    // block-parallel linear-recurrence scan (k2_mpx.cu): powers of the recurrences' state matrices,
    // computed on the host in double.  A = [[a1, a0], [1, 0]] for the pilot filter, alpha = a[0] for
    // the de-emphasis pole.  P[l] = A^(8 * 2^l) (l = 0..5: thread strides 1..16, then one warp),
    // Q[j] = A^(8 j) (j = 0..31: offset of lane j inside its warp).
    float pk_P[6][4], pk_Q[32][4];
    float de_P[6], de_Q[32];
    int use_deemph;
    int n_out;                  // B/8 fm_out samples per stream
    int keep;                   // write pilot (unscaled) for the debug getters
};

struct K3Params {
    float lpf_b[2], lpf_a[2];   // filt_iir_lpf_pll_phase_error (:215-224)
    float int_KTs;              // integrator_gain * Ts (:234)
    float Kp;                   // proportional_gain (:429)
    float f_center, f_gain, mixer_KTs;   // (:229-231)
    float agc_target, agc_beta; // dsp/agc.h:9-11
    float integ_safe;           // filled by launch_k3: |integ| below this after a group => no clamp acted inside it
    int n;                      // B/8
    int n_streams;
    int keep;
    int exact;                  // 0: fast pass, helper-warp version when the grid fits rec_sms; 1: the exact body only; 2 / 3: force the one-warp / helper-warp fast pass
    int rec_sms;                // SMs of the partition the kernel runs on (0: unknown)
};

struct K4Params {
    float taps_lpr[K4_NN];
    float taps_lmr[K4_NN];
    float taps_rds[K4_NN];
    float harmonic_lmr, harmonic_rds;
    float stereo_mix;
    int audio_out_mode;         // 0 LPR, 1 LMR, 2 STEREO
    int n;                      // B/8 fm_out-rate samples per stream
    int n_tiles;
    int parity;
    int n_streams;
    int keep;
    int balanced;               // 1: FIR taps as constant-bank operands (production); 0: from shared memory (first version, A/B)
};

struct K5Params {
    float ted_b[2], ted_a[2];
    float pll_b[2], pll_a[2];
    float ted_Kp, pll_Kp;
    float int_ted_KTs, int_pll_KTs;
    float dump_KTs;
    float ted_KTs, ted_fcenter, ted_fgain;
    float mixer_KTs, mixer_fgain;
    float agc_target, agc_beta;
    int cooldown_N;
    int n;                      // B/64 samples per stream
    int n_tiles_k4;             // number of power partials per stream
    int n_streams;
    int keep;
    int literal;                // 1: the reference's per-sample loop instead of the symbol-wise loop (same bits; A/B aid)
};

// indices into the [field][stream] SoA state arrays of the per-stream recurrences
enum PllState { PLL_LPF_X1 = 0, PLL_LPF_Y1, PLL_INT, PLL_T, PLL_E_PREV, PLL_AGC_GAIN, PLL_TH_PREV, PLL_STATE_N };
enum BpskState {
    BP_LPF_PLL_X1 = 0, BP_LPF_PLL_Y1, BP_INT_PLL, BP_MIX_T, BP_PLL_PREV_ERR,
    BP_ZCD_XN, BP_COOLDOWN, BP_TED_YN, BP_TED_PHASE_ERR, BP_TED_PREV_ERR,
    BP_LPF_TED_X1, BP_LPF_TED_Y1, BP_INT_TED, BP_DUMP_RE, BP_DUMP_IM, BP_AGC_GAIN, BP_STATE_N
};
// per-stream scalars of K2 (AoS, one CTA per stream)
enum K2Scal { K2_DEEMPH_X1 = 0, K2_DEEMPH_Y1, K2_PK_X1R, K2_PK_X1I, K2_PK_X2R, K2_PK_X2I,
              K2_PK_Y1R, K2_PK_Y1I, K2_PK_Y2R, K2_PK_Y2I, K2_SCAL_N = 16 };

struct K5Debug {
    float2* rds; float2* raw_sym; float2* pll_sym; uint8_t* zcd; uint8_t* dump_trig;
    float* ted_raw; float* ted_pi; float* pll_raw; float* pll_pi; float2* dump_filter;
};

struct K1TParams {
    const int8_t* bimg;        // [2 chunks][96 rows][128 B]: G digit planes, shared-memory (swizzled) image
    const int* ptab;           // [32 lanes][6]: dp4a operands of the output before a tile (re, im per plane)
    int off[3];       // 127 * sum of the plane's digits
    float w[3];       // plane weights 65536 / S, 256 / S, 1 / S
    float discrim_gain;
    int n_rows;                // 128-byte rows per stream per block = block_size / 64
    int tiles_per_stream, n_tiles, n_streams;
    const float* theta_in;     // [S] the discriminator's prev_theta carried from the previous block (fm_demod.cpp:41-44)
    float* theta_out;          // [S] ... for the next block (ping-pong with theta_in)
    float2* dbg_fm_in;         // keep_intermediates: [S][n_out] FIR outputs before the discriminator, else null
    int shape;                 // 0: 4-deep ring, 2 accumulator buffers, 2 CTAs/SM; 1: 2-deep ring, 1 buffer, 3 CTAs/SM
};

// K1 on the tensor cores (k1_toeplitz_i8.cu)
void k1t_build_tables(const float* taps, std::vector<int8_t>& bimg, std::vector<int>& ptab, int off[3], float w[3]);
cudaError_t launch_k1t(const uint8_t* iq, const uint8_t* hist_in, uint8_t* hist_out, float2* hist_f32_out, float* fm_demod,
                       const K1TParams& p, int n_ctas, cudaStream_t st);

// launchers (one per .cu file)
cudaError_t launch_k1(bool u8, const void* iq, const float2* hist_in, float2* hist_out, float* fm_demod,
                      const K1Params& p, cudaStream_t st);
cudaError_t launch_k2(const float* fm_demod, float* hist_demod, float* hist_out, float* scal,
                      float2* fm_out_iq, float* theta, float* power, float2* pilot_dbg,
                      const K2Params& p, int n_streams, cudaStream_t st);
cudaError_t launch_k3(const float* theta, const float* power, float* state, float* pll_dt,
                      float* dbg_raw, float* dbg_pi, const K3Params& p, cudaStream_t st);
cudaError_t launch_k4(const float2* fm_out_iq, const float* pll_dt,
                      const float* hist_x_in, const float2* hist_m2_in, const float2* hist_m3_in,
                      float* hist_x_out, float2* hist_m2_out, float2* hist_m3_out,
                      float* lmr_phase, float2* audio_out, float2* rds_out, float* est_partial,
                      float* rds_power_partial, float* dbg_lpr, float* dbg_lmr, const K4Params& p, cudaStream_t st);
struct K7Entry { int j0; float k; };     // read position of one output frame of Resample(): in[j0]*(1-k) + in[j0+1]*k
void k7_build_table(int n_in, int n_out, K7Entry* out);
cudaError_t launch_k7(const float2* audio, const K7Entry* table, float2* pcm_f32, short2* pcm_s16,
                      int n_in, int n_out, int n_streams, cudaStream_t st);
cudaError_t launch_frames_to_s16(const float2* frames, short2* out, size_t n, cudaStream_t st);
cudaError_t launch_polyphase_us(const float* ext, const float* bp, float* y, int L, int K, int n_in,
                                int is_complex, cudaStream_t st);
cudaError_t launch_k5(const float2* rds_in, const float* rds_power_partial, float* state, float* pred_sym,
                      int* sym_count, const K5Debug& d, const K5Params& p, cudaStream_t st);
// K6: RDS bit path on the device (k6_rds.cu); state is an array of rds::State, glog of fmgpu_rds_group
cudaError_t launch_k6(const float* pred_sym, const int* sym_count, void* state, const void* tables, void* glog, uint8_t* blog,
                      int n64, int gcap, int bcap, int n_streams, cudaStream_t st);
cudaError_t launch_k6_init(void* state, int n_streams, cudaStream_t st);
size_t k6_state_bytes();
// debug-only finalisation of the GUI buffers: pilot *= gain, pll = (S(t+1/4), S(t))
cudaError_t launch_fft(const float* in, int in_is_real, float2* work0, float2* work1, int n, int fftshift,
                       float2** result, cudaStream_t st);
cudaError_t launch_kdbg_audio_iq(const float2* fm_out_iq, const float* pll_dt, const float2* hist_iq, const float2* hist_m2,
                                 const float* lmr_phase_used, float2* lpr_iq, float2* lmr_iq, const K4Params& p, cudaStream_t st);
cudaError_t launch_kdbg(float2* pilot, const float* pll_state, const float* pll_dt, float2* pll_out,
                        int n, int n_streams, cudaStream_t st);
// stand-alone polyphase decimator (dsp/polyphase_filter.h:41-64) for the dsp API surface
cudaError_t launch_iir_seq(const float* x, float* y, int n, int K, const float* b, const float* a, float* state, int is_complex, cudaStream_t st);
cudaError_t launch_hilbert_pack(const float* ext, const float* fir, float2* y, int n, int mid, cudaStream_t st);
cudaError_t launch_agc_power_seq(const float2* x, int n, float* sum, cudaStream_t st);
cudaError_t launch_agc_scale(const float2* x, float2* y, int n, float g, cudaStream_t st);
cudaError_t launch_polyphase_ds(const float* ext, const float* taps, float* y, int M, int NN, int n_out,
                                int is_complex, cudaStream_t st);

} // namespace fm
