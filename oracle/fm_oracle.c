/* TEST INFRASTRUCTURE ONLY -- see fm_oracle.h.  Plain scalar C, fp32 arithmetic in the reference's
 * order of operations, compiled with -ffp-contract=off so results do not depend on the compiler's
 * FMA choices.  All file:line citations are relative to /root/reference/src. */
#define _USE_MATH_DEFINES
#include "fm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define PI_F ((float)M_PI)

typedef struct { float re, im; } c32;

static inline c32 c32_mul(c32 a, c32 b) { c32 y = { a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re }; return y; }
static inline float clampf(float x, float lo, float hi) {      /* dsp/clamp.h:4-8 */
    float y = x;
    y = (y > lo) ? y : lo;
    y = (y > hi) ? hi : y;
    return y;
}

/* ------------------------------------------------------------------------------------------
 * dsp/simd/chebyshev_sine.h:13-41 : sin(2 pi x) on [-0.5, 0.5], degree-11 odd polynomial
 * ---------------------------------------------------------------------------------------- */
static float chebyshev_sine(float x) {
    const float A0 = -25.13274193f, A1 = 64.83583069f, A2 = -67.07687378f;
    const float A3 = 38.50016403f, A4 = -14.07150173f, A5 = 3.20396066f;
    const float z = x*x;
    const float b5 = A5;
    const float b4 = b5*z + A4;
    const float b3 = b4*z + A3;
    const float b2 = b3*z + A2;
    const float b1 = b2*z + A1;
    const float b0 = b1*z + A0;
    return b0 * (z-0.25f) * x;
}

/* ------------------------------------------------------------------------------------------
 * dsp/filter_designer.cpp
 * ---------------------------------------------------------------------------------------- */
static float window_hamming(float x) { return 0.53836f - 0.46164f*cosf(x); }   /* window_functions.h:11-14 */
static float sincf(float x) {                                                  /* filter_designer.cpp:19-25 */
    if (fabsf(x) <= 1e-6f) return 1.0f;
    return sinf(PI_F*x)/(PI_F*x);
}
static float prewarp(float Kd) { return 2.0f/PI_F * tanf(PI_F/2.0f * Kd); }     /* :42-66 */

/* All designers write through ReverseArray (:27-39): b[(N-1)-i] = h[i]. */
void fmo_create_fir_lpf(float* b, int N, float k) {                            /* :84-107 */
    const float M = (float)(N-1);
    for (int i = 0; i < N; i++) {
        const float t0 = 2.0f*PI_F*(float)(i)/M;
        const float t1 = (float)i - M/2.0f;
        b[(N-1)-i] = window_hamming(t0) * (k*sincf(k*t1));
    }
}
void fmo_create_fir_hpf(float* b, int N, float k) {                            /* :109-129 */
    const float M = (float)(N-1);
    for (int i = 0; i < N; i++) {
        const float t0 = 2*PI_F*(float)(i)/M;
        const float t1 = (float)(i) - M/2.0f;
        b[(N-1)-i] = window_hamming(t0) * (sincf(t1) - k*sincf(k*t1));
    }
}
void fmo_create_fir_bpf(float* b, int N, float k1, float k2) {                 /* :131-155 */
    const float M = (float)N-1;
    for (int i = 0; i < N; i++) {
        const float t0 = 2*PI_F*(float)(i)/M;
        const float t1 = (float)(i) - M/2.0f;
        b[(N-1)-i] = window_hamming(t0) * (k2*sincf(k2*t1) - k1*sincf(k1*t1));
    }
}
void fmo_create_iir_single_pole_lpf(float* b, float* a, float k) {             /* :158-200 */
    const float k_warp = prewarp(k);
    const float A = 1.0f/(PI_F*k_warp);
    const float B0 = 1.0f + 2.0f*A;
    const float B1 = 1.0f - 2.0f*A;
    const float b0 = 1.0f/B0;
    const float a0 = B1/B0;
    b[1] = b0; b[0] = b0;       /* _b[0], _b[1] through ReverseArray with N = 2 */
    a[1] = 1.0f; a[0] = -a0;
}
static float cabs2f(float re, float im) { return sqrtf(re*re + im*im); }
void fmo_create_iir_notch_filter(float* b, float* a, float k, float r) {       /* :202-258 */
    const float wn = PI_F*k;
    const float a0 = 2.0f*cosf(wn);
    const float r2 = r*r;
    const float k_z = (k > 0.5f) ? 0.0f : 1.0f;
    /* H(z) = (z-z0)(z-z1) / ((z-r z0)(z-r z1)) evaluated at z = exp(j pi k_z) */
    const c32 z  = { cosf(PI_F*k_z), sinf(PI_F*k_z) };
    const c32 z0 = { cosf(PI_F*k), sinf(PI_F*k) };
    const c32 z1 = { cosf(-PI_F*k), sinf(-PI_F*k) };
    const c32 n0 = { z.re-z0.re, z.im-z0.im }, n1 = { z.re-z1.re, z.im-z1.im };
    const c32 d0 = { z.re-r*z0.re, z.im-r*z0.im }, d1 = { z.re-r*z1.re, z.im-r*z1.im };
    const c32 num = c32_mul(n0, n1), den = c32_mul(d0, d1);
    const float K = 1.0f/(cabs2f(num.re, num.im)/cabs2f(den.re, den.im));
    b[2] = K*1.0f; b[1] = K*(-a0); b[0] = K*1.0f;
    a[2] = 1.0f; a[1] = a0*r; a[0] = -r2;
}
void fmo_create_iir_peak_1_filter(float* b, float* a, float k, float r) {      /* :260-310 */
    const float wn = PI_F*k;
    const float a0 = 2.0f*cosf(wn);
    const float r2 = r*r;
    /* K = 1/|H(k)|, H(z) = 1/((z-r z0)(z-r z1)) at z = z0 */
    const c32 z  = { cosf(PI_F*k), sinf(PI_F*k) };
    const c32 z0 = z;
    const c32 z1 = { cosf(-PI_F*k), sinf(-PI_F*k) };
    const c32 d0 = { z.re-r*z0.re, z.im-r*z0.im }, d1 = { z.re-r*z1.re, z.im-r*z1.im };
    const c32 den = c32_mul(d0, d1);
    const float K = 1.0f/(1.0f/cabs2f(den.re, den.im));
    b[2] = K*0.0f; b[1] = K*0.0f; b[0] = K*1.0f;   /* _b[0]=0,_b[1]=0,_b[2]=1 -> b[2],b[1],b[0] */
    a[2] = 1.0f; a[1] = r*a0; a[0] = -r2;
}
/* filter_designer.cpp:312-367 (method 2: zero and pole placement), normalised with the call's own parameters */
void fmo_create_iir_peak_2_filter(float* b, float* a, float k, float r, float A_db) {
    const float A = powf(10.0f, A_db/20.0f);
    const float rc = 1.0f-r;
    const float rc_scale = rc*2.0f;
    const float r0 = 1.0f - rc_scale;
    const float r1 = 1.0f - rc_scale/A;
    const float wn = PI_F*k;
    const float a0 = 2.0f*cosf(wn);
    const c32 z  = { cosf(PI_F*k), sinf(PI_F*k) };
    const c32 z0 = z;
    const c32 z1 = { cosf(-PI_F*k), sinf(-PI_F*k) };
    const c32 n0 = { z.re-r0*z0.re, z.im-r0*z0.im }, n1 = { z.re-r0*z1.re, z.im-r0*z1.im };
    const c32 d0 = { z.re-r1*z0.re, z.im-r1*z0.im }, d1 = { z.re-r1*z1.re, z.im-r1*z1.im };
    const c32 num = c32_mul(n0, n1), den = c32_mul(d0, d1);
    const float K = 1.0f/(cabs2f(num.re, num.im)/cabs2f(den.re, den.im));
    b[2] = K*1.0f; b[1] = K*(-r0*a0); b[0] = K*(r0*r0);
    a[2] = 1.0f; a[1] = r1*a0; a[0] = -r1*r1;
}
/* dsp/window_functions.h:11-38, window id: 0 hamming, 1 hann, 2 blackman, 3 blackman-harris */
static float window_by_id(int id, float x) {
    switch (id) {
    case 1: { const float s = sinf(x/2.0f); return s*s; }
    case 2: return 0.42659f - 0.49656f*cosf(x) + 0.076849f*cosf(2.0f*x);
    case 3: return 0.35875f - 0.48829f*cosf(x) + 0.14128f*cosf(2.0f*x) - 0.01168f*cosf(3.0f*x);
    default: return 0.53836f - 0.46164f*cosf(x);
    }
}
void fmo_create_fir_lpf_window(float* b, int N, float k, int window_id) {      /* :84-107 with the window argument */
    const float M = (float)(N-1);
    for (int i = 0; i < N; i++) {
        const float t0 = 2.0f*PI_F*(float)i/M;
        const float t1 = (float)i - M/2.0f;
        b[(N-1)-i] = window_by_id(window_id, t0) * (k*sincf(k*t1));
    }
}
void fmo_create_fir_hilbert(float* b, int N) {                                 /* :369-383 */
    const int M = (N-1)/2;
    for (int i = 0; i < N; i++) {
        const int n = i-M;
        const int is_even = (n % 2) == 0;
        b[(N-1)-i] = is_even ? 0.0f : 2.0f/(PI_F*(float)n);
    }
}

/* ------------------------------------------------------------------------------------------
 * dsp/polyphase_filter.h:9-87  PolyphaseDownsampler<T>:
 *   y[i] = sum_{k<NN} b[k] * X[(i+1)M - NN + k], X = (last NN inputs of previous calls) ++ x.
 * The reference's head/tail split (M0/M1, :41-64) is equivalent to filtering over the
 * concatenation, which is what is restated here.
 * ---------------------------------------------------------------------------------------- */
typedef struct { int M, K, NN; float* b; float* hist; int is_complex; float* ext; int ext_cap; } polyds;

static void polyds_init(polyds* f, int M, int K, int is_complex) {
    f->M = M; f->K = K; f->NN = M*K; f->is_complex = is_complex;
    f->b = (float*)calloc(f->NN, sizeof(float));
    f->hist = (float*)calloc((size_t)f->NN*(is_complex ? 2 : 1), sizeof(float));
    f->ext = NULL; f->ext_cap = 0;
}
static void polyds_free(polyds* f) { free(f->b); free(f->hist); free(f->ext); }
static void polyds_process(polyds* f, const float* x, float* y, int N) {
    const int C = f->is_complex ? 2 : 1;
    const int NN = f->NN, M = f->M;
    const int n_in = N*M;
    if (f->ext_cap < (NN+n_in)*C) { free(f->ext); f->ext_cap = (NN+n_in)*C; f->ext = (float*)malloc(sizeof(float)*f->ext_cap); }
    float* e = f->ext;
    memcpy(e, f->hist, sizeof(float)*NN*C);
    memcpy(e + NN*C, x, sizeof(float)*n_in*C);
    for (int i = 0; i < N; i++) {
        const float* w = e + (size_t)(i+1)*M*C;
        if (C == 1) {
            float acc = 0.0f;
            for (int k = 0; k < NN; k++) acc += w[k]*f->b[k];
            y[i] = acc;
        } else {
            float ar = 0.0f, ai = 0.0f;
            for (int k = 0; k < NN; k++) { ar += w[2*k]*f->b[k]; ai += w[2*k+1]*f->b[k]; }
            y[2*i] = ar; y[2*i+1] = ai;
        }
    }
    memcpy(f->hist, e + (size_t)n_in*C, sizeof(float)*NN*C);
}

void fmo_polyphase_ds_f32(int M, int K, const float* b, const float* x, float* y, int N_out, int n_calls) {
    polyds f; polyds_init(&f, M, K, 0); memcpy(f.b, b, sizeof(float)*M*K);
    for (int c = 0; c < n_calls; c++) polyds_process(&f, x + (size_t)c*N_out*M, y + (size_t)c*N_out, N_out);
    polyds_free(&f);
}
void fmo_polyphase_ds_cf32(int M, int K, const float* b, const float* x, float* y, int N_out, int n_calls) {
    polyds f; polyds_init(&f, M, K, 1); memcpy(f.b, b, sizeof(float)*M*K);
    for (int c = 0; c < n_calls; c++) polyds_process(&f, x + (size_t)c*N_out*M*2, y + (size_t)c*N_out*2, N_out);
    polyds_free(&f);
}
/* dsp/fir_filter.h:9-88  FIR_Filter<T>(K): y[i] = sum_k b[k] X[i - (K-1) + k] over X = (last K inputs) ++ x, i.e. the
 * decimator above with M = 1 (its head / tail split, :30-57, is again filtering over the concatenation). */
void fmo_fir_f32(int K, const float* b, const float* x, float* y, int N, int n_calls) { fmo_polyphase_ds_f32(1, K, b, x, y, N, n_calls); }
void fmo_fir_cf32(int K, const float* b, const float* x, float* y, int N, int n_calls) { fmo_polyphase_ds_cf32(1, K, b, x, y, N, n_calls); }
/* dsp/hilbert_fir_filter.h:13-47  Hilbert_FIR_Filter<float>(K): taps by create_fir_hilbert (:21-22);
 * y[i] = { X[i - (K-1) + (K-1)/2], FIR(X)[i] }  (real part = input delayed by (K-1)/2, :34-35, :43-44). */
void fmo_hilbert_f32(int K, const float* x, float* y, int N, int n_calls) {
    polyds f; polyds_init(&f, 1, K, 0); fmo_create_fir_hilbert(f.b, K);
    float* im = (float*)malloc(sizeof(float)*(size_t)N);
    float* hist = (float*)calloc((size_t)K, sizeof(float));
    const int M = (K-1)/2;
    for (int c = 0; c < n_calls; c++) {
        const float* xc = x + (size_t)c*N;
        polyds_process(&f, xc, im, N);
        for (int i = 0; i < N; i++) {
            const int j = i - (K-1) + M;                 /* index into this call's x; negative = history */
            y[2*((size_t)c*N + i)] = j >= 0 ? xc[j] : hist[K + j];
            y[2*((size_t)c*N + i) + 1] = im[i];
        }
        memcpy(hist, f.hist, sizeof(float)*(size_t)K);
    }
    free(im); free(hist); polyds_free(&f);
}
/* dsp/iir_filter.h:5-89  IIR_Filter<T>(K), direct form I, one sample at a time (:40-46):
 *   push_x; y = sum_{i<K} (xn[i] b[i] + yn[i] a[i]) in that order; push_y (yn[K-1] stays 0). */
static void iir_generic(int K, int C, const float* b, const float* a, const float* x, float* y, int N, int n_calls) {
    float* xn = (float*)calloc((size_t)K*C, sizeof(float));
    float* yn = (float*)calloc((size_t)K*C, sizeof(float));
    for (size_t t = 0; t < (size_t)N*n_calls; t++) {
        for (int i = 0; i < (K-1)*C; i++) xn[i] = xn[i+C];
        for (int c = 0; c < C; c++) xn[(K-1)*C + c] = x[t*C + c];
        for (int c = 0; c < C; c++) {
            float acc = 0.0f;
            for (int i = 0; i < K; i++) acc += (xn[i*C + c]*b[i] + yn[i*C + c]*a[i]);
            y[t*C + c] = acc;
        }
        for (int i = 0; i < (K-2)*C; i++) yn[i] = yn[i+C];
        for (int c = 0; c < C; c++) yn[(K-2)*C + c] = y[t*C + c];
    }
    free(xn); free(yn);
}
void fmo_iir_f32(int K, const float* b, const float* a, const float* x, float* y, int N, int n_calls) { iir_generic(K, 1, b, a, x, y, N, n_calls); }
void fmo_iir_cf32(int K, const float* b, const float* a, const float* x, float* y, int N, int n_calls) { iir_generic(K, 2, b, a, x, y, N, n_calls); }

/* audio/resampled_pcm_player.cpp:37-54  Resample(buf_in, buf_out) on stereo Frame<float> arrays:
 * the read position j walks in float (j += step), j0 = (int)j, k = j - j0, the last frame is held;
 * out = f0*(1-k) + f1*k with one rounding per Frame operator (audio/frame.h:10-16, 42-49). */
void fmo_resample_linear(const float* in, int n_in, float* out, int n_out) {
    const float step = (float)n_in / (float)n_out;
    float j = 0.0f;
    for (int i = 0; i < n_out; i++) {
        const int j0 = (int)j;
        const int j1 = j0 + 1;
        const float* f0 = in + 2*(size_t)j0;
        const float* f1 = (j1 < n_in) ? in + 2*(size_t)j1 : f0;
        const float k = j - (float)j0;
        const float a = 1.0f - k;
        for (int c = 0; c < 2; c++) {
            const float u = f0[c]*a, v = f1[c]*k;
            out[2*(size_t)i + c] = u + v;
        }
        j += step;
    }
}

/* fm_scraper.cpp:74-78  convert_buffer[i] = Frame<int16_t>(data[i]*CONVERT_RESCALE), CONVERT_RESCALE =
 * 32767 * 0.95f; the Frame conversion is a static_cast per channel (audio/frame.h:67-74).  Out of the int16
 * range the C++ cast is undefined; the reference's x86 build truncates (cvttss2si, 0x80000000 when out of
 * the int32 range or NaN) and keeps the low 16 bits -- restated here explicitly so it does not depend on
 * this file's compiler. */
void fmo_frames_to_s16(const float* frames, size_t n_frames, int16_t* out) {
    const float scale = 32767.0f * 0.95f;
    for (size_t i = 0; i < 2*n_frames; i++) {
        const float v = frames[i] * scale;
        int32_t q;
        if (v != v || v >= 2147483648.0f || v <= -2147483904.0f) q = (int32_t)0x80000000u;
        else q = (int32_t)v;                         /* truncation toward zero */
        out[i] = (int16_t)(uint16_t)((uint32_t)q & 0xFFFFu);
    }
}

/* Display spectra.  CalculateFFT (dsp/calculate_fft.cpp:43-50) is FFTW3f in the reference --
 * an external, un-vendored library (vcpkg fftw3 >= 3.3.10, toolchains/ubuntu/install_packages.sh:3); what it
 * computes is the forward DFT X[k] = sum_n x[n] exp(-2 pi i n k / N), restated here directly in double (O(N log N)
 * recursion-free radix-2 for powers of two, the O(N^2) sum otherwise).  PARITY UNPINNED against FFTW itself (absent
 * here); pinned against numpy.fft in tests/test_spectra.py.  fftshift != 0 applies InplaceFFTShift
 * (dsp/fftshift.h:21-33). */
void fmo_fft_f64(const float* x, double* y, int n, int fftshift) {
    const double PI = 3.14159265358979323846;
    double* re = (double*)malloc(sizeof(double) * (size_t)n);
    double* im = (double*)malloc(sizeof(double) * (size_t)n);
    if ((n & (n - 1)) == 0) {
        int bits = 0; while ((1 << bits) < n) bits++;
        for (int i = 0; i < n; i++) {                       /* bit-reversal permutation */
            int r = 0; for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
            re[r] = x[2*i]; im[r] = x[2*i + 1];
        }
        for (int len = 2; len <= n; len <<= 1) {
            for (int i = 0; i < n; i += len) {
                for (int k = 0; k < len/2; k++) {
                    const double ang = -2.0 * PI * (double)k / (double)len;
                    const double wr = cos(ang), wi = sin(ang);
                    const int a = i + k, b = i + k + len/2;
                    const double tr = re[b]*wr - im[b]*wi, ti = re[b]*wi + im[b]*wr;
                    re[b] = re[a] - tr; im[b] = im[a] - ti;
                    re[a] += tr; im[a] += ti;
                }
            }
        }
    } else {
        for (int k = 0; k < n; k++) {
            double sr = 0, si = 0;
            for (int i = 0; i < n; i++) {
                const double ang = -2.0 * PI * (double)(((long long)i * k) % n) / (double)n;
                sr += x[2*i]*cos(ang) - x[2*i+1]*sin(ang);
                si += x[2*i]*sin(ang) + x[2*i+1]*cos(ang);
            }
            re[k] = sr; im[k] = si;
        }
    }
    const int M = n / 2;
    for (int i = 0; i < n; i++) {
        const int j = fftshift ? ((i < M) ? i + M : i - M) : i;   /* x[i] <-> x[i + M], n even in every use */
        y[2*j] = re[i]; y[2*j + 1] = im[i];
    }
    free(re); free(im);
}

/* Calculate_FFT_Mag::Process (dsp/calculate_fft_mag.cpp:11-45), trigger already resolved by the caller:
 * v = 20 log10(|x[i]| / (2N - 1)); mode 0 NORMAL y = v, 1 AVERAGE y += beta (v - y), 2 MAX_HOLD y = max(y, v). */
void fmo_fft_mag_process(int mode, float beta, const float* x_cf32, float* y, int n) {
    const float M = (float)(2*n - 1);
    for (int i = 0; i < n; i++) {
        const float a = hypotf(x_cf32[2*i], x_cf32[2*i + 1]);
        const float v = 20.0f * log10f(a / M);
        if (mode == 0) y[i] = v;
        else if (mode == 1) { const float d = v - y[i]; y[i] += beta * d; }
        else y[i] = (y[i] > v) ? y[i] : v;
    }
}

/* dsp/polyphase_filter.h:90-185 PolyphaseUpsampler<float>: coefficient repack (:108-117) and
 * y[i*L+phase] = sum_{j<K} X[i-(K-1)+j] * bp[phase*K + j]. */
void fmo_polyphase_us_f32(int L, int K, const float* _b, const float* x, float* y, int N_in, int n_calls) {
    const int NN = L*K;
    float* b = (float*)calloc(NN, sizeof(float));
    float* hist = (float*)calloc(K, sizeof(float));
    for (int phase = 0; phase < L; phase++) {
        const int phase_c = (L-1)-phase;
        for (int i = 0; i < K; i++) {
            const int j0 = phase_c*K + i;
            const int j1 = phase + i*L;
            b[j0] = _b[(NN-1)-j1] * (float)L;
        }
    }
    float* e = (float*)malloc(sizeof(float)*(K+N_in));
    for (int c = 0; c < n_calls; c++) {
        memcpy(e, hist, sizeof(float)*K);
        memcpy(e+K, x + (size_t)c*N_in, sizeof(float)*N_in);
        for (int i = 0; i < N_in; i++) {
            const float* w = e + i + 1;    /* X[i-(K-1)] .. X[i] */
            for (int phase = 0; phase < L; phase++) {
                float acc = 0.0f;
                for (int j = 0; j < K; j++) acc += w[j]*b[phase*K + j];
                y[((size_t)c*N_in + i)*L + phase] = acc;
            }
        }
        memcpy(hist, e + N_in, sizeof(float)*K);
    }
    free(e); free(b); free(hist);
}

/* ------------------------------------------------------------------------------------------
 * dsp/iir_filter.h:40-69, one-pole (K=2) real filter used sample-by-sample.
 *   y = xn[0]*b[0] + yn[0]*a[0] + xn[1]*b[1] + yn[1]*a[1], yn[1] == 0 always.
 * ---------------------------------------------------------------------------------------- */
typedef struct { float b[2], a[2], xn[2], yn[2]; } iir1;
static float iir1_step(iir1* f, float x) {
    f->xn[0] = f->xn[1]; f->xn[1] = x;
    float y = 0.0f;
    for (int i = 0; i < 2; i++) y += (f->xn[i]*f->b[i] + f->yn[i]*f->a[i]);
    f->yn[0] = y;
    return y;
}

/* dsp/agc.h:6-31 (block-wise AGC) */
typedef struct { float target_power, current_gain, beta; } agc_t;
static void agc_process(agc_t* g, c32* x, int N) {
    float avg_power = 0.0f;
    for (int i = 0; i < N; i++) avg_power += (x[i].re*x[i].re + x[i].im*x[i].im);
    avg_power /= (float)N;
    const float target_gain = sqrtf(g->target_power/avg_power);
    g->current_gain = g->current_gain + g->beta*(target_gain - g->current_gain);
    for (int i = 0; i < N; i++) { x[i].re = g->current_gain*x[i].re; x[i].im = g->current_gain*x[i].im; }
}

void fmo_agc_cf32(float target_power, float beta, float gain0, const float* x, float* y, int N, int n_calls, float* gains_out) {
    agc_t g = { target_power, gain0, beta };
    memcpy(y, x, sizeof(float)*2*(size_t)N*n_calls);
    for (int c = 0; c < n_calls; c++) { agc_process(&g, (c32*)y + (size_t)c*N, N); if (gains_out) gains_out[c] = g.current_gain; }
}

/* fm_demod/pll_mixer.cpp:12-21 */
typedef struct { float KTs, yn, phase_error, phase_error_gain, f_center, f_gain; } pll_mixer;
static float pll_mixer_update(pll_mixer* m) {
    float control = m->phase_error * m->phase_error_gain;
    control = clampf(control, -1.0f, 1.0f);
    float freq = m->f_center + control*m->f_gain;
    float t = m->KTs*freq + m->yn;          /* dsp/integrator.h:9-13 */
    t = t - roundf(t);
    m->yn = t;
    return t;
}

/* dsp/simd/apply_harmonic_pll.cpp:11-24 (scalar specification) */
static void apply_harmonic_pll(const float* dt, const c32* x, c32* y, int N, float harmonic, float offset) {
    for (int i = 0; i < N; i++) {
        float dt_sin = dt[i]*harmonic + offset;
        float dt_cos = dt_sin+0.25f;
        dt_sin = dt_sin - roundf(dt_sin);
        dt_cos = dt_cos - roundf(dt_cos);
        const c32 pll = { chebyshev_sine(dt_cos), chebyshev_sine(dt_sin) };
        y[i] = c32_mul(x[i], pll);
    }
}

/* ------------------------------------------------------------------------------------------
 * fm_demod/bpsk_synchroniser.{h,cpp}, ted_clock.cpp, zero_crossing_detector.cpp,
 * trigger_cooldown.cpp
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int block_size;
    float zcd_xn;                                   /* zero_crossing_detector.h:6 */
    int cooldown_N, cooldown_remain;                /* trigger_cooldown.h:6-7 */
    float ted_KTs, ted_yn, ted_phase_error, ted_fcenter, ted_fgain;   /* ted_clock.h */
    float int_ted_KTs, int_ted_yn;
    iir1 lpf_ted;
    float ted_prev_phase_error;
    float dump_KTs; c32 dump_yn;
    pll_mixer mixer;
    float int_pll_KTs, int_pll_yn;
    iir1 lpf_pll;
    float pll_prev_phase_error;
    float ted_Kp, pll_Kp;
    /* display buffers (bpsk_synchroniser.cpp:175-182) */
    c32* pll_sym; uint8_t* zcd_trig; uint8_t* dump_trig;
    float *ted_raw, *ted_pi, *pll_raw, *pll_pi; c32* dump_filter;
} bpsk_t;

static void bpsk_init(bpsk_t* s, int block_size) {       /* bpsk_synchroniser.cpp:12-92 */
    memset(s, 0, sizeof(*s));
    s->block_size = block_size;
    const float Fs = 16e3f, Fsymbol = 2e3f;             /* bpsk_synchroniser.h:18-32 */
    const float ted_max_freq_offset = 1.5e3f, pll_max_freq_offset = 10.0f;
    fmo_create_iir_single_pole_lpf(s->lpf_ted.b, s->lpf_ted.a, ted_max_freq_offset/(Fs/2.0f));
    fmo_create_iir_single_pole_lpf(s->lpf_pll.b, s->lpf_pll.a, pll_max_freq_offset/(Fs/2.0f));
    const float Ts = 1.0f/Fs;
    const int samples_per_symbol = (int)roundf(Fs/Fsymbol);
    s->cooldown_N = samples_per_symbol/2;
    s->cooldown_remain = 0;
    const float A = 0.5f * (float)samples_per_symbol * 1.0f;
    s->dump_KTs = 1.0f/A;
    s->ted_KTs = Ts; s->ted_fcenter = Fsymbol; s->ted_fgain = ted_max_freq_offset;
    s->mixer.f_center = 0.0f; s->mixer.f_gain = pll_max_freq_offset; s->mixer.KTs = Ts;
    s->mixer.phase_error_gain = 1.0f;
    const float k = Fsymbol/Fs;
    s->int_ted_KTs = 10.0f*Ts*k;
    s->int_pll_KTs = 10.0f*Ts*k;
    s->ted_Kp = 0.3f; s->pll_Kp = 0.3f;
    s->pll_sym = (c32*)calloc(block_size, sizeof(c32));
    s->dump_filter = (c32*)calloc(block_size, sizeof(c32));
    s->zcd_trig = (uint8_t*)calloc(block_size, 1);
    s->dump_trig = (uint8_t*)calloc(block_size, 1);
    s->ted_raw = (float*)calloc(block_size, sizeof(float));
    s->ted_pi = (float*)calloc(block_size, sizeof(float));
    s->pll_raw = (float*)calloc(block_size, sizeof(float));
    s->pll_pi = (float*)calloc(block_size, sizeof(float));
}
static void bpsk_free(bpsk_t* s) {
    free(s->pll_sym); free(s->dump_filter); free(s->zcd_trig); free(s->dump_trig);
    free(s->ted_raw); free(s->ted_pi); free(s->pll_raw); free(s->pll_pi);
}

static int bpsk_process(bpsk_t* s, const c32* x, c32* y) {    /* bpsk_synchroniser.cpp:94-186 */
    int total = 0;
    for (int i = 0; i < s->block_size; i++) {
        const float pll_lpf = iir1_step(&s->lpf_pll, s->pll_prev_phase_error);
        s->int_pll_yn = s->int_pll_KTs*s->pll_prev_phase_error + s->int_pll_yn;
        s->int_pll_yn = clampf(s->int_pll_yn, -1.0f, 1.0f);
        const float PI_pll_error = pll_lpf*s->pll_Kp + s->int_pll_yn;
        s->mixer.phase_error = PI_pll_error;

        const float dt_sin = pll_mixer_update(&s->mixer);
        float dt_cos = dt_sin+0.25f;
        dt_cos = dt_cos - roundf(dt_cos);
        const c32 pll = { chebyshev_sine(dt_cos), chebyshev_sine(dt_sin) };
        const c32 IQ = c32_mul(x[i], pll);

        /* zero_crossing_detector.cpp:3-8 */
        int is_zcd = (IQ.im*s->zcd_xn) < 0.0f;
        s->zcd_xn = IQ.im;
        /* trigger_cooldown.cpp:4-13 */
        if (is_zcd && (s->cooldown_remain == 0)) {
            s->cooldown_remain = s->cooldown_N;
            is_zcd = 1;
        } else {
            if (s->cooldown_remain > 0) s->cooldown_remain--;
            is_zcd = 0;
        }
        if (is_zcd) {
            /* ted_clock.cpp:18-28 get_timing_error */
            float error = 2.0f * s->ted_yn;
            if (error > 1.0f) error = error - 2.0f;
            s->ted_prev_phase_error = error;
        }

        const float ted_lpf = iir1_step(&s->lpf_ted, s->ted_prev_phase_error);
        s->int_ted_yn = s->int_ted_KTs*s->ted_prev_phase_error + s->int_ted_yn;
        s->int_ted_yn = clampf(s->int_ted_yn, -1.0f, 1.0f);
        const float PI_ted_error = s->ted_Kp*ted_lpf + s->int_ted_yn;
        s->ted_phase_error = -PI_ted_error;

        /* integrate and dump: Integrator_Block<complex>, dsp/integrator.h:9-13 */
        s->dump_yn.re = s->dump_KTs*IQ.re + s->dump_yn.re;
        s->dump_yn.im = s->dump_KTs*IQ.im + s->dump_yn.im;

        /* ted_clock.cpp:31-44 update */
        int is_ted;
        {
            float control = s->ted_phase_error * 1.0f;
            control = clampf(control, -1.0f, 1.0f);
            const float freq = s->ted_fcenter + control*s->ted_fgain;
            const float v = s->ted_KTs*freq + s->ted_yn;
            s->ted_yn = v;
            const float offset = s->ted_KTs * freq / 2.0f;
            if (v < (1.0f-offset)) is_ted = 0;
            else { s->ted_yn = 0.0f; is_ted = 1; }
        }
        if (is_ted) {
            const c32 sym = s->dump_yn;
            s->dump_yn.re = 0.0f; s->dump_yn.im = 0.0f;
            const float sym_phase = atan2f(sym.im, sym.re);
            const float MAX_PHASE_ERROR = PI_F/2.0f;
            const float est = (sym_phase > 0.0f) ? (+PI_F/2.0f - sym_phase) : (-PI_F/2.0f - sym_phase);
            s->pll_prev_phase_error = est/MAX_PHASE_ERROR;
            y[total] = sym;
            total++;
        }
        s->pll_sym[i] = IQ;
        s->zcd_trig[i] = (uint8_t)is_zcd;
        s->dump_trig[i] = (uint8_t)is_ted;
        s->ted_raw[i] = s->ted_prev_phase_error;
        s->ted_pi[i] = PI_ted_error;
        s->pll_raw[i] = s->pll_prev_phase_error;
        s->pll_pi[i] = PI_pll_error;
        s->dump_filter[i] = s->dump_yn;
    }
    return total;
}

/* ------------------------------------------------------------------------------------------
 * RDS bit path: rds_decoder/differential_manchester_decoder.h:25-59, rds_group_sync.cpp,
 * crc10.cpp, rds_constants.h, and the PI/PTY/PS/RT subset of rds_decoder.cpp (0A: :167-245,
 * 2A: :301-340) with rds_database_decoder_handler.cpp.
 * ---------------------------------------------------------------------------------------- */
#define RDS_CRC10_POLY 0x1B9u   /* 0b0110111001, rds_constants.h:15 */
static const uint16_t RDS_OFFSETS[6] = { 0x0FC, 0x198, 0x168, 0x350, 0x1B4, 0x000 }; /* A B C C1 D E1, :20-27 */

static uint16_t crc10(uint32_t x) {                  /* crc10.cpp:9-25 */
    uint16_t reg = 0;
    for (int i = 0; i < 26; i++) {
        const uint16_t bit = (uint16_t)((x & (1u << 25)) >> 25);
        x = x << 1;
        reg = (uint16_t)((reg << 1) | bit);
        if (reg & (1u << 10)) reg = reg ^ RDS_CRC10_POLY;
    }
    return reg & 0x3FF;
}
static uint32_t crc_error_from_syndrome(uint16_t s) { /* crc10.cpp:28-60: single-bit patterns only */
    for (int i = 0; i < 26; i++) {
        const uint32_t e = 1u << i;
        if (crc10(e) == s) return e;
    }
    return 0;
}

typedef struct { uint16_t data[4]; uint8_t valid[4]; uint8_t type[4]; } grp_t;

typedef struct {
    /* differential manchester */
    uint8_t buf[16]; size_t byte_index, bit_index; int is_read_bit, prev_bit;
    /* group sync (rds_group_sync.h) */
    uint32_t rd_block_buf; int rd_block_buf_bits;
    grp_t group; int curr_data_block, total_data_block_errors;
    int max_group_desyncs_for_reset, curr_groups_desync, total_bits_desync;
    int state; /* 0 = FINDING_SYNC, 1 = READ_BLOCK */
    /* outputs */
    grp_t* groups; int n_groups, cap_groups;
    uint8_t* bytes; int n_bytes, cap_bytes;
    /* database subset */
    uint16_t pi; uint8_t pty; char ps[8]; char rt[64]; uint8_t ab_flag_rt;
    fmo_db_ext ext;                          /* rds_database.h:26-53 beyond PI/PTY/PS/RT */
} rds_t;

static void rds_init(rds_t* r) {
    memset(r, 0, sizeof(*r));
    r->max_group_desyncs_for_reset = 3;      /* rds_group_sync.cpp:22 */
    r->ab_flag_rt = 4;                       /* rds_database_decoder_handler.h:11 */
    r->ext.ptyn_ab_flag = 4;                 /* rds_database_decoder_handler.h:12 */
}
static void rds_free(rds_t* r) { free(r->groups); free(r->bytes); }

static void rds_decode_group(rds_t* r, const grp_t* g) {   /* rds_decoder.cpp:82-126 */
    const uint16_t descriptor = g->data[1];
    const uint8_t group_code = (descriptor >> 12) & 0xF;
    const uint8_t version = (descriptor >> 11) & 1;
    const uint8_t program_type = (descriptor >> 5) & 31;
    if (g->valid[0]) r->pi = g->data[0];
    if (!g->valid[1]) return;
    r->pty = program_type;
    if (version) return;
    const int has_C = g->valid[2] && g->type[2] == 2;
    const int has_D = g->valid[3] && g->type[3] == 4;
    if (group_code == 0) {                                 /* OnGroup0A :159-245 */
        const uint8_t seg = descriptor & 3;
        const uint8_t tp = (descriptor >> 10) & 1, ta = (descriptor >> 4) & 1, di = (descriptor >> 2) & 1;
        r->ext.is_music = (descriptor >> 3) & 1;           /* handler.cpp:74-76 */
        r->ext.traffic_announcement = (uint8_t)((tp << 1) | ta);   /* handler.cpp:53-72, enum order rds_database.h:19-24 */
        switch (seg) {                                     /* :211-228 */
        case 0: r->ext.is_dynamic_program_type = di; break;
        case 1: r->ext.is_compressed = di; break;
        case 2: r->ext.is_artificial_head = di; break;
        default: r->ext.is_stereo = di; break;
        }
        if (has_D) {
            char c0 = (char)(g->data[3] >> 8), c1 = (char)(g->data[3] & 0xFF);
            if (c0 == '\r') c0 = 0;
            if (c1 == '\r') c1 = 0;
            r->ps[2*seg] = c0; r->ps[2*seg+1] = c1;
        }
    } else if (group_code == 2) {                          /* OnGroup2A :301-340 */
        const uint8_t ab = (descriptor >> 4) & 1;
        const uint8_t seg = descriptor & 15;
        if (ab != r->ab_flag_rt) memset(r->rt, 0, sizeof(r->rt));
        r->ab_flag_rt = ab;
        char c[4] = { (char)(g->data[2] >> 8), (char)(g->data[2] & 0xFF), (char)(g->data[3] >> 8), (char)(g->data[3] & 0xFF) };
        for (int i = 0; i < 4; i++) if (c[i] == '\r') c[i] = 0;
        if (has_C) { r->rt[4*seg] = c[0]; r->rt[4*seg+1] = c[1]; }
        if (has_D) { r->rt[4*seg+2] = c[2]; r->rt[4*seg+3] = c[3]; }
    } else if (group_code == 4) {                          /* OnGroup4A :363-405 */
        const uint32_t mjd = ((uint32_t)(descriptor & 3) << 15) | ((g->data[2] & 0xFFFE) >> 1);
        const uint8_t hour = (uint8_t)(((g->data[2] & 1) << 4) | ((g->data[3] & 0xF000) >> 12));
        const uint8_t minute = (uint8_t)((g->data[3] & 0x0FC0) >> 6);
        const int lto_sign = (g->data[3] >> 5) & 1, lto_val = g->data[3] & 31;
        if (has_C) {                                       /* modified_julian_date.h:8-24 */
            long J = (long)mjd + 2400001 + 68569;
            const long C = 4 * J / 146097;
            J = J - (146097 * C + 3) / 4;
            const long Y = 4000 * (J + 1) / 1461001;
            J = J - 1461 * Y / 4 + 31;
            const long M = 80 * J / 2447;
            r->ext.day = (uint8_t)(J - 2447 * M / 80);
            J = M / 11;
            r->ext.month = (uint8_t)(M + 2 - (12 * J));
            r->ext.year = (int32_t)(100 * (C - 49) + Y + J);
        }
        if (has_C && has_D) { r->ext.hour = hour; r->ext.minute = minute; }
        if (has_D) r->ext.local_time_offset = (int8_t)(lto_sign ? -lto_val : lto_val);
    } else if (group_code == 10) {                         /* OnGroup10A :407-441 */
        const uint8_t ab = (descriptor >> 4) & 1;
        const uint8_t seg = descriptor & 1;
        if (ab != r->ext.ptyn_ab_flag) memset(r->ext.programme_type_name, 0, 8);   /* handler.cpp:30-35 */
        r->ext.ptyn_ab_flag = ab;
        char c[4] = { (char)(g->data[2] >> 8), (char)(g->data[2] & 0xFF), (char)(g->data[3] >> 8), (char)(g->data[3] & 0xFF) };
        for (int i = 0; i < 4; i++) if (c[i] == '\r') c[i] = 0;
        if (has_C) { r->ext.programme_type_name[4*seg] = c[0]; r->ext.programme_type_name[4*seg+1] = c[1]; }
        if (has_D) { r->ext.programme_type_name[4*seg+2] = c[2]; r->ext.programme_type_name[4*seg+3] = c[3]; }
    }
}

static int rds_attempt_decode(uint32_t x, int id, grp_t* g, int slot) {   /* rds_group_sync.cpp:143-206 */
    x = x ^ RDS_OFFSETS[id];
    uint32_t corrected = x; int is_valid = 0;
    const uint16_t syndrome = crc10(x);
    if (syndrome == 0) is_valid = 1;
    else {
        const uint32_t e = crc_error_from_syndrome(syndrome);
        if (e != 0) {
            const uint32_t xc = x ^ e;
            if (crc10(xc) == 0) { corrected = xc; is_valid = 1; }
        }
    }
    g->type[slot] = (uint8_t)id;
    g->data[slot] = (uint16_t)((corrected >> 10) & 0xFFFF);
    g->valid[slot] = (uint8_t)is_valid;
    return is_valid;
}
static void rds_push_block(rds_t* r, uint32_t x) {       /* rds_group_sync.cpp:209-237 */
    const int slot = r->curr_data_block;
    r->group.valid[slot] = 0;
    switch (slot) {
    case 0: rds_attempt_decode(x, 0, &r->group, slot); break;
    case 1: rds_attempt_decode(x, 1, &r->group, slot); break;
    case 2: if (!rds_attempt_decode(x, 2, &r->group, slot)) rds_attempt_decode(x, 3, &r->group, slot); break;
    case 3: rds_attempt_decode(x, 4, &r->group, slot); break;
    }
    r->curr_data_block++;
    if (!r->group.valid[slot]) r->total_data_block_errors++;
}
static void rds_emit_group(rds_t* r) {
    if (r->n_groups == r->cap_groups) {
        r->cap_groups = r->cap_groups ? 2*r->cap_groups : 256;
        r->groups = (grp_t*)realloc(r->groups, sizeof(grp_t)*r->cap_groups);
    }
    r->groups[r->n_groups++] = r->group;
    rds_decode_group(r, &r->group);
}
static void rds_group_sync_bit(rds_t* r, int bit) {      /* rds_group_sync.cpp:29-138, one bit */
    r->rd_block_buf = ((r->rd_block_buf << 1) | (uint32_t)(bit & 1)) & 0x3FFFFFFu;
    if (r->state == 0) {
        r->total_bits_desync++;
        const uint32_t input_block = r->rd_block_buf ^ RDS_OFFSETS[0];
        if (crc10(input_block) != 0) { r->total_bits_desync++; return; }
        r->state = 1;
        r->total_bits_desync = 0;
        r->rd_block_buf_bits = 0;
        rds_push_block(r, r->rd_block_buf);
        return;
    }
    r->rd_block_buf_bits++;
    if (r->rd_block_buf_bits != 26) return;
    r->rd_block_buf_bits = 0;
    rds_push_block(r, r->rd_block_buf);
    if (r->curr_data_block < 4) return;
    rds_emit_group(r);
    const int total_errors = r->total_data_block_errors;
    r->curr_data_block = 0;
    r->total_data_block_errors = 0;
    if (total_errors == 0) { r->curr_groups_desync = 0; return; }
    r->curr_groups_desync++;
    if (r->curr_groups_desync >= r->max_group_desyncs_for_reset) {
        r->state = 0;
        r->curr_groups_desync = 0;
    }
}
static void rds_push_bytes(rds_t* r, const uint8_t* x, int n) {
    if (r->n_bytes + n > r->cap_bytes) {
        r->cap_bytes = 2*(r->cap_bytes + n);
        r->bytes = (uint8_t*)realloc(r->bytes, r->cap_bytes);
    }
    memcpy(r->bytes + r->n_bytes, x, n);
    r->n_bytes += n;
    for (int i = 0; i < n*8; i++) rds_group_sync_bit(r, (x[i/8] >> (7-(i%8))) & 1);
}
static void rds_push_symbol(rds_t* r, float x) {         /* differential_manchester_decoder.h:32-59 */
    r->is_read_bit = !r->is_read_bit;
    if (!r->is_read_bit) return;
    const int curr_bit = (x > 0.0f);
    const int bit = curr_bit ^ r->prev_bit;
    r->prev_bit = curr_bit;
    if (r->bit_index == 0) r->buf[r->byte_index] = 0;
    r->buf[r->byte_index] |= (uint8_t)((bit & 1) << (7-r->bit_index));
    r->bit_index++;
    r->byte_index += (r->bit_index / 8);
    r->bit_index = (r->bit_index % 8);
    if (r->byte_index == 16) {
        r->byte_index = 0;
        rds_push_bytes(r, r->buf, 16);
    }
}

void* fmo_rds_create(void) { rds_t* r = (rds_t*)malloc(sizeof(rds_t)); rds_init(r); return r; }
void fmo_rds_destroy(void* r) { rds_free((rds_t*)r); free(r); }
void fmo_rds_push_symbols(void* rv, const float* sym, size_t n) { for (size_t i = 0; i < n; i++) rds_push_symbol((rds_t*)rv, sym[i]); }
int fmo_rds_n_groups(void* rv) { return ((rds_t*)rv)->n_groups; }
static void copy_groups(const rds_t* r, uint16_t* data, uint8_t* valid, uint8_t* type) {
    for (int g = 0; g < r->n_groups; g++) for (int i = 0; i < 4; i++) {
        data[4*g+i] = r->groups[g].data[i]; valid[4*g+i] = r->groups[g].valid[i]; type[4*g+i] = r->groups[g].type[i];
    }
}
void fmo_rds_get_groups(void* rv, uint16_t* data, uint8_t* valid, uint8_t* type) { copy_groups((rds_t*)rv, data, valid, type); }
int fmo_rds_n_bytes(void* rv) { return ((rds_t*)rv)->n_bytes; }
void fmo_rds_get_bytes(void* rv, uint8_t* out) { rds_t* r = (rds_t*)rv; memcpy(out, r->bytes, r->n_bytes); }
static void copy_db(const rds_t* r, uint16_t* pi, char* ps8, char* rt64, uint8_t* pty) {
    *pi = r->pi; *pty = r->pty; memcpy(ps8, r->ps, 8); memcpy(rt64, r->rt, 64);
}
void fmo_rds_get_db(void* rv, uint16_t* pi, char* ps8, char* rt64, uint8_t* pty) { copy_db((rds_t*)rv, pi, ps8, rt64, pty); }
void fmo_rds_get_db_ext(void* rv, fmo_db_ext* out) { *out = ((rds_t*)rv)->ext; }

/* ------------------------------------------------------------------------------------------
 * fm_demod/broadcast_fm_demod.{h,cpp}
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int B, n_fm_in, n_fm_out, n_rds, n_audio;
    /* filters */
    polyds ds_fm_in, ds_fm_out, ds_lpr, ds_lmr, ds_rds;
    float hilbert_b[65]; float hilbert_hist[65];
    float deemph_b[2], deemph_a[2], deemph_xn[2], deemph_yn[2];
    float peak_b[3], peak_a[3]; c32 peak_xn[3], peak_yn[3];
    iir1 lpf_pll;
    agc_t agc_pilot, agc_rds;
    pll_mixer mixer; float int_pe_KTs, int_pe_yn, prev_phase_error; float Kp;
    float prev_theta;
    float audio_lmr_phase_error;
    bpsk_t bpsk;
    /* controls (broadcast_fm_demod.h:64-89) */
    int audio_out; float stereo_mix; int use_deemph;
    int deemph_Tus, lpr_cutoff, lmr_cutoff; int dirty_deemph, dirty_lpr, dirty_lmr;
    /* buffers */
    c32 *iq, *fm_in, *fm_out_iq, *pilot, *pll, *temp_pll, *temp_audio, *rds, *rds_raw_sym;
    float *fm_demod, *fm_out, *pll_dt, *pll_lpf_pe, *pll_raw_pe, *audio_lpr, *audio_lmr, *rds_pred_sym, *audio_out_buf;
    float* hilbert_ext;
    int rds_total_symbols;
    rds_t rdsdec;
} demod_t;

static void update_filters(demod_t* d) {                 /* broadcast_fm_demod.cpp:330-389 */
    const float k_min = 0.0f + 0.01f, k_max = 1.0f - 0.01f;
    if (d->dirty_deemph) {
        d->dirty_deemph = 0;
        const float Tc = (float)d->deemph_Tus * 1e-6f;
        const float Fs = 128000.0f;
        const float Fc = 1.0f/(2.0f*PI_F*Tc);
        float k = Fc/(Fs/2.0f);
        k = clampf(k, k_min, k_max);
        fmo_create_iir_single_pole_lpf(d->deemph_b, d->deemph_a, k);
    }
    if (d->dirty_lpr) {
        d->dirty_lpr = 0;
        float k = (float)d->lpr_cutoff/(128000.0f/2.0f);
        k = clampf(k, k_min, k_max);
        fmo_create_fir_lpf(d->ds_lpr.b, d->ds_lpr.NN, k);
    }
    if (d->dirty_lmr) {
        d->dirty_lmr = 0;
        float k = (float)d->lmr_cutoff/(128000.0f/2.0f);
        k = clampf(k, k_min, k_max);
        fmo_create_fir_lpf(d->ds_lmr.b, d->ds_lmr.NN, k);
    }
}

void* fmo_create(int B) {                                /* broadcast_fm_demod.cpp:59-305 */
    demod_t* d = (demod_t*)calloc(1, sizeof(demod_t));
    d->B = B; d->n_fm_in = B/4; d->n_fm_out = d->n_fm_in/2; d->n_rds = d->n_fm_out/8; d->n_audio = d->n_fm_out/4;
    const float ROLLOFF = 0.95f;
    polyds_init(&d->ds_fm_in, 4, 16, 1);
    fmo_create_fir_lpf(d->ds_fm_in.b, 64, (256000.0f/2.0f)/(1024000.0f/2.0f) * ROLLOFF);
    polyds_init(&d->ds_fm_out, 2, 32, 0);
    fmo_create_fir_lpf(d->ds_fm_out.b, 64, (128000.0f/2.0f)/(256000.0f/2.0f) * ROLLOFF);
    fmo_create_fir_hilbert(d->hilbert_b, 65);
    fmo_create_iir_peak_1_filter(d->peak_b, d->peak_a, 19000.0f/(128000.0f/2.0f), 0.9999f);
    fmo_create_iir_single_pole_lpf(d->lpf_pll.b, d->lpf_pll.a, 100.0f/(128000.0f/2.0f));
    {
        const float Ts = 1.0f/128000.0f;
        d->mixer.f_center = -19000.0f; d->mixer.f_gain = -100.0f; d->mixer.KTs = Ts; d->mixer.phase_error_gain = 1.0f;
        d->prev_phase_error = 0.0f;
        d->int_pe_KTs = 0.1f*Ts;
        d->Kp = 0.01f;
    }
    polyds_init(&d->ds_lpr, 4, 32, 1);
    polyds_init(&d->ds_lmr, 4, 32, 1);
    polyds_init(&d->ds_rds, 8, 16, 1);
    fmo_create_fir_lpf(d->ds_rds.b, 128, 2000.0f/(128000.0f/2.0f));
    d->agc_pilot.target_power = 1.0f; d->agc_pilot.current_gain = 0.1f; d->agc_pilot.beta = 0.2f;
    d->agc_rds.target_power = 0.5f; d->agc_rds.current_gain = 0.1f; d->agc_rds.beta = 0.2f;
    bpsk_init(&d->bpsk, d->n_rds);
    d->audio_out = 2; d->stereo_mix = 1.0f; d->use_deemph = 0;
    d->deemph_Tus = 1; d->dirty_deemph = 1;          /* :189 */
    d->lpr_cutoff = 15000; d->dirty_lpr = 1;         /* :248 */
    d->lmr_cutoff = 15000; d->dirty_lmr = 1;         /* :260 */
    update_filters(d);
#define ALLOC(p, T, n) d->p = (T*)calloc((size_t)(n) > 0 ? (size_t)(n) : 1, sizeof(T))
    ALLOC(iq, c32, B); ALLOC(fm_in, c32, d->n_fm_in); ALLOC(fm_demod, float, d->n_fm_in);
    ALLOC(fm_out, float, d->n_fm_out); ALLOC(fm_out_iq, c32, d->n_fm_out);
    ALLOC(pilot, c32, d->n_fm_out); ALLOC(pll_dt, float, d->n_fm_out); ALLOC(pll, c32, d->n_fm_out);
    ALLOC(pll_lpf_pe, float, d->n_fm_out); ALLOC(pll_raw_pe, float, d->n_fm_out);
    ALLOC(temp_pll, c32, d->n_fm_out); ALLOC(temp_audio, c32, d->n_audio);
    ALLOC(audio_lpr, float, d->n_audio); ALLOC(audio_lmr, float, d->n_audio);
    ALLOC(rds, c32, d->n_rds); ALLOC(rds_raw_sym, c32, d->n_rds); ALLOC(rds_pred_sym, float, d->n_rds);
    ALLOC(audio_out_buf, float, 2*d->n_audio);
    ALLOC(hilbert_ext, float, 65 + d->n_fm_out);
#undef ALLOC
    rds_init(&d->rdsdec);
    return d;
}

void fmo_destroy(void* hv) {
    demod_t* d = (demod_t*)hv;
    polyds_free(&d->ds_fm_in); polyds_free(&d->ds_fm_out); polyds_free(&d->ds_lpr); polyds_free(&d->ds_lmr); polyds_free(&d->ds_rds);
    bpsk_free(&d->bpsk); rds_free(&d->rdsdec);
    free(d->iq); free(d->fm_in); free(d->fm_demod); free(d->fm_out); free(d->fm_out_iq); free(d->pilot); free(d->pll_dt);
    free(d->pll); free(d->pll_lpf_pe); free(d->pll_raw_pe); free(d->temp_pll); free(d->temp_audio); free(d->audio_lpr);
    free(d->audio_lmr); free(d->rds); free(d->rds_raw_sym); free(d->rds_pred_sym); free(d->audio_out_buf); free(d->hilbert_ext);
    free(d);
}

static float wrap_phase(float x) {                       /* fm_demod/fm_demod.cpp:6-10 */
    if (x >= PI_F) return x - 2.0f*PI_F;
    else if (x <= -PI_F) return x + 2.0f*PI_F;
    else return x;
}

static void run_fm_demodulate(demod_t* d) {              /* broadcast_fm_demod.cpp:391-416 */
    polyds_process(&d->ds_fm_in, (const float*)d->iq, (float*)d->fm_in, d->n_fm_in);
    {                                                    /* fm_demod/fm_demod.cpp:30-45 */
        const float Fd = 75e3f, Fs = 256000.0f;
        const float Wd = Fd * 2.0f * PI_F;
        const float Ts = 1.0f / Fs;
        const float A = 1.0f/(Wd*Ts) * 0.5f;
        for (int i = 0; i < d->n_fm_in; i++) {
            const float curr_theta = atan2f(d->fm_in[i].im, d->fm_in[i].re);
            const float delta_theta = wrap_phase(curr_theta - d->prev_theta);
            d->fm_demod[i] = delta_theta * A;
            d->prev_theta = curr_theta;
        }
    }
    polyds_process(&d->ds_fm_out, d->fm_demod, d->fm_out, d->n_fm_out);
    if (d->use_deemph) {                                 /* :404-406, dsp/iir_filter.h:40-69 with K=2 */
        for (int i = 0; i < d->n_fm_out; i++) {
            d->deemph_xn[0] = d->deemph_xn[1]; d->deemph_xn[1] = d->fm_out[i];
            float y = 0.0f;
            for (int k = 0; k < 2; k++) y += (d->deemph_xn[k]*d->deemph_b[k] + d->deemph_yn[k]*d->deemph_a[k]);
            d->deemph_yn[0] = y;
            d->fm_out[i] = y;
        }
    }
    {   /* dsp/hilbert_fir_filter.h:26-46 over FIR_Filter (fir_filter.h:30-56), K = 65:
           imag[i] = sum_k b[k]*X[i-64+k], real[i] = X[i-32] */
        const int K = 65, N = d->n_fm_out;
        float* e = d->hilbert_ext;
        memcpy(e, d->hilbert_hist, sizeof(float)*K);
        memcpy(e+K, d->fm_out, sizeof(float)*N);
        for (int i = 0; i < N; i++) {
            const float* w = e + i + 1;
            float acc = 0.0f;
            for (int k = 0; k < K; k++) acc += (w[k] * d->hilbert_b[k]);
            d->fm_out_iq[i].re = w[32];
            d->fm_out_iq[i].im = acc;
        }
        memcpy(d->hilbert_hist, e+N, sizeof(float)*K);
    }
}

static void lock_onto_pilot(demod_t* d) {                /* broadcast_fm_demod.cpp:418-461 */
    const int N = d->n_fm_out;
    for (int i = 0; i < N; i++) {                        /* dsp/iir_filter.h:40-69 with K=3, complex data */
        d->peak_xn[0] = d->peak_xn[1]; d->peak_xn[1] = d->peak_xn[2]; d->peak_xn[2] = d->fm_out_iq[i];
        c32 y = { 0.0f, 0.0f };
        for (int k = 0; k < 3; k++) {
            y.re += (d->peak_xn[k].re*d->peak_b[k] + d->peak_yn[k].re*d->peak_a[k]);
            y.im += (d->peak_xn[k].im*d->peak_b[k] + d->peak_yn[k].im*d->peak_a[k]);
        }
        d->pilot[i] = y;
        d->peak_yn[0] = d->peak_yn[1]; d->peak_yn[1] = y;
    }
    agc_process(&d->agc_pilot, d->pilot, N);
    for (int i = 0; i < N; i++) {                        /* :430-456 */
        const float phase_error_lpf = iir1_step(&d->lpf_pll, d->prev_phase_error);
        d->int_pe_yn = d->int_pe_KTs*d->prev_phase_error + d->int_pe_yn;
        d->int_pe_yn = clampf(d->int_pe_yn, -1.0f, 1.0f);
        const float PI_error = phase_error_lpf*d->Kp + d->int_pe_yn;
        d->mixer.phase_error = PI_error;
        const float dt_sin = pll_mixer_update(&d->mixer);
        float dt_cos = dt_sin+0.25f;
        dt_cos = dt_cos - roundf(dt_cos);
        const c32 pll = { chebyshev_sine(dt_cos), chebyshev_sine(dt_sin) };
        const c32 residual = c32_mul(d->pilot[i], pll);
        d->prev_phase_error = atan2f(residual.im, residual.re);
        d->pll_dt[i] = dt_sin;
        d->pll[i] = pll;
        d->pll_raw_pe[i] = d->prev_phase_error;
        d->pll_lpf_pe[i] = PI_error;
    }
}

static void extract_components(demod_t* d) {             /* broadcast_fm_demod.cpp:463-536 */
    const float harmonic_audio_lmr = 38000.0f/19000.0f;
    const float harmonic_rds = 57000.0f/19000.0f;
    const int N_pilot = d->n_fm_out, N_audio = d->n_audio;
    polyds_process(&d->ds_lpr, (const float*)d->fm_out_iq, (float*)d->temp_audio, N_audio);
    for (int i = 0; i < N_audio; i++) d->audio_lpr[i] = d->temp_audio[i].re;

    apply_harmonic_pll(d->pll_dt, d->fm_out_iq, d->temp_pll, N_pilot, harmonic_audio_lmr, d->audio_lmr_phase_error);
    polyds_process(&d->ds_lmr, (const float*)d->temp_pll, (float*)d->temp_audio, N_audio);
    {                                                    /* :496-517 */
        const int stride = 10;
        float avg_phase_error = 0.0f;
        int total_samples = 0;
        for (int i = 0; i < N_audio; i += stride) {
            const float phase = atan2f(d->temp_audio[i].im, d->temp_audio[i].re);
            const float est = (phase > 0.0f) ? (+PI_F/2.0f - phase) : (-PI_F/2.0f - phase);
            avg_phase_error += est;
            total_samples++;
        }
        avg_phase_error /= (float)total_samples;
        d->audio_lmr_phase_error += 0.1f*avg_phase_error;
        d->audio_lmr_phase_error = fmodf(d->audio_lmr_phase_error, 2.0f*PI_F);
    }
    for (int i = 0; i < N_audio; i++) d->audio_lmr[i] = d->temp_audio[i].im;

    apply_harmonic_pll(d->pll_dt, d->fm_out_iq, d->temp_pll, N_pilot, harmonic_rds, 0.0f);
    polyds_process(&d->ds_rds, (const float*)d->temp_pll, (float*)d->rds, d->n_rds);
}

static void synchronise_rds(demod_t* d) {                /* broadcast_fm_demod.cpp:538-547 */
    agc_process(&d->agc_rds, d->rds, d->n_rds);
    d->rds_total_symbols = bpsk_process(&d->bpsk, d->rds, d->rds_raw_sym);
    for (int i = 0; i < d->rds_total_symbols; i++) d->rds_pred_sym[i] = d->rds_raw_sym[i].im;
}

static void mix_audio(demod_t* d) {                      /* broadcast_fm_demod.cpp:549-585 */
    const int N = d->n_audio;
    for (int i = 0; i < N; i++) {
        const float lpr = d->audio_lpr[i], lmr = d->audio_lmr[i];
        float l, r;
        if (d->audio_out == 2) { const float k = d->stereo_mix; l = lpr+k*lmr; r = lpr-k*lmr; }
        else if (d->audio_out == 1) { l = lmr; r = lmr; }
        else { l = lpr; r = lpr; }
        d->audio_out_buf[2*i] = l*2.0f;
        d->audio_out_buf[2*i+1] = r*2.0f;
    }
}

int fmo_process_cf32(void* hv, const float* iq) {        /* broadcast_fm_demod.cpp:309-328 */
    demod_t* d = (demod_t*)hv;
    if ((const float*)d->iq != iq) memcpy(d->iq, iq, sizeof(c32)*d->B);
    update_filters(d);
    run_fm_demodulate(d);
    lock_onto_pilot(d);
    extract_components(d);
    synchronise_rds(d);
    mix_audio(d);
    /* app.cpp:27-34: symbols -> differential manchester -> 16-byte packets -> group sync */
    for (int i = 0; i < d->rds_total_symbols; i++) rds_push_symbol(&d->rdsdec, d->rds_pred_sym[i]);
    return d->rds_total_symbols;
}

int fmo_process_u8(void* hv, const uint8_t* iq) {        /* app.cpp:56-65 */
    demod_t* d = (demod_t*)hv;
    for (int i = 0; i < d->B; i++) {
        d->iq[i].re = (float)iq[2*i+0] - 127.0f;
        d->iq[i].im = (float)iq[2*i+1] - 127.0f;
    }
    return fmo_process_cf32(hv, (const float*)d->iq);
}

int fmo_set_control(void* hv, const char* s, double value) {
    demod_t* d = (demod_t*)hv;
    if (!strcmp(s, "audio_out")) d->audio_out = (int)value;
    else if (!strcmp(s, "audio_stereo_mix_factor")) d->stereo_mix = (float)value;
    else if (!strcmp(s, "is_use_deemphasis_filter")) d->use_deemph = value != 0.0;
    else if (!strcmp(s, "filt_deemphasis_cutoff")) { d->deemph_Tus = (int)value; d->dirty_deemph = 1; }
    else if (!strcmp(s, "filt_audio_lpr_cutoff")) { d->lpr_cutoff = (int)value; d->dirty_lpr = 1; }
    else if (!strcmp(s, "filt_audio_lmr_cutoff")) { d->lmr_cutoff = (int)value; d->dirty_lmr = 1; }
    else return -1;
    return 0;
}

#define GIVE(ptr, n_) do { *p = (const void*)(ptr); *n = (size_t)(n_); return 0; } while (0)
int fmo_get(void* hv, const char* s, const void** p, size_t* n) {
    demod_t* d = (demod_t*)hv;
    if (!strcmp(s, "fm_in")) GIVE(d->fm_in, d->n_fm_in);
    if (!strcmp(s, "fm_demod")) GIVE(d->fm_demod, d->n_fm_in);
    if (!strcmp(s, "fm_out")) GIVE(d->fm_out, d->n_fm_out);
    if (!strcmp(s, "fm_out_iq")) GIVE(d->fm_out_iq, d->n_fm_out);
    if (!strcmp(s, "pilot")) GIVE(d->pilot, d->n_fm_out);
    if (!strcmp(s, "pll_dt")) GIVE(d->pll_dt, d->n_fm_out);
    if (!strcmp(s, "pll")) GIVE(d->pll, d->n_fm_out);
    if (!strcmp(s, "pll_raw_phase_error")) GIVE(d->pll_raw_pe, d->n_fm_out);
    if (!strcmp(s, "pll_lpf_phase_error")) GIVE(d->pll_lpf_pe, d->n_fm_out);
    if (!strcmp(s, "audio_lpr")) GIVE(d->audio_lpr, d->n_audio);
    if (!strcmp(s, "audio_lmr")) GIVE(d->audio_lmr, d->n_audio);
    if (!strcmp(s, "rds")) GIVE(d->rds, d->n_rds);
    if (!strcmp(s, "rds_raw_sym")) GIVE(d->rds_raw_sym, d->rds_total_symbols);
    if (!strcmp(s, "rds_pred_sym")) GIVE(d->rds_pred_sym, d->rds_total_symbols);
    if (!strcmp(s, "audio_out")) GIVE(d->audio_out_buf, d->n_audio);
    if (!strcmp(s, "bpsk_pll_sym")) GIVE(d->bpsk.pll_sym, d->n_rds);
    if (!strcmp(s, "bpsk_ted_raw_phase_error")) GIVE(d->bpsk.ted_raw, d->n_rds);
    if (!strcmp(s, "bpsk_ted_pi_phase_error")) GIVE(d->bpsk.ted_pi, d->n_rds);
    if (!strcmp(s, "bpsk_pll_raw_phase_error")) GIVE(d->bpsk.pll_raw, d->n_rds);
    if (!strcmp(s, "bpsk_pll_pi_phase_error")) GIVE(d->bpsk.pll_pi, d->n_rds);
    if (!strcmp(s, "bpsk_int_dump_filter")) GIVE(d->bpsk.dump_filter, d->n_rds);
    if (!strcmp(s, "bpsk_zcd")) GIVE(d->bpsk.zcd_trig, d->n_rds);
    if (!strcmp(s, "bpsk_int_dump_trigger")) GIVE(d->bpsk.dump_trig, d->n_rds);
    return -1;
}

float fmo_get_scalar(void* hv, const char* s) {
    demod_t* d = (demod_t*)hv;
    if (!strcmp(s, "audio_lmr_phase_error")) return d->audio_lmr_phase_error;
    if (!strcmp(s, "agc_pilot_gain")) return d->agc_pilot.current_gain;
    if (!strcmp(s, "agc_rds_gain")) return d->agc_rds.current_gain;
    if (!strcmp(s, "rds_total_symbols")) return (float)d->rds_total_symbols;
    return NAN;
}

static int taps_rw(demod_t* d, const char* s, float* b, float* a, int write) {
#define RW(dst, src, n) do { if (src) { if (write) memcpy(dst, src, sizeof(float)*(n)); else memcpy(src, dst, sizeof(float)*(n)); } } while (0)
    if (!strcmp(s, "fm_in")) { RW(d->ds_fm_in.b, b, 64); return 64; }
    if (!strcmp(s, "fm_out")) { RW(d->ds_fm_out.b, b, 64); return 64; }
    if (!strcmp(s, "hilbert")) { RW(d->hilbert_b, b, 65); return 65; }
    if (!strcmp(s, "audio_lpr")) { RW(d->ds_lpr.b, b, 128); return 128; }
    if (!strcmp(s, "audio_lmr")) { RW(d->ds_lmr.b, b, 128); return 128; }
    if (!strcmp(s, "rds")) { RW(d->ds_rds.b, b, 128); return 128; }
    if (!strcmp(s, "deemphasis")) { RW(d->deemph_b, b, 2); RW(d->deemph_a, a, 2); return 2; }
    if (!strcmp(s, "peak_pilot")) { RW(d->peak_b, b, 3); RW(d->peak_a, a, 3); return 3; }
    if (!strcmp(s, "pll_lpf")) { RW(d->lpf_pll.b, b, 2); RW(d->lpf_pll.a, a, 2); return 2; }
    if (!strcmp(s, "bpsk_ted_lpf")) { RW(d->bpsk.lpf_ted.b, b, 2); RW(d->bpsk.lpf_ted.a, a, 2); return 2; }
    if (!strcmp(s, "bpsk_pll_lpf")) { RW(d->bpsk.lpf_pll.b, b, 2); RW(d->bpsk.lpf_pll.a, a, 2); return 2; }
#undef RW
    return -1;
}
int fmo_get_taps(void* hv, const char* s, float* b, float* a) { return taps_rw((demod_t*)hv, s, b, a, 0); }
int fmo_set_taps(void* hv, const char* s, const float* b, const float* a) { return taps_rw((demod_t*)hv, s, (float*)b, (float*)a, 1); }

int fmo_n_groups(void* hv) { return ((demod_t*)hv)->rdsdec.n_groups; }
void fmo_get_groups(void* hv, uint16_t* data, uint8_t* valid, uint8_t* type) { copy_groups(&((demod_t*)hv)->rdsdec, data, valid, type); }
int fmo_n_rds_bytes(void* hv) { return ((demod_t*)hv)->rdsdec.n_bytes; }
void fmo_get_rds_bytes(void* hv, uint8_t* out) { demod_t* d = (demod_t*)hv; memcpy(out, d->rdsdec.bytes, d->rdsdec.n_bytes); }
void fmo_get_db(void* hv, uint16_t* pi, char* ps8, char* rt64, uint8_t* pty) { copy_db(&((demod_t*)hv)->rdsdec, pi, ps8, rt64, pty); }
void fmo_get_db_ext(void* hv, fmo_db_ext* out) { *out = ((demod_t*)hv)->rdsdec.ext; }

/* ------------------------------------------------------------------------------------------
 * Wideband channelizer oracle (BASELINE config 4).  The reference has NO channelizer (SURVEY.md
 * section 7, "No channelizer in the reference"), so this is the definition the CUDA channelizer is
 * checked against: per channel, frequency shift in float64 followed by a decimating direct-form FIR
 * in float64, in the reference's own conventions --
 *   unpack       x[n] = (u8 - 127)                                      app.cpp:56-65
 *   shift        xs[n] = x[n] * exp(-j 2 pi ph_c(n) / 2^32),  ph_c(n) = (inc_c * n) mod 2^32
 *                (inc_c = round(f_c / Fs * 2^32): the centre frequency quantised to Fs / 2^32)
 *   decimate     y[i] = sum_{k<NN} b[k] * xs[(i+1)*D - NN + k]          dsp/polyphase_filter.h:41-64
 *                (newest sample <-> b[NN-1]; samples before the first are zero = empty history)
 * iq: n_in interleaved (I,Q) u8 pairs; n0: absolute index of iq[0]; hist: the NN samples before
 * iq[0] (u8 pairs, or NULL = zeros... i.e. u8 value 127); out: n_ch x n_out complex double.
 * ---------------------------------------------------------------------------------------- */
void fmo_channelize_f64(const uint8_t* iq, size_t n_in, const uint8_t* hist, uint64_t n0, int D, int NN,
                        const float* b, const uint32_t* inc, int n_ch, double* out)
{
    const size_t n_out = n_in / (size_t)D;
    const double TWO_PI = 6.283185307179586476925286766559;
    double* xr = (double*)malloc(sizeof(double) * (n_in + (size_t)NN));
    double* xi = (double*)malloc(sizeof(double) * (n_in + (size_t)NN));
    for (int c = 0; c < n_ch; c++) {
        for (size_t j = 0; j < n_in + (size_t)NN; j++) {
            /* j = 0 is absolute sample n0 - NN */
            double re, im;
            if (j < (size_t)NN) {
                if (hist) { re = (double)hist[2*j] - 127.0; im = (double)hist[2*j+1] - 127.0; }
                else { re = 0.0; im = 0.0; }
            } else { re = (double)iq[2*(j - NN)] - 127.0; im = (double)iq[2*(j - NN) + 1] - 127.0; }
            const uint64_t n = n0 + (uint64_t)j - (uint64_t)NN;                 /* wraps like the phase does */
            const uint32_t ph = (uint32_t)((uint64_t)inc[c] * n);
            const double th = -TWO_PI * ((double)ph / 4294967296.0);
            const double cs = cos(th), sn = sin(th);
            xr[j] = re * cs - im * sn;
            xi[j] = re * sn + im * cs;
        }
        for (size_t i = 0; i < n_out; i++) {
            const double* pr = xr + (i + 1) * (size_t)D;       /* element k of the window = index (i+1)*D - NN + k, +NN for the prefix */
            const double* pi = xi + (i + 1) * (size_t)D;
            double ar = 0.0, ai = 0.0;
            for (int k = 0; k < NN; k++) { ar += (double)b[k] * pr[k]; ai += (double)b[k] * pi[k]; }
            out[2 * ((size_t)c * n_out + i)] = ar;
            out[2 * ((size_t)c * n_out + i) + 1] = ai;
        }
    }
    free(xr); free(xi);
}
