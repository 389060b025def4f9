// Host-side filter designers with the reference's API surface (src/dsp/filter_designer.h:8-35).
// Same formulas, same fp32 arithmetic, same memory order: every designer writes through the
// reference's ReverseArray view (filter_designer.cpp:27-39), i.e. b[N-1-i] = h[i], so that b[N-1]
// multiplies the newest sample in the FIR kernels.  Exported through the C-ABI as fmgpu_create_*.
#include <cmath>
#include <complex>
#include "../../include/fmgpu.h"

namespace {
constexpr float PI = 3.14159265358979323846f;

float sinc(float x) {                                                          // filter_designer.cpp:19-25
    if (std::abs(x) <= 1e-6f) return 1.0f;
    return std::sin(PI * x) / (PI * x);
}
float prewarp(float Kd) { return 2.0f / PI * std::tan(PI / 2.0f * Kd); }       // :42-66
std::complex<float> phasor(float x) { return { std::cos(x), std::sin(x) }; }
}

extern "C" {

// dsp/window_functions.h:11-38 (x = 2 pi i / (N - 1))
float fmgpu_window_hamming(float x) { return 0.53836f - 0.46164f * std::cos(x); }
float fmgpu_window_hann(float x) { const float a = std::sin(x / 2.0f); return a * a; }
float fmgpu_window_blackman(float x) { return 0.42659f - 0.49656f * std::cos(x) + 0.076849f * std::cos(2.0f * x); }
float fmgpu_window_blackman_harris(float x) {
    return 0.35875f - 0.48829f * std::cos(x) + 0.14128f * std::cos(2.0f * x) - 0.01168f * std::cos(3.0f * x);
}

// the reference's `const window_func_t window = window_hamming` argument (filter_designer.h:9-11); NULL = Hamming
void fmgpu_create_fir_lpf_window(float* b, int N, float k, fmgpu_window_func_t window) {        // :84-107
    if (!window) window = fmgpu_window_hamming;
    const float M = (float)(N - 1);
    for (int i = 0; i < N; i++) {
        const float t0 = 2.0f * PI * (float)i / M;
        const float t1 = (float)i - M / 2.0f;
        b[(N - 1) - i] = window(t0) * (k * sinc(k * t1));
    }
}

void fmgpu_create_fir_hpf_window(float* b, int N, float k, fmgpu_window_func_t window) {        // :109-129
    if (!window) window = fmgpu_window_hamming;
    const float M = (float)(N - 1);
    for (int i = 0; i < N; i++) {
        const float t0 = 2 * PI * (float)i / M;
        const float t1 = (float)i - M / 2.0f;
        b[(N - 1) - i] = window(t0) * (sinc(t1) - k * sinc(k * t1));
    }
}

void fmgpu_create_fir_bpf_window(float* b, int N, float k1, float k2, fmgpu_window_func_t window) {   // :131-155
    if (!window) window = fmgpu_window_hamming;
    const float M = (float)N - 1;
    for (int i = 0; i < N; i++) {
        const float t0 = 2 * PI * (float)i / M;
        const float t1 = (float)i - M / 2.0f;
        b[(N - 1) - i] = window(t0) * (k2 * sinc(k2 * t1) - k1 * sinc(k1 * t1));
    }
}

void fmgpu_create_fir_lpf(float* b, int N, float k) { fmgpu_create_fir_lpf_window(b, N, k, nullptr); }
void fmgpu_create_fir_hpf(float* b, int N, float k) { fmgpu_create_fir_hpf_window(b, N, k, nullptr); }
void fmgpu_create_fir_bpf(float* b, int N, float k1, float k2) { fmgpu_create_fir_bpf_window(b, N, k1, k2, nullptr); }

void fmgpu_create_fir_hilbert(float* b, int N) {                               // :369-383
    const int M = (N - 1) / 2;
    for (int i = 0; i < N; i++) {
        const int n = i - M;
        b[(N - 1) - i] = (n % 2 == 0) ? 0.0f : 2.0f / (PI * (float)n);
    }
}

void fmgpu_create_iir_single_pole_lpf(float* b, float* a, float k) {           // :158-200
    const float A = 1.0f / (PI * prewarp(k));
    const float B0 = 1.0f + 2.0f * A, B1 = 1.0f - 2.0f * A;
    const float b0 = 1.0f / B0, a0 = B1 / B0;
    b[1] = b0; b[0] = b0;
    a[1] = 1.0f; a[0] = -a0;
}

void fmgpu_create_iir_notch_filter(float* b, float* a, float k, float r) {     // :202-258
    const float a0 = 2.0f * std::cos(PI * k);
    const float k_z = (k > 0.5f) ? 0.0f : 1.0f;
    const auto z = phasor(PI * k_z), z0 = phasor(PI * k), z1 = phasor(-PI * k);
    const auto H = ((z - z0) * (z - z1)) / ((z - r * z0) * (z - r * z1));
    const float K = 1.0f / std::abs(H);
    b[2] = K; b[1] = K * (-a0); b[0] = K;
    a[2] = 1.0f; a[1] = a0 * r; a[0] = -(r * r);
}

void fmgpu_create_iir_peak_1_filter(float* b, float* a, float k, float r) {    // :260-310
    const float a0 = 2.0f * std::cos(PI * k);
    const auto z = phasor(PI * k), z0 = phasor(PI * k), z1 = phasor(-PI * k);
    const auto H = 1.0f / ((z - r * z0) * (z - r * z1));
    const float K = 1.0f / std::abs(H);
    b[2] = K * 0.0f; b[1] = K * 0.0f; b[0] = K * 1.0f;
    a[2] = 1.0f; a[1] = r * a0; a[0] = -(r * r);
}

// Method 2: zero and pole placement; A_db = attenuation outside the peak (filter_designer.cpp:312-367).  The
// reference normalises with a `static` lambda that captures (k, r0, r1) of the FIRST call in the process (:347); this
// uses the arguments of every call.
void fmgpu_create_iir_peak_2_filter(float* b, float* a, float k, float r, float A_db) {
    const float A = std::pow(10.0f, A_db / 20.0f);
    const float rc_scale = (1.0f - r) * 2.0f;
    const float r0 = 1.0f - rc_scale, r1 = 1.0f - rc_scale / A;
    const float a0 = 2.0f * std::cos(PI * k);
    const auto z = phasor(PI * k), z0 = phasor(PI * k), z1 = phasor(-PI * k);
    const auto H = (z - r0 * z0) * (z - r0 * z1) / ((z - r1 * z0) * (z - r1 * z1));
    const float K = 1.0f / std::abs(H);
    b[2] = K * 1.0f; b[1] = K * (-r0 * a0); b[0] = K * (r0 * r0);
    a[2] = 1.0f; a[1] = r1 * a0; a[0] = -(r1 * r1);
}

} // extern "C"
