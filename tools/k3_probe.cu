// Times K3 (pilot PLL recurrence) alone on a synthetic locked pilot: cycles per sample of the
// dependent chain.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../fm_radio_b200/csrc k3_probe.cu
#include "../fm_radio_b200/csrc/k3_pll.cu"
#include <cstdio>
#include <vector>
#include <cmath>

template <int WRAP, bool FAST = false>
static float run(int S, int n, const float* theta, const float* power, float* state, float* dt, const fm::K3Params& p, int reps, int nb = 1) {
    // nb > 1: rotate over nb copies of theta / pll_dt (nb x 33.5 MB each) so every launch reads DRAM-cold input,
    // as in the chain where K2's 167 MB of traffic has pushed part of theta out of the L2
    const size_t stride = (size_t)S * n;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) fm::k3_pll<false, WRAP, FAST><<<(S + 31) / 32, 32>>>(theta, power, state, dt, nullptr, nullptr, p);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) fm::k3_pll<false, WRAP, FAST><<<(S + 31) / 32, 32>>>(theta + (i % nb) * stride, power, state, dt + (i % nb) * stride, nullptr, nullptr, p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const int S = 1024, n = 8192;
    std::vector<float> th((size_t)S * n), pw(S, (float)n);
    for (int s = 0; s < S; s++)
        for (int i = 0; i < n; i++) {
            double ph = 19000.0 / 128000.0 * i + 0.001 * s + 0.01 * std::sin(i * 0.001);
            ph -= std::floor(ph + 0.5);
            th[(size_t)s * n + i] = (float)ph;
        }
    float *d_th, *d_pw, *d_st, *d_dt;
    const int NB = 8;
    cudaMalloc(&d_th, NB * th.size() * 4); cudaMalloc(&d_pw, S * 4); cudaMalloc(&d_st, fm::PLL_STATE_N * S * 4); cudaMalloc(&d_dt, NB * th.size() * 4);
    for (int b = 0; b < NB; b++) cudaMemcpy(d_th + b * th.size(), th.data(), th.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_pw, pw.data(), S * 4, cudaMemcpyHostToDevice);
    std::vector<float> st(fm::PLL_STATE_N * S, 0.0f);
    for (int s = 0; s < S; s++) st[fm::PLL_AGC_GAIN * S + s] = 0.1f;
    cudaMemcpy(d_st, st.data(), st.size() * 4, cudaMemcpyHostToDevice);
    fm::K3Params p{};
    p.lpf_b[0] = p.lpf_b[1] = 0.0024f; p.lpf_a[0] = 0.995f; p.lpf_a[1] = 0;
    p.int_KTs = 0.1f / 128000.0f; p.Kp = 0.01f; p.f_center = -19000.0f; p.f_gain = -100.0f; p.mixer_KTs = 1.0f / 128000.0f;
    p.agc_target = 1.0f; p.agc_beta = 0.2f; p.n = n; p.n_streams = S; p.keep = 0;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const float ms0 = run<0>(S, n, d_th, d_pw, d_st, d_dt, p, 20);
    const float ms1 = run<1>(S, n, d_th, d_pw, d_st, d_dt, p, 20);
    const float ms2 = run<0>(S, n, d_th, d_pw, d_st, d_dt, p, 24, NB);
    p.integ_safe = 0.9f;
    const float ms3 = run<0, true>(S, n, d_th, d_pw, d_st, d_dt, p, 40);
    const float ms4 = run<0, true>(S, n, d_th, d_pw, d_st, d_dt, p, 48, NB);
    printf("k3 FAST pass (locked pilot): %.4f ms/launch = %.1f cycles/sample; DRAM-cold input %.4f ms = %.1f cycles/sample\n", ms3, ms3 * 1e-3 * clk * 1e3 / n, ms4, ms4 * 1e-3 * clk * 1e3 / n);
    printf("k3 WRAP=0, DRAM-cold input (8 x 33.5 MB rotated): %.4f ms/launch = %.1f cycles/sample\n", ms2, ms2 * 1e-3 * clk * 1e3 / n);
    printf("k3 WRAP=0 (magic adds): %.4f ms/launch = %.1f cycles/sample @%d kHz\n", ms0, ms0 * 1e-3 * clk * 1e3 / n, clk);
    printf("k3 WRAP=1 (FRND)      : %.4f ms/launch = %.1f cycles/sample\n", ms1, ms1 * 1e-3 * clk * 1e3 / n);
    std::vector<float> dt(16);
    cudaMemcpy(dt.data(), d_dt + (size_t)5 * n + n - 16, 64, cudaMemcpyDeviceToHost);
    printf("tail dt: %f %f %f (%s)\n", dt[13], dt[14], dt[15], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
