// Measures the FP32 FMA issue ceilings of the device this runs on (the denominators of the FIR
// kernels' rooflines): FFMA reg*reg+reg, FFMA with a constant-bank operand, packed FFMA2
// (fma.rn.f32x2, sm_100+), and the mixes the FIR kernels issue (FFMA + PRMT/FADD, FFMA + LDS).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_probe ffma_probe.cu && ./ffma_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

struct Taps { float t[64]; };
constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_ffma_reg(float* out, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 0.001f + i;
    float x = a + threadIdx.x, y = b;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fmaf(acc[i], x, y);
    }
    float s = 0; for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma_const(float* out, float a, const __grid_constant__ Taps tp) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 0.001f + i;
    float x[4] = { a + threadIdx.x, a * 2 + threadIdx.x, a * 3, a * 4 };
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fmaf(x[k], tp.t[(i + 16 * k) & 63], acc[i]);
        x[0] += 1.0f;
    }
    float s = 0; for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    return ((uint64_t)__float_as_uint(hi) << 32) | __float_as_uint(lo);
}

__global__ void __launch_bounds__(256) k_ffma2_reg(float* out, float a, float b) {
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = pack2(threadIdx.x * 0.001f + i, i * 0.5f);
    const uint64_t x = pack2(a + threadIdx.x, a), y = pack2(b, b * 2);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = ffma2(acc[i], x, y);
    }
    float s = 0; for (int i = 0; i < 8; i++) s += __uint_as_float((uint32_t)acc[i]) + __uint_as_float((uint32_t)(acc[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FIR-like: acc(re,im) += x(re,im) * (tap,tap): the tap pair must be a register pair
__global__ void __launch_bounds__(256) k_ffma2_fir(float* out, float a, const __grid_constant__ Taps tp) {
    uint64_t acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = pack2(threadIdx.x * 0.001f + i, i * 0.5f);
    uint64_t x[4] = { pack2(a + threadIdx.x, a), pack2(a * 2, a + 1), pack2(a * 3, a), pack2(a * 4, a) };
    uint64_t tt[16];
#pragma unroll
    for (int i = 0; i < 16; i++) tt[i] = pack2(tp.t[i], tp.t[i]);
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = ffma2(x[k], tt[(i + 5 * k) & 15], acc[i]);
        x[0] += 1;
    }
    float s = 0; for (int i = 0; i < 16; i++) s += __uint_as_float((uint32_t)acc[i]) + __uint_as_float((uint32_t)(acc[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA(const) with 1 ALU-pipe op (PRMT) + 1 FADD per 8 FFMAs, as K1 issues them
__global__ void __launch_bounds__(256) k_ffma_const_mix(float* out, uint32_t w, const __grid_constant__ Taps tp) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 0.001f + i;
    uint32_t ww = w + threadIdx.x;
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float x[2];
#pragma unroll
            for (int h = 0; h < 2; h++)
                x[h] = __uint_as_float(__byte_perm(ww, 0x4B000000u, 0x7440u | (2 * k + h) & 3)) - 8388735.0f;
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = fmaf(x[i & 1], tp.t[(i + 16 * k) & 63], acc[i]);
        }
        ww = ww * 1664525u + 1013904223u;
    }
    float s = 0; for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static void run(const char* name, F launch, double flop_per_thread, int ctas, int threads) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) launch();
    cudaEventRecord(e0);
    const int reps = 10;
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double tf = flop_per_thread * ctas * threads * reps / (ms * 1e-3) / 1e12;
    printf("%-28s %8.3f ms/launch  %7.2f TFLOP/s  (%s)\n", name, ms / reps, tf, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s  SMs %d  clock %d kHz  nominal 128-lane peak %.2f TFLOP/s\n", pr.name, pr.multiProcessorCount, clk,
           pr.multiProcessorCount * 128.0 * 2 * clk * 1e3 / 1e12);
    const int threads = 256;
    Taps tp; for (int i = 0; i < 64; i++) tp.t[i] = 1e-3f * (i + 1);
    float* out; cudaMalloc(&out, sizeof(float) * threads * 148 * 64);
    for (int per_sm : { 2, 4, 8 }) {
        const int ctas = pr.multiProcessorCount * per_sm * 4;
        printf("-- %d CTAs (%d co-resident per SM x 4 waves), %d threads\n", ctas, per_sm, threads);
        run("FFMA reg,reg,reg", [&] { k_ffma_reg<<<ctas, threads>>>(out, 1.0001f, 0.5f); }, 2.0 * 16 * ITERS, ctas, threads);
        run("FFMA reg,const,reg", [&] { k_ffma_const<<<ctas, threads>>>(out, 1.0001f, tp); }, 2.0 * 16 * ITERS, ctas, threads);
        run("FFMA2 reg,reg,reg", [&] { k_ffma2_reg<<<ctas, threads>>>(out, 1.0001f, 0.5f); }, 2.0 * 16 * ITERS, ctas, threads);
        run("FFMA2 fir (x, tap pair)", [&] { k_ffma2_fir<<<ctas, threads>>>(out, 1.0001f, tp); }, 2.0 * 32 * ITERS, ctas, threads);
        run("FFMA const + PRMT/FADD 1:8", [&] { k_ffma_const_mix<<<ctas, threads>>>(out, 12345u, tp); }, 2.0 * 16 * ITERS, ctas, threads);
    }
    cudaFree(out);
    return 0;
}
