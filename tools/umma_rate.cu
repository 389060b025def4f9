// Measured rate of tcgen05.mma kind::i8 (u8 x s8 -> s32, cta_group::1, M = 128) on this GPU: back-to-back MMAs on resident
// shared-memory operands, one CTA per SM.  The roofline denominator of the two integer contractions (k1_toeplitz_i8, chan_mma_i8).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_rate umma_rate.cu
#include "../fm_radio_b200/csrc/tcgen05.cuh"
#include <cstdio>

template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, int* sink)
{
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                 // 128 rows x 128 B
    uint8_t* sB = smem + 16384;         // N rows x 128 B
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    for (int i = threadIdx.x; i < (16384 + N * 128) / 4; i += 128) ((uint32_t*)smem)[i] = 0x01010101u;
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::mbar_init_fence(); }
    if (threadIdx.x < 32) tc::tmem_alloc(&s_tmem, 256);
    tc::fence_async_smem(); tc::fence_before(); __syncthreads(); tc::fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = tc::idesc_i8_u8s8(128, N);
        const uint32_t a = tc::smem_u32(sA), b = tc::smem_u32(sB);
        for (int it = 0; it < iters; it++)
#pragma unroll
            for (int ks = 0; ks < 4; ks++)
                tc::mma_i8(tmem, tc::smem_desc_sw128(a + ks * 32), tc::smem_desc_sw128(b + ks * 32), idesc, (it | ks) ? 1u : 0u);
        tc::commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after();
    uint32_t r[16];
    tc::tmem_ld16(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), r);
    tc::tmem_ld_wait();
    if (r[0] == 0xdeadbeefu) sink[0] = (int)r[1];
    tc::fence_before(); __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem, 256);
}

template <int N> void run(int n_sm, int clk_khz) {
    int* sink; cudaMalloc(&sink, 4);
    const int smem = 16384 + N * 128 + 1024;
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 20000;
    rate_kernel<N><<<n_sm, 128, smem>>>(200, sink);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<N><<<n_sm, 128, smem>>>(iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double macs = (double)n_sm * iters * 4 * 128.0 * N * 32.0;
    const double cyc_per_mma = ms * 1e-3 * clk_khz * 1e3 / (iters * 4.0);
    printf("kind::i8 M128 N%-3d K32: %.3f ms, %.1f cycles per MMA, %.0f MAC/clk/SM, %.1f TOP/s (%s)\n", N, ms, cyc_per_mma,
           128.0 * N * 32.0 / cyc_per_mma, 2.0 * macs / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    int n_sm = 148, clk = 1965000;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%d SMs, %d kHz\n", n_sm, clk);
    run<64>(n_sm, clk); run<96>(n_sm, clk); run<128>(n_sm, clk); run<192>(n_sm, clk); run<256>(n_sm, clk);
    return 0;
}
