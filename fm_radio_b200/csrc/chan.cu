// Wideband channelizer (BASELINE config 4): one wideband u8 IQ capture -> C narrow-band complex
// channels at Fs/D, each feeding one demodulator stream of the chain in fmgpu.cu.
//
// The reference has no channelizer (SURVEY.md section 7 "No channelizer in the reference", section
// 8(f) rank 2; the reference's own TODO on a configurable front end is broadcast_fm_demod.cpp:67).
// The definition implemented here -- and restated in float64 by the test-side checker
// (fmo_channelize_f64) -- keeps the reference's conventions on either side of it:
//   unpack    x[n] = (float)u8 - 127                                       app.cpp:56-65
//   shift     xs[n] = x[n] * exp(-j 2 pi ph_c(n) / 2^32), ph_c(n) = inc_c * n mod 2^32
//   decimate  y_c[i] = sum_{k<NN} b[k] * xs[(i+1) D - NN + k]              dsp/polyphase_filter.h:41-64
//   taps      b = create_fir_lpf(NN, k) in ReverseArray order              dsp/filter_designer.cpp:84-107
//
// Because exp(-j ph(n - d)) = exp(-j ph(n)) exp(+j ph(d)), the shift moves into the taps:
//   y_c[i] = exp(-j ph_c(n_i)) * sum_k g_c[k] x[n_i - (NN-1-k)],   g_c[k] = b[k] exp(+j ph_c(NN-1-k)),
// n_i = the newest sample of output i.  For all channels at once that sum is a DENSE CONTRACTION
//   Y[i, (c, re|im)] = sum_{K < 2 NN} X[i, K] * G[K, (c, re|im)],  X[i, K] = byte K of the window of output i
// (rows of X are overlapping 2*NN-byte windows of the raw capture, 2*D bytes apart), which is the one
// place of this path where the north star allows tensor cores.  Two kernels:
//
//   chan_mma_i8    tcgen05.mma kind::i8 (sm_100a): X is the RAW u8 capture (unsigned 8-bit A operand,
//                  no unpack pass), G is split into three balanced signed base-256 digit planes
//                  (24-bit fixed point, as exact as the fp32 taps), accumulators are int32 in TMEM, so the
//                  contraction is EXACT integer arithmetic; the -127 offset is removed as an integer
//                  constant per column, the three planes are recombined in fp32 and the result is rotated
//                  by exp(-j ph_c(n_i)) in the epilogue.
//   chan_fir_fp32  the same sum on the FP32 FMA pipe (one lane per channel, taps in shared memory).
//                  Kept as the measured alternative and as an on-device cross-check of the MMA path.
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/fmgpu.h"
#include "tcgen05.cuh"
#include "fm_common.cuh"

extern "C" int fmgpu_set_last_error_(int code, const char* msg);            // fmgpu.cu

namespace {

#define CHK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    return fmgpu_set_last_error_(FMGPU_ERR_CUDA, (std::string(#call) + ": " + cudaGetErrorString(e_)).c_str()); } } while (0)

constexpr int CH_ROWS = 128;            // output times per MMA tile (= TMEM lanes)
constexpr int CH_SLOTS = 32;            // channel slots per group
constexpr int CH_NG = 2 * CH_SLOTS;     // accumulator columns per digit plane (re, im per slot)
constexpr int CH_PLANES = 3;
constexpr int CH_KCHUNK = 128;          // bytes of K per staged chunk = one 128-byte swizzle row
constexpr int CH_A_BYTES = CH_ROWS * CH_KCHUNK;                 // 16 KB
constexpr int CH_BSUB_BYTES = CH_NG * CH_KCHUNK;                // 8 KB: one (chunk, plane) sub-tile of G
constexpr int CH_TMEM_COLS = 256;       // 3 x 64 used

struct ChanMmaParams {
    const uint8_t* iq;          // staged capture: history ++ block; window of output i starts at byte i * row_bytes
    const int8_t* bimg;         // [group][chunk][plane][8 KB]: G digit planes in the shared-memory (swizzled) image
    const int* offs;            // [group][plane][64]: 127 * column sums of the digit planes
    const uint4* meta;          // [group][32]: { inc, inc * D, output channel or 0xffffffff, 0 }
    float2* out;                // [C][n_out]
    uint32_t n_newest0;         // (absolute index of the newest sample of output 0) mod 2^32
    int n_out, row_bytes, n_kchunks, n_tiles;
    float w0, w1, w2;           // weights of the digit planes
};

// ---------------------------------------------------------------- PTX wrappers (sm_100a): tcgen05.cuh ----------
using tc::smem_u32; using tc::mbar_init; using tc::mbar_wait; using tc::fence_async_smem;
__device__ __forceinline__ void tc_fence_before() { tc::fence_before(); }
__device__ __forceinline__ void tc_fence_after() { tc::fence_after(); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) { tc::commit(bar); }
__device__ __forceinline__ void tc_mma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { tc::mma_i8(d, a, b, idesc, acc); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) { tc::tmem_ld16(taddr, r); }
__device__ __forceinline__ void tmem_ld_wait() { tc::tmem_ld_wait(); }
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) { return tc::smem_desc_sw128(saddr); }
constexpr uint32_t CH_IDESC = tc::idesc_i8_u8s8(CH_ROWS, CH_NG);

// ---------------------------------------------------------------- tensor-core kernel -------------
// grid (ctas_per_group, n_groups), 128 threads, persistent over the time tiles of its group; two CTAs per SM
// (104 KB of shared memory and 256 TMEM columns each) so that one CTA's epilogue overlaps the other's MMAs.
__global__ void __launch_bounds__(128)
chan_mma_i8(const __grid_constant__ ChanMmaParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);     // swizzle atoms need 1024-byte alignment
    uint8_t* sB = smem;                                              // n_kchunks x 3 x 8 KB, resident
    uint8_t* sA = sB + (size_t)p.n_kchunks * CH_PLANES * CH_BSUB_BYTES;   // 2 x 16 KB ring
    __shared__ __align__(8) uint64_t bar_free[2];                    // MMAs that read ring buffer b have completed
    __shared__ __align__(8) uint64_t bar_acc;                        // the tile's accumulators are complete
    __shared__ uint32_t s_tmem;
    __shared__ int s_off[CH_PLANES * CH_NG];
    __shared__ uint4 s_meta[CH_SLOTS];

    const int tid = threadIdx.x, warp = tid >> 5;
    const int group = blockIdx.y;

    {   // G of this group: a linear copy, the image is already swizzled
        const uint4* src = (const uint4*)(p.bimg + (size_t)group * p.n_kchunks * CH_PLANES * CH_BSUB_BYTES);
        const int n16 = p.n_kchunks * CH_PLANES * CH_BSUB_BYTES / 16;
        for (int i = tid; i < n16; i += 128) ((uint4*)sB)[i] = __ldg(src + i);
    }
    for (int i = tid; i < CH_PLANES * CH_NG; i += 128) s_off[i] = p.offs[(size_t)group * CH_PLANES * CH_NG + i];
    if (tid < CH_SLOTS) s_meta[tid] = p.meta[(size_t)group * CH_SLOTS + tid];
    if (tid == 0) {
        mbar_init(&bar_free[0], 1); mbar_init(&bar_free[1], 1); mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&s_tmem)), "r"(CH_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();                   // sB was written through the generic proxy, the MMA reads it through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    int n_valid = 0;
    for (int s = 0; s < CH_SLOTS; s++) if (s_meta[s].z != 0xffffffffu) n_valid = s + 1;

    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    const int row8 = tid & 7;
    uint8_t* a_row = sA + (tid >> 3) * 1024 + row8 * 128;            // this thread's row in ring buffer 0
    uint32_t chunk = 0, acc_parity = 0;

    for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint8_t* src_row = p.iq + (size_t)(tile * CH_ROWS + tid) * p.row_bytes;
        for (int kc = 0; kc < p.n_kchunks; kc++, chunk++) {
            const uint32_t buf = chunk & 1u, use = chunk >> 1;
            if (use > 0) mbar_wait(&bar_free[buf], (use - 1) & 1u);
            // ---- im2col: 128 bytes of this thread's window -> its (swizzled) row of the ring buffer ----
            const uint2* s8 = (const uint2*)(src_row + kc * CH_KCHUNK);          // 8-byte aligned (row_bytes % 8 == 0)
            uint8_t* dst = a_row + buf * CH_A_BYTES;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint2 lo = __ldg(s8 + 2 * u), hi = __ldg(s8 + 2 * u + 1);
                *(uint4*)(dst + ((u ^ row8) << 4)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
            }
            fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t a_base = sA_addr + buf * CH_A_BYTES;
#pragma unroll
                for (int ks = 0; ks < CH_KCHUNK / 32; ks++) {                    // K = 32 per tcgen05.mma kind::i8
                    const uint64_t a_desc = smem_desc_sw128(a_base + ks * 32);
#pragma unroll
                    for (int j = 0; j < CH_PLANES; j++) {
                        const uint64_t b_desc = smem_desc_sw128(sB_addr + (kc * CH_PLANES + j) * CH_BSUB_BYTES + ks * 32);
                        tc_mma_i8(tmem + j * CH_NG, a_desc, b_desc, CH_IDESC, (kc | ks) != 0 ? 1u : 0u);
                    }
                }
                tc_commit(&bar_free[buf]);
                if (kc == p.n_kchunks - 1) tc_commit(&bar_acc);
            }
        }
        mbar_wait(&bar_acc, acc_parity);
        acc_parity ^= 1u;
        tc_fence_after();

        // ---- epilogue: TMEM lane tid = output time i; 8 channel slots (16 columns) per step ----
        const int i_out = tile * CH_ROWS + tid;
        const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int q = 0; q * 8 < n_valid; q++) {
            uint32_t a0[16], a1[16], a2[16];
            tmem_ld16(lane_addr + 0 * CH_NG + q * 16, a0);
            tmem_ld16(lane_addr + 1 * CH_NG + q * 16, a1);
            tmem_ld16(lane_addr + 2 * CH_NG + q * 16, a2);
            tmem_ld_wait();
#pragma unroll
            for (int s = 0; s < 8; s++) {
                const int slot = q * 8 + s;
                const uint4 m = s_meta[slot];
                if (m.z == 0xffffffffu) continue;
                const int c0 = 2 * slot, c1 = 2 * slot + 1;
                const float r0 = (float)((int)a0[2 * s] - s_off[c0]), r1 = (float)((int)a1[2 * s] - s_off[CH_NG + c0]),
                            r2 = (float)((int)a2[2 * s] - s_off[2 * CH_NG + c0]);
                const float q0 = (float)((int)a0[2 * s + 1] - s_off[c1]), q1 = (float)((int)a1[2 * s + 1] - s_off[CH_NG + c1]),
                            q2 = (float)((int)a2[2 * s + 1] - s_off[2 * CH_NG + c1]);
                const float re = fmaf(r0, p.w0, fmaf(r1, p.w1, r2 * p.w2));
                const float im = fmaf(q0, p.w0, fmaf(q1, p.w1, q2 * p.w2));
                const uint32_t ph = m.x * p.n_newest0 + m.y * (uint32_t)i_out;            // exact mod 2^32
                float sn, cs;
                sincospif((float)(int)ph * 4.656612873077393e-10f, &sn, &cs);             // turns * 2 -> pi units
                p.out[(size_t)m.z * p.n_out + i_out] = make_float2(fmaf(re, cs, im * sn), fmaf(im, cs, -re * sn));
            }
        }
        tc_fence_before();
        __syncthreads();                  // every lane has drained its accumulators before the next tile overwrites them
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(CH_TMEM_COLS) : "memory");
}

// ---------------------------------------------------------------- tensor-core kernel, pipelined ----
// The kernel above is synchronous: every thread copies its row of a K chunk through registers, the CTA meets at
// a barrier, one thread issues the MMAs, and the epilogue (32 x sincospif per thread) runs with the tensor pipe
// idle -- ncu: tensor pipe 4 % active, 61 us per launch for 100 stations x 65536 outputs, top stall
// long_scoreboard.  This one is the classic three-role pipeline, one CTA of 672 threads per SM:
//   warps 0-3  PRODUCERS   im2col of the raw capture into a ring of CH2_NSTG K-chunk stages with cp.async (8-byte
//                          copies: window rows are 2 D = 40 bytes apart, so they are 8- but not 16-byte aligned,
//                          which also rules out TMA), two chunks in flight ahead of the one being released;
//   warp  4    MMA ISSUER  one elected lane: 4 tcgen05.mma (K steps; N = 192 = 3 digit planes) per chunk into one of TWO
//                          accumulator buffers in TMEM, tcgen05.commit frees the stage / publishes the tile;
//   warps 5-20 EPILOGUE    four warps per TMEM lane quarter, one block of 8 channel slots each: tcgen05.ld of the finished buffer
//                          while the next tile's MMAs run; branch-free per-channel arithmetic (8 independent channels in
//                          flight per thread, packed FP32 for the (re, im) pair); the channel rotation exp(-j ph) comes
//                          from a [slot][row] table in shared memory built once per launch with the chain's own
//                          polynomial sine, times one factor per (tile, slot) that is +-1 on the fs_out / 256 raster.
//                          (History: one epilogue warp per SM sub-partition and a branch per channel ran at IPC 0.2
//                          and bounded the kernel, 51 us; two polynomial sines per output sample, 41.8 us with the
//                          epilogue issuing 2/3 of the kernel's 18.6 M warp instructions -- ncu, profiles/r2z2_ncu_summary.md.)
// The three digit planes of G sit side by side in shared memory, so each K step is ONE tcgen05.mma of N = 192 (100 clocks
// measured, tools/umma_rate.cu) instead of three of N = 64 (3 x 50: small-N MMAs have a ~50-clock floor).
// All hand-offs are mbarriers (full / empty per stage, acc_full / acc_empty per accumulator buffer).
constexpr int CH2_EPI_WARPS = 16;               // 4 per TMEM lane quarter: warp (quarter, eg) owns the block eg of 8 channel slots
constexpr int CH2_THREADS = 32 * (5 + CH2_EPI_WARPS);   // 4 producer warps, 1 MMA-issuer warp, 16 epilogue warps = 672 threads
constexpr int CH2_NSTG = 6;                     // 6 x 16 KB ring + 72 KB of G + 32 KB rotation table = 200 KB of shared memory: one CTA per SM
constexpr int CH2_TMEM_COLS = 512;              // 2 accumulator buffers, 192 of 256 columns used in each
constexpr int CH2_LAG = 2;                      // chunks of cp.async in flight behind the producers' issue point

template <bool NEED_F>                          // NEED_F: some channel's phase advances per 128-output tile by something other than 0 or 1/2 turn
__global__ void __launch_bounds__(CH2_THREADS, 1)
chan_mma_i8_pipelined(const __grid_constant__ ChanMmaParams p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;                                              // n_kchunks x 3 x 8 KB, resident
    uint8_t* sA = sB + (size_t)p.n_kchunks * CH_PLANES * CH_BSUB_BYTES;   // CH2_NSTG x 16 KB ring
    float2* sT = (float2*)(sA + CH2_NSTG * CH_A_BYTES);              // [slot][row] rotation table of this launch, 32 KB
    __shared__ __align__(8) uint64_t bar_full[CH2_NSTG], bar_empty[CH2_NSTG], bar_acc_full[2], bar_acc_empty[2], bar_g;
    __shared__ uint32_t s_tmem;
    __shared__ int2 s_off2[CH_PLANES][CH_SLOTS];                    // (re, im) column offsets of each slot, per digit plane
    __shared__ uint4 s_meta[CH_SLOTS];
    __shared__ float2* s_outp[CH_SLOTS];                            // output row of each slot's channel (nullptr: empty slot)
    __shared__ float2 s_F[CH2_EPI_WARPS][8];                             // per epilogue warp: this tile's phase factor of each slot

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform
    const int group = blockIdx.y;
    // G (72 KB) arrives by bulk copies issued right after the barrier initialisation; only the MMA issuer waits for it
    if (tid == 0) {
        for (int i = 0; i < CH2_NSTG; i++) { mbar_init(&bar_full[i], 4); mbar_init(&bar_empty[i], 1); }   // full: one arrival per producer warp
        for (int i = 0; i < 2; i++) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], CH2_EPI_WARPS); }  // acc_empty: one arrival per epilogue warp
        mbar_init(&bar_g, 1);
        tc::mbar_init_fence();
        const uint32_t g_bytes = (uint32_t)(p.n_kchunks * CH_PLANES * CH_BSUB_BYTES);
        const int8_t* src = p.bimg + (size_t)group * g_bytes;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar_g)), "r"(g_bytes) : "memory");
        for (uint32_t o = 0; o < g_bytes; o += CH_BSUB_BYTES)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(smem_u32(sB + o)), "l"(src + o), "r"((uint32_t)CH_BSUB_BYTES), "r"(smem_u32(&bar_g)) : "memory");
    }
    for (int i = tid; i < CH_PLANES * CH_NG; i += CH2_THREADS) ((int*)s_off2)[i] = p.offs[(size_t)group * CH_PLANES * CH_NG + i];
    if (tid < CH_SLOTS) {
        const uint4 m = p.meta[(size_t)group * CH_SLOTS + tid];
        s_meta[tid] = m;
        s_outp[tid] = m.z != 0xffffffffu ? p.out + (size_t)m.z * p.n_out : nullptr;
    }
    if (warp == 0) tc::tmem_alloc(&s_tmem, CH2_TMEM_COLS);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    const int n_my_tiles = (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_chunks = n_my_tiles * p.n_kchunks;

    if (warp < 4) {
        // ---------------- producers: row 32 warp + lane of every tile ----------------
        auto arrive_full = [&](int stg) {            // this warp's copies into ring stage stg have landed
            fence_async_smem();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar_full[stg])) : "memory");
        };
        // One instruction copies two whole window rows (16 lanes x 8 bytes each): the rows start 2 D bytes apart, so a warp
        // reads 2 D + 128 contiguous bytes (6 sectors) and writes two full 128-byte swizzle rows (2 wavefronts); with one
        // row per lane (stride 2 D) the same instruction touched 32 sectors and wrote with 4-way bank conflicts.
        // Everything that does not change from chunk to chunk is computed here (the loop used to spend 138 of its 154
        // instructions per chunk on divisions and address arithmetic, and the producers bounded the kernel -- ncu source view).
        const int h = lane >> 4, j = lane & 15;
        uint32_t dsto[4];                                // swizzled offset of row (2 i + h) & 7 within its 8-row group, i = 0..3
#pragma unroll
        for (int i = 0; i < 4; i++) { const int r7 = 2 * i + h; dsto[i] = (uint32_t)(r7 * 128 + (((j >> 1) ^ r7) << 4) + ((j & 1) << 3)); }
        const uint32_t dst0 = smem_u32(sA + 4 * warp * 1024);
        const size_t rb2 = 2 * (size_t)p.row_bytes, tile_stride = (size_t)gridDim.x * CH_ROWS * p.row_bytes;
        const uint8_t* tile_src = p.iq + (size_t)((int)blockIdx.x * CH_ROWS + 32 * warp + h) * p.row_bytes + 8 * j;
        int stage = 0, lag_stage = 0, c = 0;
        uint32_t empty_parity = 1u;                      // parity of the "previous use released" phase; first pass over the ring: no wait
        bool first_pass = true;
        for (int t = 0; t < n_my_tiles; t++, tile_src += tile_stride) {
            for (int kc = 0; kc < p.n_kchunks; kc++, c++) {
                if (!first_pass) mbar_wait(&bar_empty[stage], empty_parity);
                const uint8_t* src = tile_src + kc * CH_KCHUNK;
                const uint32_t dst = dst0 + (uint32_t)(stage * CH_A_BYTES);
#pragma unroll
                for (int i = 0; i < 16; i++, src += rb2)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst + (uint32_t)((i >> 2) * 1024) + dsto[i & 3]), "l"(src));
                tc::cp_async_commit();
                if (++stage == CH2_NSTG) { stage = 0; empty_parity ^= 1u; first_pass = false; }
                if (c >= CH2_LAG) {
                    tc::cp_async_wait<CH2_LAG>();
                    arrive_full(lag_stage);
                    if (++lag_stage == CH2_NSTG) lag_stage = 0;
                }
            }
        }
        for (int k = max(0, n_chunks - CH2_LAG); k < n_chunks; k++) {
            tc::cp_async_wait<0>();
            arrive_full(lag_stage);
            if (++lag_stage == CH2_NSTG) lag_stage = 0;
        }
    } else if (warp == 4) {
        // ---------------- MMA issuer ----------------
        const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
        mbar_wait(&bar_g, 0u);                                       // G has landed
        for (int t = 0; t < n_my_tiles; t++) {
            const int buf = t & 1;
            if (t >= 2) mbar_wait(&bar_acc_empty[buf], (uint32_t)(((t >> 1) - 1) & 1));      // the epilogue has drained this buffer
            tc_fence_after();
            for (int kc = 0; kc < p.n_kchunks; kc++) {
                const int c = t * p.n_kchunks + kc, stage = c % CH2_NSTG;
                mbar_wait(&bar_full[stage], (uint32_t)((c / CH2_NSTG) & 1));
                tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t a_base = sA_addr + stage * CH_A_BYTES;
#pragma unroll
                    for (int ks = 0; ks < CH_KCHUNK / 32; ks++) {
                        // the chunk's three plane sub-tiles are contiguous: one 192-row B operand, plane j in columns 64 j ..
                        const uint64_t a_desc = smem_desc_sw128(a_base + ks * 32);
                        const uint64_t b_desc = smem_desc_sw128(sB_addr + kc * CH_PLANES * CH_BSUB_BYTES + ks * 32);
                        tc_mma_i8(tmem + (uint32_t)(buf * 256), a_desc, b_desc, tc::idesc_i8_u8s8(CH_ROWS, CH_PLANES * CH_NG), (kc | ks) != 0 ? 1u : 0u);
                    }
                    tc_commit(&bar_empty[stage]);
                    if (kc == p.n_kchunks - 1) tc_commit(&bar_acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue (warps 5-20): TMEM lane = output time, a warp reads the lane quarter warp % 4; the four
        // warps of a quarter (eg = 0..3) each own one block of 8 channel slots, so the rotation table entries of a thread
        // (its row x its 8 slots) are loaded ONCE and stay in registers ----------------
        const int ew = warp & 3, eg = (warp - 5) >> 2, row = ew * 32 + lane;
        static_assert(CH2_EPI_WARPS / 4 * 8 == CH_SLOTS, "one block of 8 slots per epilogue warp");
        // Rotation table, built by the epilogue warps while the producers and the MMA issuer already run: the channel
        // rotation of output i = 128 tile + row is exp(-j 2 pi ph / 2^32), ph = inc n_newest0 + inc D i (exact mod 2^32).
        // T[slot][row] holds the row part (tile 0) by the chain's polynomial sine (dsp/simd/chebyshev_sine.h, 1.5e-7:
        // sin(2 pi t) = S(t), cos(2 pi t) = S(1/4 - |t|)); the tile part exp(-j 2 pi 128 tile inc D / 2^32) is one factor
        // per (tile, slot), and is exactly +-1 when 128 inc D = 0 mod 2^31 (centres on the fs_out / 256 = 4 kHz raster:
        // the FM band's 200 kHz raster and its half points at 1.024 MS/s are).
        for (int i = tid - 160; i < CH_SLOTS * CH_ROWS; i += 32 * CH2_EPI_WARPS) {
            const uint4 m = s_meta[i >> 7];
            const uint32_t ph = m.x * p.n_newest0 + m.y * (uint32_t)(i & (CH_ROWS - 1));
            const float tt = (float)(int)ph * 2.3283064365386963e-10f;
            sT[i] = make_float2(fm::chebyshev_sine(0.25f - fabsf(tt)), fm::chebyshev_sine(tt));      // (cos, sin)
        }
        asm volatile("bar.sync 1, %0;" :: "n"(32 * CH2_EPI_WARPS) : "memory");
        bool any = false;
#pragma unroll
        for (int s = 0; s < 8; s++) any |= s_meta[eg * 8 + s].z != 0xffffffffu;
        float2 T[8];
#pragma unroll
        for (int s = 0; s < 8; s++) T[s] = sT[(eg * 8 + s) * CH_ROWS + row];
        const uint32_t tile_step = s_meta[eg * 8 + (lane & 7)].y << 7;                 // lane & 7 = slot of the block: phase step per tile
        // half-turn steps (centres on the raster's half points: the 100-station plan is) only flip the sign on odd tiles
        const uint32_t half_mask = __ballot_sync(0xffffffffu, tile_step == 0x80000000u) & 0xffu;
        float2* my_F = s_F[warp - 5];
        const float2 w0 = make_float2(p.w0, p.w0), w1 = make_float2(p.w1, p.w1), w2 = make_float2(p.w2, p.w2);
        auto release = [&](int buf) {                    // this warp has finished reading accumulator buffer buf
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar_acc_empty[buf])) : "memory");
        };
        // one slot: digit planes -> complex sample (exact integer offsets first, then the same FMA order as the FP32
        // kernel's reference sum), rotated by (cs, sn); the store is explicit st.global (the pointer comes out of shared memory)
        auto finish = [&](int slot, uint32_t ar0, uint32_t ai0, uint32_t ar1, uint32_t ai1, uint32_t ar2, uint32_t ai2, float2 rot, int i_out) {
            const int2 o0 = s_off2[0][slot], o1 = s_off2[1][slot], o2 = s_off2[2][slot];
            const float2 v0 = make_float2((float)((int)ar0 - o0.x), (float)((int)ai0 - o0.y));
            const float2 v1 = make_float2((float)((int)ar1 - o1.x), (float)((int)ai1 - o1.y));
            const float2 v2 = make_float2((float)((int)ar2 - o2.x), (float)((int)ai2 - o2.y));
            const float2 y = __ffma2_rn(v0, w0, __ffma2_rn(v1, w1, __fmul2_rn(v2, w2)));
            float2* dst = s_outp[slot];
            if (dst) asm volatile("st.global.v2.f32 [%0], {%1, %2};" :: "l"(dst + i_out), "f"(fmaf(y.x, rot.x, y.y * rot.y)), "f"(fmaf(y.y, rot.x, -y.x * rot.y)) : "memory");
        };
        for (int t = 0; t < n_my_tiles; t++) {
            const int buf = t & 1;
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            if (NEED_F) {                                // lanes 0-7: exp(-j 2 pi tile_step tile / 2^32) of the block's slots
                const float tt = (float)(int)(tile_step * (uint32_t)tile) * 2.3283064365386963e-10f;
                __syncwarp();
                if (lane < 8) my_F[lane] = make_float2(fm::chebyshev_sine(0.25f - fabsf(tt)), fm::chebyshev_sine(tt));
                __syncwarp();
            }
            mbar_wait(&bar_acc_full[buf], (uint32_t)((t >> 1) & 1));
            tc_fence_after();
            if (!any) { release(buf); continue; }
            const int i_out = tile * CH_ROWS + row;
            const uint32_t lane_addr = tmem + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * 256);
            // two halves of 4 slots (8 columns per digit plane): 24 accumulator registers live instead of 48
#pragma unroll
            for (int hf = 0; hf < 2; hf++) {
                uint32_t a0[8], a1[8], a2[8];
                tc::tmem_ld8(lane_addr + 0 * CH_NG + eg * 16 + hf * 8, a0);
                tc::tmem_ld8(lane_addr + 1 * CH_NG + eg * 16 + hf * 8, a1);
                tc::tmem_ld8(lane_addr + 2 * CH_NG + eg * 16 + hf * 8, a2);
                float2 rot[4];
#pragma unroll
                for (int s4 = 0; s4 < 4; s4++) {
                    const int s = hf * 4 + s4;
                    if (NEED_F) {
                        const float2 F = my_F[s];            // (cos, sin) of the tile part
                        rot[s4] = make_float2(fmaf(T[s].x, F.x, -T[s].y * F.y), fmaf(T[s].y, F.x, T[s].x * F.y));
                    } else {
                        const uint32_t sg = (((tile & 1) ? half_mask : 0u) << (31 - s)) & 0x80000000u;
                        rot[s4] = make_float2(__uint_as_float(__float_as_uint(T[s].x) ^ sg), __uint_as_float(__float_as_uint(T[s].y) ^ sg));
                    }
                }
                tmem_ld_wait();
                if (hf == 1) release(buf);                   // hand the buffer back before the arithmetic
#pragma unroll
                for (int s4 = 0; s4 < 4; s4++)               // branch-free: independent channels; an empty slot only skips its store
                    finish(eg * 8 + hf * 4 + s4, a0[2 * s4], a0[2 * s4 + 1], a1[2 * s4], a1[2 * s4 + 1], a2[2 * s4], a2[2 * s4 + 1], rot[s4], i_out);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, CH2_TMEM_COLS);
}

// ---------------------------------------------------------------- FP32 FMA-pipe kernel -----------
// grid (n_out / 32, ceil(C / 32)), 128 threads: lane = channel, each warp owns 8 consecutive outputs.
constexpr int CF_R = 8, CF_TILE = 4 * CF_R;
struct ChanFp32Params {
    const uint8_t* iq; const float2* g;      // g: [NN][c_pad] tap-major
    const uint2* meta;                       // [C]: { inc, inc * D }
    float2* out; uint32_t n_newest0; int n_out, D, NN, C, c_pad;
};

__global__ void __launch_bounds__(128)
chan_fir_fp32(const __grid_constant__ ChanFp32Params p)
{
    extern __shared__ float2 s_x[];                                  // CF_TILE * D + NN - D samples
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = blockIdx.x * CF_TILE;
    const int n_stage = CF_TILE * p.D + p.NN - p.D;
    const uint8_t* src = p.iq + (size_t)i0 * p.D * 2;
    for (int j = tid; j < n_stage; j += 128) s_x[j] = make_float2((float)src[2 * j] - 127.0f, (float)src[2 * j + 1] - 127.0f);
    __syncthreads();
    const int c = blockIdx.y * 32 + lane;
    if (c >= p.C) return;
    float ar[CF_R], ai[CF_R];
#pragma unroll
    for (int r = 0; r < CF_R; r++) { ar[r] = 0.0f; ai[r] = 0.0f; }
    const float2* gp = p.g + c;
    const float2* xw = s_x + warp * CF_R * p.D;
    for (int k = 0; k < p.NN; k++) {
        const float2 g = gp[(size_t)k * p.c_pad];
#pragma unroll
        for (int r = 0; r < CF_R; r++) {
            const float2 x = xw[r * p.D + k];
            ar[r] = fmaf(g.x, x.x, ar[r]); ar[r] = fmaf(-g.y, x.y, ar[r]);
            ai[r] = fmaf(g.y, x.x, ai[r]); ai[r] = fmaf(g.x, x.y, ai[r]);
        }
    }
    const uint2 m = p.meta[c];
#pragma unroll
    for (int r = 0; r < CF_R; r++) {
        const int i_out = i0 + warp * CF_R + r;
        const uint32_t ph = m.x * p.n_newest0 + m.y * (uint32_t)i_out;
        float sn, cs;
        sincospif((float)(int)ph * 4.656612873077393e-10f, &sn, &cs);
        p.out[(size_t)c * p.n_out + i_out] = make_float2(fmaf(ar[r], cs, ai[r] * sn), fmaf(ai[r], cs, -ar[r] * sn));
    }
}

} // namespace

// ---------------------------------------------------------------- C-ABI object -------------------
struct fmgpu_chan {
    fmgpu_chan_config cfg{};
    int C = 0, D = 0, NN = 0, n_out = 0, depth = 0, device = 0, mode = 0;
    int n_groups = 0, per_group = 0, n_kchunks = 0, c_pad = 0;
    size_t hist_bytes = 0, in_bytes = 0;
    std::vector<float> b;                    // prototype taps (get_b)
    std::vector<double> fc;                  // requested centres
    std::vector<uint32_t> inc;               // quantised: f = inc * Fs / 2^32
    bool taps_dirty = true;
    unsigned long long n_abs = 0;            // absolute index of the next block's first sample
    unsigned long long step = 0;
    long long launches = 0;
    cudaStream_t st = nullptr;
    uint8_t* d_stage = nullptr;
    std::vector<float2*> d_out;
    std::vector<cudaEvent_t> ev_done;
    float2* d_g = nullptr; uint2* d_meta32 = nullptr;
    int8_t* d_bimg = nullptr; int* d_offs = nullptr; uint4* d_meta = nullptr;
    bool need_tile_factor = false;           // some channel's rotation does not repeat every CH_ROWS outputs
    float w0 = 0, w1 = 0, w2 = 0;
    float2* h_out = nullptr;                 // pinned mirror for the host entry point
};

namespace {

int chan_upload_taps(fmgpu_chan* h) {
    const int C = h->C, NN = h->NN;
    const double TWO_PI = 6.283185307179586476925286766559;
    // g_c[k] = b[k] * exp(+j 2 pi (inc_c (NN-1-k) mod 2^32) / 2^32), float64
    std::vector<double> gr((size_t)C * NN), gi((size_t)C * NN);
    double gmax = 0.0;
    for (int c = 0; c < C; c++)
        for (int k = 0; k < NN; k++) {
            const uint32_t ph = (uint32_t)((uint64_t)h->inc[c] * (uint64_t)(NN - 1 - k));
            const double th = TWO_PI * ((double)ph / 4294967296.0);
            gr[(size_t)c * NN + k] = (double)h->b[k] * std::cos(th);
            gi[(size_t)c * NN + k] = (double)h->b[k] * std::sin(th);
            gmax = std::max(gmax, std::max(std::fabs(gr[(size_t)c * NN + k]), std::fabs(gi[(size_t)c * NN + k])));
        }
    if (!(gmax > 0.0)) gmax = 1.0;
    // FP32 path: tap-major float2
    {
        std::vector<float2> g((size_t)NN * h->c_pad, make_float2(0.0f, 0.0f));
        for (int c = 0; c < C; c++)
            for (int k = 0; k < NN; k++) g[(size_t)k * h->c_pad + c] = make_float2((float)gr[(size_t)c * NN + k], (float)gi[(size_t)c * NN + k]);
        CHK(cudaMemcpy(h->d_g, g.data(), g.size() * sizeof(float2), cudaMemcpyHostToDevice));
        std::vector<uint2> m(C);
        for (int c = 0; c < C; c++) m[c] = make_uint2(h->inc[c], (uint32_t)((uint64_t)h->inc[c] * (uint64_t)h->D));
        CHK(cudaMemcpy(h->d_meta32, m.data(), m.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    }
    if (h->n_kchunks == 0) return FMGPU_OK;      // shape not eligible for the MMA kernel
    // MMA path: three balanced base-256 digit planes of round(G * 2^e), |value| < 127 * 65536
    const int e = (int)std::floor(std::log2(8.29e6 / gmax));
    const double S = std::ldexp(1.0, e);
    h->w0 = (float)(65536.0 / S); h->w1 = (float)(256.0 / S); h->w2 = (float)(1.0 / S);
    const int KB = 2 * NN;
    const size_t img_bytes = (size_t)h->n_groups * h->n_kchunks * CH_PLANES * CH_BSUB_BYTES;
    std::vector<int8_t> img(img_bytes, 0);
    std::vector<int> offs((size_t)h->n_groups * CH_PLANES * CH_NG, 0);
    std::vector<uint4> meta((size_t)h->n_groups * CH_SLOTS, make_uint4(0, 0, 0xffffffffu, 0));
    for (int c = 0; c < C; c++) {
        const int grp = c / h->per_group, slot = c % h->per_group;
        meta[(size_t)grp * CH_SLOTS + slot] = make_uint4(h->inc[c], (uint32_t)((uint64_t)h->inc[c] * (uint64_t)h->D), (uint32_t)c, 0);
        if (((uint32_t)((uint64_t)h->inc[c] * (uint64_t)h->D * (uint64_t)CH_ROWS) & 0x7fffffffu) != 0u) h->need_tile_factor = true;   // centre off the fs_out / 256 raster
        for (int K = 0; K < KB; K++) {
            const int k = K >> 1, comp = K & 1;
            // column 2 slot (re): gr * xr - gi * xi; column 2 slot + 1 (im): gi * xr + gr * xi
            const double v_re = comp == 0 ? gr[(size_t)c * NN + k] : -gi[(size_t)c * NN + k];
            const double v_im = comp == 0 ? gi[(size_t)c * NN + k] : gr[(size_t)c * NN + k];
            for (int col = 0; col < 2; col++) {
                long long v = std::llround((col == 0 ? v_re : v_im) * S);
                int d[3];
                d[2] = (int)(int8_t)(v & 0xff); v = (v - d[2]) >> 8;
                d[1] = (int)(int8_t)(v & 0xff); v = (v - d[1]) >> 8;
                d[0] = (int)v;                  // |d0| <= 127 by the choice of e
                const int n = 2 * slot + col, kc = K / CH_KCHUNK, kk = K % CH_KCHUNK;
                for (int j = 0; j < CH_PLANES; j++) {
                    const size_t base = (((size_t)grp * h->n_kchunks + kc) * CH_PLANES + j) * CH_BSUB_BYTES;
                    img[base + (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)(((kk >> 4) ^ (n & 7)) << 4) + (kk & 15)] = (int8_t)d[j];
                    offs[((size_t)grp * CH_PLANES + j) * CH_NG + n] += 127 * d[j];
                }
            }
        }
    }
    CHK(cudaMemcpy(h->d_bimg, img.data(), img.size(), cudaMemcpyHostToDevice));
    CHK(cudaMemcpy(h->d_offs, offs.data(), offs.size() * sizeof(int), cudaMemcpyHostToDevice));
    CHK(cudaMemcpy(h->d_meta, meta.data(), meta.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    return FMGPU_OK;
}

size_t chan_mma_smem(const fmgpu_chan* h) { return (size_t)h->n_kchunks * CH_PLANES * CH_BSUB_BYTES + 2 * CH_A_BYTES + 1024; }
size_t chan_mma2_smem(const fmgpu_chan* h) { return (size_t)h->n_kchunks * CH_PLANES * CH_BSUB_BYTES + CH2_NSTG * CH_A_BYTES + CH_SLOTS * CH_ROWS * sizeof(float2) + 1024; }

// staged buffer already holds history ++ block; writes slot, then rolls the history
int chan_run(fmgpu_chan* h, int slot) {
    if (h->taps_dirty) {
        CHK(cudaStreamSynchronize(h->st));
        const int rc = chan_upload_taps(h);
        if (rc != FMGPU_OK) return rc;
        h->taps_dirty = false;
    }
    const uint32_t n_newest0 = (uint32_t)(h->n_abs + (unsigned long long)h->D - 1ull);
    if (h->mode == FMGPU_CHAN_MODE_TENSOR) {
        ChanMmaParams p{};
        p.iq = h->d_stage; p.bimg = h->d_bimg; p.offs = h->d_offs; p.meta = h->d_meta; p.out = h->d_out[slot];
        p.n_newest0 = n_newest0; p.n_out = h->n_out; p.row_bytes = 2 * h->D; p.n_kchunks = h->n_kchunks;
        p.n_tiles = h->n_out / CH_ROWS; p.w0 = h->w0; p.w1 = h->w1; p.w2 = h->w2;
        int n_sm = 148;
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device);
        static const bool v1 = std::getenv("FMGPU_CHAN_V1") != nullptr;      // A/B aid: the synchronous first version
        if (v1 || chan_mma2_smem(h) > 227 * 1024) {           // (long prototypes: G no longer fits beside the stage ring)
            const int per_group = std::max(1, std::min(p.n_tiles, (2 * n_sm) / h->n_groups));
            chan_mma_i8<<<dim3(per_group, h->n_groups), 128, chan_mma_smem(h), h->st>>>(p);
        } else {
            const int per_group = std::max(1, std::min(p.n_tiles, n_sm / h->n_groups));
            if (h->need_tile_factor) chan_mma_i8_pipelined<true><<<dim3(per_group, h->n_groups), CH2_THREADS, chan_mma2_smem(h), h->st>>>(p);
            else                     chan_mma_i8_pipelined<false><<<dim3(per_group, h->n_groups), CH2_THREADS, chan_mma2_smem(h), h->st>>>(p);
        }
    } else {
        ChanFp32Params p{};
        p.iq = h->d_stage; p.g = h->d_g; p.meta = h->d_meta32; p.out = h->d_out[slot]; p.n_newest0 = n_newest0;
        p.n_out = h->n_out; p.D = h->D; p.NN = h->NN; p.C = h->C; p.c_pad = h->c_pad;
        const size_t smem = (size_t)(CF_TILE * h->D + h->NN - h->D) * sizeof(float2);
        chan_fir_fp32<<<dim3(h->n_out / CF_TILE, (h->C + 31) / 32), 128, smem, h->st>>>(p);
    }
    CHK(cudaGetLastError());
    h->launches++;
    // history for the next block = the last NN - D samples of (history ++ block)
    if (h->hist_bytes) CHK(cudaMemcpyAsync(h->d_stage, h->d_stage + h->in_bytes, h->hist_bytes, cudaMemcpyDeviceToDevice, h->st));
    CHK(cudaEventRecord(h->ev_done[slot], h->st));
    h->n_abs += (unsigned long long)h->n_out * (unsigned long long)h->D;
    h->step++;
    return FMGPU_OK;
}

void chan_free(fmgpu_chan* h) {
    if (h->st) cudaStreamSynchronize(h->st);
    auto F = [](void* p) { if (p) cudaFree(p); };
    F(h->d_stage); F(h->d_g); F(h->d_meta32); F(h->d_bimg); F(h->d_offs); F(h->d_meta);
    for (auto p : h->d_out) F(p);
    for (auto e : h->ev_done) if (e) cudaEventDestroy(e);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->st) cudaStreamDestroy(h->st);
}

} // namespace

extern "C" {

int fmgpu_chan_create(const fmgpu_chan_config* cfg, const double* centre_hz, fmgpu_chan** out) {
    if (!cfg || !centre_hz || !out) return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_create: null argument");
    *out = nullptr;
    const int D = cfg->decimation, NN = cfg->n_taps, C = cfg->n_channels, n_out = cfg->block_out;
    if (D < 1 || NN < D || NN > 1024 || C < 1 || !(cfg->fs_in_hz > 0.0))
        return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_create: need decimation >= 1, decimation <= n_taps <= 1024, n_channels >= 1, fs_in_hz > 0");
    if (n_out < CH_ROWS || n_out % CH_ROWS != 0)
        return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_create: block_out must be a positive multiple of 128");
    if ((size_t)n_out * D * 2 < (size_t)(NN - D) * 2)
        return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_create: block shorter than the filter history");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fmgpu_set_last_error_(FMGPU_ERR_CUDA, (std::string("chan_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e)).c_str());
    int dev = cfg->device;
    if (dev < 0) CHK(cudaGetDevice(&dev));
    if (dev >= n_dev) return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_create: bad device ordinal");
    CHK(cudaSetDevice(dev));
    cudaDeviceProp prop{};
    CHK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fmgpu_set_last_error_(FMGPU_ERR_CUDA, "chan_create: kernels are built for sm_100a only");
    // the tensor-core kernel needs 8-byte aligned windows and whole 128-byte K chunks
    const bool mma_ok = (D % 4 == 0) && ((2 * NN) % CH_KCHUNK == 0) && (2 * NN / CH_KCHUNK) <= 8;
    int mode = cfg->mode;
    if (mode == FMGPU_CHAN_MODE_AUTO) mode = mma_ok ? FMGPU_CHAN_MODE_TENSOR : FMGPU_CHAN_MODE_FP32;
    if (mode == FMGPU_CHAN_MODE_TENSOR && !mma_ok)
        return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_create: tensor mode needs decimation % 4 == 0 and n_taps % 64 == 0, n_taps <= 512");
    if (mode != FMGPU_CHAN_MODE_TENSOR && mode != FMGPU_CHAN_MODE_FP32) return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_create: unknown mode");

    auto* h = new fmgpu_chan();
    h->cfg = *cfg; h->cfg.device = dev; h->cfg.mode = mode;
    h->C = C; h->D = D; h->NN = NN; h->n_out = n_out; h->device = dev; h->mode = mode;
    h->depth = cfg->ring_depth > 0 ? cfg->ring_depth : 4;
    h->cfg.ring_depth = h->depth;
    h->n_groups = (C + CH_SLOTS - 1) / CH_SLOTS;
    h->per_group = (C + h->n_groups - 1) / h->n_groups;
    h->n_kchunks = mma_ok ? 2 * NN / CH_KCHUNK : 0;
    h->c_pad = (C + 31) / 32 * 32;
    h->hist_bytes = (size_t)(NN - D) * 2;
    h->in_bytes = (size_t)n_out * D * 2;
    h->fc.assign(centre_hz, centre_hz + C);
    h->inc.resize(C);
    for (int c = 0; c < C; c++) {
        const double turns = centre_hz[c] / cfg->fs_in_hz;                    // cycles per input sample
        h->inc[c] = (uint32_t)(long long)std::llround((turns - std::floor(turns)) * 4294967296.0);
    }
    h->b.assign(NN, 0.0f);
    const float k = cfg->cutoff_k > 0.0f ? cfg->cutoff_k : 0.95f / (float)D;  // the reference's ROLLOFF (broadcast_fm_demod.cpp:129)
    fmgpu_create_fir_lpf(h->b.data(), NN, k);
    int rc = FMGPU_OK;
    auto A = [&](cudaError_t err) { if (err != cudaSuccess && rc == FMGPU_OK) rc = fmgpu_set_last_error_(FMGPU_ERR_CUDA, cudaGetErrorString(err)); };
    A(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    A(cudaMalloc((void**)&h->d_stage, h->hist_bytes + h->in_bytes + 256));
    if (rc == FMGPU_OK) A(cudaMemset(h->d_stage, 127, h->hist_bytes + h->in_bytes + 256));   // empty history: x = 0 <-> u8 127
    h->d_out.assign(h->depth, nullptr); h->ev_done.assign(h->depth, nullptr);
    for (int i = 0; i < h->depth; i++) {
        A(cudaMalloc((void**)&h->d_out[i], (size_t)C * n_out * sizeof(float2)));
        A(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
    A(cudaMalloc((void**)&h->d_g, (size_t)NN * h->c_pad * sizeof(float2)));
    A(cudaMalloc((void**)&h->d_meta32, (size_t)C * sizeof(uint2)));
    if (h->n_kchunks) {
        A(cudaMalloc((void**)&h->d_bimg, (size_t)h->n_groups * h->n_kchunks * CH_PLANES * CH_BSUB_BYTES));
        A(cudaMalloc((void**)&h->d_offs, (size_t)h->n_groups * CH_PLANES * CH_NG * sizeof(int)));
        A(cudaMalloc((void**)&h->d_meta, (size_t)h->n_groups * CH_SLOTS * sizeof(uint4)));
        A(cudaFuncSetAttribute(chan_mma_i8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chan_mma_smem(h)));
        if (chan_mma2_smem(h) <= 227 * 1024)
        {
            A(cudaFuncSetAttribute(chan_mma_i8_pipelined<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chan_mma2_smem(h)));
            A(cudaFuncSetAttribute(chan_mma_i8_pipelined<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chan_mma2_smem(h)));
        }
    }
    {
        const size_t smem = (size_t)(CF_TILE * D + NN - D) * sizeof(float2);
        if (smem > 48 * 1024) A(cudaFuncSetAttribute(chan_fir_fp32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (rc != FMGPU_OK) { chan_free(h); delete h; return rc; }
    *out = h;
    return FMGPU_OK;
}

void fmgpu_chan_destroy(fmgpu_chan* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    chan_free(h);
    delete h;
}

float* fmgpu_chan_get_b(fmgpu_chan* h) { if (!h) return nullptr; h->taps_dirty = true; return h->b.data(); }

int fmgpu_chan_get_config(fmgpu_chan* h, fmgpu_chan_config* out) {
    if (!h || !out) return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_get_config: null argument");
    *out = h->cfg;
    return FMGPU_OK;
}

int fmgpu_chan_get_freqs(fmgpu_chan* h, double* hz, uint32_t* inc) {
    if (!h) return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_get_freqs: null handle");
    for (int c = 0; c < h->C; c++) {
        if (inc) inc[c] = h->inc[c];
        if (hz) {
            double f = (double)h->inc[c] / 4294967296.0;
            if (f >= 0.5) f -= 1.0;
            hz[c] = f * h->cfg.fs_in_hz;
        }
    }
    return FMGPU_OK;
}

int fmgpu_chan_enqueue_u8_device(fmgpu_chan* h, const uint8_t* iq_dev, float** out_dev) {
    if (!h || !iq_dev) return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_enqueue: null argument");
    CHK(cudaSetDevice(h->device));
    const int slot = (int)(h->step % (unsigned long long)h->depth);
    CHK(cudaMemcpyAsync(h->d_stage + h->hist_bytes, iq_dev, h->in_bytes, cudaMemcpyDeviceToDevice, h->st));
    const int rc = chan_run(h, slot);
    if (rc != FMGPU_OK) return rc;
    if (out_dev) *out_dev = (float*)h->d_out[slot];
    return FMGPU_OK;
}

int fmgpu_chan_process_u8(fmgpu_chan* h, const uint8_t* iq_host, size_t n_in_samples, float* out_host) {
    if (!h || !iq_host) return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_process: null argument");
    if (n_in_samples != (size_t)h->n_out * h->D) return fmgpu_set_last_error_(FMGPU_ERR_SIZE, "chan_process: n_in_samples != block_out * decimation");
    CHK(cudaSetDevice(h->device));
    const int slot = (int)(h->step % (unsigned long long)h->depth);
    CHK(cudaMemcpyAsync(h->d_stage + h->hist_bytes, iq_host, h->in_bytes, cudaMemcpyHostToDevice, h->st));
    const int rc = chan_run(h, slot);
    if (rc != FMGPU_OK) return rc;
    if (out_host) CHK(cudaMemcpyAsync(out_host, h->d_out[slot], (size_t)h->C * h->n_out * sizeof(float2), cudaMemcpyDeviceToHost, h->st));
    CHK(cudaStreamSynchronize(h->st));
    return FMGPU_OK;
}

int fmgpu_chan_sync(fmgpu_chan* h) {
    if (!h) return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_sync: null handle");
    CHK(cudaSetDevice(h->device));
    CHK(cudaStreamSynchronize(h->st));
    return FMGPU_OK;
}

int fmgpu_chan_wait_external_stream(fmgpu_chan* h, void* cuda_stream) {
    if (!h) return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_wait_external_stream: null handle");
    CHK(cudaSetDevice(h->device));
    cudaEvent_t ev;
    CHK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CHK(cudaEventRecord(ev, (cudaStream_t)cuda_stream));
    CHK(cudaStreamWaitEvent(h->st, ev, 0));
    CHK(cudaEventDestroy(ev));
    return FMGPU_OK;
}

void* fmgpu_chan_stream(fmgpu_chan* h) { return h ? (void*)h->st : nullptr; }
long long fmgpu_chan_launch_count(fmgpu_chan* h) { return h ? h->launches : 0; }

// One wideband block through channelizer + demodulators: the channelizer's output slot becomes the
// cf32 input of the demodulator handle (n_streams = n_channels, block_size = block_out).
int fmgpu_chan_feed_device(fmgpu_chan* h, fmgpu_demod* demod, const uint8_t* iq_dev) {
    if (!h || !demod || !iq_dev) return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_feed: null argument");
    fmgpu_config dc{};
    int rc = fmgpu_get_config(demod, &dc);
    if (rc != FMGPU_OK) return rc;
    if (dc.n_streams != h->C || dc.block_size != h->n_out || dc.pipeline_depth > h->depth)
        return fmgpu_set_last_error_(FMGPU_ERR_ARG, "chan_feed: demodulator must have n_streams = n_channels, block_size = block_out, pipeline_depth <= ring_depth");
    // the slot about to be overwritten was the input of the demodulator's block `depth` enqueues ago
    rc = fmgpu_stream_wait_input_free(demod, (void*)h->st);
    if (rc != FMGPU_OK) return rc;
    float* out = nullptr;
    rc = fmgpu_chan_enqueue_u8_device(h, iq_dev, &out);
    if (rc != FMGPU_OK) return rc;
    return fmgpu_enqueue_cf32_device(demod, out, (void*)h->st);
}

} // extern "C"
