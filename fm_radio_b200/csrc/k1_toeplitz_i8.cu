// K1 on the tensor cores: u8 IQ -> 64-tap /4 polyphase FIR -> polar discriminator -> fm_demod.
//
// Replaces the same reference code as k1_fir4_discrim.cu (file:line under /root/reference/src):
//   App::Run u8 -> cf32 unpack, (float)u8 - 127.0f           app.cpp:56-65
//   PolyphaseDownsampler<cf32>::process, M=4 K=16 NN=64      dsp/polyphase_filter.h:41-64,197-201
//   FM_Demod::Process, atan2 + wrapped first difference      fm_demod/fm_demod.cpp:30-45
//
// The FIR is the ONE stage of the chain whose input is integer (rtl-sdr bytes), and a decimating FIR over a byte
// stream is a block-Toeplitz contraction whose A operand is the raw capture itself:
//
//   row R of a stream = its bytes [128 R, 128 R + 128) = 64 IQ samples = the NEW input of outputs 16 R .. 16 R + 15
//   A[R, kappa]  = byte 128 (R - 1) + kappa,  kappa < 256        (rows R - 1 and R back to back: overlapping windows)
//   G[kappa, (j, c)] = b[(kappa - 8 j - 8) >> 1]  if (kappa & 1) == c and 0 <= kappa - 8 j - 8 < 128, else 0
//   Y[R, (j, c)] = sum_kappa A[R, kappa] G[kappa, (j, c)]  = re (c = 0) / im (c = 1) of output 16 R + j before "- 127"
//
// i.e. M = 128 rows (2048 outputs) x N = 32 columns x K = 256 bytes per tile, with G = the 64 taps laid out as a
// banded (Toeplitz) matrix.  Half of G is structural zeros, and the taps need three signed base-256 digit planes
// (24-bit fixed point, |error| <= 2^-26: as exact as the fp32 taps) -- 12x the algorithmic MACs -- but
// tcgen05.mma kind::i8 does 8192 MAC/clk/SM against the FMA pipe's 128 lanes: 8 MMAs of M128 N96 K32 = 384
// clocks per tile where the FFMA2 kernel spends ~4700.  The accumulators are int32 in tensor memory, so the
// sums are EXACT; "- 127" leaves as the integer constant 127 * sum(digits) per plane, and a silent capture
// (all bytes 127) gives exact zeros like the reference's (float)u8 - 127.0f (atan2(0, 0) = 0, and the all-127
// block that turns the reference's AGC into NaN for good).  The planes are recombined in fp32 by two FFMA2.
// With the FIR off the FMA pipe the kernel is bound by HBM (2 B in + 1 B out per IQ sample).
//
// Shared memory holds the tile ONCE (129 rows of 128 bytes, 128-byte swizzle, cp.async 16-byte copies): the
// K-chunk kappa < 128 of A is rows 0..127 of that buffer, the chunk kappa >= 128 is THE SAME buffer one row
// later (descriptor start address + 128 bytes; the swizzle is a function of the absolute shared-memory address,
// so the shifted view stays consistent).  The FIR history of a stream is its previous block's last row (128
// bytes; a new stream starts from bytes 127 = the reference's zero history).
//
// One persistent CTA (256 threads; 128 TMEM lanes = 128 rows, two warpgroups taking outputs 0..7 and 8..15 of
// every row) walks tiles round-robin through a ring of four tile buffers, software pipelined:
//   iteration i:  wait for tile i's bytes | thread 0 issues tile i's 8 MMAs into TMEM buffer i & 1
//                 | cp.async of tile i + 3 | epilogue of tile i - 1 from TMEM buffer (i - 1) & 1
// so up to three tiles of loads (the kernel is HBM-latency bound with fewer: 0.078 ms with one tile ahead), the
// MMAs and the discriminator epilogue of consecutive tiles overlap; two CTAs per SM.
// Epilogue per thread = half a row = 8 outputs: tcgen05.ld, integer offset + magic-number int->float (exact),
// plane recombination, minimax atan2 (fm_common.cuh), wrapped first difference, 2 x STG.128.
// The angle of the output before a thread's first comes through shared memory (same row's other half / previous
// row); for the tile's first row it is the last output of the previous row, which depends on buffer row 0 only:
// warp 4 computes its exact integer sums with dp4a (same digits, same recombination => the same bits as the tile
// that owns that output).  Across blocks the discriminator's prev_theta is carried as state, like the reference.
#include "fm_common.cuh"
#include "tcgen05.cuh"
#include <atomic>
#include <cmath>
#include <cstring>
#include <vector>

namespace fm {

constexpr int K1T_ROWS = 128;                        // A rows per tile = TMEM lanes
constexpr int K1T_THREADS = 256;                     // two warpgroups: outputs 0..7 and 8..15 of every row
constexpr int K1T_NCOL = 32;                         // 16 outputs x (re, im)
constexpr int K1T_PLANES = 3;
constexpr int K1T_N = K1T_NCOL * K1T_PLANES;         // 96 accumulator columns
constexpr int K1T_ABUF = 17 * 1024;                  // 129 rows x 128 B, padded to the 1024-byte swizzle atom
constexpr int K1T_BCHUNK = K1T_N * 128;              // 12 KB: one 128-byte K chunk of G
constexpr int K1T_BBYTES = 2 * K1T_BCHUNK;
constexpr float K1T_MAGIC = 12582912.0f;             // 1.5 * 2^23: as_float(0x4B400000 + v) == MAGIC + v for |v| < 2^22
// Two shapes of the pipeline (template parameters, measured against each other, launch_k1t picks by p.shape):
//   NS = 4, NACC = 2, 2 CTAs/SM  four-deep tile ring, two accumulator buffers: the MMAs of tile i overlap the epilogue of i - 1
//   NS = 2, NACC = 1, 3 CTAs/SM  two-deep ring, one accumulator buffer: no overlap inside a CTA, but 24 instead of 16 warps per SM
constexpr int k1t_smem(int NS) { return K1T_BBYTES + NS * K1T_ABUF + 1024; }

__device__ __forceinline__ int dp4a_u8s8(uint32_t a_u8x4, int b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}

// exact plane sums (integers) -> sum_k b[k] (u - 127): the same instruction sequence everywhere it is used
__device__ __forceinline__ float2 k1t_combine(uint32_t r0, uint32_t i0, uint32_t r1, uint32_t i1, uint32_t r2, uint32_t i2,
                                              const uint32_t (&K)[3], const float2 (&W)[3]) {
    const float2 nm = make_float2(-K1T_MAGIC, -K1T_MAGIC);
    const float2 f0 = __fadd2_rn(make_float2(__uint_as_float(r0 + K[0]), __uint_as_float(i0 + K[0])), nm);
    const float2 f1 = __fadd2_rn(make_float2(__uint_as_float(r1 + K[1]), __uint_as_float(i1 + K[1])), nm);
    const float2 f2 = __fadd2_rn(make_float2(__uint_as_float(r2 + K[2]), __uint_as_float(i2 + K[2])), nm);
    return __ffma2_rn(f0, W[0], __ffma2_rn(f1, W[1], __fmul2_rn(f2, W[2])));
}

// fm_demod.cpp:6-10 and :36-44 for two outputs: wrap(theta - prev) * gain.  The difference of two angles lies in (-2 pi, 2 pi), so
// the reference's two compares (x >= pi -> x - 2 pi, x <= -pi -> x + 2 pi) are x - 2 pi * rint(x / 2 pi): four packed
// instructions per pair instead of eleven scalar ones.  Same value (one rounding of x -+ 2 pi) whenever both take the same
// branch; they can differ only for |x| within an ulp of pi -- a phase step of half a turn per sample, i.e. noise, never FM.
__device__ __forceinline__ float2 k1t_discrim2(float2 d, float gain) {
    const float2 big = make_float2(12582912.0f, 12582912.0f);
    const float2 k = __fadd2_rn(__ffma2_rn(d, make_float2(INV_TWO_PI_F, INV_TWO_PI_F), big), make_float2(-12582912.0f, -12582912.0f));
    const float2 w = __ffma2_rn(k, make_float2(-TWO_PI_F, -TWO_PI_F), d);
    return __fmul2_rn(w, make_float2(gain, gain));
}

template <int K1T_NS, int NACC>
__global__ void __launch_bounds__(K1T_THREADS, NACC == 2 ? 2 : 3)
k1_toeplitz_i8(const uint8_t* __restrict__ iq, const uint8_t* __restrict__ hist_in, uint8_t* __restrict__ hist_out,
               float2* __restrict__ hist_f32_out, float* __restrict__ fm_demod, const __grid_constant__ K1TParams p)
{
    extern __shared__ uint8_t k1t_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)k1t_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sB = smem;
    uint8_t* sA = sB + K1T_BBYTES;                   // [K1T_NS][K1T_ABUF]
    __shared__ __align__(8) uint64_t bar_acc[2];     // the MMAs of the tile in TMEM buffer b have completed
    __shared__ uint32_t s_tmem;
    __shared__ float s_mid[2][K1T_ROWS];             // angle of output 7 of every row (warpgroup 0 -> 1)
    __shared__ float s_last[2][K1T_ROWS];            // angle of output 15 of every row (warpgroup 1 -> 0, next row)
    __shared__ float s_thprev[2];                    // angle of the output before the tile

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid & (K1T_ROWS - 1), half = tid >> 7;         // this thread's A row and which 8 of its 16 outputs
    for (int i = tid; i < K1T_BBYTES / 16; i += K1T_THREADS) ((uint4*)sB)[i] = __ldg((const uint4*)p.bimg + i);
    if (tid == 0) { tc::mbar_init(&bar_acc[0], 1); tc::mbar_init(&bar_acc[1], 1); tc::mbar_init_fence(); }
    constexpr uint32_t K1T_TMEM_COLS = 128 * NACC;       // accumulator buffers of 128 columns (96 used)
    if (warp == 0) tc::tmem_alloc(&s_tmem, K1T_TMEM_COLS);
    int pw[6] = { 0, 0, 0, 0, 0, 0 };
    if (warp == 4) {
#pragma unroll
        for (int q = 0; q < 6; q++) pw[q] = p.ptab[lane * 6 + q];
    }
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t sA_addr = tc::smem_u32(sA), sB_addr = tc::smem_u32(sB);
    const size_t stream_bytes = (size_t)p.n_rows * 128;
    const uint32_t Kc[3] = { 0x4B400000u - (uint32_t)p.off[0], 0x4B400000u - (uint32_t)p.off[1], 0x4B400000u - (uint32_t)p.off[2] };
    const float2 Wc[3] = { make_float2(p.w[0], p.w[0]), make_float2(p.w[1], p.w[1]), make_float2(p.w[2], p.w[2]) };
    constexpr uint32_t IDESC = tc::idesc_i8_u8s8(K1T_ROWS, K1T_N);

    // Tiles are walked without divisions: (stream, first row) of tile t + gridDim.x follows from those of tile t.
    struct TileRef { int tile, s, row0; };
    const int g_div = (int)gridDim.x / p.tiles_per_stream, g_mod = (int)gridDim.x % p.tiles_per_stream;
    auto tile_ref = [&](int tile) { TileRef r; r.tile = tile; r.s = tile / p.tiles_per_stream; r.row0 = (tile - r.s * p.tiles_per_stream) * K1T_ROWS; return r; };
    auto advance = [&](TileRef& r) {
        r.tile += (int)gridDim.x; r.s += g_div; r.row0 += g_mod * K1T_ROWS;
        if (r.row0 >= p.tiles_per_stream * K1T_ROWS) { r.row0 -= p.tiles_per_stream * K1T_ROWS; r.s++; }
    };
    // buffer row q of a tile <-> data row row0 - 1 + q of its stream (q = 0: the halo = previous row / history).
    // Thread t copies the 16-byte chunks t, t + 256, ...: chunk (t & 7) of buffer rows (t >> 3) + 32 k, whose swizzle
    // key (row & 7) does not depend on k, so source and destination both advance by 4096 bytes per copy.
    // Always commits a group (an empty one past the last tile), so the group count per iteration is uniform.
    const int st_q0 = tid >> 3, st_c = tid & 7;
    const uint32_t st_dst0 = (uint32_t)(st_q0 * 128 + ((st_c ^ (st_q0 & 7)) << 4));
    auto stage = [&](const TileRef& r, int slot) {
        if (r.tile < p.n_tiles) {
            const int n_valid = min(K1T_ROWS, p.n_rows - r.row0);
            const uint8_t* src = iq + (size_t)r.s * stream_bytes + (ptrdiff_t)(r.row0 - 1) * 128 + tid * 16;
            uint8_t* dst = sA + slot * K1T_ABUF + st_dst0;
            if (r.row0 == 0 && tid < 8) tc::cp_async16(dst, hist_in + (size_t)r.s * 128 + st_c * 16);
            else if (st_q0 <= n_valid) tc::cp_async16(dst, src);
#pragma unroll
            for (int k = 1; k < 5; k++)
                if (st_q0 + 32 * k <= n_valid) tc::cp_async16(dst + 4096 * k, src + 4096 * k);
        }
        tc::cp_async_commit();
    };

    // everything of a tile that reads its bytes from shared memory (before the ring slot is recycled)
    auto pre_epilogue = [&](const TileRef& r, int slot, int buf) {
        const int s = r.s, row0 = r.row0;
        const int n_valid = min(K1T_ROWS, p.n_rows - row0);
        const uint8_t* A = sA + slot * K1T_ABUF;
        if (warp == 4) {
            if (row0 == 0) {
                // a block's first output: the discriminator's carried prev_theta (fm_demod.cpp:41-44; 0 for a new stream)
                if (lane == 0) s_thprev[buf] = p.theta_in[s];
            } else {
                // the output before the tile = output 15 of buffer row 0: all 64 taps on that row's 128 bytes
                const uint32_t w4 = *(const uint32_t*)(A + 4 * lane);            // row 0: swizzle key 0
                int a[6];
#pragma unroll
                for (int q = 0; q < 6; q++) a[q] = __reduce_add_sync(0xffffffffu, dp4a_u8s8(w4, pw[q], 0));
                if (lane == 0) {
                    const float2 z = k1t_combine((uint32_t)a[0], (uint32_t)a[1], (uint32_t)a[2], (uint32_t)a[3], (uint32_t)a[4], (uint32_t)a[5], Kc, Wc);
                    s_thprev[buf] = fm_atan2f(z.y, z.x);
                }
            }
        } else if (warp >= 5 && row0 + n_valid == p.n_rows) {
            // history for the next block: the stream's last row, as bytes (this kernel) and as floats (cf32 kernel)
            const uint8_t* rowp = A + n_valid * 128;
            const int key = n_valid & 7;
            if (warp == 7 && lane < 8)
                *(uint4*)(hist_out + (size_t)s * 128 + lane * 16) = *(const uint4*)(rowp + ((lane ^ key) << 4));
            if (warp <= 6) {
                const int t0 = (warp - 5) * 32 + lane;                           // 64 samples
                const uint8_t* b = rowp + (((t0 >> 3) ^ key) << 4) + (t0 & 7) * 2;
                hist_f32_out[(size_t)s * K1_HIST + t0] = make_float2((float)b[0] - 127.0f, (float)b[1] - 127.0f);
            }
        }
    };

    // after_sync runs right after the epilogue's CTA barrier, before the stores (NACC == 1: the next-but-one tile's copies)
    auto epilogue = [&](const TileRef& r, int buf, int par, auto&& after_sync) {
        const int s = r.s, row0 = r.row0;
        const int n_valid = min(K1T_ROWS, p.n_rows - row0);
        const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(buf * 128 + 16 * half);
        const size_t o_base = (size_t)s * p.n_rows * 16 + (size_t)(row0 + row) * 16 + 8 * half;
        uint32_t a0[16], a1[16], a2[16];
        tc::tmem_ld16(lane_addr + 0 * K1T_NCOL, a0);
        tc::tmem_ld16(lane_addr + 1 * K1T_NCOL, a1);
        tc::tmem_ld16(lane_addr + 2 * K1T_NCOL, a2);
        tc::tmem_ld_wait();
        tc::fence_before();                                                      // this thread's TMEM reads are done
        float2 z[8];
#pragma unroll
        for (int j = 0; j < 8; j++) z[j] = k1t_combine(a0[2 * j], a0[2 * j + 1], a1[2 * j], a1[2 * j + 1], a2[2 * j], a2[2 * j + 1], Kc, Wc);
        if (p.dbg_fm_in && row < n_valid) {
            float4* d4 = (float4*)(p.dbg_fm_in + o_base);
#pragma unroll
            for (int j = 0; j < 8; j += 2) d4[j >> 1] = make_float4(z[j].x, z[j].y, z[j + 1].x, z[j + 1].y);
        }
        float2 th[4];
#pragma unroll
        for (int j = 0; j < 4; j++) th[j] = fm_atan2f_x2(z[2 * j].y, z[2 * j].x, z[2 * j + 1].y, z[2 * j + 1].x);
        // the angle before this thread's first output: output 7 of the same row (other warpgroup) or output 15 of the
        // previous row / of the row before the tile
        if (half == 0) s_mid[par][row] = th[3].y; else s_last[par][row] = th[3].y;
        __syncthreads();
        after_sync();
        const float th_before = half ? s_mid[par][row] : (row == 0 ? s_thprev[buf] : s_last[par][row - 1]);
        if (row >= n_valid) return;
        if (half == 1 && row0 + row == p.n_rows - 1) p.theta_out[s] = th[3].y;     // the block's last angle -> next block
        const float2 d0 = k1t_discrim2(make_float2(th[0].x - th_before, th[0].y - th[0].x), p.discrim_gain);
        const float2 d1 = k1t_discrim2(make_float2(th[1].x - th[0].y, th[1].y - th[1].x), p.discrim_gain);
        const float2 d2 = k1t_discrim2(make_float2(th[2].x - th[1].y, th[2].y - th[2].x), p.discrim_gain);
        const float2 d3 = k1t_discrim2(make_float2(th[3].x - th[2].y, th[3].y - th[3].x), p.discrim_gain);
        float4* dst = (float4*)(fm_demod + o_base);
        dst[0] = make_float4(d0.x, d0.y, d1.x, d1.y);
        dst[1] = make_float4(d2.x, d2.y, d3.x, d3.y);
    };

    // NACC == 2: tile it + NS - 1 is staged in iteration it, into the slot of tile it - 1.
    // NACC == 1: the slot of tile `it` is dead as soon as its MMAs have completed and every warp has left pre_epilogue, i.e.
    // right after the epilogue's CTA barrier -- tile it + NS is staged THERE, so its bytes have the rest of this epilogue plus
    // a whole iteration to arrive (staged at the top of the next iteration they had half of that, and ncu's source view
    // showed 18 % of the kernel's warp time waiting for them).
    constexpr int PRE = NACC == 2 ? K1T_NS - 1 : K1T_NS;                          // tiles staged ahead of the loop
    TileRef cur = tile_ref((int)blockIdx.x), pre = cur, prev = cur;
#pragma unroll
    for (int k = 0; k < PRE; k++) { stage(pre, k); advance(pre); }
    const bool mma_warp = __shfl_sync(0xffffffffu, warp, 0) == 0;                // warp-uniform
    int it = 0;
    for (; cur.tile < p.n_tiles; it++) {
        const int buf = NACC == 2 ? (it & 1) : 0, slot = it % K1T_NS;
        tc::cp_async_wait<PRE - 1>();                // all but the newest PRE - 1 groups have landed: tile `it` is in place
        tc::fence_async_smem();
        __syncthreads();                             // ... for every thread's copies; everyone has left iteration it - 1
        if (mma_warp && tc::elect_one()) {
            tc::fence_after();
            const uint32_t a_base = sA_addr + slot * K1T_ABUF;
#pragma unroll
            for (int kc = 0; kc < 2; kc++)
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    // K chunk 1 = the same buffer one row (128 B) later; the 128-byte swizzle follows the absolute address
                    const uint64_t a_desc = tc::smem_desc_sw128(a_base + kc * 128 + ks * 32);
                    const uint64_t b_desc = tc::smem_desc_sw128(sB_addr + kc * K1T_BCHUNK + ks * 32);
                    tc::mma_i8(tmem + (uint32_t)(buf * 128), a_desc, b_desc, IDESC, (kc | ks) != 0 ? 1u : 0u);
                }
            tc::commit(&bar_acc[buf]);
        }
        pre_epilogue(cur, slot, buf);
        if (NACC == 2) {
            if (it >= 1) {
                tc::mbar_wait(&bar_acc[buf ^ 1], (uint32_t)(((it - 1) >> 1) & 1));   // tile it - 1: accumulators complete, its bytes dead
                tc::fence_after();
            }
            stage(pre, (it + K1T_NS - 1) % K1T_NS);  // tile it + NS - 1 into the slot of tile it - 1
            advance(pre);
            if (it >= 1) epilogue(prev, buf ^ 1, it & 1, [] {});
        } else {
            tc::mbar_wait(&bar_acc[0], (uint32_t)(it & 1));
            tc::fence_after();
            epilogue(cur, 0, it & 1, [&] { stage(pre, slot); advance(pre); });   // tile it + NS into this tile's slot
        }
        prev = cur;
        advance(cur);
    }
    if (NACC == 2 && it >= 1) {
        tc::mbar_wait(&bar_acc[(it - 1) & 1], (uint32_t)(((it - 1) >> 1) & 1));
        tc::fence_after();
        epilogue(prev, (it - 1) & 1, it & 1, [] {});
    }
    tc::cp_async_wait<0>();
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, K1T_TMEM_COLS);
}

// ---- host side: digit planes of the taps -> shared-memory image of G, dp4a table, constants ----
void k1t_build_tables(const float* taps, std::vector<int8_t>& bimg, std::vector<int>& ptab, int off[3], float w[3])
{
    double gmax = 0.0;
    for (int k = 0; k < K1_NN; k++) gmax = std::max(gmax, (double)std::fabs(taps[k]));
    if (!(gmax > 0.0)) gmax = 1.0;
    // three balanced base-256 digits of round(b * 2^e), |value| < 127 * 65536
    const int e = (int)std::floor(std::log2(8.29e6 / gmax));
    const double S = std::ldexp(1.0, e);
    w[0] = (float)(65536.0 / S); w[1] = (float)(256.0 / S); w[2] = (float)(1.0 / S);
    int d[K1_NN][3];
    for (int k = 0; k < K1_NN; k++) {
        long long v = std::llround((double)taps[k] * S);
        d[k][2] = (int)(int8_t)(v & 0xff); v = (v - d[k][2]) >> 8;
        d[k][1] = (int)(int8_t)(v & 0xff); v = (v - d[k][1]) >> 8;
        d[k][0] = (int)v;
    }
    for (int pl = 0; pl < 3; pl++) { off[pl] = 0; for (int k = 0; k < K1_NN; k++) off[pl] += 127 * d[k][pl]; }
    bimg.assign(K1T_BBYTES, 0);
    for (int kappa = 0; kappa < 256; kappa++)
        for (int col = 0; col < K1T_NCOL; col++) {
            const int j = col >> 1, c = col & 1, rel = kappa - 8 * j - 8;
            if ((kappa & 1) != c || rel < 0 || rel >= 128) continue;
            const int k = rel >> 1, kc = kappa >> 7, kk = kappa & 127;
            for (int pl = 0; pl < 3; pl++) {
                const int n = pl * K1T_NCOL + col;
                bimg[(size_t)kc * K1T_BCHUNK + (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 + (size_t)(((kk >> 4) ^ (n & 7)) << 4) + (kk & 15)] = (int8_t)d[k][pl];
            }
        }
    ptab.assign(32 * 6, 0);
    for (int l = 0; l < 32; l++)
        for (int pl = 0; pl < 3; pl++) {
            const uint32_t lo = (uint32_t)(uint8_t)(int8_t)d[2 * l][pl], hi = (uint32_t)(uint8_t)(int8_t)d[2 * l + 1][pl];
            ptab[l * 6 + 2 * pl + 0] = (int)(lo | (hi << 16));               // bytes (I, Q, I, Q) x (d, 0, d', 0)
            ptab[l * 6 + 2 * pl + 1] = (int)((lo << 8) | (hi << 24));        //                   x (0, d, 0, d')
        }
}

cudaError_t launch_k1t(const uint8_t* iq, const uint8_t* hist_in, uint8_t* hist_out, float2* hist_f32_out, float* fm_demod,
                       const K1TParams& p, int n_ctas, cudaStream_t st)
{
    const int grid = std::max(1, std::min(p.n_tiles, n_ctas));
    static std::atomic<bool> configured[64];             // the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(k1_toeplitz_i8<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, k1t_smem(4));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k1_toeplitz_i8<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, k1t_smem(2));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    if (p.shape == 1) {
        const int grid3 = std::max(1, std::min(p.n_tiles, n_ctas * 3 / 2));
        k1_toeplitz_i8<2, 1><<<grid3, K1T_THREADS, k1t_smem(2), st>>>(iq, hist_in, hist_out, hist_f32_out, fm_demod, p);
    } else {
        k1_toeplitz_i8<4, 2><<<grid, K1T_THREADS, k1t_smem(4), st>>>(iq, hist_in, hist_out, hist_f32_out, fm_demod, p);
    }
    return cudaGetLastError();
}

} // namespace fm
