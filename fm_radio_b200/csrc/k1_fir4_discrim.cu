// K1: u8 (or cf32) baseband IQ -> 64-tap /4 polyphase FIR -> polar discriminator -> fm_demod.
//
// Replaces, fused in one pass (reference file:line under /root/reference/src):
//   App::Run u8 -> cf32 unpack, (float)u8 - 127.0f           app.cpp:56-65
//   PolyphaseDownsampler<cf32>::process, M=4 K=16 NN=64      dsp/polyphase_filter.h:41-64,197-201
//     y[i] = sum_k b[k] * x[(i+1)*4 - 64 + k]                (c32_f32_cum_mul.cpp:70-111 on AVX)
//   FM_Demod::Process, atan2 + wrapped first difference      fm_demod/fm_demod.cpp:30-45
//
// Two kernels:
//   k1_fir4_discrim_u8   the production path (rtl-sdr bytes).  FP32-FMA-pipe bound, 64 FLOP per
//                        input IQ sample; design notes at the kernel.
//   k1_fir4_discrim_cf32 first version (fp32 staging, scalar FFMA), kept for the cf32 entry point
//                        (Process(span<cf32>), one stream, GUI use) where the input is 4x larger and speed is irrelevant.
// The previous block's last 64 samples are the only state (ping-pong buffers, because the CTA
// that reads the history is not the CTA that writes it); the discriminator's prev_theta is
// recomputed from that history instead of being stored.
#include "fm_common.cuh"
#include <atomic>

namespace fm {

__device__ __forceinline__ float wrap_phase(float x) {          // fm_demod.cpp:6-10, as two selects
    const float lo = x - 2.0f * PI_F, hi = x + 2.0f * PI_F;
    return (x >= PI_F) ? lo : ((x <= -PI_F) ? hi : x);
}

__global__ void __launch_bounds__(K1_THREADS, 3)
k1_fir4_discrim_cf32(const float2* __restrict__ iq, const float2* __restrict__ hist_in,
                float2* __restrict__ hist_out, float* __restrict__ fm_demod,
                const __grid_constant__ K1Params p)
{
    extern __shared__ __align__(16) float smem[];
    const int t = threadIdx.x;
    const int tile = blockIdx.x;
    const int s = blockIdx.y;
    const int o0 = tile * K1_TILE;
    const int n_valid = min(K1_TILE, p.n_out - o0);      // outputs of this tile (multiple of 16)
    const int n_frames = n_valid + 16;                   // frames to stage, frame f <-> input frame o0-16+f
    const size_t n_in = (size_t)p.n_out * K1_M;          // IQ samples per stream per block

    // ---- stage the tile: frame f lives at smem[(f>>4)*K1_SEG + (f&15)*8 .. +8) ----
    const int f_first = (tile == 0) ? 16 : 0;            // tile 0 takes its halo from the history
    if (tile == 0 && t < K1_HIST) {
        const float2 v = hist_in[(size_t)s * K1_HIST + t];
        const int f = t >> 2;
        float* d = smem + (f >> 4) * K1_SEG + (f & 15) * 8 + (t & 3) * 2;
        d[0] = v.x; d[1] = v.y;
    }
    // 16 bytes = 2 IQ samples = half a frame
    const float4* src = (const float4*)(iq + (size_t)s * n_in + (ptrdiff_t)(o0 - 16) * K1_M);
    for (int hf = f_first * 2 + t; hf < n_frames * 2; hf += K1_THREADS) {
        const float4 v = __ldg(src + hf);
        const int f = hf >> 1;
        *(float4*)(smem + (f >> 4) * K1_SEG + (f & 15) * 8 + (hf & 1) * 4) = v;
    }
    __syncthreads();

    // ---- history for the next block: the last 64 IQ samples of this stream's block ----
    if (o0 + n_valid == p.n_out && t < K1_HIST) {
        const int f = n_valid + (t >> 2);
        const float* d = smem + (f >> 4) * K1_SEG + (f & 15) * 8 + (t & 3) * 2;
        hist_out[(size_t)s * K1_HIST + t] = make_float2(d[0], d[1]);
    }

    if (t * K1_R >= n_valid) return;

    // ---- 17 outputs per thread: r = -1 (previous thread's last output, for the difference) .. 15 ----
    // output o = 16t + r uses frames f = o+1 .. o+16 of the staged tile with tap group f-o-1.
    float ar[K1_R + 1], ai[K1_R + 1];
#pragma unroll
    for (int r = 0; r <= K1_R; r++) { ar[r] = 0.0f; ai[r] = 0.0f; }
    const float* base = smem + t * K1_SEG;
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const float* fp = base + ((j < 16) ? j * 8 : K1_SEG + (j - 16) * 8);
        const float4 a = *(const float4*)fp;
        const float4 b = *(const float4*)(fp + 4);
#pragma unroll
        for (int r = -1; r < K1_R; r++) {
            const int g = j - r - 1;                    // tap group, static after unrolling
            if (g >= 0 && g < 16) {
                ar[r + 1] = fmaf(a.x, p.taps[4 * g + 0], ar[r + 1]); ai[r + 1] = fmaf(a.y, p.taps[4 * g + 0], ai[r + 1]);
                ar[r + 1] = fmaf(a.z, p.taps[4 * g + 1], ar[r + 1]); ai[r + 1] = fmaf(a.w, p.taps[4 * g + 1], ai[r + 1]);
                ar[r + 1] = fmaf(b.x, p.taps[4 * g + 2], ar[r + 1]); ai[r + 1] = fmaf(b.y, p.taps[4 * g + 2], ai[r + 1]);
                ar[r + 1] = fmaf(b.z, p.taps[4 * g + 3], ar[r + 1]); ai[r + 1] = fmaf(b.w, p.taps[4 * g + 3], ai[r + 1]);
            }
        }
    }

    if (p.dbg_fm_in) {
        float2* d2 = p.dbg_fm_in + (size_t)s * p.n_out + o0 + t * K1_R;
#pragma unroll
        for (int r = 0; r < K1_R; r++) d2[r] = make_float2(ar[r + 1], ai[r + 1]);
    }
    // ---- discriminator epilogue (fm_demod.cpp:36-44) ----
    float prev = fm_atan2f(ai[0], ar[0]);
    float out[K1_R];
#pragma unroll
    for (int r = 0; r < K1_R; r++) {
        const float th = fm_atan2f(ai[r + 1], ar[r + 1]);
        out[r] = wrap_phase(th - prev) * p.discrim_gain;
        prev = th;
    }
    float4* dst = (float4*)(fm_demod + (size_t)s * p.n_out + o0 + t * K1_R);
#pragma unroll
    for (int q = 0; q < K1_R / 4; q++) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
}


// ---------------------------------------------------------------------------------------------
// Production u8 kernel.
//   * one CTA = 2048 consecutive outputs of one stream = 8192 IQ samples + 64 samples of halo,
//     staged in shared memory AS BYTES (16.5 KB) with 16-byte cp.async copies (LDGSTS.128): with
//     the fp32 staging of the first version the tile took 68 KB, 3 CTAs/SM, and ncu showed only
//     51 % of the issue slots in use (long-scoreboard stalls of the load phase, 12 warps/SM);
//     bytes let ~7 CTAs/SM co-reside so one CTA's load phase hides under the others' FFMAs;
//   * rows of 128 B (= 16 four-sample frames = the inputs of 16 outputs) are XOR-swizzled at
//     16-byte granularity (chunk c of row r lives at chunk c ^ (r & 7)), so the per-thread
//     window reads (thread t walks rows t and t+1) are bank-conflict free LDS.128;
//   * each thread owns 16 consecutive outputs: 32 register accumulators; per LDS.128 (two frames)
//     it unpacks 16 bytes with PRMT + FADD (exact (float)u8 - 127, no I2F) and issues 256 FFMAs;
//   * the 64 taps ride in the kernel parameter block (__grid_constant__, constant bank 0) and the
//     tap loop is fully unrolled, so every FFMA takes its tap as a constant operand;
//   * discriminator: own minimax atan2 (8-term odd polynomial, 1.2e-7 rad max error in fp32, one
//     MUFU.RCP) instead of libdevice's atan2f (about half the instructions); the previous
//     output's angle comes from the neighbouring lane (SHFL) / warp (shared), and only thread 0
//     recomputes the output before the tile (bit-identically to the CTA that owns it).
// ---------------------------------------------------------------------------------------------
constexpr int K1U_ROWS = K1_THREADS + 1;                 // 129 rows of 128 bytes
constexpr int K1U_SMEM = K1U_ROWS * 128;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(d), "l"(gmem_src));
}

// two frames (8 IQ samples) from one 16-byte chunk, applied to every output they feed.
// One (re, im) pair per 64-bit register pair.  The unpack is 2 PRMT per IQ sample and NOTHING on the
// FMA pipe: each byte is isolated into an otherwise zero word, i.e. the fp32 DENORMAL u * 2^-149,
// which the FMA datapath takes at full rate and multiplies exactly; the taps ride pre-scaled by 2^100
// (K1Params::taps_s), so every FFMA2 accumulates tap * u * 2^-49 with one IEEE rounding -- the same
// rounding, up to the power-of-two scale, as fma(tap, (float)u, acc).  The reference's "- 127.0f"
// (app.cpp:59-60) is linear and leaves once per OUTPUT instead of once per input sample:
// sum tap (u - 127) = sum tap u - 127 sum tap, applied together with the 2^49 rescale as ONE FFMA2 in
// the epilogue (k1u_finish).  Versus PRMT + FADD2 per sample this removes 128 of the 1152 FMA-pipe
// instructions a thread issued per 16 outputs.  Every tap costs ONE FFMA2 (re and im together) whose
// tap operand is a uniform-register scalar (SASS: FFMA2 R, R.F32x2.HI_LO, UR.F32, R.F32x2.HI_LO).
template <int J0>   // J0 = index of the chunk's first frame relative to the thread's window (even, 0 .. 30)
__device__ __forceinline__ void k1u_chunk(const uint4 w, float2 (&acc)[K1_R], float2& acc_prev, const bool warp0, const K1Params& p) {
    const uint32_t ws[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int j = J0 + h;                           // frame index: output r uses frames r+1 .. r+16
        float2 x[4];
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const uint32_t v = ws[2 * h + (m >> 1)];
            const uint32_t sel = 0x4440u | (uint32_t)((m & 1) * 2);
            x[m] = make_float2(__uint_as_float(__byte_perm(v, 0u, sel)), __uint_as_float(__byte_perm(v, 0u, sel + 1u)));
        }
#pragma unroll
        for (int r = 0; r < K1_R; r++) {
            const int g = j - r - 1;                    // tap group, static after unrolling
            if (g >= 0 && g < 16) {
#pragma unroll
                for (int m = 0; m < 4; m++) {
                    const float tap = p.taps_s[4 * g + m];
                    acc[r] = __ffma2_rn(x[m], make_float2(tap, tap), acc[r]);
                }
            }
        }
        // output r = -1 (the one before the thread's first): frames 0 .. 15 with tap group = frame index.  Only
        // thread 0 of the tile needs it (its predecessor lives in another CTA) -- as a 17th accumulator of warp 0
        // (a warp-uniform condition) it rides in the main loop's instruction stream, where the 64-deep dependent
        // chain hides completely; it used to be a single-lane pass after the loop that kept the tile's other three
        // warps waiting at the barrier (6 % of the kernel's warp time in ncu's source view).  Same operands in the
        // same order as the thread that owns this output in the previous tile, so the same bits.
        if (j < 16 && warp0) {
#pragma unroll
            for (int m = 0; m < 4; m++) {
                const float tap = p.taps_s[4 * j + m];
                acc_prev = __ffma2_rn(x[m], make_float2(tap, tap), acc_prev);
            }
        }
    }
}

// acc = 2^-49 * sum tap u  ->  sum tap (u - 127): one FFMA2 per output (exact rescale, one rounding)
__device__ __forceinline__ float2 k1u_finish(float2 acc, const K1Params& p) {
    return __ffma2_rn(acc, make_float2(K1_UNSCALE, K1_UNSCALE), make_float2(p.neg_dc, p.neg_dc));
}


__global__ void __launch_bounds__(K1_THREADS, 6)
k1_fir4_discrim_u8(const uint8_t* __restrict__ iq, const float2* __restrict__ hist_in,
                   float2* __restrict__ hist_out, float* __restrict__ fm_demod,
                   const __grid_constant__ K1Params p)
{
    __shared__ __align__(128) uint8_t s_tile[K1U_SMEM];
    __shared__ float s_theta[K1_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int tile = blockIdx.x, s = blockIdx.y;
    const int o0 = tile * K1_TILE;
    const int n_valid = min(K1_TILE, p.n_out - o0);      // outputs of this tile (multiple of 16)
    const int n_rows = (n_valid >> 4) + 1;               // 128-byte rows to stage, row 0 = halo
    const size_t n_in = (size_t)p.n_out * K1_M;

    // ---- stage: row q of the tile <-> input bytes [(o0-16)*8 + 128 q, +128) of this stream ----
    const uint8_t* src = iq + (size_t)s * n_in * 2 + (ptrdiff_t)(o0 - 16) * 8;
    const int c_first = (tile == 0) ? 8 : 0;             // tile 0 takes its halo row from the history
    for (int ci = c_first + t; ci < n_rows * 8; ci += K1_THREADS) {
        const int row = ci >> 3, c = ci & 7;
        cp_async16(s_tile + row * 128 + ((c ^ (row & 7)) << 4), src + (size_t)ci * 16);
    }
    if (tile == 0 && t < 8) {                            // history: 64 samples kept as exact floats
        uint32_t wv[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float2 a = hist_in[(size_t)s * K1_HIST + t * 8 + 2 * k];
            const float2 b = hist_in[(size_t)s * K1_HIST + t * 8 + 2 * k + 1];
            wv[k] = (uint32_t)(int)(a.x + 127.0f) | ((uint32_t)(int)(a.y + 127.0f) << 8)
                  | ((uint32_t)(int)(b.x + 127.0f) << 16) | ((uint32_t)(int)(b.y + 127.0f) << 24);
        }
        // stream start: no samples yet.  The reference's zero history adds exact zeros; byte 0 (the
        // denormal 0.0f) does the same here, and the outputs that see it take their own DC term (below)
        if (p.first_block) wv[0] = wv[1] = wv[2] = wv[3] = 0u;
        *(uint4*)(s_tile + ((t ^ 0) << 4)) = make_uint4(wv[0], wv[1], wv[2], wv[3]);   // row 0: swizzle key 0
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();

    // ---- history for the next block: the last 64 IQ samples (one row) of this stream's block ----
    if (o0 + n_valid == p.n_out && t < K1_HIST) {
        const int row = n_valid >> 4;                    // last staged row
        const int c = t >> 3;                            // 8 samples per chunk
        const uint8_t* b = s_tile + row * 128 + ((c ^ (row & 7)) << 4) + (t & 7) * 2;
        hist_out[(size_t)s * K1_HIST + t] = make_float2((float)b[0] - 127.0f, (float)b[1] - 127.0f);
    }

    const bool active = (t * K1_R) < n_valid;
    float theta_last = 0.0f;
    float out[K1_R];
    float th_prev_own = 0.0f;                            // thread 0 only: angle of the output before the tile
    if (active) {
        float2 acc[K1_R], acc_prev = make_float2(0.0f, 0.0f);
#pragma unroll
        for (int r = 0; r < K1_R; r++) acc[r] = make_float2(0.0f, 0.0f);
        const uint8_t* row0 = s_tile + t * 128;
        const uint8_t* row1 = row0 + 128;
        const int k0 = (t & 7) << 4, k1 = ((t + 1) & 7) << 4;
        // window frames j = 0..15 are row t (chunk c holds frames 2c, 2c+1), j = 16..31 row t+1
#define K1U_DO(ROWP, KEY, C, J0) k1u_chunk<J0>(*(const uint4*)((ROWP) + (((C) << 4) ^ (KEY))), acc, acc_prev, warp == 0, p)
        K1U_DO(row0, k0, 0, 0);  K1U_DO(row0, k0, 1, 2);  K1U_DO(row0, k0, 2, 4);  K1U_DO(row0, k0, 3, 6);
        K1U_DO(row0, k0, 4, 8);  K1U_DO(row0, k0, 5, 10); K1U_DO(row0, k0, 6, 12); K1U_DO(row0, k0, 7, 14);
        K1U_DO(row1, k1, 0, 16); K1U_DO(row1, k1, 1, 18); K1U_DO(row1, k1, 2, 20); K1U_DO(row1, k1, 3, 22);
        K1U_DO(row1, k1, 4, 24); K1U_DO(row1, k1, 5, 26); K1U_DO(row1, k1, 6, 28); K1U_DO(row1, k1, 7, 30);
#undef K1U_DO
        if (t == 0) {
            // stream start: the reference's prev_theta is 0 (fm_demod.cpp:41; its zero FIR history gives atan2(0, 0))
            const float2 zp = k1u_finish(acc_prev, p);
            th_prev_own = (tile == 0 && p.first_block) ? 0.0f : fm_atan2f(zp.y, zp.x);
        }
        float prev = 0.0f;
        if (tile == 0 && t == 0 && p.first_block) {
            // the first 16 outputs of a stream: only the taps that met real samples carry the 127 offset
#pragma unroll
            for (int r = 0; r < K1_R; r++)
                acc[r] = __ffma2_rn(acc[r], make_float2(K1_UNSCALE, K1_UNSCALE), make_float2(p.first_neg_dc[r], p.first_neg_dc[r]));
        } else {
#pragma unroll
            for (int r = 0; r < K1_R; r++) acc[r] = k1u_finish(acc[r], p);
        }
        if (p.dbg_fm_in) {                               // GUI mode only: fm_in_buf for the FM-in spectrum
            float4* d4 = (float4*)(p.dbg_fm_in + (size_t)s * p.n_out + o0 + t * K1_R);
#pragma unroll
            for (int r = 0; r < K1_R; r += 2) d4[r >> 1] = make_float4(acc[r].x, acc[r].y, acc[r + 1].x, acc[r + 1].y);
        }
#pragma unroll
        for (int r = 0; r < K1_R; r += 2) {
            const float2 z0 = acc[r], z1 = acc[r + 1];
            const float2 th = fm_atan2f_x2(z0.y, z0.x, z1.y, z1.x);
            out[r] = (r == 0) ? th.x : th.x - prev;      // r = 0 fixed up below
            out[r + 1] = th.y - th.x;
            prev = th.y;
        }
        theta_last = prev;
    }
    // angle of the output before this thread's first one: previous lane / previous warp / own
    float th_before = __shfl_up_sync(0xffffffffu, theta_last, 1);
    if (lane == 31) s_theta[warp] = theta_last;
    __syncthreads();
    if (lane == 0) th_before = (warp == 0) ? th_prev_own : s_theta[warp - 1];
    if (!active) return;
    out[0] = out[0] - th_before;
#pragma unroll
    for (int r = 0; r < K1_R; r++) out[r] = wrap_phase(out[r]) * p.discrim_gain;
    float4* dst = (float4*)(fm_demod + (size_t)s * p.n_out + o0 + t * K1_R);
#pragma unroll
    for (int q = 0; q < K1_R / 4; q++) dst[q] = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
}

constexpr int K1_SMEM_BYTES = (K1_THREADS + 1) * K1_SEG * (int)sizeof(float);

cudaError_t launch_k1(bool u8, const void* iq, const float2* hist_in, float2* hist_out, float* fm_demod,
                      const K1Params& p, cudaStream_t st)
{
    // the attribute is per device: one flag per ordinal (a handle may live on any device of the process)
    static std::atomic<bool> configured[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !configured[dev].load(std::memory_order_acquire)) {
        cudaError_t e = cudaFuncSetAttribute(k1_fir4_discrim_cf32, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev].store(true, std::memory_order_release);
    }
    const dim3 grid((p.n_out + K1_TILE - 1) / K1_TILE, p.n_streams);
    if (u8) k1_fir4_discrim_u8<<<grid, K1_THREADS, 0, st>>>((const uint8_t*)iq, hist_in, hist_out, fm_demod, p);
    else    k1_fir4_discrim_cf32<<<grid, K1_THREADS, K1_SMEM_BYTES, st>>>((const float2*)iq, hist_in, hist_out, fm_demod, p);
    return cudaGetLastError();
}

} // namespace fm
