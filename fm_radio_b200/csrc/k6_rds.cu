// K6: RDS bit path on the device -- soft BPSK symbols of K5 -> bits -> 26-bit blocks -> groups ->
// PI / PTY / PS / RadioText, one thread per stream.
//
// Replaces the host loop  DifferentialManchesterDecoder::Process -> RDS_Decoding_Chain::Process
// (rds_decoder/differential_manchester_decoder.h:25-59, rds_decoding_chain.h:24, rds_group_sync.cpp,
// crc10.cpp, rds_decoder.cpp) that App runs per block (app.cpp:66-77): at 8192 streams x 10 000 x
// realtime that loop would have to take 19 M symbols/s off the device.  The integer program is the
// one of rds_core.h, shared with the host decoder, so device and host results are bit-identical.
//
// Per stream the kernel keeps rds::State (168 bytes) and two rings in HBM: the last `gcap` groups
// and the last `bcap` packet bytes, indexed by the running totals, so the host can collect results
// every block or every few blocks (fmgpu_rds_device_fetch).  ~150 symbols per 65536-sample block:
// the kernel is a few microseconds on stage stream D behind K5.
#include "fm_common.cuh"
#include "rds_core.h"

namespace fm {

struct RingSink {
    fmgpu_rds_group* glog; uint8_t* blog; int gcap, bcap;
    __device__ __forceinline__ void group(const fmgpu_rds_group& g, unsigned long long idx) { glog[idx % (unsigned)gcap] = g; }
    __device__ __forceinline__ void packet(const uint32_t pk[4], unsigned long long idx) {
        // bcap is a multiple of 16 and the ring is 16-byte aligned: one store, no wrap inside a packet
        *(uint4*)(blog + (idx % (unsigned)bcap)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
};

// Lanes are independent decoders, so naive code is slow twice over (measured at 1024 streams: 0.4 ms
// per launch, then 0.06 ms, now a few microseconds):
//   * each lane walks its own row of symbols, 4-byte loads 4 KB apart with the L2 latency exposed in
//     every iteration.  Only the SIGN of a symbol is used, so the warp first turns a window of up to
//     2048 symbols of its 32 rows into sign bits with coalesced 128-byte loads + ballot (one 32-bit
//     word per row per load), kept in shared memory;
//   * each lane's packet fills at its own symbol index and its 128-bit group-sync burst would run
//     while the 31 others wait.  The warp therefore alternates two lock-step phases:
//       A  every lane consumes sign bits until ITS packet is full (or its symbols are exhausted);
//       B  every lane holding a full packet runs the 128-bit burst, together.
// Each lane still executes exactly the reference's sequence (bits reach the group sync when the
// packet fills); only the interleaving between lanes changes.
constexpr int K6_WIN = 2048;                    // symbols per window
constexpr int K6_WORDS = K6_WIN / 32 + 1;       // +1: row stride 65 words, conflict-free column reads

__global__ void __launch_bounds__(32)
k6_rds(const float* __restrict__ pred_sym, const int* __restrict__ sym_count, rds::State* __restrict__ state,
       const rds::Tables* __restrict__ tables, fmgpu_rds_group* __restrict__ glog, uint8_t* __restrict__ blog,
       int n64, int gcap, int bcap, int n_streams)
{
    __shared__ rds::Tables T;
    __shared__ uint32_t signs[32][K6_WORDS];
    const int lane = threadIdx.x;
    {
        const uint4* src = (const uint4*)tables;
        uint4* dst = (uint4*)&T;
        for (int i = lane; i < (int)(sizeof(rds::Tables) / 16); i += 32) dst[i] = src[i];
    }
    const int s0 = blockIdx.x * 32;
    const int s = s0 + lane;
    const bool live = s < n_streams;
    const int sc = live ? s : s0;
    rds::Work st;
    rds::load(st, state[sc]);
    RingSink sink{ glog + (size_t)sc * gcap, blog + (size_t)sc * bcap, gcap, bcap };
    const int n = live ? min(sym_count[sc], n64) : 0;
    int n_max = n;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, off));
    const int rows = min(32, n_streams - s0);

    for (int w0 = 0; w0 < n_max; w0 += K6_WIN) {
        const int w_len = min(K6_WIN, n_max - w0);
        __syncwarp();
        // sign bits of symbols [w0, w0 + w_len) of this warp's rows (values beyond a row's own count are
        // never consumed; the buffer has n64 entries per stream, so the loads stay in bounds)
        for (int c = 0; c < (w_len + 31) / 32; c++) {
            const int col = w0 + c * 32 + lane;
            for (int r0 = 0; r0 < 32; r0 += 8) {                   // 8 independent loads in flight, then 8 ballots
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int r = min(r0 + q, rows - 1);
                    v[q] = (col < n64) ? __ldg(pred_sym + (size_t)(s0 + r) * n64 + col) : 0.0f;
                }
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const uint32_t word = __ballot_sync(0xffffffffu, v[q] > 0.0f);
                    if (lane == 0) signs[r0 + q][c] = word;
                }
            }
        }
        __syncwarp();
        const int end = min(n, w0 + w_len);
        int i = min(n, w0);
        while (true) {
            bool full = false;
            while (i < end && !full) {                                            // phase A
                const int k = i - w0;
                full = rds::push_symbol_level(st, (signs[lane][k >> 5] >> (k & 31)) & 1u);
                i++;
            }
            if (!__any_sync(0xffffffffu, full)) break;
            if (full) rds::process_packet(st, T, sink);                           // phase B
        }
    }
    if (live) rds::store(st, state[s]);
}

__global__ void k6_rds_init(rds::State* state, int n_streams) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < n_streams) { rds::State st; rds::init(st); state[s] = st; }
}

cudaError_t launch_k6(const float* pred_sym, const int* sym_count, void* state, const void* tables, void* glog, uint8_t* blog,
                      int n64, int gcap, int bcap, int n_streams, cudaStream_t st)
{
    k6_rds<<<(n_streams + 31) / 32, 32, 0, st>>>(pred_sym, sym_count, (rds::State*)state, (const rds::Tables*)tables,
                                                 (fmgpu_rds_group*)glog, blog, n64, gcap, bcap, n_streams);
    return cudaGetLastError();
}

cudaError_t launch_k6_init(void* state, int n_streams, cudaStream_t st) {
    k6_rds_init<<<(n_streams + 127) / 128, 128, 0, st>>>((rds::State*)state, n_streams);
    return cudaGetLastError();
}

size_t k6_state_bytes() { return sizeof(rds::State); }
static_assert(sizeof(rds::Tables) % 16 == 0, "tables are staged with 16-byte copies");

} // namespace fm
