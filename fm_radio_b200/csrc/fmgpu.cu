// C-ABI implementation (include/fmgpu.h): handle, per-stream device state, the stage pipeline.
//
// Execution model: six stage streams per handle.  Block k of the batch flows
//     H (host->device copy) -> A (K1, K2) -> B (K3 pilot PLL) -> C (K4, K4b) -> D (K5 BPSK) -> E (K6 RDS bits)
//                                                                               D -> O (device->host)
// through ring slot k % depth; stage X of block k+1 follows stage X of block k on the same stream
// (all cross-block filter/loop state is owned by exactly one stage), and CUDA events order stage
// X(k) after X-1(k) and after the last reader of the slot's buffers.  The two latency-bound
// recurrences (K3, K5: one thread per stream) therefore overlap with the FMA-bound kernels of the
// neighbouring blocks instead of serialising the chain.
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/fmgpu.h"
#include "fm_common.cuh"
#include "rds_core.h"

extern "C" const void* fmgpu_rds_tables_(size_t* bytes);     // rds_host.cpp


namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    return fail(FMGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)

struct HostTaps {
    float fm_in[64], fm_out[64], hilbert[65], lpr[128], lmr[128], rds[128];
    float deemph_b[2], deemph_a[2], peak_b[3], peak_a[3], pll_b[2], pll_a[2];
    float ted_b[2], ted_a[2], bpll_b[2], bpll_a[2];
};

struct Slot {
    uint8_t* in_u8 = nullptr;      // staged input (host-buffer path)
    float* fm_demod = nullptr;     // [S][B/4]
    float2* fm_out_iq = nullptr;   // [S][B/8]
    float* theta = nullptr;        // [S][B/8]
    float* power = nullptr;        // [S]
    float* pll_dt = nullptr;       // [S][B/8]
    float2* audio = nullptr;       // [S][B/32]
    float2* rds = nullptr;         // [S][B/64] (before AGC)
    float* est_partial = nullptr;  // [S][tiles]
    float* rds_pw_partial = nullptr;
    float* pred_sym = nullptr;     // [S][B/64]
    int* sym_count = nullptr;      // [S]
    float2* pcm_f32 = nullptr;     // [S][pcm_n] audio at the PCM rate (K7; allocated when the stage is switched on)
    short2* pcm_s16 = nullptr;     // [S][pcm_n]
    cudaEvent_t ev_H, ev_A, ev_B, ev_C, ev_D, ev_E, ev_O, ev_P, ev_K1;
    bool fetched_in_graph = false; // the slot's outputs were copied to the host by the CUDA graph that produced them
};

struct DebugBufs {                 // keep_intermediates only (single set, not ringed)
    float2* pilot = nullptr; float2* pll = nullptr; float* pll_raw = nullptr; float* pll_pi = nullptr;
    float* lpr = nullptr; float* lmr = nullptr; float2* fm_in = nullptr;
    float2* lpr_iq = nullptr; float2* lmr_iq = nullptr; float2* hist_iq = nullptr; float* lmr_phase_used = nullptr;   // GUI audio spectra
    fm::K5Debug k5{};
};

struct HostMirror {                // pinned
    float2* audio = nullptr; float* pred_sym = nullptr; int* sym_count = nullptr;
    short2* pcm_s16 = nullptr;     // K7 on: the int16 PCM block is fetched with the audio
};

} // namespace

namespace { struct GraphEntry { cudaGraphExec_t exec = nullptr; const void* in_ptr = nullptr; bool u8 = false; unsigned long long gen = 0; }; }

struct fmgpu_demod {
    fmgpu_config cfg{};
    int B = 0, S = 0, n4 = 0, n8 = 0, n32 = 0, n64 = 0, depth = 0, k4_tiles = 0;
    int device = 0;
    cudaStream_t stH = nullptr, stA = nullptr, stA2 = nullptr, stP = nullptr, stB = nullptr, stC = nullptr, stD = nullptr, stE = nullptr, stO = nullptr;
    // SM partition (green contexts): the recurrence stages B, D get their own SMs
    CUgreenCtx gctx_rec = nullptr, gctx_fir = nullptr;
    int sms_rec = 0, sms_fir = 0;
    int n_sm_fir = 148;            // SMs that run the FIR stages (the FIR partition, or the whole device)
    std::vector<Slot> slots;
    std::vector<HostMirror> mirrors;
    DebugBufs dbg;
    // per-stream state
    float2* k1_hist[2] = { nullptr, nullptr };
    // K1 on the tensor cores (k1_toeplitz_i8.cu): byte history (ping-pong), G image, dp4a table, digit-plane constants
    uint8_t* k1t_hist[2] = { nullptr, nullptr };
    float* k1t_theta[2] = { nullptr, nullptr };
    int8_t* k1t_bimg = nullptr; int* k1t_ptab = nullptr;
    float k1t_taps[64] = { 0 }; bool k1t_ready = false, use_k1t = true;
    int k1t_off[3] = { 0, 0, 0 }; float k1t_w[3] = { 0, 0, 0 };
    int last_input_kind = 0;       // 0 none yet, 1 u8, 2 cf32
    int k1t_shape = 1;             // 3 CTAs/SM (measured 0.0574 vs 0.0627 ms for the 2-CTA shape, tools/k1t_probe.cu)
    bool k5_literal = false, k3_exact = false, k4_v1 = false, k3_single = false, k3_duo = false;
    // CUDA-graph replay of small blocks (enqueue_chain)
    cudaStream_t stG = nullptr;
    bool use_graph = false, last_was_graph = false;
    unsigned long long graph_gen = 1;
    std::vector<GraphEntry> graphs;
    unsigned fetch_mask = FMGPU_FETCH_ALL, last_fetch_mask = 0;
    float* k2_hist_demod = nullptr; float* k2_hist_out = nullptr; float* k2_scal = nullptr;
    float* pll_state = nullptr;
    float* k4_hist_x[2] = { nullptr, nullptr };
    float2* k4_hist_m2[2] = { nullptr, nullptr };
    float2* k4_hist_m3[2] = { nullptr, nullptr };
    float* lmr_phase = nullptr;
    float* bpsk_state = nullptr;
    // K6: device RDS decoders ([S] rds::State) and their result rings
    void* rds_state = nullptr; void* rds_glog = nullptr; uint8_t* rds_blog = nullptr; void* rds_tables = nullptr;
    int rds_gcap = 0, rds_bcap = 0;
    std::vector<rds::State> rds_host_state;            // fmgpu_rds_device_fetch copies
    std::vector<fmgpu_rds_group> rds_host_glog;
    std::vector<uint8_t> rds_host_blog;
    float2* in_f32 = nullptr;      // cf32 path staging (lazy)
    // host-side configuration
    HostTaps taps{};
    int ctl_audio_out = 2; float ctl_stereo_mix = 1.0f; int ctl_use_deemph = 0;
    int ctl_deemph_tus = 1, ctl_lpr_hz = 15000, ctl_lmr_hz = 15000;
    // K7 audio output stage: 0 = off; else output frames per block = (int)((rate / 32000.f) * n32)
    int ctl_pcm_rate = 0, pcm_rate_built = 0, pcm_n = 0;
    fm::K7Entry* pcm_table = nullptr;
    bool dirty_deemph = true, dirty_lpr = true, dirty_lmr = true;
    // bookkeeping
    unsigned long long step = 0;   // blocks enqueued so far
    long long launches = 0;
    int last_fetched_slot = -1;
    bool dbg_valid = false;
    std::vector<uint8_t> dbg_host; // host copies for the debug getters
    std::vector<float> scalar_host;
};

namespace {

template <typename T>
cudaError_t dalloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemset(*p, 0, n * sizeof(T));
}

void design_default_taps(fmgpu_demod* h) {
    // broadcast_fm_demod.cpp:129-274 and bpsk_synchroniser.cpp:27-48
    const float ROLLOFF = 0.95f;
    HostTaps& t = h->taps;
    fmgpu_create_fir_lpf(t.fm_in, 64, (256000.0f / 2.0f) / (1024000.0f / 2.0f) * ROLLOFF);
    fmgpu_create_fir_lpf(t.fm_out, 64, (128000.0f / 2.0f) / (256000.0f / 2.0f) * ROLLOFF);
    fmgpu_create_fir_hilbert(t.hilbert, 65);
    fmgpu_create_iir_peak_1_filter(t.peak_b, t.peak_a, 19000.0f / (128000.0f / 2.0f), 0.9999f);
    fmgpu_create_iir_single_pole_lpf(t.pll_b, t.pll_a, 100.0f / (128000.0f / 2.0f));
    fmgpu_create_fir_lpf(t.rds, 128, 2000.0f / (128000.0f / 2.0f));
    fmgpu_create_iir_single_pole_lpf(t.ted_b, t.ted_a, 1.5e3f / (16e3f / 2.0f));
    fmgpu_create_iir_single_pole_lpf(t.bpll_b, t.bpll_a, 10.0f / (16e3f / 2.0f));
}

float clampk(float k) { const float lo = 0.01f, hi = 0.99f; return k > lo ? (k > hi ? hi : k) : lo; }

// Broadcast_FM_Demod::UpdateFilters (broadcast_fm_demod.cpp:330-389)
void update_filters(fmgpu_demod* h) {
    const float PI = 3.14159265358979323846f;
    if (h->dirty_deemph) {
        h->dirty_deemph = false;
        const float Tc = (float)h->ctl_deemph_tus * 1e-6f;
        const float Fc = 1.0f / (2.0f * PI * Tc);
        fmgpu_create_iir_single_pole_lpf(h->taps.deemph_b, h->taps.deemph_a, clampk(Fc / (128000.0f / 2.0f)));
    }
    if (h->dirty_lpr) {
        h->dirty_lpr = false;
        fmgpu_create_fir_lpf(h->taps.lpr, 128, clampk((float)h->ctl_lpr_hz / (128000.0f / 2.0f)));
    }
    if (h->dirty_lmr) {
        h->dirty_lmr = false;
        fmgpu_create_fir_lpf(h->taps.lmr, 128, clampk((float)h->ctl_lmr_hz / (128000.0f / 2.0f)));
    }
}

// Powers of the state matrices of the two linear recurrences of K2, in double (see K2Params).
void fill_scan_matrices(fm::K2Params& p) {
    auto matpow = [](double a1, double a0, int n, float out[4]) {
        double m[4] = { 1, 0, 0, 1 };                       // row-major 2x2
        for (int i = 0; i < n; i++) {
            const double r0 = a1 * m[0] + a0 * m[2], r1 = a1 * m[1] + a0 * m[3];
            m[2] = m[0]; m[3] = m[1]; m[0] = r0; m[1] = r1;  // m = A * m, A = [[a1, a0], [1, 0]]
        }
        for (int i = 0; i < 4; i++) out[i] = (float)m[i];
    };
    const double a1 = p.peak_a[1], a0 = p.peak_a[0];
    for (int l = 0; l < 6; l++) matpow(a1, a0, 8 << l, p.pk_P[l]);
    for (int j = 0; j < 32; j++) matpow(a1, a0, 8 * j, p.pk_Q[j]);
    const double al = p.deemph_a[0];
    for (int l = 0; l < 6; l++) p.de_P[l] = (float)std::pow(al, (double)(8 << l));
    for (int j = 0; j < 32; j++) p.de_Q[j] = (float)std::pow(al, (double)(8 * j));
}

int init_state(fmgpu_demod* h) {
    // initial values that are not zero: AGC gains 0.1 (dsp/agc.h:10)
    std::vector<float> pll(fm::PLL_STATE_N * (size_t)h->S, 0.0f), bp(fm::BP_STATE_N * (size_t)h->S, 0.0f);
    for (int s = 0; s < h->S; s++) {
        pll[(size_t)fm::PLL_AGC_GAIN * h->S + s] = 0.1f;
        bp[(size_t)fm::BP_AGC_GAIN * h->S + s] = 0.1f;
    }
    CU(cudaMemcpy(h->pll_state, pll.data(), pll.size() * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->bpsk_state, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice));
    return FMGPU_OK;
}

// Splits the device's SMs into {recurrence partition, FIR partition} with the driver's green-context API
// (cuda.h "Green Contexts"; entry points fetched through the runtime, so libcuda is not a link dependency)
// and creates stage streams B, D on the first and A, C on the second.  Returns false, with nothing
// created, when partitioning is disabled or unsupported.
bool create_partitioned_streams(fmgpu_demod* h, int prio_hi) {
    if (std::getenv("FMGPU_NO_PARTITION")) return false;
    unsigned want = 16;                                   // 64 recurrence warps at 1024 streams = one per SM sub-partition
    if (const char* e = std::getenv("FMGPU_RECURRENCE_SMS")) want = (unsigned)std::atoi(e);
    if (want == 0) return false;
#define DRV(name) decltype(&name) p_##name = nullptr; { void* fp = nullptr; cudaDriverEntryPointQueryResult qr; \
        if (cudaGetDriverEntryPoint(#name, &fp, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fp) { cudaGetLastError(); return false; } \
        p_##name = (decltype(&name))fp; }
    DRV(cuDeviceGet) DRV(cuDeviceGetDevResource) DRV(cuDevSmResourceSplitByCount) DRV(cuDevResourceGenerateDesc)
    DRV(cuGreenCtxCreate) DRV(cuGreenCtxDestroy) DRV(cuGreenCtxStreamCreate)
#undef DRV
    CUdevice dev;
    if (p_cuDeviceGet(&dev, h->device) != CUDA_SUCCESS) return false;
    CUdevResource all{}, rec{}, rest{};
    if (p_cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    unsigned groups = 1;
    if (p_cuDevSmResourceSplitByCount(&rec, &groups, &all, &rest, 0, want) != CUDA_SUCCESS || groups != 1) return false;
    if (rec.sm.smCount == 0 || rest.sm.smCount == 0) return false;
    CUdevResourceDesc d_rec = nullptr, d_rest = nullptr;
    if (p_cuDevResourceGenerateDesc(&d_rec, &rec, 1) != CUDA_SUCCESS) return false;
    if (p_cuDevResourceGenerateDesc(&d_rest, &rest, 1) != CUDA_SUCCESS) return false;
    CUgreenCtx g_rec = nullptr, g_rest = nullptr;
    if (p_cuGreenCtxCreate(&g_rec, d_rec, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    if (p_cuGreenCtxCreate(&g_rest, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) { p_cuGreenCtxDestroy(g_rec); return false; }
    CUstream a = nullptr, a2 = nullptr, pp = nullptr, b = nullptr, c = nullptr, d = nullptr, e2 = nullptr;
    const bool ok = p_cuGreenCtxStreamCreate(&a, g_rest, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS
                 && p_cuGreenCtxStreamCreate(&a2, g_rest, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS
                 && p_cuGreenCtxStreamCreate(&pp, std::getenv("FMGPU_K7_ON_REC") ? g_rec : g_rest, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS
                 && p_cuGreenCtxStreamCreate(&c, g_rest, CU_STREAM_NON_BLOCKING, 0) == CUDA_SUCCESS
                 && p_cuGreenCtxStreamCreate(&b, g_rec, CU_STREAM_NON_BLOCKING, prio_hi) == CUDA_SUCCESS
                 && p_cuGreenCtxStreamCreate(&d, g_rec, CU_STREAM_NON_BLOCKING, prio_hi) == CUDA_SUCCESS
                 && p_cuGreenCtxStreamCreate(&e2, std::getenv("FMGPU_K6_ON_FIR") ? g_rest : g_rec, CU_STREAM_NON_BLOCKING, prio_hi) == CUDA_SUCCESS;
    if (!ok) {
        for (CUstream st : { a, a2, pp, b, c, d, e2 }) if (st) cudaStreamDestroy((cudaStream_t)st);
        p_cuGreenCtxDestroy(g_rec); p_cuGreenCtxDestroy(g_rest);
        return false;
    }
    h->stA = (cudaStream_t)a; h->stA2 = (cudaStream_t)a2; h->stP = (cudaStream_t)pp; h->stB = (cudaStream_t)b; h->stC = (cudaStream_t)c; h->stD = (cudaStream_t)d; h->stE = (cudaStream_t)e2;
    h->gctx_rec = g_rec; h->gctx_fir = g_rest;
    h->sms_rec = (int)rec.sm.smCount; h->sms_fir = (int)rest.sm.smCount;
    return true;
}

void destroy_partition(fmgpu_demod* h) {
    if (!h->gctx_rec && !h->gctx_fir) return;
    void* fp = nullptr; cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuGreenCtxDestroy", &fp, cudaEnableDefault, &qr) == cudaSuccess && fp) {
        auto destroy = (decltype(&cuGreenCtxDestroy))fp;
        if (h->gctx_rec) destroy(h->gctx_rec);
        if (h->gctx_fir) destroy(h->gctx_fir);
    }
    h->gctx_rec = h->gctx_fir = nullptr;
}

int alloc_all(fmgpu_demod* h) {
    const size_t S = h->S;
    CU(cudaStreamCreateWithFlags(&h->stH, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->stO, cudaStreamNonBlocking));
    // The two latency-bound recurrences (stage B: K3; stage D: K5, K6 -- one thread per stream, 32 + 32
    // one-warp CTAs at 1024 streams) lose 1.5x when their warps share SM sub-partitions with the FFMA
    // streams of K1/K2/K4 (measured: K3 0.278 ms alone, 0.428 ms inside the pipeline), and they are the
    // pipeline's critical stages.  So the handle splits the GPU with green contexts: a small SM
    // partition runs only stages B and D, the rest runs the FIR stages A and C.
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (!create_partitioned_streams(h, prio_hi)) {
        // no partition (FMGPU_NO_PARTITION, or the driver refused): plain streams; the recurrences at
        // least get the highest CTA-scheduling priority
        CU(cudaStreamCreateWithFlags(&h->stA, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->stA2, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&h->stP, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithPriority(&h->stB, cudaStreamNonBlocking, prio_hi));
        CU(cudaStreamCreateWithFlags(&h->stC, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithPriority(&h->stD, cudaStreamNonBlocking, prio_hi));
        CU(cudaStreamCreateWithPriority(&h->stE, cudaStreamNonBlocking, prio_hi));
    }
    h->use_k1t = std::getenv("FMGPU_K1_FP32") == nullptr;
    CU(cudaStreamCreateWithFlags(&h->stG, cudaStreamNonBlocking));
    h->graphs.assign((size_t)h->depth * 2, GraphEntry{});
    // launch-bound regime: with few streams and short blocks every kernel lasts a few microseconds and the ~60 API calls
    // of the multi-stream pipeline dominate; large batches keep the pipeline (its cross-block overlap hides the recurrences)
    h->use_graph = !std::getenv("FMGPU_NO_GRAPH") && (size_t)h->S * (size_t)h->B <= (size_t)1 << 20 && h->B <= 16384;
    if (const char* e = std::getenv("FMGPU_GRAPH")) h->use_graph = std::atoi(e) != 0;
    h->k5_literal = std::getenv("FMGPU_K5_LITERAL") != nullptr;
    h->k3_exact = std::getenv("FMGPU_K3_EXACT") != nullptr;
    h->k3_single = std::getenv("FMGPU_K3_SINGLE") != nullptr;
    h->k3_duo = std::getenv("FMGPU_K3_DUO") != nullptr;
    h->k4_v1 = std::getenv("FMGPU_K4_V1") != nullptr;
    if (const char* e = std::getenv("FMGPU_K1T_SHAPE")) h->k1t_shape = std::atoi(e);
    {
        int n_sm = 148;
        CU(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device));
        h->n_sm_fir = h->sms_fir > 0 ? h->sms_fir : n_sm;
    }
    for (int i = 0; i < 2; i++) {
        CU(cudaMalloc((void**)&h->k1t_hist[i], S * 128));
        CU(cudaMemset(h->k1t_hist[i], 127, S * 128));       // a new stream: zero FIR history <-> bytes 127
        CU(dalloc(&h->k1t_theta[i], S));
    }
    CU(cudaMalloc((void**)&h->k1t_bimg, 2 * 96 * 128));
    CU(cudaMalloc((void**)&h->k1t_ptab, 32 * 6 * sizeof(int)));
    for (int i = 0; i < 2; i++) {
        CU(dalloc(&h->k1_hist[i], S * fm::K1_HIST));
        CU(dalloc(&h->k4_hist_x[i], S * fm::K4_NN));
        CU(dalloc(&h->k4_hist_m2[i], S * fm::K4_NN));
        CU(dalloc(&h->k4_hist_m3[i], S * fm::K4_NN));
    }
    CU(dalloc(&h->k2_hist_demod, S * fm::K2_NN));
    CU(dalloc(&h->k2_hist_out, S * 64));
    CU(dalloc(&h->k2_scal, S * fm::K2_SCAL_N));
    CU(dalloc(&h->pll_state, S * fm::PLL_STATE_N));
    CU(dalloc(&h->bpsk_state, S * fm::BP_STATE_N));
    CU(dalloc(&h->lmr_phase, S));
    // rings sized for what one block can produce (n64 chips -> n64/2 bits) plus slack, so a fetch
    // every block never loses anything; at least 64 groups / 1 KiB of packets
    h->rds_gcap = std::max(64, h->n64 / 2 / 104 * 2 + 8);
    h->rds_bcap = std::max(1024, ((h->n64 / 16 * 2 + 64) + 15) / 16 * 16);
    CU(cudaMalloc(&h->rds_state, S * fm::k6_state_bytes()));
    CU(cudaMalloc(&h->rds_glog, S * h->rds_gcap * sizeof(fmgpu_rds_group)));
    CU(cudaMemset(h->rds_glog, 0, S * h->rds_gcap * sizeof(fmgpu_rds_group)));
    CU(dalloc(&h->rds_blog, S * h->rds_bcap));
    {
        size_t tb = 0;
        const void* src = fmgpu_rds_tables_(&tb);
        CU(cudaMalloc(&h->rds_tables, tb));
        CU(cudaMemcpy(h->rds_tables, src, tb, cudaMemcpyHostToDevice));
    }
    CU(fm::launch_k6_init(h->rds_state, h->S, h->stD));
    CU(cudaStreamSynchronize(h->stD));
    h->slots.resize(h->depth);
    h->mirrors.resize(h->depth);
    for (int i = 0; i < h->depth; i++) {
        Slot& sl = h->slots[i];
        CU(dalloc(&sl.fm_demod, S * h->n4));
        CU(dalloc(&sl.fm_out_iq, S * h->n8));
        CU(dalloc(&sl.theta, S * h->n8));
        CU(dalloc(&sl.power, S));
        CU(dalloc(&sl.pll_dt, S * h->n8));
        CU(dalloc(&sl.audio, S * h->n32));
        CU(dalloc(&sl.rds, S * h->n64));
        CU(dalloc(&sl.est_partial, S * h->k4_tiles));
        CU(dalloc(&sl.rds_pw_partial, S * h->k4_tiles));
        CU(dalloc(&sl.pred_sym, S * h->n64));
        CU(dalloc(&sl.sym_count, S));
        cudaEvent_t* evs[9] = { &sl.ev_H, &sl.ev_A, &sl.ev_B, &sl.ev_C, &sl.ev_D, &sl.ev_E, &sl.ev_O, &sl.ev_P, &sl.ev_K1 };
        for (auto* ev : evs) CU(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
        HostMirror& m = h->mirrors[i];
        CU(cudaMallocHost((void**)&m.audio, S * h->n32 * sizeof(float2)));
        CU(cudaMallocHost((void**)&m.pred_sym, S * h->n64 * sizeof(float)));
        CU(cudaMallocHost((void**)&m.sym_count, S * sizeof(int)));
        std::memset(m.sym_count, 0, S * sizeof(int));
    }
    if (h->cfg.keep_intermediates) {
        DebugBufs& d = h->dbg;
        CU(dalloc(&d.pilot, S * h->n8)); CU(dalloc(&d.pll, S * h->n8));
        CU(dalloc(&d.pll_raw, S * h->n8)); CU(dalloc(&d.pll_pi, S * h->n8));
        CU(dalloc(&d.lpr, S * h->n32)); CU(dalloc(&d.lmr, S * h->n32)); CU(dalloc(&d.fm_in, S * h->n4));
        CU(dalloc(&d.lpr_iq, S * h->n32)); CU(dalloc(&d.lmr_iq, S * h->n32)); CU(dalloc(&d.hist_iq, S * fm::K4_NN)); CU(dalloc(&d.lmr_phase_used, S));
        CU(dalloc(&d.k5.rds, S * h->n64)); CU(dalloc(&d.k5.raw_sym, S * h->n64)); CU(dalloc(&d.k5.pll_sym, S * h->n64));
        CU(dalloc(&d.k5.zcd, S * h->n64)); CU(dalloc(&d.k5.dump_trig, S * h->n64));
        CU(dalloc(&d.k5.ted_raw, S * h->n64)); CU(dalloc(&d.k5.ted_pi, S * h->n64));
        CU(dalloc(&d.k5.pll_raw, S * h->n64)); CU(dalloc(&d.k5.pll_pi, S * h->n64));
        CU(dalloc(&d.k5.dump_filter, S * h->n64));
    }
    return init_state(h);
}

void free_all(fmgpu_demod* h) {
    // only this handle's streams are drained: other handles and streams of the device keep running
    {
        cudaStream_t sts0[10] = { h->stH, h->stA, h->stA2, h->stP, h->stB, h->stC, h->stD, h->stE, h->stO, h->stG };
        for (auto st : sts0) if (st) cudaStreamSynchronize(st);
        for (auto& ge : h->graphs) if (ge.exec) cudaGraphExecDestroy(ge.exec);
        if (h->stG) cudaStreamDestroy(h->stG);
    }
    auto F = [](void* p) { if (p) cudaFree(p); };
    for (int i = 0; i < 2; i++) { F(h->k1t_hist[i]); F(h->k1t_theta[i]); }
    F(h->k1t_bimg); F(h->k1t_ptab);
    for (int i = 0; i < 2; i++) { F(h->k1_hist[i]); F(h->k4_hist_x[i]); F(h->k4_hist_m2[i]); F(h->k4_hist_m3[i]); }
    F(h->k2_hist_demod); F(h->k2_hist_out); F(h->k2_scal); F(h->pll_state); F(h->bpsk_state); F(h->lmr_phase); F(h->in_f32);
    F(h->rds_state); F(h->rds_glog); F(h->rds_blog); F(h->rds_tables); F(h->pcm_table);
    for (auto& sl : h->slots) {
        F(sl.in_u8); F(sl.fm_demod); F(sl.fm_out_iq); F(sl.theta); F(sl.power); F(sl.pll_dt); F(sl.audio); F(sl.rds);
        F(sl.est_partial); F(sl.rds_pw_partial); F(sl.pred_sym); F(sl.sym_count); F(sl.pcm_f32); F(sl.pcm_s16);
        cudaEvent_t evs[9] = { sl.ev_H, sl.ev_A, sl.ev_B, sl.ev_C, sl.ev_D, sl.ev_E, sl.ev_O, sl.ev_P, sl.ev_K1 };
        for (auto ev : evs) if (ev) cudaEventDestroy(ev);
    }
    for (auto& m : h->mirrors) {
        if (m.audio) cudaFreeHost(m.audio);
        if (m.pred_sym) cudaFreeHost(m.pred_sym);
        if (m.sym_count) cudaFreeHost(m.sym_count);
        if (m.pcm_s16) cudaFreeHost(m.pcm_s16);
    }
    DebugBufs& d = h->dbg;
    F(d.pilot); F(d.pll); F(d.pll_raw); F(d.pll_pi); F(d.lpr); F(d.lmr); F(d.fm_in); F(d.lpr_iq); F(d.lmr_iq); F(d.hist_iq); F(d.lmr_phase_used);
    F(d.k5.rds); F(d.k5.raw_sym); F(d.k5.pll_sym); F(d.k5.zcd); F(d.k5.dump_trig);
    F(d.k5.ted_raw); F(d.k5.ted_pi); F(d.k5.pll_raw); F(d.k5.pll_pi); F(d.k5.dump_filter);
    cudaStream_t sts[9] = { h->stH, h->stA, h->stA2, h->stP, h->stB, h->stC, h->stD, h->stE, h->stO };
    for (auto st : sts) if (st) cudaStreamDestroy(st);
    destroy_partition(h);
}

int sync_all(fmgpu_demod* h);
int fetch_copies(fmgpu_demod* h, int slot, unsigned mask, cudaStream_t st);

// K7's buffers and read-position table, (re)built when FMGPU_CTL_AUDIO_PCM_RATE_HZ changes.
int prepare_pcm(fmgpu_demod* h) {
    if (h->pcm_rate_built == h->ctl_pcm_rate) return FMGPU_OK;
    if (sync_all(h) != FMGPU_OK) return FMGPU_ERR_CUDA;
    // Resampled_PCM_Player::ConsumeBuffer (audio/resampled_pcm_player.cpp:22-24): L = out / in, M = (int)(L * N)
    const float L = (float)h->ctl_pcm_rate / 32000.0f;
    const int M = (int)(L * (float)h->n32);
    const size_t S = h->S;
    for (int i = 0; i < h->depth; i++) {
        Slot& sl = h->slots[i];
        HostMirror& m = h->mirrors[i];
        if (sl.pcm_f32) cudaFree(sl.pcm_f32);
        if (sl.pcm_s16) cudaFree(sl.pcm_s16);
        if (m.pcm_s16) cudaFreeHost(m.pcm_s16);
        sl.pcm_f32 = nullptr; sl.pcm_s16 = nullptr; m.pcm_s16 = nullptr;
        CU(dalloc(&sl.pcm_f32, S * M));
        CU(dalloc(&sl.pcm_s16, S * M));
        CU(cudaMallocHost((void**)&m.pcm_s16, S * M * sizeof(short2)));
    }
    std::vector<fm::K7Entry> tab((size_t)M);
    fm::k7_build_table(h->n32, M, tab.data());
    if (M > 0 && tab.back().j0 >= h->n32) return fail(FMGPU_ERR_ARG, "audio PCM rate: Resample()'s read position leaves the block (the reference would read out of bounds)");
    if (h->pcm_table) cudaFree(h->pcm_table);
    h->pcm_table = nullptr;
    CU(cudaMalloc((void**)&h->pcm_table, sizeof(fm::K7Entry) * (size_t)M));
    CU(cudaMemcpy(h->pcm_table, tab.data(), sizeof(fm::K7Entry) * (size_t)M, cudaMemcpyHostToDevice));
    h->pcm_n = M;
    h->pcm_rate_built = h->ctl_pcm_rate;
    return FMGPU_OK;
}

// Enqueue the chain for one block whose input is already ordered on stA (or signalled by ev_H).
// K1's tensor-core tables follow the fm_in taps (fmgpu_upload_taps may change them); synchronous, so never inside a capture
int ensure_k1t_tables(fmgpu_demod* h) {
    if (h->k1t_ready && std::memcmp(h->k1t_taps, h->taps.fm_in, sizeof(h->k1t_taps)) == 0) return FMGPU_OK;
    std::vector<int8_t> bimg; std::vector<int> ptab;
    fm::k1t_build_tables(h->taps.fm_in, bimg, ptab, h->k1t_off, h->k1t_w);
    if (sync_all(h) != FMGPU_OK) return FMGPU_ERR_CUDA;
    CU(cudaMemcpy(h->k1t_bimg, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->k1t_ptab, ptab.data(), ptab.size() * sizeof(int), cudaMemcpyHostToDevice));
    std::memcpy(h->k1t_taps, h->taps.fm_in, sizeof(h->k1t_taps));
    h->k1t_ready = true;
    return FMGPU_OK;
}

// The kernels of one block.  single == nullptr: the production pipeline -- every stage on its own stream, ordered by the
// slot's events, blocks overlapping.  single != nullptr: every kernel on that one stream in chain order, no events --
// the form that is captured into a CUDA graph for small, launch-bound blocks (enqueue_block below).
int enqueue_chain_on(fmgpu_demod* h, const void* iq_dev, bool u8, bool wait_H, cudaEvent_t* prof, cudaStream_t single) {
    cudaStream_t stA = single ? single : h->stA, stA2 = single ? single : h->stA2, stP = single ? single : h->stP,
                 stB = single ? single : h->stB, stC = single ? single : h->stC, stD = single ? single : h->stD, stE = single ? single : h->stE;
#define EV_WAIT(st, ev) do { if (!single) CU(cudaStreamWaitEvent(st, ev, 0)); } while (0)
#define EV_REC(ev, st) do { if (!single) CU(cudaEventRecord(ev, st)); } while (0)
    const int slot = (int)(h->step % (unsigned long long)h->depth);
    const int parity = (int)(h->step & 1ull);
    Slot& sl = h->slots[slot];
    const bool keep = h->cfg.keep_intermediates != 0;

    // ---- stage A: K1 + K2 ----
    EV_WAIT(stA, sl.ev_C);        // previous user of this slot's A/B buffers
    if (wait_H) EV_WAIT(stA, sl.ev_H);
    {
        fm::K1Params p{};
        std::memcpy(p.taps, h->taps.fm_in, sizeof(p.taps));
        // DC term of the u8 kernel: the float the kernel's own FMA sequence yields for a window of bytes 127
        // (each step one correctly rounded fma, reproduced here in double: a 24-bit float plus an exact
        // 31-bit product fits 53 bits).  A silent capture (all bytes 127) then gives EXACT zeros like the
        // reference's (float)u8 - 127.0f, and so does the zero history of a stream's first block.
        auto dc127 = [&](int k0) {
            float a = 0.0f;
            for (int k = k0; k < fm::K1_NN; k++) a = (float)((double)a + (double)p.taps_s[k] * (127.0 * 0x1p-149));
            return -(a * fm::K1_UNSCALE);
        };
        for (int k = 0; k < fm::K1_NN; k++) p.taps_s[k] = p.taps[k] * fm::K1_TAP_SCALE;
        p.neg_dc = dc127(0);
        for (int r = 0; r < fm::K1_R; r++) p.first_neg_dc[r] = dc127(fm::K1_NN - 4 * (r + 1));
        // fm_demod.cpp:36-39 with Fd = 75 kHz, Fs = 256 kHz (broadcast_fm_demod.cpp:396-398)
        const float Wd = 75e3f * 2.0f * 3.14159265358979323846f;
        const float Ts = 1.0f / 256000.0f;
        p.discrim_gain = 1.0f / (Wd * Ts) * 0.5f;
        p.n_out = h->n4; p.parity = parity; p.n_streams = h->S; p.first_block = (h->step == 0); p.dbg_fm_in = keep ? h->dbg.fm_in : nullptr;
        if (prof) CU(cudaEventRecord(prof[0], stA));
        if (u8 && h->use_k1t) {
            // tensor-core path (tables ensured by the caller: ensure_k1t_tables)
            fm::K1TParams t{};
            t.bimg = h->k1t_bimg; t.ptab = h->k1t_ptab;
            for (int i = 0; i < 3; i++) { t.off[i] = h->k1t_off[i]; t.w[i] = h->k1t_w[i]; }
            t.discrim_gain = p.discrim_gain;
            t.n_rows = h->B / 64; t.tiles_per_stream = (t.n_rows + 127) / 128; t.n_tiles = t.tiles_per_stream * h->S; t.n_streams = h->S;
            t.theta_in = h->k1t_theta[parity]; t.theta_out = h->k1t_theta[parity ^ 1]; t.dbg_fm_in = p.dbg_fm_in;
            t.shape = h->k1t_shape;
            CU(fm::launch_k1t((const uint8_t*)iq_dev, h->k1t_hist[parity], h->k1t_hist[parity ^ 1], h->k1_hist[parity ^ 1], sl.fm_demod,
                              t, 2 * h->n_sm_fir, stA));
        } else {
            CU(fm::launch_k1(u8, iq_dev, h->k1_hist[parity], h->k1_hist[parity ^ 1], sl.fm_demod, p, stA));
        }
    }
    {
        fm::K2Params p{};
        std::memcpy(p.taps_fm_out, h->taps.fm_out, sizeof(p.taps_fm_out));
        std::memcpy(p.taps_hilbert, h->taps.hilbert, sizeof(p.taps_hilbert));
        std::memcpy(p.deemph_b, h->taps.deemph_b, 8); std::memcpy(p.deemph_a, h->taps.deemph_a, 8);
        std::memcpy(p.peak_b, h->taps.peak_b, 12); std::memcpy(p.peak_a, h->taps.peak_a, 12);
        p.use_deemph = h->ctl_use_deemph; p.n_out = h->n8; p.keep = keep;
        fill_scan_matrices(p);
        if (prof) CU(cudaEventRecord(prof[1], stA));
        // K2 runs on its own stream of the FIR partition: it is one wave of long-lived CTAs (one per stream pair,
        // 3.9 per SM) that leaves half of the FMA pipe idle, and K1 of the NEXT block -- which only needs a free
        // slot, not K2 -- fills those issue slots instead of queueing behind it.
        static const bool k2_own = std::getenv("FMGPU_K2_SAME_STREAM") == nullptr;
        cudaStream_t st2 = k2_own ? stA2 : stA;
        if (k2_own && !single) { EV_REC(sl.ev_K1, stA); EV_WAIT(st2, sl.ev_K1); }
        CU(fm::launch_k2(sl.fm_demod, h->k2_hist_demod, h->k2_hist_out, h->k2_scal, sl.fm_out_iq, sl.theta, sl.power,
                         keep ? h->dbg.pilot : nullptr, p, h->S, st2));
        if (prof) CU(cudaEventRecord(prof[2], st2));
        EV_REC(sl.ev_A, st2);
    }

    // ---- stage B: K3 ----
    EV_WAIT(stB, sl.ev_A);
    {
        fm::K3Params p{};
        std::memcpy(p.lpf_b, h->taps.pll_b, 8); std::memcpy(p.lpf_a, h->taps.pll_a, 8);
        const float Ts = 1.0f / 128000.0f;
        p.int_KTs = 0.1f * Ts; p.Kp = 0.01f;
        p.f_center = -19000.0f; p.f_gain = -100.0f; p.mixer_KTs = Ts;
        p.agc_target = 1.0f; p.agc_beta = 0.2f;
        p.n = h->n8; p.n_streams = h->S; p.keep = keep; p.exact = h->k3_exact ? 1 : (h->k3_single ? 2 : (h->k3_duo ? 3 : 0)); p.rec_sms = h->sms_rec > 0 ? h->sms_rec : h->n_sm_fir;
        if (prof) CU(cudaEventRecord(prof[3], stB));
        CU(fm::launch_k3(sl.theta, sl.power, h->pll_state, sl.pll_dt, h->dbg.pll_raw, h->dbg.pll_pi, p, stB));
        if (prof) CU(cudaEventRecord(prof[4], stB));
        if (keep) CU(fm::launch_kdbg(h->dbg.pilot, h->pll_state, sl.pll_dt, h->dbg.pll, h->n8, h->S, stB));
    }
    EV_REC(sl.ev_B, stB);

    // ---- stage C: K4 + K4b ----
    EV_WAIT(stC, sl.ev_B);
    EV_WAIT(stC, sl.ev_D);        // K5 of the slot's previous block read sl.rds
    EV_WAIT(stC, sl.ev_O);        // ... and the fetch read sl.audio
    EV_WAIT(stC, sl.ev_P);        // ... and so did K7
    {
        fm::K4Params p{};
        std::memcpy(p.taps_lpr, h->taps.lpr, sizeof(p.taps_lpr));
        std::memcpy(p.taps_lmr, h->taps.lmr, sizeof(p.taps_lmr));
        std::memcpy(p.taps_rds, h->taps.rds, sizeof(p.taps_rds));
        p.harmonic_lmr = 38000.0f / 19000.0f; p.harmonic_rds = 57000.0f / 19000.0f;
        p.stereo_mix = h->ctl_stereo_mix; p.audio_out_mode = h->ctl_audio_out;
        p.n = h->n8; p.n_tiles = h->k4_tiles; p.parity = parity; p.n_streams = h->S; p.keep = keep;
        p.balanced = h->k4_v1 ? 0 : 1;
        if (keep) CU(cudaMemcpyAsync(h->dbg.lmr_phase_used, h->lmr_phase, (size_t)h->S * sizeof(float), cudaMemcpyDeviceToDevice, stC));
        if (prof) CU(cudaEventRecord(prof[5], stC));
        CU(fm::launch_k4(sl.fm_out_iq, sl.pll_dt, h->k4_hist_x[parity], h->k4_hist_m2[parity], h->k4_hist_m3[parity],
                         h->k4_hist_x[parity ^ 1], h->k4_hist_m2[parity ^ 1], h->k4_hist_m3[parity ^ 1],
                         h->lmr_phase, sl.audio, sl.rds, sl.est_partial, sl.rds_pw_partial,
                         h->dbg.lpr, h->dbg.lmr, p, stC));
        if (keep) {
            // GUI mode: the complex decimator outputs behind the two audio spectra, then this block's last 128
            // fm_out_iq samples as the next block's history (before ev_C lets K2 overwrite the slot)
            CU(fm::launch_kdbg_audio_iq(sl.fm_out_iq, sl.pll_dt, h->dbg.hist_iq, h->k4_hist_m2[parity], h->dbg.lmr_phase_used,
                                        h->dbg.lpr_iq, h->dbg.lmr_iq, p, stC));
            CU(cudaMemcpy2DAsync(h->dbg.hist_iq, fm::K4_NN * sizeof(float2), sl.fm_out_iq + (h->n8 - fm::K4_NN), (size_t)h->n8 * sizeof(float2),
                                 fm::K4_NN * sizeof(float2), (size_t)h->S, cudaMemcpyDeviceToDevice, stC));
        }
    }
    if (prof) CU(cudaEventRecord(prof[6], stC));
    EV_REC(sl.ev_C, stC);
    // ---- K7 (audio output stage, off by default): on stage stream C behind ev_C, so the RDS stages do not wait for
    //      it.  Measured alternatives (1024 streams, ms per step): its own stream of the FIR partition
    //      (FMGPU_K7_OWN_STREAM) 0.291 vs 0.2805 here -- a fourth concurrent FIR-partition kernel only takes CTA slots
    //      from K1; on the recurrence partition (FMGPU_K7_ON_REC as well) it runs 0.068 ms instead of 0.015 and slows
    //      K5 by 10 %: 0.32. ----
    if (h->ctl_pcm_rate > 0) {
        const int rc7 = prepare_pcm(h);
        if (rc7 != FMGPU_OK) return rc7;
        static const bool k7_own = std::getenv("FMGPU_K7_OWN_STREAM") != nullptr;
        cudaStream_t st7 = k7_own ? stP : stC;
        if (k7_own && !single) {
            EV_WAIT(st7, sl.ev_C);
            EV_WAIT(st7, sl.ev_O);   // the fetch of the slot's previous block read sl.pcm_s16
        }
        if (prof) CU(cudaEventRecord(prof[11], st7));
        CU(fm::launch_k7(sl.audio, h->pcm_table, sl.pcm_f32, sl.pcm_s16, h->n32, h->pcm_n, h->S, st7));
        if (prof) CU(cudaEventRecord(prof[10], st7));
        EV_REC(sl.ev_P, st7);
    } else if (prof) {
        CU(cudaEventRecord(prof[11], stC));
        CU(cudaEventRecord(prof[10], stC));
    }

    // ---- stage D: K5 ----
    EV_WAIT(stD, sl.ev_C);
    EV_WAIT(stD, sl.ev_O);        // the fetch read sl.pred_sym / sl.sym_count
    EV_WAIT(stD, sl.ev_E);        // ... and so did K6 of the slot's previous block
    {
        fm::K5Params p{};
        std::memcpy(p.ted_b, h->taps.ted_b, 8); std::memcpy(p.ted_a, h->taps.ted_a, 8);
        std::memcpy(p.pll_b, h->taps.bpll_b, 8); std::memcpy(p.pll_a, h->taps.bpll_a, 8);
        // bpsk_synchroniser.cpp:51-91 with the defaults of bpsk_synchroniser.h:18-32
        const float Fs = 16e3f, Fsym = 2e3f, Ts = 1.0f / Fs;
        const int sps = (int)std::round(Fs / Fsym);
        p.cooldown_N = sps / 2;
        p.dump_KTs = 1.0f / (0.5f * (float)sps * 1.0f);
        p.ted_KTs = Ts; p.ted_fcenter = Fsym; p.ted_fgain = 1.5e3f;
        p.mixer_KTs = Ts; p.mixer_fgain = 10.0f;
        const float k = Fsym / Fs;
        p.int_ted_KTs = 10.0f * Ts * k; p.int_pll_KTs = 10.0f * Ts * k;
        p.ted_Kp = 0.3f; p.pll_Kp = 0.3f;
        p.agc_target = 0.5f; p.agc_beta = 0.2f;
        p.n = h->n64; p.n_tiles_k4 = h->k4_tiles; p.n_streams = h->S; p.keep = keep; p.literal = h->k5_literal ? 1 : 0;
        if (prof) CU(cudaEventRecord(prof[7], stD));
        CU(fm::launch_k5(sl.rds, sl.rds_pw_partial, h->bpsk_state, sl.pred_sym, sl.sym_count, h->dbg.k5, p, stD));
    }
    if (prof) CU(cudaEventRecord(prof[8], stD));
    EV_REC(sl.ev_D, stD);
    // ---- stage E: K6 RDS bit path (symbols -> groups -> PI/PS/RT), per-stream state on the device.  Its own
    //      stream, so that K6 of block k overlaps K5 of block k+1 (both are latency-bound one-warp CTAs) ----
    EV_WAIT(stE, sl.ev_D);
    static const bool no_k6 = std::getenv("FMGPU_NO_K6") != nullptr;      // measurement aid
    if (!no_k6) CU(fm::launch_k6(sl.pred_sym, sl.sym_count, h->rds_state, h->rds_tables, h->rds_glog, h->rds_blog, h->n64, h->rds_gcap, h->rds_bcap, h->S, stE));
    if (prof) CU(cudaEventRecord(prof[9], stE));
    EV_REC(sl.ev_E, stE);
#undef EV_WAIT
#undef EV_REC
    return slot;
}

// One block: the multi-stream pipeline, or -- for small launch-bound blocks (h->use_graph) -- the same kernels replayed
// as one CUDA graph per (ring slot, history parity) on a single stream: ~10 API calls instead of ~60 per block.  A graph
// is re-captured when its input pointer, input kind or anything baked into the kernel parameters (controls, taps,
// options: h->graph_gen) has changed.
int enqueue_chain(fmgpu_demod* h, const void* iq_dev, bool u8, bool wait_H, cudaEvent_t* prof = nullptr) {
    // K1's FIR history is kept as bytes by the u8 kernels (exact) and as floats by the cf32 kernel; the u8 kernels also
    // write the float copy, so cf32 may follow u8, but arbitrary floats cannot become bytes again
    if (u8 && h->last_input_kind == 2)
        return fail(FMGPU_ERR_STATE, "enqueue: a u8 block cannot follow a cf32 block on the same handle (the FIR history would be truncated to bytes)");
    h->last_input_kind = u8 ? 1 : 2;
    update_filters(h);
    if (u8 && h->use_k1t) { const int rc = ensure_k1t_tables(h); if (rc != FMGPU_OK) return rc; }
    if (h->ctl_pcm_rate > 0) { const int rc = prepare_pcm(h); if (rc != FMGPU_OK) return rc; }
    const int slot = (int)(h->step % (unsigned long long)h->depth);
    const bool graph = h->use_graph && !prof && h->step > 0;
    if (graph != h->last_was_graph) {                   // cross-block state is ordered by stream order within one mode only
        if (sync_all(h) != FMGPU_OK) return FMGPU_ERR_CUDA;
        h->last_was_graph = graph;
    }
    int n_launch = 7 + (h->ctl_pcm_rate > 0 ? 1 : 0) + (h->cfg.keep_intermediates ? 2 : 0);
    if (!graph) {
        const int rc = enqueue_chain_on(h, iq_dev, u8, wait_H, prof, nullptr);
        if (rc < 0) return rc;
        h->slots[slot].fetched_in_graph = false;
    } else {
        Slot& sl = h->slots[slot];
        GraphEntry& ge = h->graphs[(size_t)slot * 2 + (size_t)(h->step & 1ull)];
        if (!ge.exec || ge.in_ptr != iq_dev || ge.u8 != u8 || ge.gen != h->graph_gen) {
            if (ge.exec) { cudaGraphExecDestroy(ge.exec); ge.exec = nullptr; }
            cudaGraph_t g = nullptr;
            CU(cudaStreamBeginCapture(h->stG, cudaStreamCaptureModeThreadLocal));
            int rc = enqueue_chain_on(h, iq_dev, u8, false, nullptr, h->stG);
            if (rc >= 0 && fetch_copies(h, slot, h->fetch_mask, h->stG) != FMGPU_OK) rc = FMGPU_ERR_CUDA;   // the fetch rides in the graph
            const cudaError_t ec = cudaStreamEndCapture(h->stG, &g);
            if (rc < 0) { if (g) cudaGraphDestroy(g); return rc; }
            if (ec != cudaSuccess) return fail(FMGPU_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ec));
            const cudaError_t ei = cudaGraphInstantiate(&ge.exec, g, 0);
            cudaGraphDestroy(g);
            if (ei != cudaSuccess) { ge.exec = nullptr; return fail(FMGPU_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ei)); }
            ge.in_ptr = iq_dev; ge.u8 = u8; ge.gen = h->graph_gen;
        }
        // previous users of the slot's buffers (the fetch of its previous block), and this block's input
        CU(cudaStreamWaitEvent(h->stG, sl.ev_O, 0));
        if (wait_H) CU(cudaStreamWaitEvent(h->stG, sl.ev_H, 0));
        CU(cudaGraphLaunch(ge.exec, h->stG));
        cudaEvent_t evs[7] = { sl.ev_A, sl.ev_B, sl.ev_C, sl.ev_D, sl.ev_E, sl.ev_P, sl.ev_K1 };
        for (auto ev : evs) CU(cudaEventRecord(ev, h->stG));
        sl.fetched_in_graph = true;
    }
    h->launches += n_launch;
    h->step++;
    h->dbg_valid = false;
    return slot;
}

// The device -> host copies of one slot's outputs (what fmgpu_set_fetch_mask selects), on stream st.
int fetch_copies(fmgpu_demod* h, int slot, unsigned mask, cudaStream_t st) {
    Slot& sl = h->slots[slot];
    HostMirror& m = h->mirrors[slot];
    const size_t S = h->S;
    if (mask & FMGPU_FETCH_AUDIO_F32)
        CU(cudaMemcpyAsync(m.audio, sl.audio, S * h->n32 * sizeof(float2), cudaMemcpyDeviceToHost, st));
    if (mask & FMGPU_FETCH_RDS_SYMBOLS) {
        // only the part of each stream's row that can hold symbols: the timing clock runs at <= f_center + f_gain =
        // 3500 Hz (ted_clock.cpp:31-44), i.e. at most n64 * 3500 / 16000 + 1 dumps per block (~152 at lock)
        const size_t cap = std::min<size_t>((size_t)h->n64, (size_t)h->n64 * 7 / 32 + 2);
        CU(cudaMemcpy2DAsync(m.pred_sym, (size_t)h->n64 * sizeof(float), sl.pred_sym, (size_t)h->n64 * sizeof(float),
                             cap * sizeof(float), S, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaMemcpyAsync(m.sym_count, sl.sym_count, S * sizeof(int), cudaMemcpyDeviceToHost, st));
    const bool pcm_on = h->pcm_rate_built > 0 && h->ctl_pcm_rate == h->pcm_rate_built;
    if (pcm_on && (mask & FMGPU_FETCH_PCM_S16))
        CU(cudaMemcpyAsync(m.pcm_s16, sl.pcm_s16, S * h->pcm_n * sizeof(short2), cudaMemcpyDeviceToHost, st));
    return FMGPU_OK;
}

int fetch_slot(fmgpu_demod* h, int slot) {
    Slot& sl = h->slots[slot];
    const unsigned mask = h->fetch_mask;
    if (sl.fetched_in_graph) {
        // the block was replayed as a CUDA graph whose last nodes are these copies (same mask: it is part of graph_gen)
        sl.fetched_in_graph = false;
        CU(cudaStreamWaitEvent(h->stO, sl.ev_D, 0));
        CU(cudaEventRecord(sl.ev_O, h->stO));
    } else {
        CU(cudaStreamWaitEvent(h->stO, sl.ev_D, 0));
        const bool pcm_on = h->pcm_rate_built > 0 && h->ctl_pcm_rate == h->pcm_rate_built;
        if (pcm_on && (mask & FMGPU_FETCH_PCM_S16)) CU(cudaStreamWaitEvent(h->stO, sl.ev_P, 0));
        const int rc = fetch_copies(h, slot, mask, h->stO);
        if (rc != FMGPU_OK) return rc;
        CU(cudaEventRecord(sl.ev_O, h->stO));
    }
    h->last_fetched_slot = slot;
    h->last_fetch_mask = mask;
    return FMGPU_OK;
}

int sync_all(fmgpu_demod* h) {
    CU(cudaStreamSynchronize(h->stH));
    CU(cudaStreamSynchronize(h->stA));
    CU(cudaStreamSynchronize(h->stA2));
    CU(cudaStreamSynchronize(h->stP));
    CU(cudaStreamSynchronize(h->stB));
    CU(cudaStreamSynchronize(h->stC));
    CU(cudaStreamSynchronize(h->stD));
    CU(cudaStreamSynchronize(h->stE));
    CU(cudaStreamSynchronize(h->stO));
    if (h->stG) CU(cudaStreamSynchronize(h->stG));
    return FMGPU_OK;
}

struct BufInfo { size_t elem; int per_stream; };   // element bytes, elements per stream

} // namespace

extern "C" {

const char* fmgpu_last_error(void) { return g_last_error.c_str(); }
const char* fmgpu_version(void) { return "fm-radio-b200 0.1 (sm_100a)"; }

int fmgpu_create(const fmgpu_config* cfg, fmgpu_demod** out) {
    if (!cfg || !out) return fail(FMGPU_ERR_ARG, "fmgpu_create: null argument");
    *out = nullptr;
    const int B = cfg->block_size;
    if (B < 1024 || (B & (B - 1)) != 0) return fail(FMGPU_ERR_ARG, "fmgpu_create: block_size must be a power of two >= 1024");
    if (cfg->n_streams < 1) return fail(FMGPU_ERR_ARG, "fmgpu_create: n_streams must be >= 1");
    if (cfg->n_streams > 65535) return fail(FMGPU_ERR_ARG, "fmgpu_create: n_streams must be <= 65535 per handle (streams ride in gridDim.y)");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(FMGPU_ERR_CUDA, std::string("fmgpu_create: no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    int dev = cfg->device;
    if (dev < 0) CU(cudaGetDevice(&dev));
    if (dev >= n_dev) return fail(FMGPU_ERR_ARG, "fmgpu_create: bad device ordinal");
    CU(cudaSetDevice(dev));
    cudaDeviceProp prop{};
    CU(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(FMGPU_ERR_CUDA, "fmgpu_create: kernels are built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
    // Round 1 measured (driver 580.159, tools/bisect_bench.py) that the FIRST handle whose green contexts are created in a
    // process ran its FIR partition 6-7 % slower than every later one, and worked around it by building and dropping a
    // scratch handle first.  Round 2's bisection (FMGPU_PRIME_MODE, profiles/r2_prime_modes.md): the scratch handle need
    // not be destroyed, green contexts + streams alone give part of the effect -- and with the round-2 kernels (K1 on the
    // tensor cores) the effect is gone altogether: 0.2469 ms per step without priming, 0.2468 with.  Priming is therefore
    // OFF by default; FMGPU_PRIME=1 (or an explicit FMGPU_PRIME_MODE) brings the old behaviour back for measurements.
    static std::atomic<bool> primed[64];                    // per device ordinal (zero-initialised)
    if ((std::getenv("FMGPU_PRIME") || std::getenv("FMGPU_PRIME_MODE")) && !std::getenv("FMGPU_NO_PARTITION") && dev < 64 && !primed[dev].exchange(true)) {
        // FMGPU_PRIME_MODE (measurement aid for the bisection in DESIGN.md section 9): which part of the scratch handle matters
        const char* pm = std::getenv("FMGPU_PRIME_MODE");
        const std::string mode = pm ? pm : "full";
        if (mode == "gctx" || mode == "gctx_launch" || mode == "gctx_alloc" || mode == "gctx_events") {
            // green contexts + stage streams (+ one kernel on each partition / + device allocations / + events and pinned memory)
            fmgpu_demod tmp{};
            tmp.device = dev;
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            if (create_partitioned_streams(&tmp, hi)) {
                if (mode == "gctx_launch") {
                    void* st8 = nullptr;
                    if (cudaMalloc(&st8, 64 * fm::k6_state_bytes()) == cudaSuccess) {
                        fm::launch_k6_init(st8, 64, tmp.stD);
                        fm::launch_k6_init(st8, 64, tmp.stA);
                        cudaStreamSynchronize(tmp.stD); cudaStreamSynchronize(tmp.stA);
                        cudaFree(st8);
                    }
                }
                if (mode == "gctx_alloc") {
                    std::vector<void*> ps;
                    for (int i = 0; i < 40; i++) { void* q = nullptr; if (cudaMalloc(&q, 4096 + 4096 * i) == cudaSuccess) { cudaMemset(q, 0, 4096); ps.push_back(q); } }
                    cudaDeviceSynchronize();
                    for (void* q : ps) cudaFree(q);
                }
                if (mode == "gctx_events") {
                    std::vector<cudaEvent_t> evs(36);
                    for (auto& e : evs) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
                    std::vector<void*> hp(12, nullptr);
                    for (auto& q : hp) cudaMallocHost(&q, 4096);
                    for (auto& e : evs) { cudaEventRecord(e, tmp.stA); cudaEventDestroy(e); }
                    cudaStreamSynchronize(tmp.stA);
                    for (auto& q : hp) if (q) cudaFreeHost(q);
                }
                cudaStream_t sts[7] = { tmp.stA, tmp.stA2, tmp.stP, tmp.stB, tmp.stC, tmp.stD, tmp.stE };
                for (auto st : sts) if (st) cudaStreamDestroy(st);
                destroy_partition(&tmp);
            }
        } else if (mode == "alloc") {                       // device + pinned allocations only
            void* d = nullptr; void* hp = nullptr;
            if (cudaMalloc(&d, 64 << 20) == cudaSuccess) { cudaMemset(d, 0, 64 << 20); cudaDeviceSynchronize(); cudaFree(d); }
            if (cudaMallocHost(&hp, 1 << 20) == cudaSuccess) cudaFreeHost(hp);
        } else if (mode != "none") {
            fmgpu_config sc{};
            sc.block_size = 1024; sc.n_streams = 1; sc.device = dev; sc.pipeline_depth = 1;
            fmgpu_demod* scratch = nullptr;
            if (fmgpu_create(&sc, &scratch) == FMGPU_OK && mode != "nodestroy") fmgpu_destroy(scratch);
        }
        CU(cudaSetDevice(dev));
    }
    auto* h = new fmgpu_demod();
    h->cfg = *cfg; h->cfg.device = dev;
    h->device = dev;
    h->B = B; h->S = cfg->n_streams;
    h->n4 = B / 4; h->n8 = B / 8; h->n32 = B / 32; h->n64 = B / 64;
    h->depth = cfg->pipeline_depth > 0 ? cfg->pipeline_depth : 4;
    // the GUI buffers exist once (not per ring slot): with them on, blocks are not overlapped
    if (cfg->keep_intermediates) h->depth = 1;
    h->cfg.pipeline_depth = h->depth;
    h->k4_tiles = (h->n8 + fm::K4_TS - 1) / fm::K4_TS;
    design_default_taps(h);
    update_filters(h);
    const int rc = alloc_all(h);
    if (rc != FMGPU_OK) { const std::string msg = g_last_error; free_all(h); delete h; g_last_error = msg; return rc; }
    *out = h;
    return FMGPU_OK;
}

void fmgpu_destroy(fmgpu_demod* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    free_all(h);
    delete h;
}

static int process_host(fmgpu_demod* h, const void* iq_host, size_t n_samples, bool u8) {
    if (!h || !iq_host) return fail(FMGPU_ERR_ARG, "process: null argument");
    if (n_samples != (size_t)h->B) return fail(FMGPU_ERR_SIZE, "process: n_samples != block_size");
    CU(cudaSetDevice(h->device));
    const int slot = (int)(h->step % (unsigned long long)h->depth);
    Slot& sl = h->slots[slot];
    const void* dev_in;
    if (u8) {
        const size_t bytes = (size_t)h->S * h->B * 2;
        if (!sl.in_u8) CU(cudaMalloc((void**)&sl.in_u8, bytes));
        CU(cudaStreamWaitEvent(h->stH, sl.ev_A, 0));     // K1 of the slot's previous block read in_u8
        CU(cudaMemcpyAsync(sl.in_u8, iq_host, bytes, cudaMemcpyHostToDevice, h->stH));
        dev_in = sl.in_u8;
    } else {
        const size_t bytes = (size_t)h->S * h->B * sizeof(float2);
        if (!h->in_f32) CU(cudaMalloc((void**)&h->in_f32, bytes));
        CU(cudaStreamSynchronize(h->stA));
        CU(cudaMemcpyAsync(h->in_f32, iq_host, bytes, cudaMemcpyHostToDevice, h->stH));
        dev_in = h->in_f32;
    }
    CU(cudaEventRecord(sl.ev_H, h->stH));
    const int rc = enqueue_chain(h, dev_in, u8, true);
    if (rc < 0) return rc;
    return FMGPU_OK;
}

int fmgpu_process_u8(fmgpu_demod* h, const uint8_t* iq_host, size_t n_samples) {
    int rc = process_host(h, iq_host, n_samples, true);
    if (rc != FMGPU_OK) return rc;
    rc = fetch_slot(h, (int)((h->step - 1) % (unsigned long long)h->depth));
    if (rc != FMGPU_OK) return rc;
    return sync_all(h);
}

int fmgpu_process_cf32(fmgpu_demod* h, const float* iq_host, size_t n_samples) {
    int rc = process_host(h, iq_host, n_samples, false);
    if (rc != FMGPU_OK) return rc;
    rc = fetch_slot(h, (int)((h->step - 1) % (unsigned long long)h->depth));
    if (rc != FMGPU_OK) return rc;
    return sync_all(h);
}

int fmgpu_enqueue_u8_host(fmgpu_demod* h, const uint8_t* iq_pinned_host) {
    if (!h) return fail(FMGPU_ERR_ARG, "enqueue: null handle");
    return process_host(h, iq_pinned_host, (size_t)h->B, true);
}

int fmgpu_enqueue_u8_device(fmgpu_demod* h, const uint8_t* iq_dev) {
    if (!h || !iq_dev) return fail(FMGPU_ERR_ARG, "enqueue: null argument");
    CU(cudaSetDevice(h->device));
    const int rc = enqueue_chain(h, iq_dev, true, false);
    return rc < 0 ? rc : FMGPU_OK;
}

// cf32 input already on the device (the channelizer's output): K1's cf32 variant reads it in place.
// after_stream: CUDA stream whose queued work produces iq_dev; NULL = nothing to wait for (for the legacy
// default stream pass cudaStreamLegacy, not 0).
int fmgpu_enqueue_cf32_device(fmgpu_demod* h, const float* iq_dev, void* after_stream) {
    if (!h || !iq_dev) return fail(FMGPU_ERR_ARG, "enqueue: null argument");
    CU(cudaSetDevice(h->device));
    if (after_stream) {
        cudaEvent_t ev;
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CU(cudaEventRecord(ev, (cudaStream_t)after_stream));
        CU(cudaStreamWaitEvent(h->stA, ev, 0));
        CU(cudaStreamWaitEvent(h->stG, ev, 0));             // the block may be replayed as a graph on stG
        CU(cudaEventDestroy(ev));
    }
    const int rc = enqueue_chain(h, iq_dev, false, false);
    return rc < 0 ? rc : FMGPU_OK;
}

// Makes `cuda_stream` wait until the input of the block enqueued pipeline_depth enqueues ago has been
// consumed (its stage A is complete), i.e. until a producer that recycles its output buffers with the
// same period may overwrite the oldest one.
int fmgpu_stream_wait_input_free(fmgpu_demod* h, void* cuda_stream) {
    if (!h) return fail(FMGPU_ERR_ARG, "stream_wait_input_free: null handle");
    CU(cudaSetDevice(h->device));
    const int slot = (int)(h->step % (unsigned long long)h->depth);
    CU(cudaStreamWaitEvent((cudaStream_t)cuda_stream, h->slots[slot].ev_A, 0));
    return FMGPU_OK;
}

int fmgpu_set_last_error_(int code, const char* msg) { return fail(code, msg ? msg : ""); }

int fmgpu_profile_stages(fmgpu_demod* h, const uint8_t* iq_dev, int n_blocks, float ms[6]) {
    float m7[7];
    const int rc = fmgpu_profile_stages7(h, iq_dev, n_blocks, m7);
    if (rc == FMGPU_OK && ms) std::memcpy(ms, m7, 6 * sizeof(float));
    return rc;
}

int fmgpu_profile_stages7(fmgpu_demod* h, const uint8_t* iq_dev, int n_blocks, float ms[7]) {
    if (!h || !iq_dev || !ms || n_blocks == 0) return fail(FMGPU_ERR_ARG, "profile_stages: bad argument");
    CU(cudaSetDevice(h->device));
    // n_blocks > 0: blocks one at a time (kernel times in isolation).  n_blocks < 0: |n_blocks| blocks
    // enqueued back to back like the production path, so the times are those of the kernels while they
    // share the GPU with the other stages' kernels (averaged over the second half of the run).
    const bool piped = n_blocks < 0;
    const int n = piped ? -n_blocks : n_blocks;
    constexpr int NE = 12;
    std::vector<cudaEvent_t> ev((size_t)NE * n);
    for (auto& e : ev) CU(cudaEventCreate(&e));
    double acc[7] = { 0, 0, 0, 0, 0, 0, 0 };
    const int pairs[7][2] = { { 0, 1 }, { 1, 2 }, { 3, 4 }, { 5, 6 }, { 7, 8 }, { 8, 9 }, { 11, 10 } };
    if (sync_all(h) != FMGPU_OK) return FMGPU_ERR_CUDA;
    for (int b = 0; b < n; b++) {
        const int rc = enqueue_chain(h, iq_dev, true, false, &ev[(size_t)NE * b]);
        if (rc < 0) return rc;
        if (!piped && sync_all(h) != FMGPU_OK) return FMGPU_ERR_CUDA;
    }
    if (sync_all(h) != FMGPU_OK) return FMGPU_ERR_CUDA;
    const int b0 = piped ? n / 2 : 0;
    for (int b = b0; b < n; b++)
        for (int i = 0; i < 7; i++) {
            float t = 0.0f;
            CU(cudaEventElapsedTime(&t, ev[(size_t)NE * b + pairs[i][0]], ev[(size_t)NE * b + pairs[i][1]]));
            acc[i] += t;
        }
    for (int i = 0; i < 7; i++) ms[i] = (float)(acc[i] / (n - b0));
    for (auto& e : ev) cudaEventDestroy(e);
    return FMGPU_OK;
}

int fmgpu_sync(fmgpu_demod* h) {
    if (!h) return fail(FMGPU_ERR_ARG, "sync: null handle");
    CU(cudaSetDevice(h->device));
    return sync_all(h);
}

int fmgpu_fetch_outputs(fmgpu_demod* h, int slot) {
    if (!h || slot < 0 || slot >= h->depth) return fail(FMGPU_ERR_ARG, "fetch: bad slot");
    CU(cudaSetDevice(h->device));
    return fetch_slot(h, slot);
}

int fmgpu_set_fetch_mask(fmgpu_demod* h, unsigned mask) {
    if (!h) return fail(FMGPU_ERR_ARG, "set_fetch_mask: null handle");
    if (mask & ~(unsigned)FMGPU_FETCH_ALL) return fail(FMGPU_ERR_ARG, "set_fetch_mask: unknown bits");
    h->fetch_mask = mask;
    h->graph_gen++;
    return FMGPU_OK;
}

int fmgpu_wait_external_stream(fmgpu_demod* h, void* cuda_stream) {
    if (!h) return fail(FMGPU_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    cudaEvent_t ev;
    CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CU(cudaEventRecord(ev, (cudaStream_t)cuda_stream));
    cudaStream_t sts[10] = { h->stH, h->stA, h->stA2, h->stP, h->stB, h->stC, h->stD, h->stE, h->stO, h->stG };
    for (auto st : sts) CU(cudaStreamWaitEvent(st, ev, 0));
    CU(cudaEventDestroy(ev));
    return FMGPU_OK;
}

int fmgpu_signal_external_stream(fmgpu_demod* h, void* cuda_stream) {
    if (!h) return fail(FMGPU_ERR_ARG, "null handle");
    CU(cudaSetDevice(h->device));
    cudaStream_t sts[10] = { h->stH, h->stA, h->stA2, h->stP, h->stB, h->stC, h->stD, h->stE, h->stO, h->stG };
    for (auto st : sts) {
        cudaEvent_t ev;
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CU(cudaEventRecord(ev, st));
        CU(cudaStreamWaitEvent((cudaStream_t)cuda_stream, ev, 0));
        CU(cudaEventDestroy(ev));
    }
    return FMGPU_OK;
}

static bool buf_info(const fmgpu_demod* h, fmgpu_buffer b, BufInfo* bi, const void** dev, int slot) {
    const Slot& sl = h->slots[slot];
    const DebugBufs& d = h->dbg;
    const bool keep = h->cfg.keep_intermediates != 0;
    switch (b) {
    case FMGPU_BUF_AUDIO_OUT: *bi = { 8, h->n32 }; *dev = sl.audio; return true;
    case FMGPU_BUF_RDS_PRED_SYM: *bi = { 4, h->n64 }; *dev = sl.pred_sym; return true;
    case FMGPU_BUF_RDS_SYM_COUNT: *bi = { 4, 1 }; *dev = sl.sym_count; return true;
    case FMGPU_BUF_FM_DEMOD: *bi = { 4, h->n4 }; *dev = sl.fm_demod; return true;
    case FMGPU_BUF_FM_OUT_IQ: *bi = { 8, h->n8 }; *dev = sl.fm_out_iq; return true;
    case FMGPU_BUF_PLL_DT: *bi = { 4, h->n8 }; *dev = sl.pll_dt; return true;
    case FMGPU_BUF_AUDIO_PCM_F32: if (!sl.pcm_f32) return false; *bi = { 8, h->pcm_n }; *dev = sl.pcm_f32; return true;
    case FMGPU_BUF_AUDIO_PCM_S16: if (!sl.pcm_s16) return false; *bi = { 4, h->pcm_n }; *dev = sl.pcm_s16; return true;
    default: break;
    }
    if (!keep) return false;
    switch (b) {
    case FMGPU_BUF_PILOT: *bi = { 8, h->n8 }; *dev = d.pilot; return true;
    case FMGPU_BUF_PLL: *bi = { 8, h->n8 }; *dev = d.pll; return true;
    case FMGPU_BUF_PLL_RAW_PHASE_ERROR: *bi = { 4, h->n8 }; *dev = d.pll_raw; return true;
    case FMGPU_BUF_PLL_LPF_PHASE_ERROR: *bi = { 4, h->n8 }; *dev = d.pll_pi; return true;
    case FMGPU_BUF_AUDIO_LPR: *bi = { 4, h->n32 }; *dev = d.lpr; return true;
    case FMGPU_BUF_AUDIO_LMR: *bi = { 4, h->n32 }; *dev = d.lmr; return true;
    case FMGPU_BUF_FM_IN: *bi = { 8, h->n4 }; *dev = d.fm_in; return true;
    case FMGPU_BUF_AUDIO_LPR_IQ: *bi = { 8, h->n32 }; *dev = d.lpr_iq; return true;
    case FMGPU_BUF_AUDIO_LMR_IQ: *bi = { 8, h->n32 }; *dev = d.lmr_iq; return true;
    case FMGPU_BUF_RDS: *bi = { 8, h->n64 }; *dev = d.k5.rds; return true;
    case FMGPU_BUF_RDS_RAW_SYM: *bi = { 8, h->n64 }; *dev = d.k5.raw_sym; return true;
    case FMGPU_BUF_BPSK_PLL_SYM: *bi = { 8, h->n64 }; *dev = d.k5.pll_sym; return true;
    case FMGPU_BUF_BPSK_ZCD: *bi = { 1, h->n64 }; *dev = d.k5.zcd; return true;
    case FMGPU_BUF_BPSK_INT_DUMP_TRIGGER: *bi = { 1, h->n64 }; *dev = d.k5.dump_trig; return true;
    case FMGPU_BUF_BPSK_TED_RAW_PHASE_ERROR: *bi = { 4, h->n64 }; *dev = d.k5.ted_raw; return true;
    case FMGPU_BUF_BPSK_TED_PI_PHASE_ERROR: *bi = { 4, h->n64 }; *dev = d.k5.ted_pi; return true;
    case FMGPU_BUF_BPSK_PLL_RAW_PHASE_ERROR: *bi = { 4, h->n64 }; *dev = d.k5.pll_raw; return true;
    case FMGPU_BUF_BPSK_PLL_PI_PHASE_ERROR: *bi = { 4, h->n64 }; *dev = d.k5.pll_pi; return true;
    case FMGPU_BUF_BPSK_INT_DUMP_FILTER: *bi = { 8, h->n64 }; *dev = d.k5.dump_filter; return true;
    default: return false;
    }
}

int fmgpu_get_buffer(fmgpu_demod* h, int stream, fmgpu_buffer buf, const void** host_ptr, size_t* n_elems) {
    if (!h || !host_ptr || !n_elems) return fail(FMGPU_ERR_ARG, "get_buffer: null argument");
    if (stream < 0 || stream >= h->S) return fail(FMGPU_ERR_ARG, "get_buffer: bad stream");
    if (h->last_fetched_slot < 0) return fail(FMGPU_ERR_STATE, "get_buffer: nothing processed yet");
    const int slot = h->last_fetched_slot;
    HostMirror& m = h->mirrors[slot];
    const int count = m.sym_count[stream];
    switch (buf) {
    case FMGPU_BUF_AUDIO_OUT:
        if (!(h->last_fetch_mask & FMGPU_FETCH_AUDIO_F32)) return fail(FMGPU_ERR_STATE, "get_buffer: the last fetch left the f32 audio on the device (fmgpu_set_fetch_mask)");
        *host_ptr = m.audio + (size_t)stream * h->n32; *n_elems = h->n32; return FMGPU_OK;
    case FMGPU_BUF_RDS_PRED_SYM:
        if (!(h->last_fetch_mask & FMGPU_FETCH_RDS_SYMBOLS)) return fail(FMGPU_ERR_STATE, "get_buffer: the last fetch left the RDS symbols on the device (fmgpu_set_fetch_mask)");
        *host_ptr = m.pred_sym + (size_t)stream * h->n64; *n_elems = (size_t)count; return FMGPU_OK;
    case FMGPU_BUF_RDS_SYM_COUNT: *host_ptr = m.sym_count + stream; *n_elems = 1; return FMGPU_OK;
    case FMGPU_BUF_AUDIO_PCM_S16:
        if (!m.pcm_s16 || h->pcm_rate_built <= 0) return fail(FMGPU_ERR_STATE, "get_buffer: the audio output stage is off (FMGPU_CTL_AUDIO_PCM_RATE_HZ)");
        if (!(h->last_fetch_mask & FMGPU_FETCH_PCM_S16)) return fail(FMGPU_ERR_STATE, "get_buffer: the last fetch left the int16 PCM on the device (fmgpu_set_fetch_mask)");
        *host_ptr = m.pcm_s16 + (size_t)stream * h->pcm_n; *n_elems = (size_t)h->pcm_n; return FMGPU_OK;
    default: break;
    }
    // Everything else is copied on demand from the device (GUI / test path, not the hot path).
    BufInfo bi{}; const void* dev = nullptr;
    if (!buf_info(h, buf, &bi, &dev, slot))
        return fail(FMGPU_ERR_STATE, "get_buffer: buffer needs keep_intermediates = 1 (GUI buffers) or FMGPU_CTL_AUDIO_PCM_RATE_HZ > 0 (PCM buffers)");
    CU(cudaSetDevice(h->device));
    if (sync_all(h) != FMGPU_OK) return FMGPU_ERR_CUDA;
    const size_t bytes = bi.elem * (size_t)bi.per_stream;
    // host scratch: one region per buffer id, sized lazily
    const size_t max_bytes = 8ull * (size_t)h->n4;
    if (h->dbg_host.size() < (size_t)FMGPU_BUF__COUNT * max_bytes) h->dbg_host.resize((size_t)FMGPU_BUF__COUNT * max_bytes);
    uint8_t* dst = h->dbg_host.data() + (size_t)buf * max_bytes;
    CU(cudaMemcpy(dst, (const uint8_t*)dev + (size_t)stream * bytes, bytes, cudaMemcpyDeviceToHost));
    *host_ptr = dst;
    *n_elems = (buf == FMGPU_BUF_RDS_RAW_SYM) ? (size_t)count : (size_t)bi.per_stream;
    return FMGPU_OK;
}

int fmgpu_get_device_buffer(fmgpu_demod* h, int slot, fmgpu_buffer buf, void** dev_ptr, size_t* n_elems_per_stream) {
    if (!h || !dev_ptr || slot < 0 || slot >= h->depth) return fail(FMGPU_ERR_ARG, "get_device_buffer: bad argument");
    BufInfo bi{}; const void* dev = nullptr;
    if (!buf_info(h, buf, &bi, &dev, slot)) return fail(FMGPU_ERR_STATE, "get_device_buffer: buffer needs keep_intermediates = 1");
    *dev_ptr = (void*)dev;
    if (n_elems_per_stream) *n_elems_per_stream = (size_t)bi.per_stream;
    return FMGPU_OK;
}

int fmgpu_get_scalar(fmgpu_demod* h, int stream, fmgpu_scalar which, float* out) {
    if (!h || !out || stream < 0 || stream >= h->S) return fail(FMGPU_ERR_ARG, "get_scalar: bad argument");
    CU(cudaSetDevice(h->device));
    if (sync_all(h) != FMGPU_OK) return FMGPU_ERR_CUDA;
    const float* src = nullptr;
    switch (which) {
    case FMGPU_SCALAR_AUDIO_LMR_PHASE_ERROR: src = h->lmr_phase + stream; break;
    case FMGPU_SCALAR_AGC_PILOT_GAIN: src = h->pll_state + (size_t)fm::PLL_AGC_GAIN * h->S + stream; break;
    case FMGPU_SCALAR_AGC_RDS_GAIN: src = h->bpsk_state + (size_t)fm::BP_AGC_GAIN * h->S + stream; break;
    default: return fail(FMGPU_ERR_ARG, "get_scalar: unknown id");
    }
    CU(cudaMemcpy(out, src, sizeof(float), cudaMemcpyDeviceToHost));
    return FMGPU_OK;
}

int fmgpu_set_control(fmgpu_demod* h, fmgpu_control which, double value) {
    if (!h) return fail(FMGPU_ERR_ARG, "set_control: null handle");
    switch (which) {
    case FMGPU_CTL_AUDIO_OUT:
        if ((int)value < 0 || (int)value > 2) return fail(FMGPU_ERR_ARG, "set_control: audio_out must be 0..2");
        h->ctl_audio_out = (int)value; break;
    case FMGPU_CTL_AUDIO_STEREO_MIX_FACTOR: h->ctl_stereo_mix = (float)value; break;
    case FMGPU_CTL_USE_DEEMPHASIS: h->ctl_use_deemph = value != 0.0; break;
    case FMGPU_CTL_DEEMPHASIS_TUS: h->ctl_deemph_tus = (int)value; h->dirty_deemph = true; break;
    case FMGPU_CTL_AUDIO_LPR_CUTOFF_HZ: h->ctl_lpr_hz = (int)value; h->dirty_lpr = true; break;
    case FMGPU_CTL_AUDIO_LMR_CUTOFF_HZ: h->ctl_lmr_hz = (int)value; h->dirty_lmr = true; break;
    case FMGPU_CTL_AUDIO_PCM_RATE_HZ:
        if ((int)value != 0 && ((int)value < 8000 || (int)value > 192000)) return fail(FMGPU_ERR_ARG, "set_control: audio PCM rate must be 0 (off) or 8000..192000 Hz");
        h->ctl_pcm_rate = (int)value; break;
    default: return fail(FMGPU_ERR_ARG, "set_control: unknown id");
    }
    h->graph_gen++;
    return FMGPU_OK;
}

static bool taps_ptr(fmgpu_demod* h, fmgpu_filter which, float** b, float** a, int* n) {
    HostTaps& t = h->taps;
    *a = nullptr;
    switch (which) {
    case FMGPU_FILT_FM_IN: *b = t.fm_in; *n = 64; return true;
    case FMGPU_FILT_FM_OUT: *b = t.fm_out; *n = 64; return true;
    case FMGPU_FILT_HILBERT: *b = t.hilbert; *n = 65; return true;
    case FMGPU_FILT_AUDIO_LPR: *b = t.lpr; *n = 128; return true;
    case FMGPU_FILT_AUDIO_LMR: *b = t.lmr; *n = 128; return true;
    case FMGPU_FILT_RDS: *b = t.rds; *n = 128; return true;
    case FMGPU_FILT_DEEMPHASIS: *b = t.deemph_b; *a = t.deemph_a; *n = 2; return true;
    case FMGPU_FILT_PEAK_PILOT: *b = t.peak_b; *a = t.peak_a; *n = 3; return true;
    case FMGPU_FILT_PLL_LPF: *b = t.pll_b; *a = t.pll_a; *n = 2; return true;
    case FMGPU_FILT_BPSK_TED_LPF: *b = t.ted_b; *a = t.ted_a; *n = 2; return true;
    case FMGPU_FILT_BPSK_PLL_LPF: *b = t.bpll_b; *a = t.bpll_a; *n = 2; return true;
    default: return false;
    }
}

int fmgpu_upload_taps(fmgpu_demod* h, fmgpu_filter which, const float* b, const float* a, int n) {
    if (!h || !b) return fail(FMGPU_ERR_ARG, "upload_taps: null argument");
    float *hb, *ha; int hn;
    if (!taps_ptr(h, which, &hb, &ha, &hn)) return fail(FMGPU_ERR_ARG, "upload_taps: unknown filter");
    if (n != hn) return fail(FMGPU_ERR_ARG, "upload_taps: wrong length");
    if (ha && !a) return fail(FMGPU_ERR_ARG, "upload_taps: IIR filter needs a[]");
    update_filters(h);          // settle pending redesigns first so the upload is not overwritten
    std::memcpy(hb, b, sizeof(float) * n);
    if (ha) std::memcpy(ha, a, sizeof(float) * n);
    h->graph_gen++;
    return FMGPU_OK;
}

int fmgpu_download_taps(fmgpu_demod* h, fmgpu_filter which, float* b, float* a, int n) {
    if (!h || !b) return fail(FMGPU_ERR_ARG, "download_taps: null argument");
    float *hb, *ha; int hn;
    if (!taps_ptr(h, which, &hb, &ha, &hn)) return fail(FMGPU_ERR_ARG, "download_taps: unknown filter");
    if (n != hn) return fail(FMGPU_ERR_ARG, "download_taps: wrong length");
    update_filters(h);
    std::memcpy(b, hb, sizeof(float) * n);
    if (ha && a) std::memcpy(a, ha, sizeof(float) * n);
    return FMGPU_OK;
}

int fmgpu_get_rates(fmgpu_demod* h, int rates_hz[5]) {
    if (!h || !rates_hz) return fail(FMGPU_ERR_ARG, "get_rates: null argument");
    rates_hz[0] = 1024000; rates_hz[1] = 256000; rates_hz[2] = 128000; rates_hz[3] = 16000; rates_hz[4] = 32000;
    return FMGPU_OK;
}

int fmgpu_get_config(fmgpu_demod* h, fmgpu_config* out) {
    if (!h || !out) return fail(FMGPU_ERR_ARG, "get_config: null argument");
    *out = h->cfg;
    return FMGPU_OK;
}

long long fmgpu_launch_count(fmgpu_demod* h) { return h ? h->launches : 0; }

// Implementation switches for A/B measurements and cross-checks (the defaults are the production path):
//   "k1_fp32"     1: the u8 FIR + discriminator on the FP32 FMA pipe (k1_fir4_discrim_u8) instead of the tensor cores
//   "k5_literal"  1: the BPSK synchroniser's per-sample loop instead of the symbol-wise loop (identical bits)
//   "k3_exact"    1: the pilot PLL's exact body only, without the fast pass (k3_pll.cu)
//   "k3_single"   1: always the fast pass in one warp;  "k3_duo" 1: always recurrence warp + helper warp (bit-identical; by default
//                 the helper-warp version runs when its grid fits the recurrence partition one CTA per SM, k3_pll.cu)
//   "graph"       1 / 0: force the CUDA-graph replay of the per-block chain on / off (default: on for small launch-bound blocks)
//   "k4_v1"       1: the mixdown + FIR kernel reads its FIR taps from shared memory (first version) instead of the constant bank
int fmgpu_set_option(fmgpu_demod* h, const char* name, int value) {
    if (!h || !name) return fail(FMGPU_ERR_ARG, "set_option: null argument");
    const std::string n = name;
    if (n == "k1_fp32") {
        if (h->step != 0 && h->use_k1t != (value == 0)) return fail(FMGPU_ERR_STATE, "set_option: k1_fp32 must be chosen before the first block");
        h->use_k1t = value == 0;
    } else if (n == "k5_literal") h->k5_literal = value != 0;
    else if (n == "k3_exact") h->k3_exact = value != 0;
    else if (n == "k3_single") h->k3_single = value != 0;
    else if (n == "k3_duo") h->k3_duo = value != 0;
    else if (n == "k4_v1") h->k4_v1 = value != 0;
    else if (n == "graph") h->use_graph = value != 0;
    else return fail(FMGPU_ERR_ARG, "set_option: unknown option");
    h->graph_gen++;
    return FMGPU_OK;
}

int fmgpu_get_partition(fmgpu_demod* h, int sms[2]) {
    if (!h || !sms) return fail(FMGPU_ERR_ARG, "get_partition: null argument");
    sms[0] = h->sms_rec; sms[1] = h->sms_fir;
    return FMGPU_OK;
}

// ---- device RDS decoders (K6) ----
int fmgpu_rds_device_fetch(fmgpu_demod* h) {
    if (!h) return fail(FMGPU_ERR_ARG, "rds_device_fetch: null handle");
    CU(cudaSetDevice(h->device));
    if (sync_all(h) != FMGPU_OK) return FMGPU_ERR_CUDA;
    const size_t S = h->S;
    h->rds_host_state.resize(S);
    h->rds_host_glog.resize(S * h->rds_gcap);
    h->rds_host_blog.resize(S * h->rds_bcap);
    CU(cudaMemcpy(h->rds_host_state.data(), h->rds_state, S * sizeof(rds::State), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->rds_host_glog.data(), h->rds_glog, S * h->rds_gcap * sizeof(fmgpu_rds_group), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h->rds_host_blog.data(), h->rds_blog, S * h->rds_bcap, cudaMemcpyDeviceToHost));
    return FMGPU_OK;
}

static const rds::State* rds_fetched(fmgpu_demod* h, int stream) {
    if (!h || stream < 0 || stream >= h->S) { fail(FMGPU_ERR_ARG, "rds_device: bad argument"); return nullptr; }
    if (h->rds_host_state.empty()) { fail(FMGPU_ERR_STATE, "rds_device: call fmgpu_rds_device_fetch first"); return nullptr; }
    return &h->rds_host_state[stream];
}

int fmgpu_rds_device_counts(fmgpu_demod* h, int stream, unsigned long long* n_groups, unsigned long long* n_bytes, int* ring_caps) {
    const rds::State* st = rds_fetched(h, stream);
    if (!st) return FMGPU_ERR_STATE;
    if (n_groups) *n_groups = st->n_groups;
    if (n_bytes) *n_bytes = st->n_bytes;
    if (ring_caps) { ring_caps[0] = h->rds_gcap; ring_caps[1] = h->rds_bcap; }
    return FMGPU_OK;
}

int fmgpu_rds_device_get_groups(fmgpu_demod* h, int stream, unsigned long long first, fmgpu_rds_group* out, int max_groups) {
    const rds::State* st = rds_fetched(h, stream);
    if (!st || !out) return FMGPU_ERR_STATE;
    const unsigned long long total = st->n_groups, cap = (unsigned long long)h->rds_gcap;
    if (first > total || total - first > cap) return fail(FMGPU_ERR_STATE, "rds_device_get_groups: range no longer in the ring");
    int n = 0;
    for (unsigned long long g = first; g < total && n < max_groups; g++, n++)
        out[n] = h->rds_host_glog[(size_t)stream * cap + (size_t)(g % cap)];
    return n;
}

int fmgpu_rds_device_get_bytes(fmgpu_demod* h, int stream, unsigned long long first, uint8_t* out, int max_bytes) {
    const rds::State* st = rds_fetched(h, stream);
    if (!st || !out) return FMGPU_ERR_STATE;
    const unsigned long long total = st->n_bytes, cap = (unsigned long long)h->rds_bcap;
    if (first > total || total - first > cap) return fail(FMGPU_ERR_STATE, "rds_device_get_bytes: range no longer in the ring");
    int n = 0;
    for (unsigned long long b = first; b < total && n < max_bytes; b++, n++)
        out[n] = h->rds_host_blog[(size_t)stream * cap + (size_t)(b % cap)];
    return n;
}

int fmgpu_rds_device_get_db(fmgpu_demod* h, int stream, uint16_t* pi, char ps8[8], char rt64[64], uint8_t* pty) {
    const rds::State* st = rds_fetched(h, stream);
    if (!st) return FMGPU_ERR_STATE;
    if (pi) *pi = st->pi;
    if (pty) *pty = st->pty;
    if (ps8) std::memcpy(ps8, st->ps, 8);
    if (rt64) std::memcpy(rt64, st->rt, 64);
    return FMGPU_OK;
}

int fmgpu_rds_device_get_db_ext(fmgpu_demod* h, int stream, fmgpu_rds_db_ext* out) {
    const rds::State* st = rds_fetched(h, stream);
    if (!st) return FMGPU_ERR_STATE;
    if (!out) return fail(FMGPU_ERR_ARG, "rds_device_get_db_ext: null out");
    *out = st->ext;
    return FMGPU_OK;
}

// ---- stand-alone polyphase decimator (dsp/polyphase_filter.h:9-87) ----
struct fmgpu_polyphase {
    int M, K, NN, is_complex;
    int up = 0;                    // 1: PolyphaseUpsampler (M holds L; hist holds the last K inputs)
    std::vector<float> b;          // host taps (get_b)
    std::vector<float> hist;       // last NN inputs
    float* d_ext = nullptr; float* d_taps = nullptr; float* d_y = nullptr;
    size_t cap_ext = 0, cap_y = 0;
};

int fmgpu_polyphase_ds_create(int M, int K, int is_complex, fmgpu_polyphase** out) {
    if (!out || M < 1 || K < 1) return fail(FMGPU_ERR_ARG, "polyphase_ds_create: bad argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(FMGPU_ERR_CUDA, "polyphase_ds_create: no CUDA device (there is no CPU fallback)");
    auto* f = new fmgpu_polyphase();
    f->M = M; f->K = K; f->NN = M * K; f->is_complex = is_complex ? 1 : 0;
    f->b.assign(f->NN, 0.0f);
    f->hist.assign((size_t)f->NN * (is_complex ? 2 : 1), 0.0f);
    cudaError_t e = cudaMalloc((void**)&f->d_taps, f->NN * sizeof(float));
    if (e != cudaSuccess) { delete f; return fail(FMGPU_ERR_CUDA, cudaGetErrorString(e)); }
    *out = f;
    return FMGPU_OK;
}

void fmgpu_polyphase_destroy(fmgpu_polyphase* f) {
    if (!f) return;
    if (f->d_ext) cudaFree(f->d_ext);
    if (f->d_taps) cudaFree(f->d_taps);
    if (f->d_y) cudaFree(f->d_y);
    delete f;
}

float* fmgpu_polyphase_get_b(fmgpu_polyphase* f) { return f ? f->b.data() : nullptr; }

int fmgpu_polyphase_ds_process(fmgpu_polyphase* f, const float* x_host, float* y_host, int n_out) {
    if (!f || !x_host || !y_host || n_out < 0) return fail(FMGPU_ERR_ARG, "polyphase_ds_process: bad argument");
    if (n_out == 0) return FMGPU_OK;
    const int C = f->is_complex ? 2 : 1;
    const size_t n_in = (size_t)n_out * f->M;
    const size_t ext_floats = ((size_t)f->NN + n_in) * C, y_floats = (size_t)n_out * C;
    if (f->cap_ext < ext_floats) { if (f->d_ext) cudaFree(f->d_ext); CU(cudaMalloc((void**)&f->d_ext, ext_floats * 4)); f->cap_ext = ext_floats; }
    if (f->cap_y < y_floats) { if (f->d_y) cudaFree(f->d_y); CU(cudaMalloc((void**)&f->d_y, y_floats * 4)); f->cap_y = y_floats; }
    CU(cudaMemcpy(f->d_taps, f->b.data(), f->NN * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(f->d_ext, f->hist.data(), (size_t)f->NN * C * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(f->d_ext + (size_t)f->NN * C, x_host, n_in * C * 4, cudaMemcpyHostToDevice));
    CU(fm::launch_polyphase_ds(f->d_ext, f->d_taps, f->d_y, f->M, f->NN, n_out, f->is_complex, 0));
    CU(cudaMemcpy(y_host, f->d_y, y_floats * 4, cudaMemcpyDeviceToHost));
    // new history = last NN samples of (hist ++ x)
    CU(cudaMemcpy(f->hist.data(), f->d_ext + n_in * C, (size_t)f->NN * C * 4, cudaMemcpyDeviceToHost));
    return FMGPU_OK;
}

// ---- FIR_Filter / Hilbert_FIR_Filter / IIR_Filter / AGC_Filter stand-alone (dsp/fir_filter.h, hilbert_fir_filter.h, iir_filter.h, agc.h) ----
struct fmgpu_dsp_filter {
    int kind, K, is_complex;
    std::vector<float> b, a;       // host taps (get_b / get_a)
    std::vector<float> state;      // FIR kinds: the last K inputs; IIR: xn[K] ++ yn[K]
    float* d_x = nullptr; float* d_y = nullptr; float* d_tmp = nullptr; float* d_coef = nullptr; float* d_state = nullptr;
    size_t cap = 0;
};

int fmgpu_dsp_filter_create(int kind, int K, int is_complex, fmgpu_dsp_filter** out) {
    if (!out || K < 1 || kind < FMGPU_FILTER_FIR || kind > FMGPU_FILTER_IIR) return fail(FMGPU_ERR_ARG, "filter_create: bad argument");
    if (kind == FMGPU_FILTER_IIR && K < 2) return fail(FMGPU_ERR_ARG, "filter_create: IIR_Filter needs K >= 2 (the reference indexes yn[K-2])");
    if (kind == FMGPU_FILTER_HILBERT && is_complex) return fail(FMGPU_ERR_ARG, "filter_create: Hilbert_FIR_Filter takes a real input");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(FMGPU_ERR_CUDA, "filter_create: no CUDA device (there is no CPU fallback)");
    auto* f = new fmgpu_dsp_filter();
    f->kind = kind; f->K = K; f->is_complex = is_complex ? 1 : 0;
    const int C = f->is_complex ? 2 : 1;
    f->b.assign(K, 0.0f);
    if (kind == FMGPU_FILTER_IIR) f->a.assign(K, 0.0f);
    if (kind == FMGPU_FILTER_HILBERT) fmgpu_create_fir_hilbert(f->b.data(), K);          // the constructor designs the taps (:21-22)
    f->state.assign((size_t)(kind == FMGPU_FILTER_IIR ? 2 : 1) * K * C, 0.0f);
    cudaError_t e = cudaMalloc((void**)&f->d_coef, 2 * (size_t)K * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&f->d_state, f->state.size() * sizeof(float));
    if (e != cudaSuccess) { fmgpu_dsp_filter_destroy(f); return fail(FMGPU_ERR_CUDA, cudaGetErrorString(e)); }
    *out = f;
    return FMGPU_OK;
}

void fmgpu_dsp_filter_destroy(fmgpu_dsp_filter* f) {
    if (!f) return;
    for (float* p : { f->d_x, f->d_y, f->d_tmp, f->d_coef, f->d_state }) if (p) cudaFree(p);
    delete f;
}

float* fmgpu_dsp_filter_get_b(fmgpu_dsp_filter* f) { return f ? f->b.data() : nullptr; }
float* fmgpu_dsp_filter_get_a(fmgpu_dsp_filter* f) { return (f && !f->a.empty()) ? f->a.data() : nullptr; }
int fmgpu_dsp_filter_get_K(const fmgpu_dsp_filter* f) { return f ? f->K : 0; }

int fmgpu_dsp_filter_process(fmgpu_dsp_filter* f, const float* x_host, float* y_host, int n) {
    if (!f || !x_host || !y_host || n < 0) return fail(FMGPU_ERR_ARG, "filter_process: bad argument");
    if (n == 0) return FMGPU_OK;
    const int C = f->is_complex ? 2 : 1, K = f->K;
    const size_t need = ((size_t)K + n) * 2;                       // floats per buffer, enough for every kind
    if (f->cap < need) {
        for (float** p : { &f->d_x, &f->d_y, &f->d_tmp }) { if (*p) cudaFree(*p); *p = nullptr; CU(cudaMalloc((void**)p, need * sizeof(float))); }
        f->cap = need;
    }
    CU(cudaMemcpy(f->d_coef, f->b.data(), K * sizeof(float), cudaMemcpyHostToDevice));
    if (f->kind == FMGPU_FILTER_IIR) {
        CU(cudaMemcpy(f->d_coef + K, f->a.data(), K * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(f->d_state, f->state.data(), f->state.size() * sizeof(float), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(f->d_x, x_host, (size_t)n * C * sizeof(float), cudaMemcpyHostToDevice));
        CU(fm::launch_iir_seq(f->d_x, f->d_y, n, K, f->d_coef, f->d_coef + K, f->d_state, f->is_complex, 0));
        CU(cudaMemcpy(y_host, f->d_y, (size_t)n * C * sizeof(float), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(f->state.data(), f->d_state, f->state.size() * sizeof(float), cudaMemcpyDeviceToHost));
        return FMGPU_OK;
    }
    // FIR kinds: ext = (last K inputs) ++ x, y[i] = sum_k b[k] ext[i + 1 + k]  (the polyphase kernel with M = 1, NN = K)
    CU(cudaMemcpy(f->d_x, f->state.data(), (size_t)K * C * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(f->d_x + (size_t)K * C, x_host, (size_t)n * C * sizeof(float), cudaMemcpyHostToDevice));
    if (f->kind == FMGPU_FILTER_FIR) {
        CU(fm::launch_polyphase_ds(f->d_x, f->d_coef, f->d_y, 1, K, n, f->is_complex, 0));
        CU(cudaMemcpy(y_host, f->d_y, (size_t)n * C * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        CU(fm::launch_polyphase_ds(f->d_x, f->d_coef, f->d_tmp, 1, K, n, 0, 0));
        CU(fm::launch_hilbert_pack(f->d_x, f->d_tmp, (float2*)f->d_y, n, (K - 1) / 2, 0));
        CU(cudaMemcpy(y_host, f->d_y, (size_t)n * 2 * sizeof(float), cudaMemcpyDeviceToHost));
    }
    CU(cudaMemcpy(f->state.data(), f->d_x + (size_t)n * C, (size_t)K * C * sizeof(float), cudaMemcpyDeviceToHost));
    return FMGPU_OK;
}

void fmgpu_agc_init(fmgpu_agc* g) { if (g) { g->target_power = 1.0f; g->current_gain = 0.1f; g->beta = 0.2f; } }

int fmgpu_agc_process(fmgpu_agc* g, const float* x_host, float* y_host, int n) {
    if (!g || !x_host || !y_host || n <= 0) return fail(FMGPU_ERR_ARG, "agc_process: bad argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(FMGPU_ERR_CUDA, "agc_process: no CUDA device (there is no CPU fallback)");
    float2* d_x = nullptr; float* d_sum = nullptr;
    CU(cudaMalloc((void**)&d_x, (size_t)n * sizeof(float2)));
    cudaError_t e = cudaMalloc((void**)&d_sum, sizeof(float));
    if (e != cudaSuccess) { cudaFree(d_x); return fail(FMGPU_ERR_CUDA, cudaGetErrorString(e)); }
    auto done = [&](cudaError_t err) { cudaFree(d_x); cudaFree(d_sum); return err == cudaSuccess ? FMGPU_OK : fail(FMGPU_ERR_CUDA, cudaGetErrorString(err)); };
    if ((e = cudaMemcpy(d_x, x_host, (size_t)n * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess) return done(e);
    if ((e = fm::launch_agc_power_seq(d_x, n, d_sum, 0)) != cudaSuccess) return done(e);
    float sum = 0.0f;
    if ((e = cudaMemcpy(&sum, d_sum, sizeof(float), cudaMemcpyDeviceToHost)) != cudaSuccess) return done(e);
    const float avg_power = sum / (float)n;                                              // agc.h:28
    const float target_gain = std::sqrt(g->target_power / avg_power);                   // :15 (unguarded: 0 power -> inf, as the reference)
    g->current_gain = g->current_gain + g->beta * (target_gain - g->current_gain);      // :16
    if ((e = fm::launch_agc_scale(d_x, d_x, n, g->current_gain, 0)) != cudaSuccess) return done(e);
    e = cudaMemcpy(y_host, d_x, (size_t)n * sizeof(float2), cudaMemcpyDeviceToHost);
    return done(e);
}

// ---- PolyphaseUpsampler<T> (dsp/polyphase_filter.h:90-185) ----
int fmgpu_polyphase_us_create(const float* b, int L, int K, int is_complex, fmgpu_polyphase** out) {
    if (!out || !b || L < 1 || K < 1) return fail(FMGPU_ERR_ARG, "polyphase_us_create: bad argument");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return fail(FMGPU_ERR_CUDA, "polyphase_us_create: no CUDA device (there is no CPU fallback)");
    auto* f = new fmgpu_polyphase();
    f->M = L; f->K = K; f->NN = L * K; f->is_complex = is_complex ? 1 : 0; f->up = 1;
    f->b.assign(f->NN, 0.0f);
    // the constructor's repack (:106-116): phase-contiguous taps, gain L
    for (int phase = 0; phase < L; phase++) {
        const int phase_c = (L - 1) - phase;
        for (int i = 0; i < K; i++) f->b[(size_t)phase_c * K + i] = b[(f->NN - 1) - (phase + i * L)] * (float)L;
    }
    f->hist.assign((size_t)K * (is_complex ? 2 : 1), 0.0f);
    cudaError_t e = cudaMalloc((void**)&f->d_taps, f->NN * sizeof(float));
    if (e != cudaSuccess) { delete f; return fail(FMGPU_ERR_CUDA, cudaGetErrorString(e)); }
    *out = f;
    return FMGPU_OK;
}

int fmgpu_polyphase_us_process(fmgpu_polyphase* f, const float* x_host, float* y_host, int n_in) {
    if (!f || !f->up || !x_host || !y_host || n_in < 0) return fail(FMGPU_ERR_ARG, "polyphase_us_process: bad argument");
    if (n_in == 0) return FMGPU_OK;
    const int C = f->is_complex ? 2 : 1, L = f->M, K = f->K;
    const size_t ext_floats = ((size_t)K + n_in) * C, y_floats = (size_t)n_in * L * C;
    if (f->cap_ext < ext_floats) { if (f->d_ext) cudaFree(f->d_ext); CU(cudaMalloc((void**)&f->d_ext, ext_floats * 4)); f->cap_ext = ext_floats; }
    if (f->cap_y < y_floats) { if (f->d_y) cudaFree(f->d_y); CU(cudaMalloc((void**)&f->d_y, y_floats * 4)); f->cap_y = y_floats; }
    CU(cudaMemcpy(f->d_taps, f->b.data(), f->NN * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(f->d_ext, f->hist.data(), (size_t)K * C * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(f->d_ext + (size_t)K * C, x_host, (size_t)n_in * C * 4, cudaMemcpyHostToDevice));
    CU(fm::launch_polyphase_us(f->d_ext, f->d_taps, f->d_y, L, K, n_in, f->is_complex, 0));
    CU(cudaMemcpy(y_host, f->d_y, y_floats * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(f->hist.data(), f->d_ext + (size_t)n_in * C, (size_t)K * C * 4, cudaMemcpyDeviceToHost));   // last K of (hist ++ x)
    return FMGPU_OK;
}

// ---- Resample() and the scraper's int16 conversion, stand-alone (host buffers) ----
int fmgpu_resample_linear(const float* frames_in_host, int n_in, float* frames_out_host, int n_out) {
    if (!frames_in_host || !frames_out_host || n_in < 1 || n_out < 0) return fail(FMGPU_ERR_ARG, "resample_linear: bad argument");
    if (n_out == 0) return FMGPU_OK;
    std::vector<fm::K7Entry> tab((size_t)n_out);
    fm::k7_build_table(n_in, n_out, tab.data());
    if (tab.back().j0 >= n_in) return fail(FMGPU_ERR_ARG, "resample_linear: read position leaves the input (the reference would read out of bounds)");
    float2* d_in = nullptr; float2* d_out = nullptr; fm::K7Entry* d_tab = nullptr;
    auto done = [&](int rc) { if (d_in) cudaFree(d_in); if (d_out) cudaFree(d_out); if (d_tab) cudaFree(d_tab); return rc; };
#define CUF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(FMGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); } while (0)
    CUF(cudaMalloc((void**)&d_in, (size_t)n_in * 8)); CUF(cudaMalloc((void**)&d_out, (size_t)n_out * 8));
    CUF(cudaMalloc((void**)&d_tab, (size_t)n_out * sizeof(fm::K7Entry)));
    CUF(cudaMemcpy(d_in, frames_in_host, (size_t)n_in * 8, cudaMemcpyHostToDevice));
    CUF(cudaMemcpy(d_tab, tab.data(), (size_t)n_out * sizeof(fm::K7Entry), cudaMemcpyHostToDevice));
    CUF(fm::launch_k7(d_in, d_tab, d_out, nullptr, n_in, n_out, 1, 0));
    CUF(cudaMemcpy(frames_out_host, d_out, (size_t)n_out * 8, cudaMemcpyDeviceToHost));
    return done(FMGPU_OK);
}

int fmgpu_frames_to_s16(const float* frames_host, size_t n_frames, int16_t* out_host) {
    if (!frames_host || !out_host) return fail(FMGPU_ERR_ARG, "frames_to_s16: bad argument");
    if (n_frames == 0) return FMGPU_OK;
    float2* d_in = nullptr; short2* d_out = nullptr;
    auto done = [&](int rc) { if (d_in) cudaFree(d_in); if (d_out) cudaFree(d_out); return rc; };
    CUF(cudaMalloc((void**)&d_in, n_frames * 8)); CUF(cudaMalloc((void**)&d_out, n_frames * 4));
    CUF(cudaMemcpy(d_in, frames_host, n_frames * 8, cudaMemcpyHostToDevice));
    CUF(fm::launch_frames_to_s16(d_in, d_out, n_frames, 0));
    CUF(cudaMemcpy(out_host, d_out, n_frames * 4, cudaMemcpyDeviceToHost));
    return done(FMGPU_OK);
#undef CUF
}

// ---- display spectra: CalculateFFT + InplaceFFTShift (k_fft.cu) ----
static int fft_run(const float* d_in, int in_is_real, int n, int fftshift, float* y_host) {
    float2* w0 = nullptr; float2* w1 = nullptr; float2* res = nullptr;
    auto done = [&](int rc) { if (w0) cudaFree(w0); if (w1) cudaFree(w1); return rc; };
#define CUF(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return done(fail(FMGPU_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); } while (0)
    CUF(cudaMalloc((void**)&w0, (size_t)n * 8)); CUF(cudaMalloc((void**)&w1, (size_t)n * 8));
    CUF(fm::launch_fft(d_in, in_is_real, w0, w1, n, fftshift, &res, 0));
    CUF(cudaMemcpy(y_host, res, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return done(FMGPU_OK);
#undef CUF
}

int fmgpu_calculate_fft(const float* x_host, float* y_host, int n, int fftshift) {
    if (!x_host || !y_host) return fail(FMGPU_ERR_ARG, "calculate_fft: null argument");
    if (n < 2 || (n & (n - 1)) != 0) return fail(FMGPU_ERR_ARG, "calculate_fft: n must be a power of two >= 2 (every block size of the chain is)");
    float* d_in = nullptr;
    CU(cudaMalloc((void**)&d_in, (size_t)n * 8));
    cudaError_t e = cudaMemcpy(d_in, x_host, (size_t)n * 8, cudaMemcpyHostToDevice);
    const int rc = e == cudaSuccess ? fft_run(d_in, 0, n, fftshift, y_host) : fail(FMGPU_ERR_CUDA, cudaGetErrorString(e));
    cudaFree(d_in);
    return rc;
}

int fmgpu_get_fft(fmgpu_demod* h, int stream, fmgpu_buffer buf, int fftshift, float* y_host, size_t* n_out) {
    if (!h || !y_host) return fail(FMGPU_ERR_ARG, "get_fft: null argument");
    if (stream < 0 || stream >= h->S) return fail(FMGPU_ERR_ARG, "get_fft: bad stream");
    if (h->last_fetched_slot < 0) return fail(FMGPU_ERR_STATE, "get_fft: nothing processed yet");
    BufInfo bi{}; const void* dev = nullptr;
    if (!buf_info(h, buf, &bi, &dev, h->last_fetched_slot))
        return fail(FMGPU_ERR_STATE, "get_fft: buffer needs keep_intermediates = 1");
    if (buf == FMGPU_BUF_PLL_DT || buf == FMGPU_BUF_RDS_PRED_SYM || buf == FMGPU_BUF_RDS_RAW_SYM || (bi.elem != 8 && bi.elem != 4)
        || buf == FMGPU_BUF_AUDIO_OUT || buf == FMGPU_BUF_AUDIO_PCM_F32 || buf == FMGPU_BUF_AUDIO_PCM_S16 || buf == FMGPU_BUF_RDS_SYM_COUNT)
        return fail(FMGPU_ERR_ARG, "get_fft: not a signal buffer (complex, or real f32, of fixed length)");
    CU(cudaSetDevice(h->device));
    if (sync_all(h) != FMGPU_OK) return FMGPU_ERR_CUDA;
    const int n = bi.per_stream;
    if (n_out) *n_out = (size_t)n;
    const float* d_in = (const float*)((const uint8_t*)dev + (size_t)stream * bi.elem * (size_t)n);
    return fft_run(d_in, bi.elem == 4, n, fftshift, y_host);
}

} // extern "C"

