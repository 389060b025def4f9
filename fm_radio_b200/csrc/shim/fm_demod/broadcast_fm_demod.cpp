// Implementation of the Broadcast_FM_Demod / BPSK_Synchroniser shims over the C-ABI.
// See broadcast_fm_demod.h in this directory; reference behaviour cited per method.
#include "broadcast_fm_demod.h"
#include "bpsk_synchroniser.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "fmgpu.h"

static void die_if(int rc, const char* what) {
    // The reference has no error channel (no exceptions, no return codes); failing to reach the GPU
    // is not something it can express, and silently producing nothing would look like a CPU fallback.
    if (rc != FMGPU_OK && rc != FMGPU_ERR_SIZE) {
        fprintf(stderr, "[fmgpu shim] %s failed (%d): %s\n", what, rc, fmgpu_last_error());
        abort();
    }
}

Broadcast_FM_Demod::Broadcast_FM_Demod(const int _block_size)
: block_size(_block_size)
{
    fmgpu_config cfg{};
    cfg.block_size = block_size;
    cfg.n_streams = 1;
    cfg.device = -1;
    // FMGPU_LEAN=1 drops the GUI-only buffers (what fm_demod_benchmark needs); default keeps them all
    const char* lean = getenv("FMGPU_LEAN");
    cfg.keep_intermediates = (lean && lean[0] == '1') ? 0 : 1;
    cfg.pipeline_depth = 1;
    die_if(fmgpu_create(&cfg, &handle), "fmgpu_create");
    bpsk_sync = std::make_unique<BPSK_Synchroniser>(*this, block_size / 64);
    audio_out_buf.resize((size_t)block_size / 32);
    rds_pred_sym_buf.resize((size_t)block_size / 64);
    const size_t fft_sizes[8] = { (size_t)block_size, (size_t)block_size / 4, (size_t)block_size / 8, (size_t)block_size / 8,
                                  (size_t)block_size / 8, (size_t)block_size / 32, (size_t)block_size / 32, (size_t)block_size / 64 };
    for (int i = 0; i < 8; i++) fft_mag_bufs[i].assign(fft_sizes[i], 0.0f);
    // broadcast_fm_demod.cpp:189,248,260: the constructor marks the three editable cutoffs dirty
    controls.filt_deemphasis_cutoff.SetValue(params.Tus_min_deemphasis);
    controls.filt_audio_lpr_cutoff.SetValue(params.F_audio_lpr);
    controls.filt_audio_lmr_cutoff.SetValue(params.F_audio_lmr_bandwidth);
}

Broadcast_FM_Demod::~Broadcast_FM_Demod() { fmgpu_destroy(handle); }

// UpdateFilters + the plain control fields (broadcast_fm_demod.cpp:330-389, 555-560), latched at entry
void Broadcast_FM_Demod::LatchControls() {
    if ((int)controls.audio_out != sent_audio_out) {
        sent_audio_out = (int)controls.audio_out;
        die_if(fmgpu_set_control(handle, FMGPU_CTL_AUDIO_OUT, (double)sent_audio_out), "set audio_out");
    }
    if (controls.audio_stereo_mix_factor != sent_mix) {
        sent_mix = controls.audio_stereo_mix_factor;
        die_if(fmgpu_set_control(handle, FMGPU_CTL_AUDIO_STEREO_MIX_FACTOR, (double)sent_mix), "set mix");
    }
    if ((int)controls.is_use_deemphasis_filter != sent_deemph) {
        sent_deemph = (int)controls.is_use_deemphasis_filter;
        die_if(fmgpu_set_control(handle, FMGPU_CTL_USE_DEEMPHASIS, (double)sent_deemph), "set deemphasis");
    }
    if (controls.filt_deemphasis_cutoff.IsDirty()) {
        controls.filt_deemphasis_cutoff.ClearDirty();
        die_if(fmgpu_set_control(handle, FMGPU_CTL_DEEMPHASIS_TUS, (double)controls.filt_deemphasis_cutoff.GetValue()), "set deemphasis cutoff");
    }
    if (controls.filt_audio_lpr_cutoff.IsDirty()) {
        controls.filt_audio_lpr_cutoff.ClearDirty();
        die_if(fmgpu_set_control(handle, FMGPU_CTL_AUDIO_LPR_CUTOFF_HZ, (double)controls.filt_audio_lpr_cutoff.GetValue()), "set lpr cutoff");
    }
    if (controls.filt_audio_lmr_cutoff.IsDirty()) {
        controls.filt_audio_lmr_cutoff.ClearDirty();
        die_if(fmgpu_set_control(handle, FMGPU_CTL_AUDIO_LMR_CUTOFF_HZ, (double)controls.filt_audio_lmr_cutoff.GetValue()), "set lmr cutoff");
    }
}

// broadcast_fm_demod.cpp:325-327: emit audio and RDS symbols to the observers, synchronously
void Broadcast_FM_Demod::AfterProcess() {
    generation++;
    const void* p = nullptr; size_t n = 0;
    die_if(fmgpu_get_buffer(handle, 0, FMGPU_BUF_AUDIO_OUT, &p, &n), "get audio");
    std::memcpy(audio_out_buf.data(), p, n * sizeof(Frame<float>));
    die_if(fmgpu_get_buffer(handle, 0, FMGPU_BUF_RDS_PRED_SYM, &p, &n), "get symbols");
    rds_total_symbols = (int)n;
    std::memcpy(rds_pred_sym_buf.data(), p, n * sizeof(float));
    obs_on_audio_block.Notify(audio_out_buf, GetAudioSampleRate());
    obs_on_rds_symbols.Notify(tcb::span<const float>(rds_pred_sym_buf).first((size_t)rds_total_symbols));
}

// UpdateFFTCalc (broadcast_fm_demod.cpp:27-40) for the spectra whose source buffer exists on the device:
// CalculateFFT + InplaceFFTShift there, Calculate_FFT_Mag::Process (the reference's class) here.
void Broadcast_FM_Demod::UpdateSpectra(const float* baseband_cf32) {
    fft_tmp.resize((size_t)block_size);
    float* tmp = reinterpret_cast<float*>(fft_tmp.data());
    if (baseband_cf32 && calc_fft_mag[0].IsAwaitingUpdate()) {                       // :413
        die_if(fmgpu_calculate_fft(baseband_cf32, tmp, block_size, 1), "fmgpu_calculate_fft");
        calc_fft_mag[0].Process(tcb::span<const std::complex<float>>(fft_tmp.data(), (size_t)block_size), fft_mag_bufs[0]);
    }
    const struct { int idx; fmgpu_buffer buf; } src[7] = { { 1, FMGPU_BUF_FM_IN }, { 2, FMGPU_BUF_FM_OUT_IQ },     // :414, :415
                                                           { 3, FMGPU_BUF_PILOT }, { 4, FMGPU_BUF_PLL },           // :459, :460
                                                           { 5, FMGPU_BUF_AUDIO_LPR_IQ }, { 6, FMGPU_BUF_AUDIO_LMR_IQ },   // :481, :523
                                                           { 7, FMGPU_BUF_RDS } };                                 // :535
    for (const auto& e : src) {
        if (!calc_fft_mag[e.idx].IsAwaitingUpdate()) continue;
        size_t n = 0;
        if (fmgpu_get_fft(handle, 0, e.buf, 1, tmp, &n) != FMGPU_OK) continue;      // FMGPU_LEAN: the GUI buffers do not exist
        calc_fft_mag[e.idx].Process(tcb::span<const std::complex<float>>(fft_tmp.data(), n), fft_mag_bufs[e.idx]);
    }
}

void Broadcast_FM_Demod::Process(tcb::span<const std::complex<float>> x) {
    if (x.size() != (size_t)block_size) return;                  // broadcast_fm_demod.cpp:311-313
    LatchControls();
    die_if(fmgpu_process_cf32(handle, reinterpret_cast<const float*>(x.data()), x.size()), "fmgpu_process_cf32");
    UpdateSpectra(reinterpret_cast<const float*>(x.data()));
    AfterProcess();
}

void Broadcast_FM_Demod::ProcessU8(tcb::span<const std::complex<uint8_t>> x) {
    if (x.size() != (size_t)block_size) return;
    LatchControls();
    die_if(fmgpu_process_u8(handle, reinterpret_cast<const uint8_t*>(x.data()), x.size()), "fmgpu_process_u8");
    if (calc_fft_mag[0].IsAwaitingUpdate()) {                    // App::Run's unpack (app.cpp:56-65), only for the display
        std::vector<float> f(2 * (size_t)block_size);
        const uint8_t* b = reinterpret_cast<const uint8_t*>(x.data());
        for (size_t i = 0; i < f.size(); i++) f[i] = (float)b[i] - 127.0f;
        UpdateSpectra(f.data());
    } else {
        UpdateSpectra(nullptr);
    }
    AfterProcess();
}

template <typename T>
tcb::span<T> Broadcast_FM_Demod::Fetch(int buf) {
    Mirror& m = mirrors[buf];
    if (m.fetched_at != generation) {
        const void* p = nullptr; size_t n = 0;
        if (generation == 0) { m.bytes.assign(sizeof(T) * (size_t)block_size, 0); m.n = 0; }
        else {
            die_if(fmgpu_get_buffer(handle, 0, (fmgpu_buffer)buf, &p, &n), "fmgpu_get_buffer");
            m.bytes.resize(n * sizeof(T) + sizeof(T));
            std::memcpy(m.bytes.data(), p, n * sizeof(T));
            m.n = n;
        }
        m.fetched_at = generation;
    }
    return tcb::span<T>(reinterpret_cast<T*>(m.bytes.data()), m.n);
}

tcb::span<std::complex<float>> Broadcast_FM_Demod::GetFMOutIQ() { return Fetch<std::complex<float>>(FMGPU_BUF_FM_OUT_IQ); }
tcb::span<std::complex<float>> Broadcast_FM_Demod::GetPilotOutput() { return Fetch<std::complex<float>>(FMGPU_BUF_PILOT); }
tcb::span<std::complex<float>> Broadcast_FM_Demod::GetPLLOutput() { return Fetch<std::complex<float>>(FMGPU_BUF_PLL); }
tcb::span<float> Broadcast_FM_Demod::Get_PLL_Raw_Phase_Error_Output() { return Fetch<float>(FMGPU_BUF_PLL_RAW_PHASE_ERROR); }
tcb::span<float> Broadcast_FM_Demod::Get_PLL_LPF_Phase_Error_Output() { return Fetch<float>(FMGPU_BUF_PLL_LPF_PHASE_ERROR); }
tcb::span<float> Broadcast_FM_Demod::GetLPRAudioOutput() { return Fetch<float>(FMGPU_BUF_AUDIO_LPR); }
tcb::span<float> Broadcast_FM_Demod::GetLMRAudioOutput() { return Fetch<float>(FMGPU_BUF_AUDIO_LMR); }
tcb::span<std::complex<float>> Broadcast_FM_Demod::GetRDSOutput() { return Fetch<std::complex<float>>(FMGPU_BUF_RDS); }
tcb::span<std::complex<float>> Broadcast_FM_Demod::GetRDSRawSymbols() { return Fetch<std::complex<float>>(FMGPU_BUF_RDS_RAW_SYM); }

float Broadcast_FM_Demod::GetAudioLMRPhaseError() {
    float v = 0.0f;
    die_if(fmgpu_get_scalar(handle, 0, FMGPU_SCALAR_AUDIO_LMR_PHASE_ERROR, &v), "fmgpu_get_scalar");
    return v;
}

// ---- BPSK_Synchroniser getters (bpsk_synchroniser.h:76-85) ----
template <typename T>
tcb::span<const T> BPSK_Synchroniser::Fetch(int slot, int buf) const {
    Mirror& m = mirrors[slot];
    const unsigned long long gen = owner.GetGeneration();
    if (m.fetched_at != gen) {
        m.bytes.assign(sizeof(T) * (size_t)block_size, 0);
        if (gen != 0) {
            const void* p = nullptr; size_t n = 0;
            die_if(fmgpu_get_buffer(owner.GetHandle(), 0, (fmgpu_buffer)buf, &p, &n), "fmgpu_get_buffer");
            std::memcpy(m.bytes.data(), p, n * sizeof(T));
        }
        m.fetched_at = gen;
    }
    return tcb::span<const T>(reinterpret_cast<const T*>(m.bytes.data()), (size_t)block_size);
}

tcb::span<const std::complex<float>> BPSK_Synchroniser::GetPLLSymbols() const { return Fetch<std::complex<float>>(0, FMGPU_BUF_BPSK_PLL_SYM); }
tcb::span<const bool> BPSK_Synchroniser::GetZeroCrossings() const { return Fetch<bool>(1, FMGPU_BUF_BPSK_ZCD); }
tcb::span<const bool> BPSK_Synchroniser::GetIntDumpTriggers() const { return Fetch<bool>(2, FMGPU_BUF_BPSK_INT_DUMP_TRIGGER); }
tcb::span<const float> BPSK_Synchroniser::GetTEDRawPhaseError() const { return Fetch<float>(3, FMGPU_BUF_BPSK_TED_RAW_PHASE_ERROR); }
tcb::span<const float> BPSK_Synchroniser::GetTEDPIPhaseError() const { return Fetch<float>(4, FMGPU_BUF_BPSK_TED_PI_PHASE_ERROR); }
tcb::span<const float> BPSK_Synchroniser::GetPLLRawPhaseError() const { return Fetch<float>(5, FMGPU_BUF_BPSK_PLL_RAW_PHASE_ERROR); }
tcb::span<const float> BPSK_Synchroniser::GetPLLPIPhaseError() const { return Fetch<float>(6, FMGPU_BUF_BPSK_PLL_PI_PHASE_ERROR); }
tcb::span<const std::complex<float>> BPSK_Synchroniser::GetIntDumpFilter() const { return Fetch<std::complex<float>>(7, FMGPU_BUF_BPSK_INT_DUMP_FILTER); }
