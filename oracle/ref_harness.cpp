// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from the product path.
//
// C-ABI harness around the UNMODIFIED reference classes, compiled in place from
// /root/reference/src by oracle/Makefile into oracle/_ref/libfmref.so.  It re-creates the
// wiring of the reference's App (src/app.cpp:9-35, 56-65): u8 IQ -> (float)u8 - 127.0f ->
// Broadcast_FM_Demod::Process -> DifferentialManchesterDecoder -> RDS_Decoding_Chain, and
// additionally records every intermediate buffer and every RDS group so that tests can pin
// oracle/fm_oracle.c (the CPU restatement) and the CUDA path against the real thing.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load the resulting library.
//
// `#define private public` is used so the harness can read buffers that the reference keeps
// private without getters (fm_in_buf, fm_demod_buf, fm_out_buf, pll_dt_buf).  Access specifiers
// do not change the Itanium-ABI layout, so the reference translation units stay unmodified.
#include <complex>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define private public
#include "fm_demod/broadcast_fm_demod.h"
#include "fm_demod/bpsk_synchroniser.h"
#undef private
#include "dsp/filter_designer.h"
#include "dsp/fir_filter.h"
#include "dsp/hilbert_fir_filter.h"
#include "dsp/iir_filter.h"
#include "dsp/agc.h"
#include "audio/frame.h"
#include "dsp/calculate_fft_mag.h"
#include "dsp/fftshift.h"
#include <limits>
void Resample(tcb::span<const Frame<float>> buf_in, tcb::span<Frame<float>> buf_out);   // audio/resampled_pcm_player.cpp:3
#include "rds_decoder/differential_manchester_decoder.h"
#include "rds_decoder/rds_decoding_chain.h"

namespace {

struct RefGroup {
    uint16_t data[4];
    uint8_t valid[4];
    uint8_t type[4];
};

struct RefRdsChain {
    std::vector<uint8_t> bytes_buf;
    std::unique_ptr<DifferentialManchesterDecoder> manchester;
    std::unique_ptr<RDS_Decoding_Chain> chain;
    std::vector<RefGroup> groups;
    std::vector<uint8_t> rds_bytes;
    RefRdsChain() : bytes_buf(16) {
        manchester = std::make_unique<DifferentialManchesterDecoder>(tcb::span<uint8_t>(bytes_buf));
        chain = std::make_unique<RDS_Decoding_Chain>();
        chain->group_sync.OnGroup().Attach([this](rds_group_t g) {
            RefGroup r;
            for (int i = 0; i < 4; i++) {
                r.data[i] = g[i].data;
                r.valid[i] = g[i].is_valid ? 1 : 0;
                r.type[i] = (uint8_t)g[i].block_type;
            }
            groups.push_back(r);
        });
        manchester->OnBytes().Attach([this](tcb::span<const uint8_t> x) {
            rds_bytes.insert(rds_bytes.end(), x.begin(), x.end());
            chain->Process(x);
        });
    }
};

struct RefHandle {
    int block_size;
    std::vector<std::complex<float>> iq_f32;
    std::unique_ptr<Broadcast_FM_Demod> demod;
    RefRdsChain rds;
    std::vector<float> symbols_last;   // symbols of the most recent block
    std::vector<float> bool_scratch;
    explicit RefHandle(int B) : block_size(B), iq_f32(B) {
        demod = std::make_unique<Broadcast_FM_Demod>(B);
        demod->OnRDSOut().Attach([this](tcb::span<const float> x) {
            symbols_last.assign(x.begin(), x.end());
            rds.manchester->Process(x);
        });
    }
};

template <typename T>
int give(tcb::span<T> s, const void** p, size_t* n) {
    *p = (const void*)s.data();
    *n = s.size();
    return 0;
}

} // namespace

extern "C" {

void* fmref_create(int block_size) { return new RefHandle(block_size); }
void fmref_destroy(void* h) { delete (RefHandle*)h; }

// src/app.cpp:56-65 (App::Run): (float)u8 - 127.0f per component, then Process.
int fmref_process_u8(void* hv, const uint8_t* iq) {
    auto* h = (RefHandle*)hv;
    for (int i = 0; i < h->block_size; i++) {
        h->iq_f32[i] = { (float)iq[2*i+0] - 127.0f, (float)iq[2*i+1] - 127.0f };
    }
    h->symbols_last.clear();
    h->demod->Process(h->iq_f32);
    return 0;
}

int fmref_process_cf32(void* hv, const float* iq) {
    auto* h = (RefHandle*)hv;
    h->symbols_last.clear();
    h->demod->Process(tcb::span<const std::complex<float>>((const std::complex<float>*)iq, (size_t)h->block_size));
    return 0;
}

// controls: broadcast_fm_demod.h:64-89
int fmref_set_control(void* hv, const char* name, double value) {
    auto* h = (RefHandle*)hv;
    auto& c = h->demod->GetControls();
    const std::string s(name);
    if (s == "audio_out") c.audio_out = (Broadcast_FM_Demod_Controls::AudioOut)(int)value;
    else if (s == "audio_stereo_mix_factor") c.audio_stereo_mix_factor = (float)value;
    else if (s == "is_use_deemphasis_filter") c.is_use_deemphasis_filter = value != 0.0;
    else if (s == "filt_deemphasis_cutoff") c.filt_deemphasis_cutoff.SetValue((int)value);
    else if (s == "filt_audio_lpr_cutoff") c.filt_audio_lpr_cutoff.SetValue((int)value);
    else if (s == "filt_audio_lmr_cutoff") c.filt_audio_lmr_cutoff.SetValue((int)value);
    else return -1;
    return 0;
}

// Named views of the demodulator's buffers after the most recent Process().
// Returns 0 and (*ptr,*n_elems) in units of the buffer's element type; -1 for unknown names.
int fmref_get(void* hv, const char* name, const void** p, size_t* n) {
    auto* h = (RefHandle*)hv;
    auto& d = *h->demod;
    auto& b = d.GetBPSKSync();
    const std::string s(name);
    if (s == "fm_in") return give(d.fm_in_buf, p, n);                 // cf32
    if (s == "fm_demod") return give(d.fm_demod_buf, p, n);           // f32
    if (s == "fm_out") return give(d.fm_out_buf, p, n);               // f32
    if (s == "fm_out_iq") return give(d.fm_out_iq_buf, p, n);         // cf32
    if (s == "pilot") return give(d.pilot_buf, p, n);                 // cf32
    if (s == "pll_dt") return give(d.pll_dt_buf, p, n);               // f32
    if (s == "pll") return give(d.pll_buf, p, n);                     // cf32
    if (s == "pll_raw_phase_error") return give(d.pll_raw_phase_error, p, n);
    if (s == "pll_lpf_phase_error") return give(d.pll_lpf_phase_error, p, n);
    if (s == "audio_lpr") return give(d.audio_lpr_buf, p, n);
    if (s == "audio_lmr") return give(d.audio_lmr_buf, p, n);
    if (s == "rds") return give(d.rds_buf, p, n);                     // cf32 (after AGC)
    if (s == "rds_raw_sym") return give(d.GetRDSRawSymbols(), p, n);  // cf32[count]
    if (s == "rds_pred_sym") return give(d.GetRDSPredSymbols(), p, n);// f32[count]
    if (s == "audio_out") return give(d.audio_out_buf, p, n);         // Frame<float> = 2 x f32
    if (s == "bpsk_pll_sym") return give(b.GetPLLSymbols(), p, n);
    if (s == "bpsk_ted_raw_phase_error") return give(b.GetTEDRawPhaseError(), p, n);
    if (s == "bpsk_ted_pi_phase_error") return give(b.GetTEDPIPhaseError(), p, n);
    if (s == "bpsk_pll_raw_phase_error") return give(b.GetPLLRawPhaseError(), p, n);
    if (s == "bpsk_pll_pi_phase_error") return give(b.GetPLLPIPhaseError(), p, n);
    if (s == "bpsk_int_dump_filter") return give(b.GetIntDumpFilter(), p, n);
    if (s == "bpsk_zcd") return give(b.GetZeroCrossings(), p, n);        // bool
    if (s == "bpsk_int_dump_trigger") return give(b.GetIntDumpTriggers(), p, n); // bool
    return -1;
}

float fmref_get_scalar(void* hv, const char* name) {
    auto* h = (RefHandle*)hv;
    auto& d = *h->demod;
    const std::string s(name);
    if (s == "audio_lmr_phase_error") return d.GetAudioLMRPhaseError();
    if (s == "agc_pilot_gain") return d.agc_pilot.current_gain;
    if (s == "agc_rds_gain") return d.agc_rds.current_gain;
    if (s == "rds_total_symbols") return (float)d.rds_total_symbols;
    return 0.0f/0.0f;
}

// Filter taps as designed by the reference at construction / UpdateFilters
// (broadcast_fm_demod.cpp:133-274, 330-389).  Returns number of taps written into b (and a).
int fmref_get_taps(void* hv, const char* name, float* b, float* a) {
    auto* h = (RefHandle*)hv;
    auto& d = *h->demod;
    const std::string s(name);
    auto copy = [](float* dst, const float* src, int n) { if (dst) std::memcpy(dst, src, sizeof(float)*n); return n; };
    if (s == "fm_in") return copy(b, d.filt_poly_ds_lpf_fm_in->get_b(), d.filt_poly_ds_lpf_fm_in->get_K());
    if (s == "fm_out") return copy(b, d.filt_poly_ds_lpf_fm_out->get_b(), d.filt_poly_ds_lpf_fm_out->get_K());
    if (s == "hilbert") return copy(b, d.filt_hilbert_transform->get_b(), d.filt_hilbert_transform->get_K());
    if (s == "audio_lpr") return copy(b, d.filt_poly_ds_lpf_audio_lpr->get_b(), d.filt_poly_ds_lpf_audio_lpr->get_K());
    if (s == "audio_lmr") return copy(b, d.filt_poly_ds_lpf_audio_lmr->get_b(), d.filt_poly_ds_lpf_audio_lmr->get_K());
    if (s == "rds") return copy(b, d.filt_poly_ds_lpf_rds->get_b(), d.filt_poly_ds_lpf_rds->get_K());
    if (s == "deemphasis") { copy(a, d.filt_iir_lpf_fm_deemphasis->get_a(), 2); return copy(b, d.filt_iir_lpf_fm_deemphasis->get_b(), 2); }
    if (s == "peak_pilot") { copy(a, d.filt_iir_peak_pilot->get_a(), 3); return copy(b, d.filt_iir_peak_pilot->get_b(), 3); }
    if (s == "pll_lpf") { copy(a, d.filt_iir_lpf_pll_phase_error->get_a(), 2); return copy(b, d.filt_iir_lpf_pll_phase_error->get_b(), 2); }
    if (s == "bpsk_ted_lpf") { auto& f = d.GetBPSKSync().filt_iir_lpf_ted_phase_error; copy(a, f->get_a(), 2); return copy(b, f->get_b(), 2); }
    if (s == "bpsk_pll_lpf") { auto& f = d.GetBPSKSync().filt_iir_lpf_pll_phase_error; copy(a, f->get_a(), 2); return copy(b, f->get_b(), 2); }
    return -1;
}

// RDS results accumulated since creation.
int fmref_n_groups(void* hv) { return (int)((RefHandle*)hv)->rds.groups.size(); }
void fmref_get_groups(void* hv, uint16_t* data, uint8_t* valid, uint8_t* type) {
    auto* h = (RefHandle*)hv;
    size_t k = 0;
    for (auto& g : h->rds.groups) {
        for (int i = 0; i < 4; i++, k++) { data[k] = g.data[i]; valid[k] = g.valid[i]; type[k] = g.type[i]; }
    }
}
int fmref_n_rds_bytes(void* hv) { return (int)((RefHandle*)hv)->rds.rds_bytes.size(); }
void fmref_get_rds_bytes(void* hv, uint8_t* out) {
    auto* h = (RefHandle*)hv;
    std::memcpy(out, h->rds.rds_bytes.data(), h->rds.rds_bytes.size());
}
void fmref_get_db(void* hv, uint16_t* pi, char* ps8, char* rt64, uint8_t* pty) {
    auto* h = (RefHandle*)hv;
    auto& db = h->rds.chain->db;
    *pi = db.PI_code;
    *pty = db.programme_type;
    std::memcpy(ps8, db.service_name, 8);
    std::memcpy(rt64, db.radio_text, 64);
}

// RDS_Database beyond PI / PTY / PS / RT, in the flat layout of fm_oracle.h's fmo_db_ext (24 bytes).
// ptyn_ab_flag is private decoder state in the reference (rds_database_decoder_handler.h:12) and is
// reported as 0xFF here: comparisons skip it.
struct RefDbExt {
    char programme_type_name[8];
    int32_t year;
    uint8_t day, month, hour, minute;
    int8_t local_time_offset;
    uint8_t traffic_announcement;
    uint8_t is_stereo, is_music, is_artificial_head, is_compressed, is_dynamic_program_type;
    uint8_t ptyn_ab_flag;
};
static void fill_db_ext(const RDS_Database& db, RefDbExt* o) {
    std::memcpy(o->programme_type_name, db.programme_type_name, 8);
    o->year = db.datetime.year; o->day = (uint8_t)db.datetime.day; o->month = (uint8_t)db.datetime.month;
    o->hour = db.datetime.hour; o->minute = db.datetime.minute;
    o->local_time_offset = db.local_time_offset;
    o->traffic_announcement = (uint8_t)db.traffic_announcement;
    o->is_stereo = db.is_stereo; o->is_music = db.is_music; o->is_artificial_head = db.is_artificial_head;
    o->is_compressed = db.is_compressed; o->is_dynamic_program_type = db.is_dynamic_program_type;
    o->ptyn_ab_flag = 0xFF;
}
void fmref_get_db_ext(void* hv, RefDbExt* out) { fill_db_ext(((RefHandle*)hv)->rds.chain->db, out); }

// Stand-alone RDS bit path (differential_manchester_decoder.h:25-59 -> rds_group_sync.cpp:29):
// feeds arbitrary soft symbols through the reference decoder; used to judge symbols produced
// by the CUDA path with the reference's own decoder.
void* fmref_rds_create() { return new RefRdsChain(); }
void fmref_rds_destroy(void* r) { delete (RefRdsChain*)r; }
void fmref_rds_push_symbols(void* rv, const float* sym, size_t n) {
    auto* r = (RefRdsChain*)rv;
    r->manchester->Process(tcb::span<const float>(sym, n));
}
int fmref_rds_n_groups(void* rv) { return (int)((RefRdsChain*)rv)->groups.size(); }
void fmref_rds_get_groups(void* rv, uint16_t* data, uint8_t* valid, uint8_t* type) {
    auto* r = (RefRdsChain*)rv;
    size_t k = 0;
    for (auto& g : r->groups) {
        for (int i = 0; i < 4; i++, k++) { data[k] = g.data[i]; valid[k] = g.valid[i]; type[k] = g.type[i]; }
    }
}
int fmref_rds_n_bytes(void* rv) { return (int)((RefRdsChain*)rv)->rds_bytes.size(); }
void fmref_rds_get_bytes(void* rv, uint8_t* out) {
    auto* r = (RefRdsChain*)rv;
    std::memcpy(out, r->rds_bytes.data(), r->rds_bytes.size());
}
void fmref_rds_get_db(void* rv, uint16_t* pi, char* ps8, char* rt64, uint8_t* pty) {
    auto* r = (RefRdsChain*)rv;
    auto& db = r->chain->db;
    *pi = db.PI_code;
    *pty = db.programme_type;
    std::memcpy(ps8, db.service_name, 8);
    std::memcpy(rt64, db.radio_text, 64);
}

void fmref_rds_get_db_ext(void* rv, RefDbExt* out) { fill_db_ext(((RefRdsChain*)rv)->chain->db, out); }

// Filter designers (dsp/filter_designer.h:8-35) for pinning the host-side re-implementation.
void fmref_create_fir_lpf(float* b, int N, float k) { create_fir_lpf(b, N, k); }
void fmref_create_fir_hpf(float* b, int N, float k) { create_fir_hpf(b, N, k); }
void fmref_create_fir_bpf(float* b, int N, float k1, float k2) { create_fir_bpf(b, N, k1, k2); }
void fmref_create_fir_hilbert(float* b, int N) { create_fir_hilbert(b, N); }
void fmref_create_iir_single_pole_lpf(float* b, float* a, float k) { create_iir_single_pole_lpf(b, a, k); }
void fmref_create_iir_notch_filter(float* b, float* a, float k, float r) { create_iir_notch_filter(b, a, k, r); }
void fmref_create_iir_peak_1_filter(float* b, float* a, float k, float r) { create_iir_peak_1_filter(b, a, k, r); }
void fmref_create_iir_peak_2_filter(float* b, float* a, float k, float r, float A_db) { create_iir_peak_2_filter(b, a, k, r, A_db); }
void fmref_create_fir_lpf_window(float* b, int N, float k, int window_id) {
    const window_func_t w[4] = { window_hamming, window_hann, window_blackman, window_blackman_harris };
    create_fir_lpf(b, N, k, w[window_id & 3]);
}

// dsp/polyphase_filter.h:9-87 stand-alone, for pinning the generic resampler kernels.
void fmref_polyphase_ds_f32(int M, int K, const float* b, const float* x, float* y, int N_out, int n_calls) {
    PolyphaseDownsampler<float> f(M, K);
    std::memcpy(f.get_b(), b, sizeof(float)*M*K);
    for (int c = 0; c < n_calls; c++) f.process(x + (size_t)c*N_out*M, y + (size_t)c*N_out, N_out);
}
void fmref_polyphase_ds_cf32(int M, int K, const float* b, const float* x, float* y, int N_out, int n_calls) {
    PolyphaseDownsampler<std::complex<float>> f(M, K);
    std::memcpy(f.get_b(), b, sizeof(float)*M*K);
    auto* xc = (const std::complex<float>*)x;
    auto* yc = (std::complex<float>*)y;
    for (int c = 0; c < n_calls; c++) f.process(xc + (size_t)c*N_out*M, yc + (size_t)c*N_out, N_out);
}
void fmref_polyphase_us_f32(int L, int K, const float* b, const float* x, float* y, int N_in, int n_calls) {
    PolyphaseUpsampler<float> f(b, L, K);
    for (int c = 0; c < n_calls; c++) f.process(x + (size_t)c*N_in, y + (size_t)c*N_in*L, N_in);
}

// audio/resampled_pcm_player.cpp:37-54: the reference's own Resample(), compiled from its source file.
// dsp/fir_filter.h, hilbert_fir_filter.h, iir_filter.h, agc.h stand-alone: n_calls consecutive blocks through ONE object.
void fmref_fir_f32(int K, const float* b, const float* x, float* y, int N, int n_calls) {
    FIR_Filter<float> f(K);
    std::memcpy(f.get_b(), b, sizeof(float)*K);
    for (int c = 0; c < n_calls; c++) f.process(x + (size_t)c*N, y + (size_t)c*N, N);
}
void fmref_fir_cf32(int K, const float* b, const float* x, float* y, int N, int n_calls) {
    FIR_Filter<std::complex<float>> f(K);
    std::memcpy(f.get_b(), b, sizeof(float)*K);
    auto* xc = (const std::complex<float>*)x; auto* yc = (std::complex<float>*)y;
    for (int c = 0; c < n_calls; c++) f.process(xc + (size_t)c*N, yc + (size_t)c*N, N);
}
void fmref_hilbert_f32(int K, const float* x, float* y, int N, int n_calls) {
    Hilbert_FIR_Filter<float> f(K);
    auto* yc = (std::complex<float>*)y;
    for (int c = 0; c < n_calls; c++) f.process(x + (size_t)c*N, yc + (size_t)c*N, N);
}
void fmref_iir_f32(int K, const float* b, const float* a, const float* x, float* y, int N, int n_calls) {
    IIR_Filter<float> f(K);
    std::memcpy(f.get_b(), b, sizeof(float)*K); std::memcpy(f.get_a(), a, sizeof(float)*K);
    for (int c = 0; c < n_calls; c++) f.process(x + (size_t)c*N, y + (size_t)c*N, N);
}
void fmref_iir_cf32(int K, const float* b, const float* a, const float* x, float* y, int N, int n_calls) {
    IIR_Filter<std::complex<float>> f(K);
    std::memcpy(f.get_b(), b, sizeof(float)*K); std::memcpy(f.get_a(), a, sizeof(float)*K);
    auto* xc = (const std::complex<float>*)x; auto* yc = (std::complex<float>*)y;
    for (int c = 0; c < n_calls; c++) f.process(xc + (size_t)c*N, yc + (size_t)c*N, N);
}
void fmref_agc_cf32(float target_power, float beta, float gain0, const float* x, float* y, int N, int n_calls, float* gains_out) {
    AGC_Filter<std::complex<float>> g;
    g.target_power = target_power; g.beta = beta; g.current_gain = gain0;
    auto* xc = (const std::complex<float>*)x; auto* yc = (std::complex<float>*)y;
    for (int c = 0; c < n_calls; c++) { g.process(xc + (size_t)c*N, yc + (size_t)c*N, N); if (gains_out) gains_out[c] = g.current_gain; }
}

void fmref_resample_linear(const float* in, int n_in, float* out, int n_out) {
    Resample(tcb::span<const Frame<float>>((const Frame<float>*)in, (size_t)n_in),
             tcb::span<Frame<float>>((Frame<float>*)out, (size_t)n_out));
}
// fm_scraper.cpp:74-78, the statement itself (Audio_Scraper needs a filesystem target, so the loop is lifted).
void fmref_frames_to_s16(const float* frames, size_t n_frames, int16_t* out) {
    constexpr float CONVERT_RESCALE = float(std::numeric_limits<int16_t>::max()) * 0.95f;
    auto* data = (const Frame<float>*)frames;
    auto* dst = (Frame<int16_t>*)out;
    for (size_t i = 0; i < n_frames; i++) dst[i] = Frame<int16_t>(data[i]*CONVERT_RESCALE);
}

// dsp/calculate_fft_mag.cpp:11-45 and dsp/fftshift.h:21-33: the reference's own class and template.  The trigger is
// SINGLE + raised, i.e. exactly one update per call, as the GUI drives it (gui/render_fm_demod.cpp:399).
void fmref_fft_mag_process(int mode, float beta, const float* x_cf32, float* y, int n) {
    Calculate_FFT_Mag calc;
    calc.SetMode((Calculate_FFT_Mag::Mode)mode);
    calc.GetAverageBeta() = beta;
    calc.SetTrigger(Calculate_FFT_Mag::Trigger::SINGLE);
    calc.RaiseSingleTrigger();
    calc.Process(tcb::span<const std::complex<float>>((const std::complex<float>*)x_cf32, (size_t)n), tcb::span<float>(y, (size_t)n));
}
void fmref_fftshift_inplace(float* x_cf32, int n) {
    InplaceFFTShift(tcb::span<std::complex<float>>((std::complex<float>*)x_cf32, (size_t)n));
}

} // extern "C"
