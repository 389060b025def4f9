// Header-compatible replacement of the reference's src/fm_demod/broadcast_fm_demod.h:91-299.
//
// Drop this file, broadcast_fm_demod.cpp and bpsk_synchroniser.h over the reference's files of the
// same names (or put this directory first in an overlay of src/) and link
// libfmgpu.so: src/app.cpp, src/fm_demod_benchmark.cpp, src/fm_demod_no_tuner.cpp and src/gui/*
// compile unchanged.  Same class name, same public methods, same argument meaning, same error
// behaviour (a wrong-sized block is silently ignored, broadcast_fm_demod.cpp:311-313); every
// method forwards to the C-ABI of include/fmgpu.h.  There is no CPU implementation behind it.
//
// Differences, all additive:
//   * ProcessU8(span<const complex<uint8_t>>) feeds rtl-sdr bytes straight to the fused first
//     kernel, saving App::Run's host-side unpack (src/app.cpp:56-65);
//   * buffers are fetched from the device lazily, when a getter is called, into host mirrors that
//     stay valid until the next Process (the lifetime the reference's spans have);
//   * magnitude spectra (display only; FFTW in the reference): UpdateFFTCalc (broadcast_fm_demod.cpp:27-40)
//     runs when a spectrum's trigger is raised -- the DFT + FFT shift on the device (fmgpu_get_fft /
//     fmgpu_calculate_fft), then the reference's own Calculate_FFT_Mag::Process on the host -- for the
//     all eight spectra.  (GUI mode only: fm_in_buf and the complex decimator outputs behind the FM-in and
//     audio spectra are extra stores / a small extra kernel that lean mode, FMGPU_LEAN=1, does not run.)
#pragma once

#include <complex>
#include <memory>
#include <vector>

#include "dsp/calculate_fft_mag.h"
#include "audio/frame.h"
#include "utility/observable.h"
#include "utility/span.h"

struct fmgpu_demod;
class BPSK_Synchroniser;

// broadcast_fm_demod.h:27-40
struct Broadcast_FM_Demod_Analog_Parameters {
    float F_wbfm_deviation = 75e3f;
    int F_audio_lpr = 15000;
    int F_pilot = 19000;
    int F_pilot_deviation = 100;
    int F_audio_lmr_center = 38000;
    int F_audio_lmr_bandwidth = 15000;
    int F_rds_center = 57000;
    int F_rds_bandwidth = 2000;
    int Tus_min_deemphasis = 1;
    int Tus_max_deemphasis = 100;
};

// broadcast_fm_demod.h:64-89
struct Broadcast_FM_Demod_Controls {
    template <typename T>
    struct EditableControl {
    private:
        bool is_dirty = false;
        T value = 0;
    public:
        void SetValue(T v) { value = v; is_dirty = true; }
        void ClearDirty() { is_dirty = false; }
        auto GetValue() const { return value; }
        auto IsDirty() const { return is_dirty; }
    };
    enum AudioOut { LPR, LMR, STEREO };

    AudioOut audio_out = AudioOut::STEREO;
    float audio_stereo_mix_factor = 1.0f;
    bool is_use_deemphasis_filter = false;
    EditableControl<int> filt_deemphasis_cutoff;
    EditableControl<int> filt_audio_lpr_cutoff;
    EditableControl<int> filt_audio_lmr_cutoff;
};

class Broadcast_FM_Demod
{
private:
    const int block_size;
    Broadcast_FM_Demod_Analog_Parameters params;
    Broadcast_FM_Demod_Controls controls;
    fmgpu_demod* handle = nullptr;
    std::unique_ptr<BPSK_Synchroniser> bpsk_sync;
    int rds_total_symbols = 0;

    // host mirrors, refreshed on demand after each Process
    unsigned long long generation = 0;
    struct Mirror { std::vector<unsigned char> bytes; unsigned long long fetched_at = ~0ull; size_t n = 0; };
    Mirror mirrors[32];
    std::vector<Frame<float>> audio_out_buf;
    std::vector<float> rds_pred_sym_buf;
    std::vector<float> fft_mag_bufs[8];
    std::vector<std::complex<float>> fft_tmp;
    Calculate_FFT_Mag calc_fft_mag[8];

    // controls as last sent to the device
    int sent_audio_out = -1;
    float sent_mix = -1.0f;
    int sent_deemph = -1;

    Observable<tcb::span<const Frame<float>>, int> obs_on_audio_block;
    Observable<tcb::span<const float>> obs_on_rds_symbols;
public:
    Broadcast_FM_Demod(const int _block_size);
    ~Broadcast_FM_Demod();
    void Process(tcb::span<const std::complex<float>> x);
    // additive fast path: rtl-sdr bytes, unpack fused on the device
    void ProcessU8(tcb::span<const std::complex<uint8_t>> x);
private:
    void LatchControls();
    void AfterProcess();
    void UpdateSpectra(const float* baseband_cf32);
    template <typename T> tcb::span<T> Fetch(int buf);
public:
    // 1. FM demodulation
    tcb::span<std::complex<float>> GetFMOutIQ();
    // 2. Lock onto pilot
    tcb::span<std::complex<float>> GetPilotOutput();
    tcb::span<std::complex<float>> GetPLLOutput();
    tcb::span<float> Get_PLL_Raw_Phase_Error_Output();
    tcb::span<float> Get_PLL_LPF_Phase_Error_Output();
    // 3. Extract components
    tcb::span<float> GetLPRAudioOutput();
    tcb::span<float> GetLMRAudioOutput();
    tcb::span<std::complex<float>> GetRDSOutput();
    // 4. RDS synchronisation
    tcb::span<float> GetRDSPredSymbols() { return tcb::span<float>(rds_pred_sym_buf).first((size_t)rds_total_symbols); }
    tcb::span<std::complex<float>> GetRDSRawSymbols();
    // 5. Audio mixing
    tcb::span<Frame<float>> GetAudioOut() { return audio_out_buf; }
    // 6. FFT (display only: see header comment)
    tcb::span<float> GetBasebandMagnitudeSpectrum() { return fft_mag_bufs[0]; }
    tcb::span<float> GetFMInMagnitudeSpectrum() { return fft_mag_bufs[1]; }
    tcb::span<float> GetFMOutMagnitudeSpectrum() { return fft_mag_bufs[2]; }
    tcb::span<float> GetPilotMagnitudeSpectrum() { return fft_mag_bufs[3]; }
    tcb::span<float> GetPLLPilotMagnitudeSpectrum() { return fft_mag_bufs[4]; }
    tcb::span<float> GetAudioLPRMagnitudeSpectrum() { return fft_mag_bufs[5]; }
    tcb::span<float> GetAudioLMRMagnitudeSpectrum() { return fft_mag_bufs[6]; }
    tcb::span<float> GetRDSMagnitudeSpectrum() { return fft_mag_bufs[7]; }

    auto& GetBasebandMagnitudeSpectrumControls() { return calc_fft_mag[0]; }
    auto& GetFMInputMagnitudeSpectrumControls() { return calc_fft_mag[1]; }
    auto& GetSignalMagnitudeSpectrumControls() { return calc_fft_mag[2]; }
    auto& GetPilotMagnitudeSpectrumControls() { return calc_fft_mag[3]; }
    auto& GetPLLPilotMagnitudeSpectrumControls() { return calc_fft_mag[4]; }
    auto& GetAudioLPRMagnitudeSpectrumControls() { return calc_fft_mag[5]; }
    auto& GetAudioLMRMagnitudeSpectrumControls() { return calc_fft_mag[6]; }
    auto& GetRDSMagnitudeSpectrumControls() { return calc_fft_mag[7]; }

    // Sample rates (broadcast_fm_demod.cpp:62-72)
    auto GetBasebandSampleRate() const { return 1024000; }
    auto GetFMInSampleRate() const { return 256000; }
    auto GetFMOutSampleRate() const { return 128000; }
    auto GetRDSSampleRate() const { return 16000; }
    auto GetAudioSampleRate() const { return 32000; }

    float GetAudioLMRPhaseError();
    auto& GetBPSKSync() { return *(bpsk_sync.get()); }
    auto& GetAnalogParams() { return params; }
    auto& GetControls() { return controls; }

    auto& OnAudioOut() { return obs_on_audio_block; }
    auto& OnRDSOut() { return obs_on_rds_symbols; }

    // used by the BPSK_Synchroniser shim
    fmgpu_demod* GetHandle() { return handle; }
    unsigned long long GetGeneration() const { return generation; }
};
