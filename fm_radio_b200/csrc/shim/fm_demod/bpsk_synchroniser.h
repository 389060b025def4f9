// Header-compatible replacement of the reference's src/fm_demod/bpsk_synchroniser.h:34-85 for the
// GUI (src/gui/render_bpsk_sync.cpp:12-65): the same getters, served from the device's display
// arrays of kernel K5 through include/fmgpu.h.  Process() is not exposed: the synchroniser runs
// inside the fused chain.
#pragma once

#include <complex>
#include <vector>
#include "utility/span.h"

class Broadcast_FM_Demod;

struct BPSK_Synchroniser_Config {        // bpsk_synchroniser.h:18-32
    float F_sample_rate = 16e3f;
    float F_symbol_rate = 2e3f;
    struct { float integrator_gain = 10.0f; float proportional_gain = 0.3f; } ted_phase_error;
    struct { float integrator_gain = 10.0f; float proportional_gain = 0.3f; } pll_phase_error;
    float ted_max_freq_offset = 1.5e3f;
    float pll_max_freq_offset = 10.0f;
    float agc_target_power = 0.5f;
};

class BPSK_Synchroniser
{
private:
    Broadcast_FM_Demod& owner;
    const int block_size;
    BPSK_Synchroniser_Config cfg;
    struct Mirror { std::vector<unsigned char> bytes; unsigned long long fetched_at = ~0ull; };
    mutable Mirror mirrors[8];
    template <typename T> tcb::span<const T> Fetch(int slot, int buf) const;
public:
    BPSK_Synchroniser(Broadcast_FM_Demod& _owner, const int _block_size) : owner(_owner), block_size(_block_size) {}
    const auto& GetConfig() const { return cfg; }
    tcb::span<const std::complex<float>> GetPLLSymbols() const;
    tcb::span<const bool> GetZeroCrossings() const;
    tcb::span<const bool> GetIntDumpTriggers() const;
    tcb::span<const float> GetTEDRawPhaseError() const;
    tcb::span<const float> GetTEDPIPhaseError() const;
    tcb::span<const float> GetPLLRawPhaseError() const;
    tcb::span<const float> GetPLLPIPhaseError() const;
    tcb::span<const std::complex<float>> GetIntDumpFilter() const;
};
