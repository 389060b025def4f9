// Test infrastructure: drives the Broadcast_FM_Demod SHIM (fm_radio_b200/csrc/shim) the way the reference's GUI does
// (gui/render_fm_demod.cpp:399 raises a spectrum's single trigger every frame while its window is open) and prints,
// per spectrum, the bin and level of its maximum after the last block.  Built by build_overlay.sh next to the
// reference's own driver; read by tests/test_spectra.py.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <complex>
#include "fm_demod/broadcast_fm_demod.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s capture.u8 block_size [n_blocks]\n", argv[0]); return 2; }
    const int B = atoi(argv[2]);
    const int n_blocks = argc > 3 ? atoi(argv[3]) : 60;
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) { perror("open"); return 2; }
    Broadcast_FM_Demod demod(B);
    Calculate_FFT_Mag* calcs[8] = { &demod.GetBasebandMagnitudeSpectrumControls(), &demod.GetFMInputMagnitudeSpectrumControls(),
        &demod.GetSignalMagnitudeSpectrumControls(), &demod.GetPilotMagnitudeSpectrumControls(),
        &demod.GetPLLPilotMagnitudeSpectrumControls(), &demod.GetAudioLPRMagnitudeSpectrumControls(),
        &demod.GetAudioLMRMagnitudeSpectrumControls(), &demod.GetRDSMagnitudeSpectrumControls() };
    std::vector<uint8_t> raw(2 * (size_t)B);
    std::vector<std::complex<float>> x((size_t)B);
    for (int k = 0; k < n_blocks; k++) {
        if (fread(raw.data(), 1, raw.size(), fp) != raw.size()) break;
        for (int i = 0; i < B; i++) x[i] = { (float)raw[2 * i] - 127.0f, (float)raw[2 * i + 1] - 127.0f };   // App::Run, app.cpp:56-65
        for (auto* c : calcs) c->RaiseSingleTrigger();
        demod.Process(x);
    }
    tcb::span<float> specs[8] = { demod.GetBasebandMagnitudeSpectrum(), demod.GetFMInMagnitudeSpectrum(), demod.GetFMOutMagnitudeSpectrum(),
        demod.GetPilotMagnitudeSpectrum(), demod.GetPLLPilotMagnitudeSpectrum(), demod.GetAudioLPRMagnitudeSpectrum(),
        demod.GetAudioLMRMagnitudeSpectrum(), demod.GetRDSMagnitudeSpectrum() };
    const char* names[8] = { "baseband", "fm_in", "fm_out", "pilot", "pll", "audio_lpr", "audio_lmr", "rds" };
    for (int i = 0; i < 8; i++) {
        size_t arg = 0; float mx = -1e30f, mn = 1e30f;
        for (size_t j = 0; j < specs[i].size(); j++) { if (specs[i][j] > mx) { mx = specs[i][j]; arg = j; } if (specs[i][j] < mn) mn = specs[i][j]; }
        printf("%s n=%zu argmax=%zu max=%.3f min=%.3f\n", names[i], specs[i].size(), arg, mx, mn);
    }
    fclose(fp);
    return 0;
}
