// TEST INFRASTRUCTURE ONLY.
// Stand-in for /root/reference/src/dsp/calculate_fft.cpp, the only reference file that needs
// FFTW3f (not installed here).  The two functions (dsp/calculate_fft.h:8-14) are reached only
// when a GUI spectrum trigger is raised (broadcast_fm_demod.cpp:27-40), which never happens in
// fm_demod_benchmark or in the oracle harness; abort loudly if that assumption ever breaks.
#include <complex>
#include <cstdio>
#include <cstdlib>
#include "dsp/calculate_fft.h"

void CalculateFFT(tcb::span<const std::complex<float>>, tcb::span<std::complex<float>>) {
    fprintf(stderr, "oracle/fft_stub: CalculateFFT called (FFTW is not available)\n");
    abort();
}
void CalculateIFFT(tcb::span<const std::complex<float>>, tcb::span<std::complex<float>>) {
    fprintf(stderr, "oracle/fft_stub: CalculateIFFT called (FFTW is not available)\n");
    abort();
}
