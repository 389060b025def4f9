"""fm-radio-b200: B200-native (sm_100a) broadcast-FM stereo + RDS demodulation chain.

The product is the CUDA library behind include/fmgpu.h (fm_radio_b200/libfmgpu.so, built by
fm_radio_b200/build.py); this package is its host-side mirror of the reference's
Broadcast_FM_Demod / src/dsp interfaces.  Nothing here imports oracle/.
"""
from .api import (FMDemod, Channelizer, ChanMode, RDSDecoder, PolyphaseDownsampler, PolyphaseUpsampler, FIRFilter, HilbertFIRFilter, IIRFilter, AGCFilter, resample_linear, frames_to_s16, calculate_fft, Buf, Scalar, Control, Filter, FMGPUError,
                  create_fir_lpf, create_fir_hpf, create_fir_bpf, create_fir_hilbert,
                  create_iir_single_pole_lpf, create_iir_notch_filter, create_iir_peak_1_filter, create_iir_peak_2_filter,
                  create_fir_lpf_window, create_fir_hpf_window, create_fir_bpf_window, WINDOWS, lib)
from . import synth

__all__ = ["FMDemod", "Channelizer", "ChanMode", "RDSDecoder", "PolyphaseDownsampler", "PolyphaseUpsampler", "FIRFilter", "HilbertFIRFilter", "IIRFilter", "AGCFilter", "resample_linear", "frames_to_s16", "calculate_fft", "Buf", "Scalar", "Control", "Filter", "FMGPUError",
           "create_fir_lpf", "create_fir_hpf", "create_fir_bpf", "create_fir_hilbert",
           "create_iir_single_pole_lpf", "create_iir_notch_filter", "create_iir_peak_1_filter", "create_iir_peak_2_filter",
           "create_fir_lpf_window", "create_fir_hpf_window", "create_fir_bpf_window", "WINDOWS", "lib", "synth"]
