"""Generates tests/golden/golden_spectra.npz from the UNMODIFIED reference (oracle/_ref/libfmref.so):
Calculate_FFT_Mag::Process (dsp/calculate_fft_mag.cpp:11-45) in its three modes over three consecutive updates,
and InplaceFFTShift (dsp/fftshift.h:21-33).  The DFT inputs come from numpy.fft (the reference's FFT is FFTW3f,
not installed here).

    python tests/golden/make_golden_spectra.py      (needs `make -C oracle ref`, i.e. /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bind  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    L = bind.lib("ref")
    rng = np.random.default_rng(21)
    n = 2048
    t = np.arange(n)
    out = {"beta": np.float32(0.1), "y0": np.full(n, -60.0, np.float32)}
    for rep in range(3):
        x = 80 * np.exp(2j * np.pi * 0.148 * t) + 5 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
        out[f"X{rep}"] = np.fft.fftshift(np.fft.fft(x)).astype(np.complex64)
    for mode in (0, 1, 2):
        y = out["y0"].copy()
        for rep in range(3):
            X = np.ascontiguousarray(out[f"X{rep}"])
            L.fft_mag_process(mode, 0.1, X.ctypes.data, y.ctypes.data, n)
            out[f"mag_m{mode}_r{rep}"] = y.copy()
    z = np.fft.ifftshift(out["X0"]).copy()          # un-shifted copy, so that the reference's shift gives X0 back ...
    out["X0"] = z.copy()                            # ... and the fixture stores (un-shifted, shifted-by-the-reference)
    L.fftshift_inplace(z.ctypes.data, n)
    out["shifted"] = z
    # X0 changed meaning above: recompute the mode results that used it, with the stored X0
    for mode in (0, 1, 2):
        y = out["y0"].copy()
        for rep in range(3):
            X = np.ascontiguousarray(out[f"X{rep}"])
            L.fft_mag_process(mode, 0.1, X.ctypes.data, y.ctypes.data, n)
            out[f"mag_m{mode}_r{rep}"] = y.copy()
    np.savez_compressed(os.path.join(HERE, "golden_spectra.npz"), **out)
    print("wrote golden_spectra.npz")


if __name__ == "__main__":
    main()
