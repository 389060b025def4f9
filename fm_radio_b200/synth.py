"""Deterministic synthetic broadcast-FM captures (stereo + 19 kHz pilot + 57 kHz RDS) as rtl-sdr u8 IQ.

This is the workload generator of SURVEY.md section 8(d): the reference ships no sample capture, so
every parity test and the benchmark use this recipe.  Two back ends share the same formulas:

* ``synth_u8_numpy``  - float64 numpy, bit-reproducible from the seed; used by tests and the oracle.
* ``synth_u8_torch``  - the same maths on a torch device, used by ``bench.py`` to fill HBM with many
  distinct streams quickly (different RNG, so not byte-identical to the numpy version).

Signal (Fs = 1.024 MS/s, the rate hard-wired in the reference, broadcast_fm_demod.cpp:62-72):

    mpx(t) = 0.40*(L+R)/1.4 + 0.10*sin(wp t + p0) + 0.40*(L-R)/1.4*sin(2(wp t + p0))
           + 0.06*rds(t)*sin(3(wp t + p0))                                   wp = 2 pi 19 kHz
    iq(t)  = A*exp(j*(2 pi 75e3 * cumsum(mpx)/Fs + 2 pi f_off t)) + complex AWGN at `snr_db`
    u8     = clip(round(127 + iq), 0, 255), I then Q interleaved (what App::Run expects,
             src/app.cpp:56-65)

RDS (EN 50067): groups alternate 0A (PI, PTY, one PS segment, AF pair) and 2A (one RadioText
segment); each 26-bit block is 16 data bits + (CRC-10 xor offset word), MSB first; bits are
differentially encoded and sent as biphase chips at 2375 chip/s, each chip shaped by a Hann pulse
of one chip length.
"""
from __future__ import annotations

import dataclasses
import numpy as np

FS_BASEBAND = 1_024_000
F_PILOT = 19_000.0
F_DEVIATION = 75_000.0
RDS_CHIP_RATE = 2375.0  # = 57000/24, two chips per 1187.5 bit/s data bit

# rds_constants.h:15-28 (EN 50067 clause 2.3 / annex A)
RDS_CRC10_POLY_FULL = 0b10110111001  # x^10 + x^8 + x^7 + x^5 + x^4 + x^3 + 1
RDS_OFFSET = {"A": 0b0011111100, "B": 0b0110011000, "C": 0b0101101000, "C1": 0b1101010000, "D": 0b0110110100}


@dataclasses.dataclass
class StreamParams:
    seed: int = 0
    pi_code: int = 0x1234
    pty: int = 10
    ps: str = "B200-FM!"
    radiotext: str = "Blackwell B200 native FM stereo + RDS demodulator test signal....."
    af_pair: tuple = (0x10, 0x20)
    snr_db: float = 40.0
    amplitude: float = 100.0
    pilot_phase: float = 0.0      # radians
    f_offset_hz: float = 0.0      # carrier (tuning) offset
    tones_left: tuple = ((1000.0, 0.5), (3500.0, 0.2))
    tones_right: tuple = ((1700.0, 0.5), (5200.0, 0.2))
    # group 0A flag bits (EN 50067 3.1.5.1) and, when extended_groups is set, one 10A (programme type name)
    # and one 4A (clock-time and date) group after every six 0A/2A groups
    tp: int = 0
    ta: int = 0
    ms: int = 1
    di: int = 0                   # d3 d2 d1 d0 = dynamic PTY, compressed, artificial head, stereo
    extended_groups: bool = False
    ptyn: str = "B200 PTY"
    mjd: int = 61331              # 2026-10-18
    hour: int = 13
    minute: int = 37
    lto: int = -7                 # local time offset in half hours

    @staticmethod
    def for_stream(s: int) -> "StreamParams":
        """Per-stream variation of config 3: seed s, PI 0x1000+s, pilot phase,
        carrier offset within +-2 kHz and SNR within 30..50 dB all derived from s."""
        rng = np.random.default_rng(1_000_003 * (s + 1))
        ps = f"S{s:05d}FM"[:8]
        return StreamParams(
            seed=s, pi_code=(0x1000 + s) & 0xFFFF, pty=(s % 31) + 1, ps=ps,
            radiotext=(f"stream {s:05d} on B200 " * 4)[:64],
            snr_db=float(rng.uniform(30.0, 50.0)),
            pilot_phase=float(rng.uniform(0.0, 2 * np.pi)),
            f_offset_hz=float(rng.uniform(-2000.0, 2000.0)),
        )


def crc10(data16: int) -> int:
    """Remainder of data16 * x^10 modulo g(x) (EN 50067 clause 2.3)."""
    reg = data16 << 10
    for bit in range(25, 9, -1):
        if reg & (1 << bit):
            reg ^= RDS_CRC10_POLY_FULL << (bit - 10)
    return reg & 0x3FF


def rds_block(data16: int, offset_name: str) -> int:
    return ((data16 & 0xFFFF) << 10) | (crc10(data16 & 0xFFFF) ^ RDS_OFFSET[offset_name])


def rds_group_words(p: StreamParams, index: int) -> list[int]:
    """The four 16-bit data words of the index-th transmitted group (0A and 2A alternate; with
    p.extended_groups every 7th is a 10A and every 8th a 4A)."""
    pty_tp = ((p.tp & 1) << 10) | ((p.pty & 31) << 5)
    if p.extended_groups:
        cyc, pos = divmod(index, 8)
        if pos == 6:                                            # 10A, figure 31
            seg = cyc % 2
            name = p.ptyn.ljust(8)[:8].encode("latin-1")
            b = (10 << 12) | pty_tp | (0 << 4) | seg
            return [p.pi_code & 0xFFFF, b, (name[4 * seg] << 8) | name[4 * seg + 1], (name[4 * seg + 2] << 8) | name[4 * seg + 3]]
        if pos == 7:                                            # 4A, figure 20
            minute = (p.minute + cyc) % 60
            b = (4 << 12) | pty_tp | ((p.mjd >> 15) & 3)
            c = ((p.mjd & 0x7FFF) << 1) | ((p.hour >> 4) & 1)
            d = ((p.hour & 15) << 12) | (minute << 6) | ((1 if p.lto < 0 else 0) << 5) | (abs(p.lto) & 31)
            return [p.pi_code & 0xFFFF, b, c, d]
        index = cyc * 6 + pos
    k = index // 2
    if index % 2 == 0:
        seg = k % 4
        b = (0 << 12) | (0 << 11) | pty_tp | ((p.ta & 1) << 4) | ((p.ms & 1) << 3) | (((p.di >> (3 - seg)) & 1) << 2) | seg
        c = ((p.af_pair[0] & 0xFF) << 8) | (p.af_pair[1] & 0xFF)
        ps = p.ps.ljust(8)[:8].encode("latin-1")
        d = (ps[2 * seg] << 8) | ps[2 * seg + 1]
    else:
        seg = k % 16
        b = (2 << 12) | (0 << 11) | pty_tp | (0 << 4) | seg
        rt = p.radiotext.ljust(64)[:64].encode("latin-1")
        c = (rt[4 * seg] << 8) | rt[4 * seg + 1]
        d = (rt[4 * seg + 2] << 8) | rt[4 * seg + 3]
    return [p.pi_code & 0xFFFF, b, c, d]


def rds_bits(p: StreamParams, n_bits: int) -> np.ndarray:
    """First n_bits of the (un-encoded) RDS bit stream, MSB first per 26-bit block."""
    n_groups = (n_bits + 103) // 104
    out = np.zeros(n_groups * 104, dtype=np.uint8)
    pos = 0
    for g in range(n_groups):
        for word, name in zip(rds_group_words(p, g), ("A", "B", "C", "D")):
            blk = rds_block(word, name)
            for bit in range(25, -1, -1):
                out[pos] = (blk >> bit) & 1
                pos += 1
    return out[:n_bits]


def rds_chips(p: StreamParams, n_chips: int) -> np.ndarray:
    """Differentially encoded biphase chips (+-1), two per data bit."""
    n_bits = (n_chips + 1) // 2
    b = rds_bits(p, n_bits)
    d = np.bitwise_xor.accumulate(b)            # d[n] = b[n] ^ d[n-1], d[-1] = 0
    sym = d.astype(np.float64) * 2.0 - 1.0      # 1 -> +1, 0 -> -1
    chips = np.empty(n_bits * 2, dtype=np.float64)
    chips[0::2] = sym
    chips[1::2] = -sym
    return chips[:n_chips]


def _audio(t, tones, xp):
    y = 0.0
    for f, a in tones:
        y = y + a * xp.sin(2 * np.pi * f * t)
    return y


def synth_u8_numpy(n_samples: int, p: StreamParams | None = None) -> np.ndarray:
    """uint8 array of length 2*n_samples (I,Q interleaved)."""
    p = p or StreamParams()
    fs = float(FS_BASEBAND)
    n = np.arange(n_samples, dtype=np.float64)
    t = n / fs
    left = _audio(t, p.tones_left, np)
    right = _audio(t, p.tones_right, np)
    wp = 2 * np.pi * F_PILOT * t + p.pilot_phase
    n_chips = int(np.ceil(n_samples * RDS_CHIP_RATE / fs)) + 2
    chips = rds_chips(p, n_chips)
    u = t * RDS_CHIP_RATE
    ci = np.floor(u).astype(np.int64)
    rds = chips[ci] * np.sin(np.pi * (u - ci)) ** 2
    mpx = (0.40 * (left + right) / 1.4 + 0.10 * np.sin(wp)
           + 0.40 * (left - right) / 1.4 * np.sin(2 * wp) + 0.06 * rds * np.sin(3 * wp))
    phi = 2 * np.pi * F_DEVIATION * np.cumsum(mpx) / fs + 2 * np.pi * p.f_offset_hz * t
    rng = np.random.default_rng(p.seed)
    sigma = p.amplitude * 10.0 ** (-p.snr_db / 20.0) * np.sqrt(0.5)
    noise = rng.standard_normal((n_samples, 2)) * sigma
    iq = np.empty((n_samples, 2), dtype=np.float64)
    iq[:, 0] = p.amplitude * np.cos(phi) + noise[:, 0]
    iq[:, 1] = p.amplitude * np.sin(phi) + noise[:, 1]
    return np.clip(np.rint(127.0 + iq), 0, 255).astype(np.uint8).reshape(-1)


def synth_u8_torch(n_samples: int, params: list[StreamParams], device):
    """torch.uint8 tensor [len(params), 2*n_samples] on `device`, one row per stream: the same
    recipe as synth_u8_numpy evaluated with torch ops (float64 phase, torch's own noise generator)."""
    import torch
    fs = float(FS_BASEBAND)
    dev = torch.device(device)
    t = torch.arange(n_samples, dtype=torch.float64, device=dev) / fs
    out = torch.empty((len(params), 2 * n_samples), dtype=torch.uint8, device=dev)
    n_chips = int(np.ceil(n_samples * RDS_CHIP_RATE / fs)) + 2
    u = t * RDS_CHIP_RATE
    ci = torch.floor(u).to(torch.int64)
    shape = torch.sin(np.pi * (u - ci)) ** 2
    gen = torch.Generator(device=dev)
    for s, p in enumerate(params):
        left = _audio(t, p.tones_left, torch)
        right = _audio(t, p.tones_right, torch)
        wp = 2 * np.pi * F_PILOT * t + p.pilot_phase
        chips = torch.from_numpy(rds_chips(p, n_chips)).to(dev)
        rds = chips[ci] * shape
        mpx = (0.40 * (left + right) / 1.4 + 0.10 * torch.sin(wp)
               + 0.40 * (left - right) / 1.4 * torch.sin(2 * wp) + 0.06 * rds * torch.sin(3 * wp))
        phi = 2 * np.pi * F_DEVIATION * torch.cumsum(mpx, 0) / fs + 2 * np.pi * p.f_offset_hz * t
        gen.manual_seed(p.seed)
        sigma = p.amplitude * 10.0 ** (-p.snr_db / 20.0) * np.sqrt(0.5)
        noise = torch.randn((n_samples, 2), generator=gen, device=dev, dtype=torch.float32) * sigma
        iq = torch.stack((p.amplitude * torch.cos(phi), p.amplitude * torch.sin(phi)), dim=1).to(torch.float32) + noise
        out[s] = torch.clamp(torch.round(127.0 + iq), 0, 255).to(torch.uint8).reshape(-1)
    return out


# ---------------------------------------------------------------------------------------------
# Wideband capture (BASELINE config 4): many stations on a raster in one complex u8 stream.
# ---------------------------------------------------------------------------------------------
FS_WIDEBAND = 20_480_000          # = 20 x the chain's 1.024 MS/s: "20 MHz" of spectrum, integer decimation
RASTER_HZ = 200_000.0


def wideband_centres(n_stations: int = 100, spacing_hz: float = RASTER_HZ) -> np.ndarray:
    """Centre frequencies (Hz, relative to the capture's centre) of n stations on a raster, symmetric about 0
    and never at 0 Hz for even n (100 stations: -9.9 MHz ... +9.9 MHz)."""
    return (np.arange(n_stations, dtype=np.float64) - (n_stations - 1) / 2.0) * spacing_hz


def wideband_amplitude(n_stations: int) -> float:
    """Per-station amplitude that puts the 4-sigma point of the sum at the u8 rails."""
    return 127.0 / (4.0 * np.sqrt(max(n_stations, 1) / 2.0))


def _mpx_chunk(xp, t, p: StreamParams, chips, fs):
    left = _audio(t, p.tones_left, xp)
    right = _audio(t, p.tones_right, xp)
    wp = 2 * np.pi * F_PILOT * t + p.pilot_phase
    u = t * RDS_CHIP_RATE
    ci = xp.floor(u)
    frac = u - ci
    ci = ci.astype(np.int64) if xp is np else ci.to(chips.device).long()
    rds = chips[ci] * xp.sin(np.pi * frac) ** 2
    return (0.40 * (left + right) / 1.4 + 0.10 * xp.sin(wp)
            + 0.40 * (left - right) / 1.4 * xp.sin(2 * wp) + 0.06 * rds * xp.sin(3 * wp))


def synth_wideband_u8(n_samples: int, centres_hz, params: list[StreamParams], fs: float = FS_WIDEBAND,
                      amplitude: float | None = None, device=None, chunk: int = 1 << 21):
    """uint8 [2*n_samples] (I,Q interleaved): sum over stations of A exp(j(phi_s(t) + 2 pi (f_c + f_off) t)),
    phi_s the FM phase of the same stereo+RDS multiplex as synth_u8_numpy evaluated at the wideband
    rate, quantised like an SDR front end: clip(round(127 + .), 0, 255).  device=None: float64 numpy
    (bit-reproducible); otherwise a torch device (same formulas, returns a torch tensor there)."""
    centres_hz = np.asarray(centres_hz, np.float64)
    assert len(centres_hz) == len(params)
    amp = wideband_amplitude(len(params)) if amplitude is None else float(amplitude)
    n_chips = int(np.ceil(n_samples * RDS_CHIP_RATE / fs)) + 2
    if device is None:
        xp = np
        out = np.empty(2 * n_samples, np.uint8)
        chips = [rds_chips(p, n_chips) for p in params]
        carry = [0.0] * len(params)
    else:
        import torch
        xp = torch
        dev = torch.device(device)
        out = torch.empty(2 * n_samples, dtype=torch.uint8, device=dev)
        chips = [torch.from_numpy(rds_chips(p, n_chips)).to(dev) for p in params]
        carry = [0.0] * len(params)
    for n0 in range(0, n_samples, chunk):
        n1 = min(n0 + chunk, n_samples)
        if device is None:
            t = np.arange(n0, n1, dtype=np.float64) / fs
            re = np.zeros(n1 - n0); im = np.zeros(n1 - n0)
        else:
            t = torch.arange(n0, n1, dtype=torch.float64, device=dev) / fs
            re = torch.zeros(n1 - n0, dtype=torch.float64, device=dev); im = torch.zeros_like(re)
        for s, p in enumerate(params):
            mpx = _mpx_chunk(xp, t, p, chips[s], fs)
            phi = carry[s] + 2 * np.pi * F_DEVIATION * xp.cumsum(mpx, 0) / fs
            carry[s] = float(phi[-1])
            ang = phi + 2 * np.pi * (centres_hz[s] + p.f_offset_hz) * t
            re += amp * xp.cos(ang)
            im += amp * xp.sin(ang)
        if device is None:
            out[2 * n0:2 * n1:2] = np.clip(np.rint(127.0 + re), 0, 255).astype(np.uint8)
            out[2 * n0 + 1:2 * n1:2] = np.clip(np.rint(127.0 + im), 0, 255).astype(np.uint8)
        else:
            out[2 * n0:2 * n1:2] = torch.clamp(torch.round(127.0 + re), 0, 255).to(torch.uint8)
            out[2 * n0 + 1:2 * n1:2] = torch.clamp(torch.round(127.0 + im), 0, 255).to(torch.uint8)
    return out
