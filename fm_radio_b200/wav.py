"""Host side of the audio scraper: the 16-bit stereo WAV container of fm_scraper.cpp:55-160
(Audio_Scraper::create_wav_file / on_audio_data / update_wav_header / close_wav_file).

The samples come from the device already converted (kernel K7, Buf.AUDIO_PCM_S16, or fm.frames_to_s16); this
class only lays the reference's 44-byte header in front of them and keeps its size fields current after every
write, as the reference does.  `reference_sizes=True` reproduces the reference's header byte for byte: it
passes the number of FRAMES written where the format wants bytes (total_bytes_written counts fwrite's element
count, fm_scraper.cpp:84-89), so its data-size fields are 4x too small; the default writes correct sizes.
"""
from __future__ import annotations

import struct

import numpy as np


def wav_header(sample_rate: int, n_data_bytes: int) -> bytes:
    """struct WavHeader of fm_scraper.cpp:107-146: PCM, 2 channels, 16 bits."""
    channels, bits = 2, 16
    return struct.pack("<4si4s4sihhiihh4si", b"RIFF", 36 + n_data_bytes, b"WAVE", b"fmt ", 16, 1, channels, sample_rate,
                       sample_rate * channels * bits // 8, channels * bits // 8, bits, b"data", n_data_bytes)


class WavWriter:
    def __init__(self, path: str, sample_rate: int, reference_sizes: bool = False):
        self.fp = open(path, "wb+")
        self.sample_rate, self.reference_sizes = int(sample_rate), bool(reference_sizes)
        self.frames = 0
        self.fp.write(wav_header(self.sample_rate, 0))

    def write(self, frames_s16: np.ndarray) -> None:
        """frames_s16: [n, 2] int16 (L, R), e.g. FMDemod.get(Buf.AUDIO_PCM_S16).reshape(-1, 2)."""
        a = np.ascontiguousarray(frames_s16, dtype="<i2").reshape(-1, 2)
        self.fp.write(a.tobytes())
        self.frames += a.shape[0]
        self._update_header()

    def _update_header(self) -> None:
        n = self.frames if self.reference_sizes else self.frames * 4
        self.fp.seek(4); self.fp.write(struct.pack("<i", 36 + n))      # ChunkSize
        self.fp.seek(40); self.fp.write(struct.pack("<i", n))          # Subchunk2Size
        self.fp.seek(0, 2)

    def close(self) -> None:
        if self.fp:
            self._update_header()
            self.fp.close()
            self.fp = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
