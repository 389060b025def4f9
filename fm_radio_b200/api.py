"""ctypes binding of the C-ABI (include/fmgpu.h) plus thin host-side classes.

``FMDemod`` mirrors the reference's ``Broadcast_FM_Demod`` block interface
(src/fm_demod/broadcast_fm_demod.h:228-299) for a batch of independent streams;
``RDSDecoder`` is the host RDS bit path; ``PolyphaseDownsampler`` and the ``create_*`` designers
mirror src/dsp.  There is no CPU fallback: loading fails loudly if libfmgpu.so is missing and every
compute call fails if no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FMGPU_LIB") or os.path.join(HERE, "libfmgpu.so")     # FMGPU_LIB: an A/B build (fm_radio_b200/build.py)


class FMGPUError(RuntimeError):
    pass


class Buf(enum.IntEnum):
    AUDIO_OUT = 0
    RDS_PRED_SYM = 1
    RDS_SYM_COUNT = 2
    FM_DEMOD = 3
    FM_OUT_IQ = 4
    PILOT = 5
    PLL_DT = 6
    PLL = 7
    PLL_RAW_PHASE_ERROR = 8
    PLL_LPF_PHASE_ERROR = 9
    AUDIO_LPR = 10
    AUDIO_LMR = 11
    RDS = 12
    RDS_RAW_SYM = 13
    BPSK_PLL_SYM = 14
    BPSK_ZCD = 15
    BPSK_INT_DUMP_TRIGGER = 16
    BPSK_TED_RAW_PHASE_ERROR = 17
    BPSK_TED_PI_PHASE_ERROR = 18
    BPSK_PLL_RAW_PHASE_ERROR = 19
    BPSK_PLL_PI_PHASE_ERROR = 20
    BPSK_INT_DUMP_FILTER = 21
    AUDIO_PCM_F32 = 22
    AUDIO_PCM_S16 = 23
    FM_IN = 24
    AUDIO_LPR_IQ = 25
    AUDIO_LMR_IQ = 26


_BUF_DTYPE = {
    Buf.AUDIO_OUT: (np.float32, 2), Buf.RDS_PRED_SYM: (np.float32, 1), Buf.RDS_SYM_COUNT: (np.int32, 1),
    Buf.FM_DEMOD: (np.float32, 1), Buf.FM_OUT_IQ: (np.complex64, 1), Buf.PILOT: (np.complex64, 1),
    Buf.PLL_DT: (np.float32, 1), Buf.PLL: (np.complex64, 1), Buf.PLL_RAW_PHASE_ERROR: (np.float32, 1),
    Buf.PLL_LPF_PHASE_ERROR: (np.float32, 1), Buf.AUDIO_LPR: (np.float32, 1), Buf.AUDIO_LMR: (np.float32, 1),
    Buf.RDS: (np.complex64, 1), Buf.RDS_RAW_SYM: (np.complex64, 1), Buf.BPSK_PLL_SYM: (np.complex64, 1),
    Buf.BPSK_ZCD: (np.uint8, 1), Buf.BPSK_INT_DUMP_TRIGGER: (np.uint8, 1),
    Buf.BPSK_TED_RAW_PHASE_ERROR: (np.float32, 1), Buf.BPSK_TED_PI_PHASE_ERROR: (np.float32, 1),
    Buf.BPSK_PLL_RAW_PHASE_ERROR: (np.float32, 1), Buf.BPSK_PLL_PI_PHASE_ERROR: (np.float32, 1),
    Buf.BPSK_INT_DUMP_FILTER: (np.complex64, 1),
    Buf.AUDIO_PCM_F32: (np.float32, 2), Buf.AUDIO_PCM_S16: (np.int16, 2), Buf.FM_IN: (np.complex64, 1),
    Buf.AUDIO_LPR_IQ: (np.complex64, 1), Buf.AUDIO_LMR_IQ: (np.complex64, 1),
}


class Scalar(enum.IntEnum):
    AUDIO_LMR_PHASE_ERROR = 0
    AGC_PILOT_GAIN = 1
    AGC_RDS_GAIN = 2


class Control(enum.IntEnum):
    AUDIO_OUT = 0
    AUDIO_STEREO_MIX_FACTOR = 1
    USE_DEEMPHASIS = 2
    DEEMPHASIS_TUS = 3
    AUDIO_LPR_CUTOFF_HZ = 4
    AUDIO_LMR_CUTOFF_HZ = 5
    AUDIO_PCM_RATE_HZ = 6


class Filter(enum.IntEnum):
    FM_IN = 0
    FM_OUT = 1
    HILBERT = 2
    AUDIO_LPR = 3
    AUDIO_LMR = 4
    RDS = 5
    DEEMPHASIS = 6
    PEAK_PILOT = 7
    PLL_LPF = 8
    BPSK_TED_LPF = 9
    BPSK_PLL_LPF = 10


FILTER_LEN = {Filter.FM_IN: 64, Filter.FM_OUT: 64, Filter.HILBERT: 65, Filter.AUDIO_LPR: 128,
              Filter.AUDIO_LMR: 128, Filter.RDS: 128, Filter.DEEMPHASIS: 2, Filter.PEAK_PILOT: 3,
              Filter.PLL_LPF: 2, Filter.BPSK_TED_LPF: 2, Filter.BPSK_PLL_LPF: 2}
FILTER_IS_IIR = {Filter.DEEMPHASIS, Filter.PEAK_PILOT, Filter.PLL_LPF, Filter.BPSK_TED_LPF, Filter.BPSK_PLL_LPF}


class _Config(C.Structure):
    _fields_ = [("block_size", C.c_int), ("n_streams", C.c_int), ("device", C.c_int),
                ("keep_intermediates", C.c_int), ("pipeline_depth", C.c_int)]


class _ChanConfig(C.Structure):
    _fields_ = [("fs_in_hz", C.c_double), ("decimation", C.c_int), ("n_taps", C.c_int), ("cutoff_k", C.c_float),
                ("n_channels", C.c_int), ("block_out", C.c_int), ("device", C.c_int), ("mode", C.c_int),
                ("ring_depth", C.c_int)]


class ChanMode(enum.IntEnum):
    AUTO = 0
    TENSOR = 1
    FP32 = 2


class RDSGroup(C.Structure):
    _fields_ = [("data", C.c_uint16 * 4), ("valid", C.c_uint8 * 4), ("type", C.c_uint8 * 4)]


# every symbol include/fmgpu.h declares; tests check the library exports all of them
EXPORTED_SYMBOLS = [
    "fmgpu_create", "fmgpu_destroy", "fmgpu_process_u8", "fmgpu_process_cf32", "fmgpu_enqueue_u8_device",
    "fmgpu_enqueue_u8_host", "fmgpu_sync", "fmgpu_fetch_outputs", "fmgpu_wait_external_stream",
    "fmgpu_signal_external_stream", "fmgpu_get_buffer", "fmgpu_get_device_buffer", "fmgpu_get_scalar",
    "fmgpu_set_control", "fmgpu_upload_taps", "fmgpu_download_taps", "fmgpu_get_rates", "fmgpu_get_config",
    "fmgpu_launch_count", "fmgpu_profile_stages", "fmgpu_create_fir_lpf", "fmgpu_create_fir_hpf", "fmgpu_create_fir_bpf",
    "fmgpu_create_fir_hilbert", "fmgpu_create_iir_single_pole_lpf", "fmgpu_create_iir_notch_filter",
    "fmgpu_create_iir_peak_1_filter", "fmgpu_polyphase_ds_create", "fmgpu_polyphase_destroy",
    "fmgpu_polyphase_get_b", "fmgpu_polyphase_ds_process", "fmgpu_rds_create", "fmgpu_rds_destroy",
    "fmgpu_rds_push_symbols", "fmgpu_rds_n_groups", "fmgpu_rds_get_groups", "fmgpu_rds_n_bytes",
    "fmgpu_rds_get_bytes", "fmgpu_rds_get_db", "fmgpu_last_error", "fmgpu_version",
    "fmgpu_rds_device_fetch", "fmgpu_rds_device_counts", "fmgpu_rds_device_get_groups",
    "fmgpu_rds_device_get_bytes", "fmgpu_rds_device_get_db", "fmgpu_get_partition",
    "fmgpu_rds_get_db_ext", "fmgpu_rds_device_get_db_ext",
    "fmgpu_dsp_filter_create", "fmgpu_dsp_filter_destroy", "fmgpu_dsp_filter_get_b", "fmgpu_dsp_filter_get_a", "fmgpu_dsp_filter_get_K",
    "fmgpu_dsp_filter_process", "fmgpu_agc_init", "fmgpu_agc_process",
    "fmgpu_enqueue_cf32_device", "fmgpu_stream_wait_input_free",
    "fmgpu_chan_create", "fmgpu_chan_destroy", "fmgpu_chan_get_b", "fmgpu_chan_get_config", "fmgpu_chan_get_freqs",
    "fmgpu_chan_process_u8", "fmgpu_chan_enqueue_u8_device", "fmgpu_chan_feed_device",
    "fmgpu_chan_wait_external_stream", "fmgpu_chan_sync", "fmgpu_chan_stream", "fmgpu_chan_launch_count",
    "fmgpu_profile_stages7", "fmgpu_polyphase_us_create", "fmgpu_polyphase_us_process", "fmgpu_resample_linear",
    "fmgpu_frames_to_s16", "fmgpu_calculate_fft", "fmgpu_get_fft", "fmgpu_set_option", "fmgpu_set_fetch_mask",
    "fmgpu_create_iir_peak_2_filter", "fmgpu_window_hamming", "fmgpu_window_hann", "fmgpu_window_blackman",
    "fmgpu_window_blackman_harris", "fmgpu_create_fir_lpf_window", "fmgpu_create_fir_hpf_window", "fmgpu_create_fir_bpf_window",
]

_lib = None


class RdsDbExt(C.Structure):
    """fmgpu_rds_db_ext (include/fmgpu.h): RDS_Database beyond PI / PTY / PS / RT (rds_database.h:26-53)."""
    _fields_ = [("programme_type_name", C.c_char * 8), ("year", C.c_int32), ("day", C.c_uint8), ("month", C.c_uint8),
                ("hour", C.c_uint8), ("minute", C.c_uint8), ("local_time_offset", C.c_int8),
                ("traffic_announcement", C.c_uint8), ("is_stereo", C.c_uint8), ("is_music", C.c_uint8),
                ("is_artificial_head", C.c_uint8), ("is_compressed", C.c_uint8), ("is_dynamic_program_type", C.c_uint8),
                ("ptyn_ab_flag", C.c_uint8)]

    TRAFFIC = ("NONE", "EON_INFO", "AWAIT_EON_ANNOUNCE", "NOW_EON_ANNOUNCE")    # rds_database.h:19-24

    def as_dict(self) -> dict:
        d = {name: getattr(self, name) for name, _ in self._fields_ if name not in ("programme_type_name", "ptyn_ab_flag")}
        d["programme_type_name"] = bytes(bytearray(self)[:8])
        return d


def lib():
    """Loads libfmgpu.so (no fallback: raises if the CUDA extension has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FMGPUError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, ci, cs = C.c_void_p, C.c_int, C.c_size_t
    L.fmgpu_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
    L.fmgpu_destroy.argtypes = [vp]
    L.fmgpu_destroy.restype = None
    L.fmgpu_process_u8.argtypes = [vp, vp, cs]
    L.fmgpu_process_cf32.argtypes = [vp, vp, cs]
    L.fmgpu_enqueue_u8_device.argtypes = [vp, vp]
    L.fmgpu_enqueue_u8_host.argtypes = [vp, vp]
    L.fmgpu_sync.argtypes = [vp]
    L.fmgpu_fetch_outputs.argtypes = [vp, ci]
    L.fmgpu_wait_external_stream.argtypes = [vp, vp]
    L.fmgpu_signal_external_stream.argtypes = [vp, vp]
    L.fmgpu_get_buffer.argtypes = [vp, ci, ci, C.POINTER(vp), C.POINTER(cs)]
    L.fmgpu_get_device_buffer.argtypes = [vp, ci, ci, C.POINTER(vp), C.POINTER(cs)]
    L.fmgpu_get_scalar.argtypes = [vp, ci, ci, C.POINTER(C.c_float)]
    L.fmgpu_set_control.argtypes = [vp, ci, C.c_double]
    L.fmgpu_upload_taps.argtypes = [vp, ci, vp, vp, ci]
    L.fmgpu_download_taps.argtypes = [vp, ci, vp, vp, ci]
    L.fmgpu_get_rates.argtypes = [vp, C.POINTER(ci * 5)]
    L.fmgpu_get_config.argtypes = [vp, C.POINTER(_Config)]
    L.fmgpu_launch_count.argtypes = [vp]
    L.fmgpu_profile_stages.argtypes = [vp, vp, ci, C.POINTER(C.c_float * 6)]
    L.fmgpu_profile_stages7.argtypes = [vp, vp, ci, C.POINTER(C.c_float * 7)]
    L.fmgpu_launch_count.restype = C.c_longlong
    for name in ("lpf", "hpf"):
        getattr(L, f"fmgpu_create_fir_{name}").argtypes = [vp, ci, C.c_float]
        getattr(L, f"fmgpu_create_fir_{name}").restype = None
    L.fmgpu_create_fir_bpf.argtypes = [vp, ci, C.c_float, C.c_float]
    L.fmgpu_create_fir_bpf.restype = None
    L.fmgpu_create_fir_hilbert.argtypes = [vp, ci]
    L.fmgpu_create_fir_hilbert.restype = None
    L.fmgpu_create_iir_single_pole_lpf.argtypes = [vp, vp, C.c_float]
    L.fmgpu_create_iir_single_pole_lpf.restype = None
    L.fmgpu_create_iir_notch_filter.argtypes = [vp, vp, C.c_float, C.c_float]
    L.fmgpu_create_iir_notch_filter.restype = None
    L.fmgpu_create_iir_peak_1_filter.argtypes = [vp, vp, C.c_float, C.c_float]
    L.fmgpu_create_iir_peak_1_filter.restype = None
    L.fmgpu_create_iir_peak_2_filter.argtypes = [vp, vp, C.c_float, C.c_float, C.c_float]
    L.fmgpu_create_iir_peak_2_filter.restype = None
    L.fmgpu_polyphase_ds_create.argtypes = [ci, ci, ci, C.POINTER(vp)]
    L.fmgpu_polyphase_destroy.argtypes = [vp]
    L.fmgpu_polyphase_destroy.restype = None
    L.fmgpu_polyphase_get_b.argtypes = [vp]
    L.fmgpu_polyphase_get_b.restype = C.POINTER(C.c_float)
    L.fmgpu_polyphase_ds_process.argtypes = [vp, vp, vp, ci]
    L.fmgpu_polyphase_us_create.argtypes = [vp, ci, ci, ci, C.POINTER(vp)]
    L.fmgpu_polyphase_us_process.argtypes = [vp, vp, vp, ci]
    L.fmgpu_resample_linear.argtypes = [vp, ci, vp, ci]
    L.fmgpu_frames_to_s16.argtypes = [vp, cs, vp]
    L.fmgpu_calculate_fft.argtypes = [vp, vp, ci, ci]
    L.fmgpu_get_fft.argtypes = [vp, ci, ci, ci, vp, C.POINTER(cs)]
    L.fmgpu_rds_create.restype = vp
    L.fmgpu_rds_destroy.argtypes = [vp]
    L.fmgpu_rds_destroy.restype = None
    L.fmgpu_rds_push_symbols.argtypes = [vp, vp, cs]
    L.fmgpu_rds_push_symbols.restype = None
    L.fmgpu_rds_n_groups.argtypes = [vp]
    L.fmgpu_rds_get_groups.argtypes = [vp, C.POINTER(RDSGroup), ci]
    L.fmgpu_rds_n_bytes.argtypes = [vp]
    L.fmgpu_rds_get_bytes.argtypes = [vp, vp, ci]
    L.fmgpu_rds_get_db.argtypes = [vp, C.POINTER(C.c_uint16), vp, vp, C.POINTER(C.c_uint8)]
    L.fmgpu_rds_get_db.restype = None
    L.fmgpu_rds_get_db_ext.argtypes = [vp, vp]
    L.fmgpu_dsp_filter_create.argtypes = [ci, ci, ci, C.POINTER(vp)]
    L.fmgpu_dsp_filter_destroy.argtypes = [vp]
    L.fmgpu_dsp_filter_destroy.restype = None
    L.fmgpu_dsp_filter_get_b.argtypes = [vp]
    L.fmgpu_dsp_filter_get_b.restype = C.POINTER(C.c_float)
    L.fmgpu_dsp_filter_get_a.argtypes = [vp]
    L.fmgpu_dsp_filter_get_a.restype = C.POINTER(C.c_float)
    L.fmgpu_dsp_filter_get_K.argtypes = [vp]
    L.fmgpu_dsp_filter_process.argtypes = [vp, vp, vp, ci]
    L.fmgpu_agc_init.argtypes = [vp]
    L.fmgpu_agc_init.restype = None
    L.fmgpu_agc_process.argtypes = [vp, vp, vp, ci]
    L.fmgpu_rds_get_db_ext.restype = None
    L.fmgpu_rds_device_get_db_ext.argtypes = [vp, ci, vp]
    L.fmgpu_get_partition.argtypes = [vp, C.POINTER(ci * 2)]
    L.fmgpu_set_option.argtypes = [vp, C.c_char_p, ci]
    L.fmgpu_set_fetch_mask.argtypes = [vp, C.c_uint]
    L.fmgpu_rds_device_fetch.argtypes = [vp]
    L.fmgpu_rds_device_counts.argtypes = [vp, ci, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(ci * 2)]
    L.fmgpu_rds_device_get_groups.argtypes = [vp, ci, C.c_ulonglong, C.POINTER(RDSGroup), ci]
    L.fmgpu_rds_device_get_bytes.argtypes = [vp, ci, C.c_ulonglong, vp, ci]
    L.fmgpu_rds_device_get_db.argtypes = [vp, ci, C.POINTER(C.c_uint16), vp, vp, C.POINTER(C.c_uint8)]
    L.fmgpu_enqueue_cf32_device.argtypes = [vp, vp, vp]
    L.fmgpu_stream_wait_input_free.argtypes = [vp, vp]
    L.fmgpu_chan_create.argtypes = [C.POINTER(_ChanConfig), C.POINTER(C.c_double), C.POINTER(vp)]
    L.fmgpu_chan_destroy.argtypes = [vp]
    L.fmgpu_chan_destroy.restype = None
    L.fmgpu_chan_get_b.argtypes = [vp]
    L.fmgpu_chan_get_b.restype = C.POINTER(C.c_float)
    L.fmgpu_chan_get_config.argtypes = [vp, C.POINTER(_ChanConfig)]
    L.fmgpu_chan_get_freqs.argtypes = [vp, vp, vp]
    L.fmgpu_chan_process_u8.argtypes = [vp, vp, cs, vp]
    L.fmgpu_chan_enqueue_u8_device.argtypes = [vp, vp, C.POINTER(vp)]
    L.fmgpu_chan_feed_device.argtypes = [vp, vp, vp]
    L.fmgpu_chan_wait_external_stream.argtypes = [vp, vp]
    L.fmgpu_chan_sync.argtypes = [vp]
    L.fmgpu_chan_stream.argtypes = [vp]
    L.fmgpu_chan_stream.restype = vp
    L.fmgpu_chan_launch_count.argtypes = [vp]
    L.fmgpu_chan_launch_count.restype = C.c_longlong
    L.fmgpu_last_error.restype = C.c_char_p
    L.fmgpu_version.restype = C.c_char_p
    _lib = L
    return L


def _check(rc: int, what: str):
    if rc != 0:
        raise FMGPUError(f"{what} failed ({rc}): {lib().fmgpu_last_error().decode()}")


def _ptr(x) -> int:
    """Address of a numpy array, torch tensor (host or device) or raw int."""
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous()
        return x.data_ptr()
    raise TypeError(type(x))


class FMDemod:
    """A batch of `n_streams` independent broadcast-FM demodulators on one GPU.

    Mirrors Broadcast_FM_Demod(block_size) / Process / the Get* getters of the reference
    (broadcast_fm_demod.h:228-299); stream index selects the demodulator."""

    def __init__(self, block_size: int = 65536, n_streams: int = 1, device: int = -1,
                 keep_intermediates: bool = False, pipeline_depth: int = 0):
        self.L = lib()
        cfg = _Config(block_size, n_streams, device, int(keep_intermediates), pipeline_depth)
        h = C.c_void_p()
        _check(self.L.fmgpu_create(C.byref(cfg), C.byref(h)), "fmgpu_create")
        self.h = h
        out = _Config()
        self.L.fmgpu_get_config(self.h, C.byref(out))
        self.block_size, self.n_streams = out.block_size, out.n_streams
        self.depth, self.device = out.pipeline_depth, out.device
        self.keep_intermediates = bool(out.keep_intermediates)
        self.blocks_enqueued = 0
        self.pcm_rate = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.fmgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- synchronous (host buffers) ----
    def process_u8(self, iq) -> None:
        """iq: uint8 [n_streams, 2*block_size] (I,Q interleaved), numpy or pinned torch CPU tensor."""
        n = iq.size if isinstance(iq, np.ndarray) else iq.numel()
        if n != 2 * self.block_size * self.n_streams:
            return  # the reference silently ignores wrong-sized blocks (broadcast_fm_demod.cpp:311-313)
        _check(self.L.fmgpu_process_u8(self.h, _ptr(iq), self.block_size), "fmgpu_process_u8")
        self.blocks_enqueued += 1

    def process_cf32(self, iq) -> None:
        """iq: complex64 [n_streams, block_size]."""
        n = iq.size if isinstance(iq, np.ndarray) else iq.numel()
        if n != self.block_size * self.n_streams:
            return
        _check(self.L.fmgpu_process_cf32(self.h, _ptr(iq), self.block_size), "fmgpu_process_cf32")
        self.blocks_enqueued += 1

    # ---- asynchronous ----
    def enqueue_u8_device(self, iq_dev) -> int:
        """Queues one block whose input lives in device memory; returns the ring slot it uses."""
        slot = self.blocks_enqueued % self.depth
        _check(self.L.fmgpu_enqueue_u8_device(self.h, _ptr(iq_dev)), "fmgpu_enqueue_u8_device")
        self.blocks_enqueued += 1
        return slot

    def enqueue_u8_host(self, iq_host) -> int:
        slot = self.blocks_enqueued % self.depth
        _check(self.L.fmgpu_enqueue_u8_host(self.h, _ptr(iq_host)), "fmgpu_enqueue_u8_host")
        self.blocks_enqueued += 1
        return slot

    def enqueue_cf32_device(self, iq_dev, after_stream=None) -> int:
        """Queues one block of complex64 [n_streams, block_size] device input (the channelizer's output).
        after_stream: handle of the CUDA stream whose queued work produces iq_dev; None = nothing to wait for.
        Handle 0 (the legacy default stream) is passed to the C-ABI as cudaStreamLegacy, because there NULL means "none"."""
        slot = self.blocks_enqueued % self.depth
        st = None if after_stream is None else (int(after_stream) or 1)
        _check(self.L.fmgpu_enqueue_cf32_device(self.h, _ptr(iq_dev), st), "fmgpu_enqueue_cf32_device")
        self.blocks_enqueued += 1
        return slot

    def fetch_outputs(self, slot: int) -> None:
        _check(self.L.fmgpu_fetch_outputs(self.h, slot), "fmgpu_fetch_outputs")

    def sync(self) -> None:
        _check(self.L.fmgpu_sync(self.h), "fmgpu_sync")

    def wait_external_stream(self, cuda_stream: int) -> None:
        _check(self.L.fmgpu_wait_external_stream(self.h, cuda_stream), "fmgpu_wait_external_stream")

    def signal_external_stream(self, cuda_stream: int) -> None:
        _check(self.L.fmgpu_signal_external_stream(self.h, cuda_stream), "fmgpu_signal_external_stream")

    # ---- getters ----
    def get(self, buf: Buf, stream: int = 0) -> np.ndarray:
        p = C.c_void_p()
        n = C.c_size_t()
        _check(self.L.fmgpu_get_buffer(self.h, stream, int(buf), C.byref(p), C.byref(n)), f"fmgpu_get_buffer({buf.name})")
        dt, mult = _BUF_DTYPE[Buf(buf)]
        dt = np.dtype(dt)
        count = n.value * mult
        if count == 0:
            return np.zeros(0, dt)
        raw = (C.c_char * (count * dt.itemsize)).from_address(p.value)
        return np.frombuffer(raw, dtype=dt, count=count).copy()

    def device_buffer(self, buf: Buf, slot: int):
        p = C.c_void_p()
        n = C.c_size_t()
        _check(self.L.fmgpu_get_device_buffer(self.h, slot, int(buf), C.byref(p), C.byref(n)), "fmgpu_get_device_buffer")
        return p.value, n.value

    def fft(self, buf: Buf, stream: int = 0, fftshift: bool = True) -> np.ndarray:
        """CalculateFFT (+ InplaceFFTShift) of one stream's signal buffer of the last block, computed on the device
        (UpdateFFTCalc, fm_demod/broadcast_fm_demod.cpp:27-40, without the dB step)."""
        y = np.zeros(self.block_size, np.complex64)
        n = C.c_size_t()
        _check(self.L.fmgpu_get_fft(self.h, stream, int(buf), int(fftshift), y.ctypes.data, C.byref(n)), f"fmgpu_get_fft({buf.name})")
        return y[:n.value].copy()

    def scalar(self, which: Scalar, stream: int = 0) -> float:
        v = C.c_float()
        _check(self.L.fmgpu_get_scalar(self.h, stream, int(which), C.byref(v)), "fmgpu_get_scalar")
        return float(v.value)

    def set_control(self, which: Control, value: float) -> None:
        _check(self.L.fmgpu_set_control(self.h, int(which), float(value)), "fmgpu_set_control")
        if which == Control.AUDIO_PCM_RATE_HZ:
            self.pcm_rate = int(value)

    def upload_taps(self, which: Filter, b, a=None) -> None:
        b = np.ascontiguousarray(b, np.float32)
        a = None if a is None else np.ascontiguousarray(a, np.float32)
        _check(self.L.fmgpu_upload_taps(self.h, int(which), b.ctypes.data, None if a is None else a.ctypes.data, b.size),
               "fmgpu_upload_taps")

    def download_taps(self, which: Filter):
        n = FILTER_LEN[Filter(which)]
        b = np.zeros(n, np.float32)
        a = np.zeros(n, np.float32)
        _check(self.L.fmgpu_download_taps(self.h, int(which), b.ctypes.data, a.ctypes.data, n), "fmgpu_download_taps")
        return (b, a) if Filter(which) in FILTER_IS_IIR else (b, None)

    def rates(self) -> dict:
        r = (C.c_int * 5)()
        _check(self.L.fmgpu_get_rates(self.h, C.byref(r)), "fmgpu_get_rates")
        return dict(zip(("baseband", "fm_in", "fm_out", "rds", "audio"), list(r)))

    def profile_stages(self, iq_dev, n_blocks: int = 4) -> dict:
        """Average device ms per kernel with blocks run one at a time (CUDA events inside the library).
        k7_audio_pcm is present when the audio output stage is on (Control.AUDIO_PCM_RATE_HZ)."""
        ms = (C.c_float * 7)()
        _check(self.L.fmgpu_profile_stages7(self.h, _ptr(iq_dev), n_blocks, C.byref(ms)), "fmgpu_profile_stages7")
        self.blocks_enqueued += abs(n_blocks)
        out = dict(zip(("k1_fir4_discrim", "k2_mpx", "k3_pll", "k4_mix_fir", "k5_bpsk", "k6_rds", "k7_audio_pcm"), [float(x) for x in ms]))
        if not self.pcm_rate:
            del out["k7_audio_pcm"]
        return out

    FETCH_AUDIO_F32, FETCH_PCM_S16, FETCH_RDS_SYMBOLS, FETCH_ALL = 1, 2, 4, 7

    def set_fetch_mask(self, mask: int) -> None:
        """Which outputs fetch_outputs copies to the host (include/fmgpu.h FMGPU_FETCH_*); counts always travel."""
        _check(self.L.fmgpu_set_fetch_mask(self.h, int(mask)), "fmgpu_set_fetch_mask")

    def set_option(self, name: str, value: int) -> None:
        """Implementation switches for A/B measurements: "k1_fp32", "k5_literal" (include/fmgpu.h)."""
        _check(self.L.fmgpu_set_option(self.h, name.encode(), int(value)), "fmgpu_set_option")

    def partition(self):
        """(SMs reserved for the recurrence stages, SMs of the FIR stages); (0, 0) = not partitioned."""
        sms = (C.c_int * 2)()
        _check(self.L.fmgpu_get_partition(self.h, C.byref(sms)), "fmgpu_get_partition")
        return int(sms[0]), int(sms[1])

    @property
    def launch_count(self) -> int:
        return int(self.L.fmgpu_launch_count(self.h))

    # ---- RDS decoded on the device (kernel K6) ----
    def rds_fetch(self) -> None:
        """Synchronises and copies every stream's decoder state + result rings to the host."""
        _check(self.L.fmgpu_rds_device_fetch(self.h), "fmgpu_rds_device_fetch")

    def rds_counts(self, stream: int = 0):
        """(groups, packet bytes) decoded since creation, as of the last rds_fetch()."""
        g, b = C.c_ulonglong(0), C.c_ulonglong(0)
        _check(self.L.fmgpu_rds_device_counts(self.h, stream, C.byref(g), C.byref(b), None), "fmgpu_rds_device_counts")
        return int(g.value), int(b.value)

    def rds_ring_caps(self):
        caps = (C.c_int * 2)()
        _check(self.L.fmgpu_rds_device_counts(self.h, 0, None, None, C.byref(caps)), "fmgpu_rds_device_counts")
        return int(caps[0]), int(caps[1])

    def rds_groups(self, stream: int = 0, first: int = 0):
        """Groups [first, total) of `stream` as (data[n,4] u16, valid[n,4] u8, type[n,4] u8)."""
        total, _ = self.rds_counts(stream)
        n = max(total - first, 0)
        arr = (RDSGroup * max(n, 1))()
        got = self.L.fmgpu_rds_device_get_groups(self.h, stream, first, arr, n)
        if got < 0:
            _check(got, "fmgpu_rds_device_get_groups")
        raw = np.frombuffer(arr, dtype=np.uint8, count=16 * got).reshape(got, 16) if got else np.zeros((0, 16), np.uint8)
        data = raw[:, :8].copy().view(np.uint16).reshape(got, 4)
        return data, raw[:, 8:12].copy(), raw[:, 12:16].copy()

    def rds_bytes(self, stream: int = 0, first: int = 0) -> bytes:
        _, total = self.rds_counts(stream)
        n = max(total - first, 0)
        out = np.zeros(max(n, 1), np.uint8)
        got = self.L.fmgpu_rds_device_get_bytes(self.h, stream, first, out.ctypes.data, n)
        if got < 0:
            _check(got, "fmgpu_rds_device_get_bytes")
        return out[:got].tobytes()

    def rds_db(self, stream: int = 0) -> dict:
        pi, pty = C.c_uint16(0), C.c_uint8(0)
        ps, rt = C.create_string_buffer(8), C.create_string_buffer(64)
        _check(self.L.fmgpu_rds_device_get_db(self.h, stream, C.byref(pi), ps, rt, C.byref(pty)), "fmgpu_rds_device_get_db")
        return {"pi": pi.value, "pty": pty.value, "ps": ps.raw, "rt": rt.raw}

    def rds_db_ext(self, stream: int = 0) -> dict:
        """The rest of RDS_Database (rds_database.h:26-53): TA/TP, M/S, DI flags, clock, programme type name."""
        e = RdsDbExt()
        _check(self.L.fmgpu_rds_device_get_db_ext(self.h, stream, C.byref(e)), "fmgpu_rds_device_get_db_ext")
        return e.as_dict()


class Channelizer:
    """Wideband channelizer (include/fmgpu.h, fmgpu_chan_*): one u8 IQ capture at `fs_in_hz` ->
    len(centres_hz) complex channels at fs_in_hz / decimation, laid out as the cf32 input of an
    FMDemod(block_out, n_channels).  New component (the reference has none)."""

    def __init__(self, fs_in_hz: float, centres_hz, decimation: int = 20, n_taps: int = 192, block_out: int = 65536,
                 device: int = -1, mode: ChanMode = ChanMode.AUTO, ring_depth: int = 0, cutoff_k: float = 0.0):
        self.L = lib()
        centres = np.ascontiguousarray(centres_hz, np.float64)
        cfg = _ChanConfig(float(fs_in_hz), decimation, n_taps, float(cutoff_k), centres.size, block_out, device, int(mode), ring_depth)
        h = C.c_void_p()
        _check(self.L.fmgpu_chan_create(C.byref(cfg), centres.ctypes.data_as(C.POINTER(C.c_double)), C.byref(h)), "fmgpu_chan_create")
        self.h = h
        out = _ChanConfig()
        self.L.fmgpu_chan_get_config(self.h, C.byref(out))
        self.fs_in_hz, self.decimation, self.n_taps = out.fs_in_hz, out.decimation, out.n_taps
        self.n_channels, self.block_out, self.mode, self.depth = out.n_channels, out.block_out, ChanMode(out.mode), out.ring_depth
        self.block_in = self.block_out * self.decimation
        self.blocks = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.fmgpu_chan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def get_b(self) -> np.ndarray:
        """The prototype taps (reference order); writes take effect at the next block."""
        return np.ctypeslib.as_array(self.L.fmgpu_chan_get_b(self.h), shape=(self.n_taps,))

    def freqs(self):
        """(quantised centre frequencies in Hz, phase increments per input sample in turns * 2^32)."""
        hz = np.zeros(self.n_channels, np.float64)
        inc = np.zeros(self.n_channels, np.uint32)
        _check(self.L.fmgpu_chan_get_freqs(self.h, hz.ctypes.data, inc.ctypes.data), "fmgpu_chan_get_freqs")
        return hz, inc

    def process_u8(self, iq) -> np.ndarray:
        """iq: uint8 [2 * block_in] host array -> complex64 [n_channels, block_out]."""
        iq = np.ascontiguousarray(iq, np.uint8)
        out = np.zeros((self.n_channels, self.block_out), np.complex64)
        _check(self.L.fmgpu_chan_process_u8(self.h, iq.ctypes.data, iq.size // 2, out.ctypes.data), "fmgpu_chan_process_u8")
        self.blocks += 1
        return out

    def enqueue_u8_device(self, iq_dev) -> int:
        """Asynchronous; returns the device address of the [n_channels, block_out] complex64 output."""
        p = C.c_void_p()
        _check(self.L.fmgpu_chan_enqueue_u8_device(self.h, _ptr(iq_dev), C.byref(p)), "fmgpu_chan_enqueue_u8_device")
        self.blocks += 1
        return p.value

    def feed(self, demod: "FMDemod", iq_dev) -> int:
        """One wideband block through the channelizer and `demod` (asynchronous); returns demod's ring slot."""
        slot = demod.blocks_enqueued % demod.depth
        _check(self.L.fmgpu_chan_feed_device(self.h, demod.h, _ptr(iq_dev)), "fmgpu_chan_feed_device")
        self.blocks += 1
        demod.blocks_enqueued += 1
        return slot

    def wait_external_stream(self, cuda_stream: int) -> None:
        _check(self.L.fmgpu_chan_wait_external_stream(self.h, cuda_stream), "fmgpu_chan_wait_external_stream")

    def sync(self) -> None:
        _check(self.L.fmgpu_chan_sync(self.h), "fmgpu_chan_sync")

    @property
    def stream(self) -> int:
        return int(self.L.fmgpu_chan_stream(self.h) or 0)

    @property
    def launch_count(self) -> int:
        return int(self.L.fmgpu_chan_launch_count(self.h))


class RDSDecoder:
    """Host RDS bit path: soft symbols -> groups -> PI / PS / RadioText."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.fmgpu_rds_create())

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fmgpu_rds_destroy(self.h)
            self.h = None

    def push_symbols(self, sym) -> None:
        sym = np.ascontiguousarray(sym, np.float32)
        self.L.fmgpu_rds_push_symbols(self.h, sym.ctypes.data, sym.size)

    def groups(self):
        n = self.L.fmgpu_rds_n_groups(self.h)
        arr = (RDSGroup * max(n, 1))()
        n = self.L.fmgpu_rds_get_groups(self.h, arr, n)
        data = np.array([[g.data[i] for i in range(4)] for g in arr[:n]], np.uint16).reshape(n, 4)
        valid = np.array([[g.valid[i] for i in range(4)] for g in arr[:n]], np.uint8).reshape(n, 4)
        typ = np.array([[g.type[i] for i in range(4)] for g in arr[:n]], np.uint8).reshape(n, 4)
        return data, valid, typ

    def rds_bytes(self) -> bytes:
        n = self.L.fmgpu_rds_n_bytes(self.h)
        out = np.zeros(max(n, 1), np.uint8)
        n = self.L.fmgpu_rds_get_bytes(self.h, out.ctypes.data, n)
        return out[:n].tobytes()

    def db(self) -> dict:
        pi, pty = C.c_uint16(0), C.c_uint8(0)
        ps, rt = C.create_string_buffer(8), C.create_string_buffer(64)
        self.L.fmgpu_rds_get_db(self.h, C.byref(pi), ps, rt, C.byref(pty))
        return {"pi": pi.value, "pty": pty.value, "ps": ps.raw, "rt": rt.raw}

    def db_ext(self) -> dict:
        e = RdsDbExt()
        self.L.fmgpu_rds_get_db_ext(self.h, C.byref(e))
        return e.as_dict()


class _DspFilter:
    """Common part of FIR_Filter / Hilbert_FIR_Filter / IIR_Filter (include/fmgpu.h fmgpu_dsp_filter_*)."""
    KIND = 0

    def __init__(self, K: int, is_complex: bool = False):
        self.L = lib()
        self.K, self.is_complex = K, bool(is_complex)
        h = C.c_void_p()
        _check(self.L.fmgpu_dsp_filter_create(self.KIND, K, int(self.is_complex), C.byref(h)), "fmgpu_dsp_filter_create")
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fmgpu_dsp_filter_destroy(self.h)
            self.h = None

    def get_b(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.fmgpu_dsp_filter_get_b(self.h), shape=(self.K,))

    def get_K(self) -> int:
        return self.L.fmgpu_dsp_filter_get_K(self.h)

    def _process(self, x, out_complex: bool) -> np.ndarray:
        x = np.ascontiguousarray(x, np.complex64 if self.is_complex else np.float32)
        y = np.zeros(x.size, np.complex64 if out_complex else np.float32)
        _check(self.L.fmgpu_dsp_filter_process(self.h, x.ctypes.data, y.ctypes.data, x.size), "fmgpu_dsp_filter_process")
        return y


class FIRFilter(_DspFilter):
    """dsp/fir_filter.h:9-88 FIR_Filter<T>(K): get_b(), get_K(), process(x) -> y."""
    KIND = 0

    def process(self, x) -> np.ndarray:
        return self._process(x, self.is_complex)


class HilbertFIRFilter(_DspFilter):
    """dsp/hilbert_fir_filter.h:13-47 Hilbert_FIR_Filter<float>(K): process(x real) -> complex { delayed x, Hilbert(x) }."""
    KIND = 1

    def __init__(self, K: int):
        super().__init__(K, False)

    def process(self, x) -> np.ndarray:
        return self._process(x, True)


class IIRFilter(_DspFilter):
    """dsp/iir_filter.h:5-89 IIR_Filter<T>(K): get_b(), get_a(), process(x) -> y (direct form I, the reference's operation order)."""
    KIND = 2

    def get_a(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.L.fmgpu_dsp_filter_get_a(self.h), shape=(self.K,))

    def process(self, x) -> np.ndarray:
        return self._process(x, self.is_complex)


class _Agc(C.Structure):
    _fields_ = [("target_power", C.c_float), ("current_gain", C.c_float), ("beta", C.c_float)]


class AGCFilter:
    """dsp/agc.h:6-31 AGC_Filter<std::complex<float>>: public fields target_power / current_gain / beta, process(x) -> y."""

    def __init__(self):
        self.L = lib()
        self._g = _Agc()
        self.L.fmgpu_agc_init(C.byref(self._g))

    target_power = property(lambda self: self._g.target_power, lambda self, v: setattr(self._g, "target_power", v))
    current_gain = property(lambda self: self._g.current_gain, lambda self, v: setattr(self._g, "current_gain", v))
    beta = property(lambda self: self._g.beta, lambda self, v: setattr(self._g, "beta", v))

    def process(self, x) -> np.ndarray:
        x = np.ascontiguousarray(x, np.complex64)
        y = np.zeros(x.size, np.complex64)
        _check(self.L.fmgpu_agc_process(C.byref(self._g), x.ctypes.data, y.ctypes.data, x.size), "fmgpu_agc_process")
        return y


class PolyphaseDownsampler:
    """dsp/polyphase_filter.h:9-87 on the GPU: PolyphaseDownsampler<T>(M, K), get_b(), process()."""

    def __init__(self, M: int, K: int, is_complex: bool):
        self.L = lib()
        self.M, self.K, self.NN, self.is_complex = M, K, M * K, bool(is_complex)
        h = C.c_void_p()
        _check(self.L.fmgpu_polyphase_ds_create(M, K, int(is_complex), C.byref(h)), "fmgpu_polyphase_ds_create")
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fmgpu_polyphase_destroy(self.h)
            self.h = None

    def get_b(self) -> np.ndarray:
        p = self.L.fmgpu_polyphase_get_b(self.h)
        return np.ctypeslib.as_array(p, shape=(self.NN,))

    def get_K(self) -> int:
        return self.NN

    def process(self, x: np.ndarray, n_out: int) -> np.ndarray:
        dt = np.complex64 if self.is_complex else np.float32
        x = np.ascontiguousarray(x, dt)
        assert x.size == n_out * self.M
        y = np.zeros(n_out, dt)
        _check(self.L.fmgpu_polyphase_ds_process(self.h, x.ctypes.data, y.ctypes.data, n_out), "fmgpu_polyphase_ds_process")
        return y


class PolyphaseUpsampler:
    """dsp/polyphase_filter.h:90-185 on the GPU: PolyphaseUpsampler<T>(b, L, K), process(x, y, N)."""

    def __init__(self, b: np.ndarray, L: int, K: int, is_complex: bool):
        self.lib = lib()
        self.Lf, self.K, self.is_complex = L, K, bool(is_complex)
        b = np.ascontiguousarray(b, np.float32)
        assert b.size == L * K
        h = C.c_void_p()
        _check(self.lib.fmgpu_polyphase_us_create(b.ctypes.data, L, K, int(is_complex), C.byref(h)), "fmgpu_polyphase_us_create")
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.fmgpu_polyphase_destroy(self.h)
            self.h = None

    def process(self, x: np.ndarray) -> np.ndarray:
        dt = np.complex64 if self.is_complex else np.float32
        x = np.ascontiguousarray(x, dt)
        y = np.zeros(x.size * self.Lf, dt)
        _check(self.lib.fmgpu_polyphase_us_process(self.h, x.ctypes.data, y.ctypes.data, x.size), "fmgpu_polyphase_us_process")
        return y


def resample_linear(frames: np.ndarray, n_out: int) -> np.ndarray:
    """Resample() of audio/resampled_pcm_player.cpp:37-54 on the GPU: frames [n_in, 2] f32 -> [n_out, 2]."""
    frames = np.ascontiguousarray(frames, np.float32).reshape(-1, 2)
    out = np.zeros((n_out, 2), np.float32)
    _check(lib().fmgpu_resample_linear(frames.ctypes.data, frames.shape[0], out.ctypes.data, n_out), "fmgpu_resample_linear")
    return out


def calculate_fft(x: np.ndarray, fftshift: bool = False) -> np.ndarray:
    """CalculateFFT (dsp/calculate_fft.cpp:43-50) [+ InplaceFFTShift, dsp/fftshift.h:21-33] on the GPU."""
    x = np.ascontiguousarray(x, np.complex64)
    y = np.zeros(x.size, np.complex64)
    _check(lib().fmgpu_calculate_fft(x.ctypes.data, y.ctypes.data, x.size, int(fftshift)), "fmgpu_calculate_fft")
    return y


def frames_to_s16(frames: np.ndarray) -> np.ndarray:
    """Audio_Scraper::on_audio_data's float -> int16 conversion (fm_scraper.cpp:74-78) on the GPU."""
    frames = np.ascontiguousarray(frames, np.float32).reshape(-1, 2)
    out = np.zeros((frames.shape[0], 2), np.int16)
    _check(lib().fmgpu_frames_to_s16(frames.ctypes.data, frames.shape[0], out.ctypes.data), "fmgpu_frames_to_s16")
    return out


def _designer(name, n_b, n_a, *args):
    b = np.zeros(n_b, np.float32)
    if n_a:
        a = np.zeros(n_a, np.float32)
        getattr(lib(), name)(b.ctypes.data, a.ctypes.data, *args)
        return b, a
    getattr(lib(), name)(b.ctypes.data, *args)
    return b


def create_fir_lpf(N: int, k: float): return _designer("fmgpu_create_fir_lpf", N, 0, N, k)
def create_fir_hpf(N: int, k: float): return _designer("fmgpu_create_fir_hpf", N, 0, N, k)
def create_fir_bpf(N: int, k1: float, k2: float): return _designer("fmgpu_create_fir_bpf", N, 0, N, k1, k2)
def create_fir_hilbert(N: int): return _designer("fmgpu_create_fir_hilbert", N, 0, N)
def create_iir_single_pole_lpf(k: float): return _designer("fmgpu_create_iir_single_pole_lpf", 2, 2, k)
def create_iir_notch_filter(k: float, r: float): return _designer("fmgpu_create_iir_notch_filter", 3, 3, k, r)
def create_iir_peak_1_filter(k: float, r: float): return _designer("fmgpu_create_iir_peak_1_filter", 3, 3, k, r)
def create_iir_peak_2_filter(k: float, r: float, A_db: float): return _designer("fmgpu_create_iir_peak_2_filter", 3, 3, k, r, A_db)


TOTAL_TAPS_IIR_SINGLE_POLE_LPF, TOTAL_TAPS_IIR_SECOND_ORDER_NOTCH_FILTER, TOTAL_TAPS_IIR_SECOND_ORDER_PEAK_FILTER = 2, 3, 3
WINDOWS = ("hamming", "hann", "blackman", "blackman_harris")


def _windowed(name: str, window: str, N: int, *ks):
    """create_fir_{lpf,hpf,bpf}(b, N, k..., window) of dsp/filter_designer.h:9-11 with one of dsp/window_functions.h."""
    L = lib()
    fn = getattr(L, f"fmgpu_create_fir_{name}_window")
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_int] + [C.c_float] * len(ks) + [C.c_void_p]
    w = C.cast(getattr(L, f"fmgpu_window_{window}"), C.c_void_p)
    b = np.zeros(N, np.float32)
    fn(b.ctypes.data, N, *[float(k) for k in ks], w)
    return b


def create_fir_lpf_window(N: int, k: float, window: str = "hamming"): return _windowed("lpf", window, N, k)
def create_fir_hpf_window(N: int, k: float, window: str = "hamming"): return _windowed("hpf", window, N, k)
def create_fir_bpf_window(N: int, k1: float, k2: float, window: str = "hamming"): return _windowed("bpf", window, N, k1, k2)
