"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libfmref.so (the unmodified reference,
see oracle/ref_harness.cpp).  Import only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libfmref.so")
REF_BENCH = os.path.join(HERE, "_ref", "fm_demod_benchmark")

_DTYPES = {
    "fm_in": np.complex64, "fm_demod": np.float32, "fm_out": np.float32, "fm_out_iq": np.complex64,
    "pilot": np.complex64, "pll_dt": np.float32, "pll": np.complex64,
    "pll_raw_phase_error": np.float32, "pll_lpf_phase_error": np.float32,
    "audio_lpr": np.float32, "audio_lmr": np.float32, "rds": np.complex64,
    "rds_raw_sym": np.complex64, "rds_pred_sym": np.float32, "audio_out": np.float32,
    "bpsk_pll_sym": np.complex64, "bpsk_ted_raw_phase_error": np.float32,
    "bpsk_ted_pi_phase_error": np.float32, "bpsk_pll_raw_phase_error": np.float32,
    "bpsk_pll_pi_phase_error": np.float32, "bpsk_int_dump_filter": np.complex64,
    "bpsk_zcd": np.bool_, "bpsk_int_dump_trigger": np.bool_,
}
_MULT = {"audio_out": 2}


def available() -> bool:
    return os.path.exists(REF_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        L.fmref_create.restype = C.c_void_p
        L.fmref_create.argtypes = [C.c_int]
        L.fmref_destroy.argtypes = [C.c_void_p]
        L.fmref_process_u8.argtypes = [C.c_void_p, C.c_void_p]
        L.fmref_process_cf32.argtypes = [C.c_void_p, C.c_void_p]
        L.fmref_set_control.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.fmref_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.fmref_get_scalar.restype = C.c_float
        L.fmref_get_scalar.argtypes = [C.c_void_p, C.c_char_p]
        L.fmref_get_taps.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
        L.fmref_n_groups.argtypes = [C.c_void_p]
        L.fmref_get_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fmref_n_rds_bytes.argtypes = [C.c_void_p]
        L.fmref_get_rds_bytes.argtypes = [C.c_void_p, C.c_void_p]
        L.fmref_get_db.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fmref_rds_create.restype = C.c_void_p
        L.fmref_rds_destroy.argtypes = [C.c_void_p]
        L.fmref_rds_push_symbols.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.fmref_rds_n_groups.argtypes = [C.c_void_p]
        L.fmref_rds_get_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fmref_rds_n_bytes.argtypes = [C.c_void_p]
        L.fmref_rds_get_bytes.argtypes = [C.c_void_p, C.c_void_p]
        L.fmref_rds_get_db.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.fmref_create_fir_lpf.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.fmref_create_fir_hpf.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.fmref_create_fir_bpf.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.fmref_create_fir_hilbert.argtypes = [C.c_void_p, C.c_int]
        L.fmref_create_iir_single_pole_lpf.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        L.fmref_create_iir_notch_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.fmref_create_iir_peak_1_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.fmref_polyphase_ds_f32.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.fmref_polyphase_ds_cf32.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.fmref_polyphase_us_f32.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _groups(n, getter, h):
    data = np.zeros((n, 4), np.uint16)
    valid = np.zeros((n, 4), np.uint8)
    typ = np.zeros((n, 4), np.uint8)
    if n:
        getter(h, data.ctypes.data, valid.ctypes.data, typ.ctypes.data)
    return data, valid, typ


def _db(getter, h):
    pi = C.c_uint16(0)
    pty = C.c_uint8(0)
    ps = C.create_string_buffer(8)
    rt = C.create_string_buffer(64)
    getter(h, C.byref(pi), ps, rt, C.byref(pty))
    return {"pi": pi.value, "pty": pty.value, "ps": ps.raw, "rt": rt.raw}


class RefDemod:
    """The reference's App wiring (src/app.cpp) around an unmodified Broadcast_FM_Demod."""

    def __init__(self, block_size: int):
        self.L = lib()
        self.block_size = block_size
        self.h = self.L.fmref_create(block_size)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fmref_destroy(self.h)
            self.h = None

    def process_u8(self, iq_u8: np.ndarray):
        iq_u8 = np.ascontiguousarray(iq_u8, dtype=np.uint8)
        assert iq_u8.size == 2 * self.block_size
        self.L.fmref_process_u8(self.h, iq_u8.ctypes.data)

    def process_cf32(self, iq: np.ndarray):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        assert iq.size == self.block_size
        self.L.fmref_process_cf32(self.h, iq.ctypes.data)

    def set_control(self, name: str, value: float):
        assert self.L.fmref_set_control(self.h, name.encode(), float(value)) == 0

    def get(self, name: str) -> np.ndarray:
        p = C.c_void_p()
        n = C.c_size_t()
        assert self.L.fmref_get(self.h, name.encode(), C.byref(p), C.byref(n)) == 0, name
        dt = np.dtype(_DTYPES[name])
        count = n.value * _MULT.get(name, 1)
        if count == 0:
            return np.zeros(0, dt)
        buf = (C.c_char * (count * dt.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=count).copy()

    def scalar(self, name: str) -> float:
        return float(self.L.fmref_get_scalar(self.h, name.encode()))

    def taps(self, name: str):
        b = np.zeros(256, np.float32)
        a = np.zeros(8, np.float32)
        n = self.L.fmref_get_taps(self.h, name.encode(), b.ctypes.data, a.ctypes.data)
        assert n > 0, name
        return b[:n].copy(), a[:min(n, 8)].copy()

    def groups(self):
        return _groups(self.L.fmref_n_groups(self.h), self.L.fmref_get_groups, self.h)

    def rds_bytes(self) -> bytes:
        n = self.L.fmref_n_rds_bytes(self.h)
        out = np.zeros(n, np.uint8)
        if n:
            self.L.fmref_get_rds_bytes(self.h, out.ctypes.data)
        return out.tobytes()

    def db(self):
        return _db(self.L.fmref_get_db, self.h)


class RefRds:
    """The reference's RDS bit path alone (DifferentialManchesterDecoder -> RDS_Decoding_Chain)."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.fmref_rds_create()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.fmref_rds_destroy(self.h)
            self.h = None

    def push_symbols(self, sym: np.ndarray):
        sym = np.ascontiguousarray(sym, dtype=np.float32)
        self.L.fmref_rds_push_symbols(self.h, sym.ctypes.data, sym.size)

    def groups(self):
        return _groups(self.L.fmref_rds_n_groups(self.h), self.L.fmref_rds_get_groups, self.h)

    def rds_bytes(self) -> bytes:
        n = self.L.fmref_rds_n_bytes(self.h)
        out = np.zeros(n, np.uint8)
        if n:
            self.L.fmref_rds_get_bytes(self.h, out.ctypes.data)
        return out.tobytes()

    def db(self):
        return _db(self.L.fmref_rds_get_db, self.h)
