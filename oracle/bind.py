"""TEST INFRASTRUCTURE ONLY: ctypes binding shared by the two CPU checkers

* kind="ref"  -> oracle/_ref/libfmref.so, the UNMODIFIED reference compiled in place
                 (oracle/ref_harness.cpp, prefix fmref_)
* kind="port" -> oracle/libfmoracle.so, the plain-C restatement (oracle/fm_oracle.c, prefix fmo_)

Import only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs -- never from the product package."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libfmref.so")
PORT_SO = os.path.join(HERE, "libfmoracle.so")
_PATH = {"ref": REF_SO, "port": PORT_SO}
_PREFIX = {"ref": "fmref_", "port": "fmo_"}
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


def available(kind: str = "ref") -> bool:
    return os.path.exists(_PATH[kind])


class _Lib:
    """Prefix-stripping view of the shared object: L.create == fmref_create / fmo_create."""

    def __init__(self, kind):
        self._cdll = C.CDLL(_PATH[kind])
        self._prefix = _PREFIX[kind]

    def __getattr__(self, name):
        return getattr(self._cdll, self._prefix + name)


_libs = {}


def lib(kind: str = "ref"):
    if kind not in _libs:
        L = _Lib(kind)
        L.create.restype = C.c_void_p
        L.create.argtypes = [C.c_int]
        L.destroy.argtypes = [C.c_void_p]
        L.process_u8.argtypes = [C.c_void_p, C.c_void_p]
        L.process_cf32.argtypes = [C.c_void_p, C.c_void_p]
        L.set_control.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.get_scalar.restype = C.c_float
        L.get_scalar.argtypes = [C.c_void_p, C.c_char_p]
        L.get_taps.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
        L.n_groups.argtypes = [C.c_void_p]
        L.get_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.n_rds_bytes.argtypes = [C.c_void_p]
        L.get_rds_bytes.argtypes = [C.c_void_p, C.c_void_p]
        L.get_db.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.get_db_ext.argtypes = [C.c_void_p, C.c_void_p]
        L.rds_get_db_ext.argtypes = [C.c_void_p, C.c_void_p]
        L.rds_create.restype = C.c_void_p
        L.rds_destroy.argtypes = [C.c_void_p]
        L.rds_push_symbols.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.rds_n_groups.argtypes = [C.c_void_p]
        L.rds_get_groups.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.rds_n_bytes.argtypes = [C.c_void_p]
        L.rds_get_bytes.argtypes = [C.c_void_p, C.c_void_p]
        L.rds_get_db.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.create_fir_lpf.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.create_fir_hpf.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.create_fir_bpf.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.create_fir_hilbert.argtypes = [C.c_void_p, C.c_int]
        L.create_iir_single_pole_lpf.argtypes = [C.c_void_p, C.c_void_p, C.c_float]
        L.create_iir_notch_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.create_iir_peak_1_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.create_iir_peak_2_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_float]
        L.create_fir_lpf_window.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int]
        L.polyphase_ds_f32.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.polyphase_ds_cf32.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.polyphase_us_f32.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.resample_linear.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        for name in ("fir_f32", "fir_cf32"):
            getattr(L, name).argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.hilbert_f32.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        for name in ("iir_f32", "iir_cf32"):
            getattr(L, name).argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.agc_cf32.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.frames_to_s16.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.fft_mag_process.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int]
        if kind == "port":
            L.fft_f64.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        else:
            L.fftshift_inplace.argtypes = [C.c_void_p, C.c_int]
        if kind == "port":
            L.set_taps.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
        _libs[kind] = L
    return _libs[kind]


def _groups(n, getter, h):
    data = np.zeros((n, 4), np.uint16)
    valid = np.zeros((n, 4), np.uint8)
    typ = np.zeros((n, 4), np.uint8)
    if n:
        getter(h, data.ctypes.data, valid.ctypes.data, typ.ctypes.data)
    return data, valid, typ


class DbExt(C.Structure):
    """fmo_db_ext (oracle/fm_oracle.h); the harness fills the same 24 bytes from the reference's RDS_Database."""
    _fields_ = [("programme_type_name", C.c_char * 8), ("year", C.c_int32), ("day", C.c_uint8), ("month", C.c_uint8),
                ("hour", C.c_uint8), ("minute", C.c_uint8), ("local_time_offset", C.c_int8),
                ("traffic_announcement", C.c_uint8), ("is_stereo", C.c_uint8), ("is_music", C.c_uint8),
                ("is_artificial_head", C.c_uint8), ("is_compressed", C.c_uint8), ("is_dynamic_program_type", C.c_uint8),
                ("ptyn_ab_flag", C.c_uint8)]


def _db_ext(getter, h) -> dict:
    e = DbExt()
    getter(h, C.byref(e))
    d = {name: getattr(e, name) for name, _ in DbExt._fields_ if name not in ("programme_type_name", "ptyn_ab_flag")}
    d["programme_type_name"] = bytes(bytearray(e)[:8])
    return d


def _db(getter, h):
    pi = C.c_uint16(0)
    pty = C.c_uint8(0)
    ps = C.create_string_buffer(8)
    rt = C.create_string_buffer(64)
    getter(h, C.byref(pi), ps, rt, C.byref(pty))
    return {"pi": pi.value, "pty": pty.value, "ps": ps.raw, "rt": rt.raw}


class CpuDemod:
    """App wiring (src/app.cpp) around Broadcast_FM_Demod: kind="ref" is the unmodified reference,
    kind="port" the C restatement."""

    def __init__(self, block_size: int, kind: str = "ref"):
        self.kind = kind
        self.L = lib(kind)
        self.block_size = block_size
        self.h = self.L.create(block_size)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.destroy(self.h)
            self.h = None

    def process_u8(self, iq_u8: np.ndarray):
        iq_u8 = np.ascontiguousarray(iq_u8, dtype=np.uint8)
        assert iq_u8.size == 2 * self.block_size
        self.L.process_u8(self.h, iq_u8.ctypes.data)

    def process_cf32(self, iq: np.ndarray):
        iq = np.ascontiguousarray(iq, dtype=np.complex64)
        assert iq.size == self.block_size
        self.L.process_cf32(self.h, iq.ctypes.data)

    def set_control(self, name: str, value: float):
        assert self.L.set_control(self.h, name.encode(), float(value)) == 0

    def get(self, name: str) -> np.ndarray:
        p = C.c_void_p()
        n = C.c_size_t()
        assert self.L.get(self.h, name.encode(), C.byref(p), C.byref(n)) == 0, name
        dt = np.dtype(_DTYPES[name])
        count = n.value * _MULT.get(name, 1)
        if count == 0:
            return np.zeros(0, dt)
        buf = (C.c_char * (count * dt.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt, count=count).copy()

    def scalar(self, name: str) -> float:
        return float(self.L.get_scalar(self.h, name.encode()))

    def taps(self, name: str):
        b = np.zeros(256, np.float32)
        a = np.zeros(8, np.float32)
        n = self.L.get_taps(self.h, name.encode(), b.ctypes.data, a.ctypes.data)
        assert n > 0, name
        return b[:n].copy(), a[:min(n, 8)].copy()

    def set_taps(self, name: str, b, a=None):
        assert self.kind == "port"
        b = np.ascontiguousarray(b, np.float32)
        a = None if a is None else np.ascontiguousarray(a, np.float32)
        n = self.L.set_taps(self.h, name.encode(), b.ctypes.data, None if a is None else a.ctypes.data)
        assert n == b.size, (name, n, b.size)

    def groups(self):
        return _groups(self.L.n_groups(self.h), self.L.get_groups, self.h)

    def rds_bytes(self) -> bytes:
        n = self.L.n_rds_bytes(self.h)
        out = np.zeros(n, np.uint8)
        if n:
            self.L.get_rds_bytes(self.h, out.ctypes.data)
        return out.tobytes()

    def db(self):
        return _db(self.L.get_db, self.h)

    def db_ext(self) -> dict:
        return _db_ext(self.L.get_db_ext, self.h)


class CpuRds:
    """The RDS bit path alone (DifferentialManchesterDecoder -> RDS_Decoding_Chain)."""

    def __init__(self, kind: str = "ref"):
        self.L = lib(kind)
        self.h = self.L.rds_create()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.rds_destroy(self.h)
            self.h = None

    def push_symbols(self, sym: np.ndarray):
        sym = np.ascontiguousarray(sym, dtype=np.float32)
        self.L.rds_push_symbols(self.h, sym.ctypes.data, sym.size)

    def groups(self):
        return _groups(self.L.rds_n_groups(self.h), self.L.rds_get_groups, self.h)

    def rds_bytes(self) -> bytes:
        n = self.L.rds_n_bytes(self.h)
        out = np.zeros(n, np.uint8)
        if n:
            self.L.rds_get_bytes(self.h, out.ctypes.data)
        return out.tobytes()

    def db(self):
        return _db(self.L.rds_get_db, self.h)

    def db_ext(self) -> dict:
        return _db_ext(self.L.rds_get_db_ext, self.h)
