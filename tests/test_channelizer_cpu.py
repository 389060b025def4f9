"""CPU suite for the wideband front end (BASELINE config 4): the float64 channelizer checker against an
independent numpy evaluation of its definition, the wideband generator, the rank partition, and the
world_size-2 broadcast path (gloo) with stand-ins for the CUDA objects."""
import ctypes as C
import json
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from fm_radio_b200 import synth
from oracle import bind
from tests import helpers as H


def _chan_oracle(iq, n0, D, NN, b, inc, hist=None):
    L = bind.lib("port")._cdll
    L.fmo_channelize_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]
    n_in = iq.size // 2
    inc = np.ascontiguousarray(inc, np.uint32)
    out = np.zeros((len(inc), n_in // D, 2), np.float64)
    L.fmo_channelize_f64(iq.ctypes.data, n_in, None if hist is None else hist.ctypes.data, n0, D, NN,
                         np.ascontiguousarray(b, np.float32).ctypes.data, inc.ctypes.data, len(inc), out.ctypes.data)
    return out[..., 0] + 1j * out[..., 1]


def test_checker_matches_independent_numpy_definition():
    rng = np.random.default_rng(5)
    D, NN, n_out = 20, 192, 256
    iq = rng.integers(0, 256, 2 * D * n_out, dtype=np.uint8)
    b = np.zeros(NN, np.float32)
    bind.lib("port").create_fir_lpf(b.ctypes.data, NN, 0.95 / D)
    inc = np.array([0, 123456789, 2**31 + 5, 2**32 - 41943040], np.uint32)        # 0 Hz, odd, past Nyquist, -200 kHz at 20.48 MS/s
    got = _chan_oracle(iq, 0, D, NN, b, inc)
    x = (iq[0::2].astype(np.float64) - 127.0) + 1j * (iq[1::2].astype(np.float64) - 127.0)
    n = np.arange(x.size, dtype=np.uint64)
    for c, ic in enumerate(inc):
        ph = (np.uint64(ic) * n) % np.uint64(2**32)
        xs = np.concatenate([np.zeros(NN, complex), x * np.exp(-2j * np.pi * ph.astype(np.float64) / 2.0**32)])
        ref = np.array([np.dot(b.astype(np.float64), xs[(i + 1) * D:(i + 1) * D + NN]) for i in range(n_out)])
        assert np.abs(got[c] - ref).max() < 1e-9 * max(np.abs(ref).max(), 1.0)


def test_checker_is_continuous_across_blocks():
    """Two calls with carried history and absolute index == one call over the concatenation."""
    rng = np.random.default_rng(6)
    D, NN, n_out = 20, 192, 128
    iq = rng.integers(0, 256, 2 * D * n_out * 2, dtype=np.uint8)
    b = np.zeros(NN, np.float32)
    bind.lib("port").create_fir_lpf(b.ctypes.data, NN, 0.95 / D)
    inc = np.array([987654321], np.uint32)
    whole = _chan_oracle(iq, 0, D, NN, b, inc)
    half = iq.size // 2
    first = _chan_oracle(iq[:half], 0, D, NN, b, inc)
    hist = np.ascontiguousarray(iq[half - 2 * NN:half])
    second = _chan_oracle(iq[half:], D * n_out, D, NN, b, inc, hist)
    assert np.abs(np.concatenate([first, second], axis=1) - whole).max() < 1e-9


def test_wideband_generator_is_reproducible_and_selects_stations():
    cent = synth.wideband_centres(100)
    assert len(cent) == 100 and cent[0] == -9.9e6 and cent[-1] == 9.9e6 and np.allclose(np.diff(cent), 200e3)
    idx = [10, 60]
    ps = [synth.StreamParams.for_stream(s) for s in idx]
    a = synth.synth_wideband_u8(1 << 16, cent[idx], ps)
    b = synth.synth_wideband_u8(1 << 16, cent[idx], ps, chunk=1 << 13)          # chunking must not change the bytes
    assert a.dtype == np.uint8 and a.size == 2 << 16 and np.array_equal(a, b)
    # the spectrum has its energy at the two centres
    x = (a[0::2].astype(np.float64) - 127.0) + 1j * (a[1::2].astype(np.float64) - 127.0)
    f = np.fft.fftfreq(x.size, 1.0 / synth.FS_WIDEBAND)
    p = np.abs(np.fft.fft(x)) ** 2
    for fc in cent[idx]:
        assert p[np.abs(f - fc) < 150e3].sum() > 0.3 * p.sum()


def test_channel_partition():
    from fm_radio_b200 import batch
    for n, w in ((100, 8), (100, 1), (7, 2), (3, 8)):
        parts = [batch.shard_channels(n, r, w) for r in range(w)]
        assert sorted(c for p in parts for c in p) == list(range(n))
        assert all(c % w == r for r, p in enumerate(parts) for c in p)


WORKER = textwrap.dedent("""
    import os, sys, json, ctypes as C
    import numpy as np
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from fm_radio_b200 import synth
    from fm_radio_b200.batch import WidebandReceiver, gather_results
    from oracle import bind

    D, NN, B, NB = 20, 192, 16384, 75                       # 1.2 s of signal
    L = bind.lib("port")._cdll
    L.fmo_channelize_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]

    class CheckerChan:
        '''feed() stand-in for Channelizer: the float64 checker, then one CPU demodulator per channel.'''
        def __init__(self, centres):
            self.inc = np.array([int(round(((f / synth.FS_WIDEBAND) %% 1.0) * 2**32)) %% 2**32 for f in centres], np.uint32)
            self.b = np.zeros(NN, np.float32); bind.lib("port").create_fir_lpf(self.b.ctypes.data, NN, 0.95 / D)
            self.hist = None; self.n0 = 0
        def feed(self, demod, iq):
            iq = iq.numpy()
            out = np.zeros((len(self.inc), B, 2), np.float64)
            L.fmo_channelize_f64(iq.ctypes.data, B * D, None if self.hist is None else self.hist.ctypes.data, self.n0, D, NN,
                                 self.b.ctypes.data, self.inc.ctypes.data, len(self.inc), out.ctypes.data)
            self.hist = iq[-2 * NN:].copy(); self.n0 += B * D
            for c, chk in enumerate(demod.chk):
                chk.process_cf32((out[c, :, 0] + 1j * out[c, :, 1]).astype(np.complex64))
            return 0

    class CheckerDemod:
        def __init__(self, n): self.chk = [bind.CpuDemod(B, "port") for _ in range(n)]
        def rds_fetch(self): pass
        def rds_db(self, i): return self.chk[i].db()
        def rds_counts(self, i): return (len(self.chk[i].groups()[0]), 0)

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    idx = [20, 50, 81]                                       # 3 stations over 2 ranks: {0, 2} and {1}
    cent = synth.wideband_centres(100)[idx]
    params = [synth.StreamParams.for_stream(500 + s) for s in idx]
    mine = list(range(rank, len(idx), world))
    rx = WidebandReceiver(synth.FS_WIDEBAND, cent, rank, world, block_out=B, decimation=D, n_taps=NN,
                          chan=CheckerChan(cent[mine]), demod=CheckerDemod(len(mine)))
    assert rx.channel_ids == mine
    cap = synth.synth_wideband_u8(B * D * NB, cent, params, device="cpu") if rank == 0 else None
    for k in range(NB):
        blk = cap[2 * B * D * k:2 * B * D * (k + 1)].clone() if rank == 0 else torch.zeros(2 * B * D, dtype=torch.uint8)
        rx.broadcast_and_feed(blk, src=0)                    # the path's one collective
    res = gather_results(rx.results())
    if rank == 0:
        print("RESULT " + json.dumps([[r[0], r[1], r[2].decode("latin1"), r[4]] for r in res]))
    dist.destroy_process_group()
""") % H.ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_ranks_broadcast_channelize_decode_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=H.ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    assert [x[0] for x in res] == [0, 1, 2]
    for (cid, pi, ps, n_groups), s in zip(res, (20, 50, 81)):
        p = synth.StreamParams.for_stream(500 + s)
        assert pi == p.pi_code, (cid, hex(pi))
        assert n_groups >= 5
