"""CPU check of the K5 kernel's control-flow restructuring (fm_radio_b200/csrc/k5_core.h compiled for the host):
the symbol-wise loop the kernel runs must produce, bit for bit, the symbols and the loop state of the literal
per-sample loop (the reference's order, bpsk_synchroniser.cpp:94-186), for every batch length, on the RDS
baseband of a real capture including the acquisition transient; and the symbols must match the checker's."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from oracle import bind
from tests import helpers as H

SRC = os.path.join(H.ROOT, "tests", "native", "k5_core_check.cpp")


@pytest.fixture(scope="module")
def k5lib(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("k5") / "libk5check.so"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), SRC], check=True)
    L = C.CDLL(str(so))
    vp = C.c_void_p
    L.k5_check.argtypes = [vp, C.c_int, C.c_int, vp, vp, C.c_float, C.c_int, vp, vp, vp, vp, vp]
    return L


def test_symbolwise_loop_is_bit_identical_to_the_per_sample_loop(k5lib):
    B, nblk = H.B, 40
    cap = H.capture("seed0", 70)
    d = bind.CpuDemod(B, "port")
    rds, syms = [], []
    for k in range(nblk):
        d.process_u8(cap[2 * B * k:2 * B * (k + 1)])
        rds.append(d.get("rds").copy())                  # after the RDS AGC: the synchroniser's input
        syms.append(d.get("rds_raw_sym").copy())
    tb, ta = d.taps("bpsk_ted_lpf")
    pb, pa = d.taps("bpsk_pll_lpf")
    x = np.ascontiguousarray(np.concatenate(rds)).view(np.float32)
    n = B // 64
    ted = np.array([tb[0], tb[1], ta[0]], np.float32)
    pll = np.array([pb[0], pb[1], pa[0]], np.float32)
    for nb in (1, 3, 7, 8):
        sl = np.zeros(nblk * n * 2, np.float32); ss = np.zeros(nblk * n * 2, np.float32)
        cl = np.zeros(nblk, np.int32); cs = np.zeros(nblk, np.int32); sd = np.zeros(16, np.float32)
        rc = k5lib.k5_check(x.ctypes.data, n, nblk, ted.ctypes.data, pll.ctypes.data, 1.0, nb, sl.ctypes.data, cl.ctypes.data,
                            ss.ctypes.data, cs.ctypes.data, sd.ctypes.data)
        assert rc == 0, (nb, rc)
        assert np.array_equal(cl, cs) and np.all(sd == 0.0)
        assert np.array_equal(sl, ss)
    # against the checker's own synchroniser (libm atan2f, roundf, no FMA contraction): rounding noise only
    for k in range(nblk):
        assert cl[k] == len(syms[k]), k
        mine = sl[k * n * 2:k * n * 2 + 2 * cl[k]].view(np.complex64)
        assert np.abs(mine - syms[k]).max() <= 2e-5 * max(1.0, np.abs(syms[k]).max()), k
