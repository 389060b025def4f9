"""Builds fm_radio_b200/libfmgpu.so (the C-ABI of include/fmgpu.h) in-tree with nvcc for sm_100a.

The shared object is git-ignored but travels to the GPU box with the snapshot; nothing is JIT
compiled at run time."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("FMGPU_BUILD_OUT") or os.path.join(HERE, "libfmgpu.so")        # FMGPU_BUILD_OUT + FMGPU_NVCC_EXTRA: A/B builds
CU_SOURCES = ["fmgpu.cu", "k1_fir4_discrim.cu", "k1_toeplitz_i8.cu", "k2_mpx.cu", "k3_pll.cu", "k4_mix_fir.cu", "k5_bpsk.cu", "k6_rds.cu", "k7_audio_pcm.cu", "k_fft.cu", "k_misc.cu", "chan.cu"]
CPP_SOURCES = ["filter_designer.cpp", "rds_host.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources() -> list[str]:
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "fmgpu.h"))
    return [os.path.join(CSRC, f) for f in CU_SOURCES + CPP_SOURCES] + hdrs


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    extra = os.environ.get("FMGPU_NVCC_EXTRA", "").split()
    objdir = os.path.join(HERE, "build" + ("_" + "".join(c for c in "_".join(extra) if c.isalnum() or c == "_") if extra else ""))
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    log = []
    for src in CU_SOURCES + CPP_SOURCES:
        obj = os.path.join(objdir, src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
