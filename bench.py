#!/usr/bin/env python
"""Benchmark of the FM stereo + RDS demodulation hot path (contract: see the task statement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU chain

Workload = BASELINE.json configs[2] ("config 3"): 1024 independent stereo+RDS streams per GPU,
one STEP = one 65536-sample IQ block of every stream (67.1 M IQ samples per GPU per step; 128 MiB
of u8 input per step, larger than the 126 MB L2, cycling over distinct blocks).  Weak scaling:
N GPUs demodulate N x 1024 streams (config 5's 8192 at N = 8) with no data-path collective.

metric  IQ MS/s demodulated (stereo audio + RDS symbols out), whole job.
value   inputs resident in HBM, blocks pipelined through the handle's stage streams.
e2e     the same through the host-facing C-ABI: pinned host u8 in (H2D inside the timed region),
        audio + RDS symbols out to pinned host memory (D2H inside the timed region) every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 65536
STREAMS_PER_GPU = 1024
FS = 1_024_000
FLOP_PER_SAMPLE_K1 = 64.0          # 64 real taps x (re, im) x FMA=2 x 1/4 output rate
FLOP_PER_SAMPLE_CHAIN = 160.0      # SURVEY.md 8(d) headline figure
BYTES_PER_SAMPLE_FUSED = 2.26      # 2 B u8 in + 0.25 B audio + 0.01 B symbols
BYTES_PER_SAMPLE_K1 = 3.0          # K1 as built: 2 B u8 in + 4 B fm_demod out per 4 samples
N_SM, FP32_LANES = 148, 128
# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of every kernel at the default workload come from the
# `ncu --set full` capture of the same command, exported by tools/summarize_ncu.py --traffic-json (a run under a profiler is
# never timed, so the bench can only read the capture; the file and its source are named in the JSON line)
NCU_TRAFFIC_JSON = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")


def ncu_traffic():
    try:
        return json.load(open(NCU_TRAFFIC_JSON))
    except Exception:
        return {}


# algorithmic FLOP per input IQ sample of each kernel (per-stage figures; FMA = 2)
KERNEL_FLOP_PER_SAMPLE = {"k1_fir4_discrim": 64.0,      # a2 (a1 unpack 2.0 and a3 discriminator 1.0 not counted); on the tensor cores since round 2
                          "k2_mpx": 16.0 + 16.25 + 3.0 + 0.6,   # a4 + a6 + a7 + a8
                          "k3_pll": 7.5,                 # a9
                          "k4_mix_fir": 16.0 + 7.5 + 16.0 + 8.0,   # a10 + a11 + a12 + a13
                          "k5_bpsk": 2.2,                # a14-a16
                          "k6_rds": 0.0,                 # a19: integer only
                          "k7_audio_pcm": 6.0 * 48000 / FS}   # Resample(): 2 mul + 1 add per channel per 48 kHz frame
# algorithmic HBM bytes per input IQ sample of each kernel as built (reads + writes of its buffers)
KERNEL_BYTES_PER_SAMPLE = {"k1_fir4_discrim": 2.0 + 1.0, "k2_mpx": 1.0 + 1.0 + 0.5, "k3_pll": 0.5 + 0.5,
                           "k4_mix_fir": 1.0 + 0.5 + 0.25 + 0.125, "k5_bpsk": 0.125 + 0.0625,
                           "k6_rds": 0.0625,
                           "k7_audio_pcm": 8.0 * 32000 / FS + (8.0 + 4.0) * 48000 / FS}   # f32 frames in; f32 + s16 frames out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.first = index, [], None, 0

    def mark(self):
        """The timed region starts here: rows sampled before this call are not reported."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", os.environ.get("BENCH_SMI_MS", "25")], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = self.rows[self.first:] or self.rows
        sm = sorted(float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# --------------------------------------------------------------------------------------------
# CPU reference: the unmodified reference's fm_demod_benchmark (oracle/_ref, compiled in place)
# --------------------------------------------------------------------------------------------
def _write_capture(args):
    seed, n_blocks, path = args
    from fm_radio_b200 import synth
    synth.synth_u8_numpy(BLOCK * n_blocks, synth.StreamParams.for_stream(seed)).tofile(path)
    return path


def cpu_reference_run(n_procs: int, n_blocks: int, repeats: int, warmup: int, tmpdir: str):
    """n_procs concurrent `fm_demod_benchmark -b 65536 -i capture` processes (the reference DSP is
    single threaded), one capture of n_blocks blocks each; returns per-repeat wall s."""
    from concurrent.futures import ProcessPoolExecutor
    from oracle import bind
    exe = bind.REF_BENCH
    kind = "reference"
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/fm_demod_benchmark missing: run __graft_entry__.build() where /root/reference exists")
    n_unique = min(n_procs, 8)
    with ProcessPoolExecutor(max_workers=min(n_unique, os.cpu_count() or 1)) as ex:
        files = list(ex.map(_write_capture, [(s, n_blocks, os.path.join(tmpdir, f"cap{s}.u8")) for s in range(n_unique)]))
    times = []
    for it in range(warmup + repeats):
        t0 = time.perf_counter()
        procs = [subprocess.Popen([exe, "-b", str(BLOCK), "-i", files[i % n_unique]], stdout=subprocess.DEVNULL,
                                  stderr=subprocess.DEVNULL) for i in range(n_procs)]
        for p in procs:
            p.wait()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return times, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_blocks = 156                      # 10 s of signal per process (config 1's capture length)
    with tempfile.TemporaryDirectory() as td:
        times, kind = cpu_reference_run(cores, n_blocks, max(args.steps, 1), max(args.warmup, 1), td)
    samples = cores * n_blocks * BLOCK
    t = sum(times) / len(times)
    value = samples / t / 1e6
    line = {
        "impl": "reference", "metric": "IQ MS/s demodulated stereo+RDS", "value": value, "unit": "MS/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": max(args.warmup, 1), "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1024 streams x 65536-sample blocks per GPU (BASELINE config 3); reference arm = "
                               f"{cores} concurrent single-threaded fm_demod_benchmark processes, 10 s capture each",
                   "block_size": BLOCK},
        "cpu_baseline": {"value": value, "unit": "MS/s", "cores": cores, "kind": kind,
                         "sample": f"{cores} processes x {n_blocks} blocks x {BLOCK} IQ samples per step"},
        "e2e": {"value": value, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "x_realtime": value * 1e6 / FS,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------------------------
def run_cuda_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import fm_radio_b200 as fm
    from fm_radio_b200 import Buf, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if not args.no_affinity:
        # pin this rank to the CPUs local to its GPU before any pinned host memory is allocated, so the e2e
        # path's host buffers sit on the GPU's own NUMA node (no effect on a single-socket host)
        try:
            import pynvml
            pynvml.nvmlInit()
            hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            pynvml.nvmlDeviceSetCpuAffinity(hnd)
            numa = sorted(os.sched_getaffinity(0))
            numa = f"{len(numa)} cpus {numa[0]}-{numa[-1]}"
        except Exception as ex:  # noqa: BLE001
            numa = f"unavailable ({type(ex).__name__})"
    S, B = args.streams, BLOCK
    n_in = args.input_blocks

    # ---- synthetic stereo+RDS captures for this rank's streams, resident in HBM ----
    from fm_radio_b200.batch import shard_streams
    my_streams = shard_streams(S * world, rank, world)
    params = [synth.StreamParams.for_stream(s) for s in my_streams]
    cap = torch.empty((n_in, S, 2 * B), dtype=torch.uint8, device=dev)
    chunk = 64
    for s0 in range(0, S, chunk):
        piece = synth.synth_u8_torch(B * n_in, params[s0:s0 + chunk], dev)
        cap[:, s0:s0 + chunk] = piece.view(len(params[s0:s0 + chunk]), n_in, 2 * B).transpose(0, 1)
    torch.cuda.synchronize()

    demod = fm.FMDemod(B, S, device=local_rank, pipeline_depth=args.depth)
    if args.audio_pcm_rate:
        # the north star's "48 kHz audio": the audio output stage K7 runs for every block, on and off the clock
        demod.set_control(fm.Control.AUDIO_PCM_RATE_HZ, args.audio_pcm_rate)
    ext = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        demod.sync()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    demod.wait_external_stream(ext)
    # clock ramp: a box that has just been handed out idles at 120 MHz and takes ~0.3 s of load to reach
    # its boost clock (measured: the first 48 steps after idle ran 1.45x slower), so besides the W warm-up
    # steps the GPU is kept busy with untimed blocks for args.clock_warmup_ms before anything is timed
    # nvidia-smi is started here, before the clock ramp, so that its start-up (NVML attach) is over when the timed
    # region begins; only rows sampled after mark() are reported
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_w = time.perf_counter()
    n_ramp = 0
    while (time.perf_counter() - t_w) * 1e3 < args.clock_warmup_ms:
        for k in range(8):
            demod.enqueue_u8_device(cap[(n_ramp + k) % n_in])
        n_ramp += 8
        demod.sync()
    for k in range(args.warmup):
        demod.enqueue_u8_device(cap[k % n_in])
    barrier()
    if args.clock_warmup_ms < 300.0:
        time.sleep(0.3)                                  # nvidia-smi's start-up must be over
    launches0 = demod.launch_count
    sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    demod.wait_external_stream(ext)
    for k in range(args.steps):
        demod.enqueue_u8_device(cap[(args.warmup + k) % n_in])
    demod.signal_external_stream(ext)
    e1.record()
    barrier()
    t_dev = e0.elapsed_time(e1) * 1e-3
    launches = demod.launch_count - launches0
    slot = (demod.blocks_enqueued - 1) % demod.depth
    demod.fetch_outputs(slot)
    demod.sync()
    audio = demod.get(Buf.AUDIO_OUT, 0)
    assert np.isfinite(audio).all() and np.abs(audio).max() > 1e-3, "demodulated audio is empty"

    # ---- end to end through the host-facing C-ABI: pinned host in, pinned host out ----
    n_host = 2
    # what the bulk consumer of this workload takes home: 48 kHz int16 PCM (fm_scraper.cpp:74-78) + the compacted soft
    # symbols; the 32 kHz float frames stay on the device (fmgpu_set_fetch_mask)
    if args.audio_pcm_rate:
        demod.set_fetch_mask(demod.FETCH_PCM_S16 | demod.FETCH_RDS_SYMBOLS)
    host_in = [torch.empty((S, 2 * B), dtype=torch.uint8).pin_memory() for _ in range(n_host)]
    for i in range(n_host):
        host_in[i].copy_(cap[i % n_in])
    torch.cuda.synchronize()
    e2e_steps = max(4, args.steps // 2)
    for k in range(2):
        sl = demod.enqueue_u8_host(host_in[k % n_host]); demod.fetch_outputs(sl)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        sl = demod.enqueue_u8_host(host_in[k % n_host])
        demod.fetch_outputs(sl)
    demod.sync()
    t_e2e = time.perf_counter() - t0
    clocks = sampler.stop()                              # sampled across both timed regions (device-resident and e2e)
    h2d = S * 2 * B
    sym_cap = min(B // 64, (B // 64) * 7 // 32 + 2)        # compacted symbol rows (fmgpu.cu fetch_slot)
    d2h = S * sym_cap * 4 + S * 4
    if args.audio_pcm_rate:
        d2h += S * int(np.float32(args.audio_pcm_rate) / np.float32(32000.0) * np.float32(B // 32)) * 4   # int16 stereo PCM
    else:
        d2h += S * (B // 32) * 8                           # 32 kHz float frames
    demod.set_fetch_mask(demod.FETCH_ALL)

    # ---- per-kernel device times, one block at a time (events inside the library) ----
    demod.sync()
    stage_ms = demod.profile_stages(cap[0], 4)
    stage_ms_piped = demod.profile_stages(cap[1], -32)      # the same kernels while the stages overlap

    # ---- correctness inside the bench: a fresh handle demodulates the n_in CONTINUOUS blocks of every
    #      stream and the device RDS decoders (K6) must have found each stream's own PI code ----
    rds_ok = rds_locked = None
    if n_in * B >= FS:                                   # needs >= 1 s of continuous signal to lock and decode
        chk = fm.FMDemod(B, S, device=local_rank, pipeline_depth=args.depth)
        chk.wait_external_stream(ext)
        for k in range(n_in):
            chk.enqueue_u8_device(cap[k])
        chk.rds_fetch()
        rds_ok = sum(chk.rds_db(i)["pi"] == params[i].pi_code for i in range(S))
        rds_locked = sum(chk.rds_counts(i)[0] > 0 for i in range(S))
        chk.close()

    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e = float(tt[0]), float(tt[1])
    samples_step = world * S * B
    value = samples_step * args.steps / t_dev / 1e6
    e2e_value = samples_step * e2e_steps / t_e2e / 1e6

    line = None
    if rank == 0:
        peaks = measured_peaks()
        sm_max = clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        fp32_peak = N_SM * FP32_LANES * 2 * sm_max * 1e6 / 1e12             # TFLOP/s at max SM clock
        k1_ms = stage_ms["k1_fir4_discrim"]
        k1_flops = FLOP_PER_SAMPLE_K1 * S * B
        k1_bytes = BYTES_PER_SAMPLE_K1 * S * B
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic = ncu_traffic()
        default_wl = (S == STREAMS_PER_GPU and B == BLOCK)
        k1_name = "k1_fir4_discrim_u8" if os.environ.get("FMGPU_K1_FP32") else "k1_toeplitz_i8"
        k1_on_tensor = not os.environ.get("FMGPU_K1_FP32")

        def bound_of(name):
            if name in ("k3_pll", "k5_bpsk", "k6_rds"):
                return "latency (one thread per stream, dependent chain)"
            if name == "k7_audio_pcm" or (name == "k1_fir4_discrim" and k1_on_tensor):
                return "hbm"
            return "fp32"
        # K1 -- the kernel that touches every input byte -- runs its FIR on the tensor cores (tcgen05 kind::i8, exact), which
        # takes it off the FP32 roofline: its bound is HBM (2 B in + 1 B out per IQ sample).  achieved = algorithmic bytes
        # per launch / the launch's duration (CUDA events inside the library, one block at a time); peak = measured copy
        # bandwidth (MEASURED_PEAKS.json).  The FP32 figures of the FMA-pipe kernels (K2, K4) are in "kernels".
        roof = {
            "kernel": k1_name, "bound": "hbm" if k1_on_tensor else "fp32",
            "achieved": (k1_bytes / (k1_ms * 1e-3) / 1e9) if k1_on_tensor else (k1_flops / (k1_ms * 1e-3) / 1e12),
            "peak": hbm_peak if k1_on_tensor else fp32_peak, "unit": "GB/s" if k1_on_tensor else "TFLOP/s",
            "frac": (k1_bytes / (k1_ms * 1e-3) / 1e9 / hbm_peak) if k1_on_tensor else (k1_flops / (k1_ms * 1e-3) / 1e12 / fp32_peak),
            "peak_source": ("MEASURED_PEAKS.json hbm_gbs (burst copy bandwidth; the kernel is timed alone)" if "hbm_gbs" in peaks else "fallback 6650 GB/s")
                           if k1_on_tensor else f"148 SMs x 128 FP32 lanes x 2 x {sm_max:.0f} MHz",
            "traffic": (traffic.get(k1_name, {}).get("dram_bytes") if default_wl else None),
            "traffic_source": f"{os.path.relpath(NCU_TRAFFIC_JSON, ROOT)} <- {traffic.get('_source')} (ncu --set full of this command, "
                              "dram__bytes_read.sum + dram__bytes_write.sum per launch; read from the capture, not measured by this run)",
            "algorithmic_bytes_per_launch": k1_bytes, "flop_per_launch": k1_flops, "ms_per_launch": k1_ms,
            "sm_partition_note": "the kernel runs on the FIR partition (132 of 148 SMs) of the handle",
            "fp32_peak": {"value": fp32_peak, "unit": "TFLOP/s",
                          "source": f"148 SMs x 128 FP32 lanes x 2 x {sm_max:.0f} MHz (MEASURED_PEAKS.json has no fp32 figure; tools/ffma_probe.cu reached 72-73)"},
            "kernels": {name: {"ms": ms, "tflops": KERNEL_FLOP_PER_SAMPLE[name] * S * B / (ms * 1e-3) / 1e12,
                               "frac_fp32": KERNEL_FLOP_PER_SAMPLE[name] * S * B / (ms * 1e-3) / 1e12 / fp32_peak,
                               "gbs": KERNEL_BYTES_PER_SAMPLE[name] * S * B / (ms * 1e-3) / 1e9,
                               "frac_hbm": KERNEL_BYTES_PER_SAMPLE[name] * S * B / (ms * 1e-3) / 1e9 / hbm_peak,
                               "bound": bound_of(name)}
                        for name, ms in stage_ms.items()},
            "chain": {"flop_per_sample": FLOP_PER_SAMPLE_CHAIN,
                      "achieved_tflops": FLOP_PER_SAMPLE_CHAIN * value * 1e6 / world / 1e12,
                      "frac_of_fp32_peak": FLOP_PER_SAMPLE_CHAIN * value * 1e6 / world / 1e12 / fp32_peak,
                      "hbm_frac_fused_min": BYTES_PER_SAMPLE_FUSED * value * 1e6 / world / 1e9 / hbm_peak},
        }
        line = {
            "metric": "IQ MS/s demodulated stereo+RDS", "value": value, "unit": "MS/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "dtype_note": "f32 throughout, as the reference; the first FIR (K1) sums u8 samples x 24-bit fixed-point taps exactly in int32 "
                          "(tcgen05 kind::i8, three signed digit planes) and recombines in f32 -- |tap error| <= 2^-26, not a reduced precision",
            "config": {"workload": f"{S} streams x {B}-sample u8 IQ blocks per GPU (BASELINE config 3; N GPUs = config 5 sharded by stream)",
                       "streams_per_gpu": S, "block_size": B, "pipeline_depth": demod.depth,
                       "audio_out": (f"32 kHz f32 frames + {args.audio_pcm_rate} Hz f32 and int16 PCM (K7)" if args.audio_pcm_rate
                                     else "32 kHz f32 frames"), "clock_warmup_ms": args.clock_warmup_ms,
                       "sm_partition": dict(zip(("recurrence_sms", "fir_sms"), demod.partition())),
                       "l2": f"input per step {S * 2 * B / 2**20:.0f} MiB > 126 MB L2, cycling over {n_in} distinct continuous blocks per stream",
                       "parallelism": f"streams sharded over {world} GPU(s), no data-path collective"},
            "x_realtime": value * 1e6 / FS,
            "e2e": {"value": e2e_value, "unit": "MS/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "x_realtime": e2e_value * 1e6 / FS,
                    "api": "fmgpu_enqueue_u8_host + fmgpu_fetch_outputs per block, pinned host buffers",
                    "cpu_affinity": numa},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "stage_ms_serial": stage_ms,
            "stage_ms_pipelined": stage_ms_piped,
            "rds_check": {"streams_with_own_pi_decoded_on_device": rds_ok, "streams_with_groups": rds_locked, "of": S,
                          "signal_s": n_in * B / FS},
        }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cores = os.cpu_count() or 1
            with tempfile.TemporaryDirectory() as td:
                times, kind = cpu_reference_run(cores, 156, 2, 1, td)      # 10 s of signal per process (config 1's capture)
            v = cores * 156 * BLOCK / (sum(times) / len(times)) / 1e6
            line["cpu_baseline"] = {"value": v, "unit": "MS/s", "cores": cores, "kind": kind,
                                    "sample": f"{cores} concurrent fm_demod_benchmark processes x 156 blocks x {BLOCK} samples (10 s of signal each), mean of 2 after 1 warm-up",
                                    "per_core": v / cores}
        except Exception as ex:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "MS/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    demod.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# CUDA arm, wideband workload (BASELINE configs 4 / 5): one 20.48 MS/s u8 capture, 100 stations, channels
# sharded over the ranks; the only data-path collective is the NCCL broadcast of the wideband block
# --------------------------------------------------------------------------------------------
def run_wideband_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import fm_radio_b200 as fm
    from fm_radio_b200 import ChanMode, synth
    from fm_radio_b200.batch import WidebandReceiver, gather_results

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    n_st, B, D = args.stations, BLOCK, 20
    cent = synth.wideband_centres(n_st)
    ps = [synth.StreamParams.for_stream(2000 + s) for s in range(n_st)]
    n_in = B * D
    n_cap = 32                                               # 2.1 s of continuous signal, cycled
    if rank == 0:
        cap = synth.synth_wideband_u8(n_in * n_cap, cent, ps, device=dev)
        blocks = [cap[2 * n_in * k:2 * n_in * (k + 1)] for k in range(n_cap)]
        host_blocks = [b.cpu().pin_memory() for b in blocks[:2]]
    rx = WidebandReceiver(synth.FS_WIDEBAND, cent, rank, world, B, D, 192, mode=ChanMode.AUTO, device=local_rank)
    rx.demod.set_control(fm.Control.AUDIO_PCM_RATE_HZ, args.audio_pcm_rate or 48000)
    stage = torch.empty(2 * n_in, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())

    def step(k, from_host=False):
        if rank == 0:
            stage.copy_(host_blocks[k % 2] if from_host else blocks[k % n_cap], non_blocking=True)
        rx.broadcast_and_feed(stage)                         # NCCL broadcast (world > 1), channelize + demodulate own channels

    def barrier():
        rx.demod.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    with torch.cuda.stream(side):
        # clock ramp + warm-up: the SAME number of steps on every rank (each step is a collective: a time-based loop would
        # leave the ranks with different broadcast counts and hang) -- at >= 0.17 ms per step this covers clock_warmup_ms
        n_ramp = max(n_cap + args.warmup, int(args.clock_warmup_ms / 0.17))
        for k in range(n_ramp):
            step(k)
            if (k + 1) % 8 == 0:
                rx.demod.sync()
        k = n_ramp
        barrier()
        sampler.mark()
        launches0 = rx.demod.launch_count + rx.chan.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in range(args.steps):
            step(k + j)
        rx.demod.signal_external_stream(side.cuda_stream)
        e1.record()
        barrier()
        t_dev = e0.elapsed_time(e1) * 1e-3
        launches = rx.demod.launch_count + rx.chan.launch_count - launches0
        # end to end: the ingest rank's block comes from pinned host memory, every rank fetches its stations' PCM + symbols
        rx.demod.set_fetch_mask(rx.demod.FETCH_PCM_S16 | rx.demod.FETCH_RDS_SYMBOLS)
        e2e_steps = max(4, args.steps // 2)
        for j in range(2):
            step(j, True); rx.demod.fetch_outputs((rx.demod.blocks_enqueued - 1) % rx.demod.depth)
        barrier()
        t0 = time.perf_counter()
        for j in range(e2e_steps):
            step(j, True)
            rx.demod.fetch_outputs((rx.demod.blocks_enqueued - 1) % rx.demod.depth)
        rx.demod.sync()
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        # the channelizer kernel alone (own stations of this rank), device-resident
        rx.chan.sync()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cs = torch.cuda.ExternalStream(rx.chan.stream)
        c0.record(cs)
        for j in range(16):
            rx.chan.enqueue_u8_device(stage)
        c1.record(cs)
        rx.chan.sync()
        chan_ms = c0.elapsed_time(c1) / 16
    clocks = sampler.stop()
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e, chan_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, chan_ms = float(tt[0]), float(tt[1]), float(tt[2])
    # correctness inside the bench: a FRESH receiver takes the n_cap continuous blocks once (the timed receiver has seen
    # the capture wrap around, i.e. phase jumps every 2.1 s) and every station must decode its own PI on the device
    rx.demod.sync()
    chk = WidebandReceiver(synth.FS_WIDEBAND, cent, rank, world, B, D, 192, mode=ChanMode.AUTO, device=local_rank)
    with torch.cuda.stream(side):
        for k in range(n_cap):
            if rank == 0:
                stage.copy_(blocks[k], non_blocking=True)
            chk.broadcast_and_feed(stage)
        chk.demod.sync()
        torch.cuda.synchronize()
    res = gather_results(chk.results())
    chk.close()
    own_pi = sum(1 for (c, pi, _ps, _rt, _n) in res if pi == ps[c].pi_code)
    my_st = len(rx.channel_ids)
    value = n_st * B * args.steps / t_dev / 1e6              # station-rate IQ samples demodulated per second, whole job
    e2e_value = n_st * B * e2e_steps / t_e2e / 1e6
    if rank == 0:
        peaks = measured_peaks()
        int8_peak = 2.0 * peaks.get("bf16_tflops", 2250.0 / 2 * 2)          # dense int8 = 2 x dense bf16 on sm_100
        macs = 2.0 * 192 * 4 * my_st * B                     # algorithmic real ops of one launch: 192 complex taps x complex data
        pcm_n = int(np.float32(args.audio_pcm_rate or 48000) / np.float32(32000.0) * np.float32(B // 32))
        sym_cap = min(B // 64, (B // 64) * 7 // 32 + 2)
        line = {
            "metric": "IQ MS/s demodulated stereo+RDS", "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_dev / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "i8 (channelizer, int32 accumulate) + f32 (demodulators)", "data": "synthetic",
            "config": {"workload": f"wideband: one {synth.FS_WIDEBAND / 1e6:.2f} MS/s u8 capture, {n_st} stations at 200 kHz, polyphase channelizer + "
                                   f"per-station demodulator + device RDS (BASELINE configs 4 / 5), {B * D} wideband samples per step",
                       "stations": n_st, "stations_per_rank": my_st, "block_out": B, "decimation": D,
                       "parallelism": f"channels sharded c mod {world}; NCCL broadcast of the wideband u8 block from rank 0 each step"
                                      if world > 1 else "one GPU, no collective",
                       "broadcast_bytes_per_step": 2 * n_in if world > 1 else 0,
                       "l2": f"wideband block {2 * n_in / 2**20:.1f} MiB per step, cycling over {n_cap} continuous blocks"},
            "x_realtime_wideband": n_in * args.steps / t_dev / synth.FS_WIDEBAND,
            "wideband_MSps": n_in * args.steps / t_dev / 1e6,
            "e2e": {"value": e2e_value, "unit": "MS/s", "h2d_bytes_per_step": 2 * n_in,
                    "d2h_bytes_per_step": n_st * (pcm_n * 4 + sym_cap * 4 + 4), "steps": e2e_steps,
                    "api": "pinned host block -> rank 0 -> broadcast -> fmgpu_chan_feed_device; fmgpu_fetch_outputs (PCM + symbols) per rank"},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": "chan_mma_i8_pipelined", "bound": "tensor", "achieved": macs / (chan_ms * 1e-3) / 1e12, "peak": int8_peak,
                         "unit": "TOP/s", "frac": macs / (chan_ms * 1e-3) / 1e12 / int8_peak,
                         "traffic": (ncu_traffic().get("chan_mma_i8_pipelined", {}).get("dram_bytes") if (world == 1 and n_st == 100) else None),
                         "traffic_source": "profiles/r2_ncu_traffic.json (ncu --set full of tools/chan_profile.py, 100 stations x 65536 outputs; "
                                           "read from the capture, not measured by this run)",
                         "ms_per_launch": chan_ms, "ops_per_launch": macs,
                         "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (dense int8 = 2 x dense bf16 on sm_100)",
                         "note": "algorithmic ops (192 complex taps); the kernel executes 3 digit planes of them and the step is "
                                 "bound by the demodulators' recurrence latency at this stream count, not by the channelizer"},
            "rds_check": {"stations_with_own_pi_decoded_on_device": own_pi, "of": n_st},
        }
        print(json.dumps(line), flush=True)
    rx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="streams", choices=["streams", "wideband"],
                    help="streams: 1024 independent streams per GPU (config 3, the headline); wideband: 100 stations from one "
                         "20.48 MS/s capture, channels sharded over the GPUs, NCCL broadcast of the capture (configs 4 / 5)")
    ap.add_argument("--stations", type=int, default=100)
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="streams per GPU")
    ap.add_argument("--depth", type=int, default=3, help="pipeline depth = ring slots = blocks in flight (3: the chain's latency is 2.7 steps; deeper rings let the first stages run ahead and lengthen the drain of short runs)")
    ap.add_argument("--input-blocks", type=int, default=24, help="distinct, CONTINUOUS input blocks per stream kept in HBM (24 = 1.5 s of signal, 3.2 GB)")
    ap.add_argument("--audio-pcm-rate", type=int, default=48000, help="audio output stage K7: resample GetAudioOut to this rate + int16 PCM (0 = off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-affinity", action="store_true", help="do not bind the rank to its GPU's local CPUs")
    ap.add_argument("--clock-warmup-ms", type=float, default=500.0, help="untimed load before the W warm-up steps (clock ramp from idle)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference_arm(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.workload == "wideband":
        run_wideband_arm(args)
        return
    run_cuda_arm(args)


if __name__ == "__main__":
    main()
