#!/usr/bin/env python
"""Benchmark of the fingerprint + alignment hot path (BASELINE.json metric).

One step = BASELINE config[1] ("C2") applied to P independent source/CDN pairs per GPU:
  1. fingerprint both 5-min 44.1 kHz streams of every pair (1024/256 Hann STFT -> 26-mel/13-MFCC,
     spectral descriptors, FP64 short-time energy / ZCR, YIN) with the algorithms built at the real
     sample rate ("fixed-sr" mode, so the MFCC work is not degenerate);
  2. normalised cross-correlation of the two short-time-energy series over +-60 s (20,671 lags);
  3. banded DTW (r = 50) of the lag-trimmed energy series.
`value` = audio-seconds fingerprinted per second of whole-step time with the PCM resident in HBM,
`e2e`   = the same through the host-pointer C ABI (pinned host PCM, H2D + D2H inside the timing).
`alignments_per_s` (extra key) = pairs per second of the same step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]

Under torchrun every rank runs its own P pairs on its own GPU (weak scaling, no data-path
collective: SURVEY §8e); the timed region is bracketed by barrier + synchronize and the max over
ranks is reported by rank 0 as ONE JSON line.
"""
from __future__ import annotations

import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 44100
WIN, HOP = 1024, 256
MAX_LAG_S = 60.0
DTW_BAND = 50
_OUT = sys.stdout
METRIC = "audio-sec/s fingerprinted (1024/256 MFCC); 5-min pair alignments/s @60s lag"
UNIT = "audio-s/s"
ALGO_BYTES_PER_FRAME = HOP * 8 + 13 * 8  # every input sample read once, 13 MFCC written


def pkg():
    return importlib.import_module("sonido-sonar_b200")


def make_pair(synth, seconds, idx, seed0=200):
    rng = np.random.default_rng(seed0 + idx)
    off = float(rng.uniform(-0.9, 0.9)) * MAX_LAG_S * (seconds / 300.0) if idx else 7.3 * (seconds / 300.0)
    return synth.aligned_pair(seconds, offset_seconds=off, sr=SR, seed=seed0 + 2 * idx)


def trim_by_lag(ea, eb, lag, length):
    """TruncateToAlignmentPCM's convention (extractors/alignment.go:239-243): lag > 0 skips the start of stream 2."""
    if lag >= 0:
        a, b = ea, eb[lag:]
    else:
        a, b = ea[-lag:], eb
    return np.ascontiguousarray(a[:length]), np.ascontiguousarray(b[:length])


class ClockSampler:
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md), sampled through NVML in a
    thread (spawning `nvidia-smi -lms` next to the timed loop measurably slows the launches it is meant to
    observe); falls back to a low-rate nvidia-smi loop when pynvml is unavailable."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index, period=0.05):
        self.sm, self.bits, self.max_mhz, self.how = [], 0, None, "nvml"
        self._stop = threading.Event()
        self.proc = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._loop_nvml, args=(period,), daemon=True)
            self.th.start()
        except Exception:
            self.nv, self.how = None, "nvidia-smi"
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                              str(index), "-lms", "500"], stdout=subprocess.PIPE,
                                             stderr=subprocess.DEVNULL, text=True)
                self.th = threading.Thread(target=self._loop_smi, daemon=True)
                self.th.start()
            except OSError:
                self.proc = None

    def _loop_nvml(self, period):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            self._stop.wait(period)

    def _loop_smi(self):
        names = (0x8, 0x40, 0x20, 0x4)
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) < 6:
                continue
            try:
                self.sm.append(float(r[0]))
                self.max_mhz = float(r[1])
            except ValueError:
                continue
            for bit, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    self.bits |= bit

    def stop(self):
        self._stop.set()
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples"], "how": self.how}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(n for b, n in self.REASONS.items() if self.bits & b), "samples": len(self.sm),
                "how": self.how}


def workload_config(P, seconds, T, max_lag, world):
    """`config` of the JSON line: identical for both arms (--impl ours / reference), so that the driver's ratio is a
    same-config ratio.  It names the WORKLOAD; how much of it a step of either arm processes is said in that arm's own
    keys (`pairs_per_gpu` here is the GPU arm's batch, the reference arm states its per-step sample in `cpu_baseline`)."""
    NS = 2 * P
    n = int(round(seconds * SR))
    return {
        "workload": f"C2 pipeline: fingerprint both {int(seconds)} s 44.1 kHz streams of each source/CDN pair (1024/256 "
                    "Hann, 26-mel/13-MFCC, fixed-sr mode; STFT/MFCC FP32, energy/ZCR FP64, YIN), NCC over "
                    f"+-{int(MAX_LAG_S)} s ({2 * max_lag + 1} lags), banded DTW r={DTW_BAND}",
        "pairs_per_gpu": P, "streams_per_gpu": NS, "frames_per_stream": int(T),
        "l2": f"inputs {NS * n * 8 / 1e9:.2f} GB per step exceed the 126 MB L2; no flush needed",
        "timing": "CUDA events on the library stream, max over ranks",
    }


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Per-launch DRAM bytes of the roofline kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's Go path (no Go toolchain here: SURVEY §8c)
# --------------------------------------------------------------------------------------

def cpu_pipeline(ora, q, r, seconds):
    """The same step on the CPU oracle for one pair; returns detected lag."""
    p = ora.default_params(algo_sample_rate=SR)
    ea = ora.fingerprint(q, p).short_time_energy
    eb = ora.fingerprint(r, p).short_time_energy
    max_lag = int(MAX_LAG_S * (seconds / 300.0) * SR) // HOP
    _, xs, _ = ora.align_xcorr(ea, eb, max_lag, HOP, SR)
    length = min(ea.size, eb.size) - max_lag
    a, b = trim_by_lag(ea, eb, xs.peak_lag, length)
    ora.dtw(a, b, band=DTW_BAND)
    return xs.peak_lag


def cpu_baseline_single(capi, synth):
    ora = capi.SonarLib(os.path.join(ROOT, "oracle", "libsonar_oracle.so"))
    seconds = 300.0
    q, r = make_pair(synth, seconds, 0)
    t0 = time.perf_counter()
    lag = cpu_pipeline(ora, q, r, seconds)
    dt = time.perf_counter() - t0
    return {"value": 2 * seconds / dt, "unit": UNIT, "cores": 1, "kind": "port", "detected_lag_frames": int(lag),
            "sample": f"1 pair of 2x{int(seconds)} s streams (fingerprint both + NCC +-60 s + DTW r=50), "
                      f"C++ -O2 oracle port of the Go path, 1 thread, {dt:.1f} s; the Go toolchain is absent so the "
                      "reference itself cannot run here",
            "alignments_per_s": 1.0 / dt}


def run_reference(args, rank):
    """--impl reference: the CPU path on every host thread, SAME configuration as the GPU arm (5-min pairs, +-60 s lag,
    DTW r=50); a step = one pair per host thread.  Rank 0 only."""
    if rank != 0:
        return
    p = pkg()
    capi, synth = p.capi, p.synth
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    seconds = float(args.seconds)
    n = int(round(seconds * SR))
    T = (n - WIN) // HOP + 1
    max_lag = int(MAX_LAG_S * SR) // HOP
    from concurrent.futures import ThreadPoolExecutor
    oras = [capi.SonarLib(os.path.join(ROOT, "oracle", "libsonar_oracle.so")) for _ in range(cores)]
    distinct = [make_pair(synth, seconds, i) for i in range(min(cores, 4))]
    pairs = [distinct[i % len(distinct)] for i in range(cores)]

    def one(i):
        return cpu_pipeline(oras[i], pairs[i][0], pairs[i][1], seconds)

    times = []
    with ThreadPoolExecutor(max_workers=cores) as ex:
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            lags = list(ex.map(one, range(cores)))
            if it >= args.warmup:
                times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = cores * 2 * seconds / (ms / 1e3)
    sample = (f"{cores} pairs of 2x{int(seconds)} s streams per step, one pair per host thread ({cores} threads; ctypes "
              f"releases the GIL), +-{int(MAX_LAG_S)} s lag ({2 * max_lag + 1} lags), DTW r={DTW_BAND}: the GPU arm's "
              "configuration; C++ -O2 oracle port of the Go path (no Go toolchain on the box, so the reference itself "
              "cannot run)")
    cfg = workload_config(args.pairs, seconds, T, max_lag, 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "alignments_per_s": cores / (ms / 1e3), "detected_lags_frames": lags[:4],
    }
    print(json.dumps(line), file=_OUT, flush=True)


# --------------------------------------------------------------------------------------
# Extra legs: the other BASELINE.json configurations (C1, C3, C4, C5) at their named sizes
# --------------------------------------------------------------------------------------

def _roof_from_profile(kern, frames, bytes_per_frame, peak):
    k_ms, k_n = kern.get("stft_features_kernel", (0.0, 0))
    if not k_n:
        return None
    achieved = frames * bytes_per_frame / (k_ms / 1e3) / 1e9
    return {"kernel": "stft_features_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "algorithmic_bytes_per_frame": bytes_per_frame, "kernel_ms_total": k_ms,
            "launches_timed": k_n}


def extra_legs(lib, capi, synth, torch, ext, barrier, reduce_max, rank, world, peak):
    """C1 (30 s, 1024/256), C3 (1 h @16 kHz, 512/160, 40 mel), C4 (4,096 x 60 s sharded over the ranks) and C5
    (1,024 x 10-min pairs sharded over the ranks), each with the PCM resident in HBM (CUDA events) and, where the bytes
    fit the time budget, through the host-pointer C ABI.  Inputs are replicated synthetic streams: C4 / C5 re-use one
    resident chunk (their 86.7 GB / 433 GB of float64 PCM do not fit HBM at once), every stream of a chunk is distinct."""
    out = {}

    def timed(fn, reps):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            fn()
        e1.record(ext)
        barrier()
        return reduce_max(e0.elapsed_time(e1) / reps)

    def profiled(fn):
        lib.profile_read()
        lib.profile_enable(True)
        fn()
        barrier()
        lib.profile_enable(False)
        return lib.profile_read()

    # ---- C1: one 30 s stream (latency of a single GenerateFingerprint) and 64 of them resident ----
    sr = 44100
    x1 = synth.sweep_noise(30.0, seed=1)
    n1 = x1.size
    p1 = lib.default_params(algo_sample_rate=sr, call_sample_rate=sr)
    L1 = lib.fp_dev_layout(p1, n1)
    S1 = 64
    st1 = (n1 + 1) & ~1
    d1 = torch.zeros((S1, st1), dtype=torch.float64, device="cuda")
    d1[:, :n1] = torch.from_numpy(x1).cuda()
    f1 = torch.empty(S1 * L1.total, dtype=torch.float64, device="cuda")
    ms = timed(lambda: lib.fingerprint_batch_dev(d1.data_ptr(), n1, st1, S1, p1, f1.data_ptr()), 5)
    t0 = time.perf_counter()
    for _ in range(3):
        lib.fingerprint(x1, p1)
    one_ms = 1e3 * (time.perf_counter() - t0) / 3
    out["c1"] = {"workload": "GenerateFingerprint, 30 s 44.1 kHz sweep+noise, 1024/256, music, fixed-sr mode",
                 "resident_64_streams": {"value": world * S1 * 30.0 / (ms / 1e3), "unit": UNIT, "ms": ms},
                 "single_stream_host_call": {"value": 30.0 / (one_ms / 1e3), "unit": UNIT, "ms": one_ms,
                                             "note": "one blocking sonar_fingerprint_f64 call from pageable host memory"}}
    del d1, f1

    # ---- C3: 1 h @ 16 kHz, 512/160, 40 mel / 13 MFCC ----
    sr3 = 16000
    x3 = synth.speech_band_noise(600.0, sr=sr3)          # 10 min generated, tiled to the hour
    x3 = np.tile(x3, 6)
    n3 = x3.size
    p3 = lib.default_params(window_size=512, hop_size=160, energy_frame=512, energy_hop=160, algo_sample_rate=sr3,
                            call_sample_rate=sr3, n_mel=40)
    L3 = lib.fp_dev_layout(p3, n3)
    S3 = 4
    st3 = (n3 + 1) & ~1
    d3 = torch.zeros((S3, st3), dtype=torch.float64, device="cuda")
    d3[:, :n3] = torch.from_numpy(x3).cuda()
    f3 = torch.empty(S3 * L3.total, dtype=torch.float64, device="cuda")
    run3 = lambda: lib.fingerprint_batch_dev(d3.data_ptr(), n3, st3, S3, p3, f3.data_ptr())
    ms = timed(run3, 3)
    T3 = (n3 - 512) // 160 + 1
    k3 = profiled(run3)
    t0 = time.perf_counter()
    lib.fingerprint(x3, p3)
    h_ms = 1e3 * (time.perf_counter() - t0)
    out["c3"] = {"workload": "speech/news fingerprint, 1 h 16 kHz speech-band noise, 512/160, 40-mel / 13-MFCC + flux",
                 "frames_per_stream": int(T3),
                 "resident_4_streams": {"value": world * S3 * 3600.0 / (ms / 1e3), "unit": UNIT, "ms": ms},
                 "single_stream_host_call": {"value": 3600.0 / (h_ms / 1e3), "unit": UNIT, "ms": h_ms},
                 "roofline": _roof_from_profile(k3, S3 * T3, 160 * 8 + 13 * 8, peak),
                 "kernels_ms": {k: v[0] for k, v in sorted(k3.items(), key=lambda kv: -kv[1][0])[:6]}}
    del d3, f3

    # ---- C4: 4,096 x 60 s streams, sharded round-robin over the ranks; chunks of 256 streams through HBM ----
    n4 = 60 * sr
    st4 = (n4 + 1) & ~1
    per_rank = 4096 // world
    CH = 256
    base = synth.sweep_noise(60.0, seed=100 + rank)
    d4 = torch.empty((CH, st4), dtype=torch.float64, device="cuda")
    row = torch.from_numpy(base).cuda()
    for i in range(CH):  # distinct streams: a different gain and a circular shift per stream
        d4[i, :n4] = torch.roll(row, 997 * i) * (0.5 + 0.5 * (i % 7) / 7.0)
    L4 = lib.fp_dev_layout(p1, n4)
    f4 = torch.empty(CH * L4.total, dtype=torch.float64, device="cuda")
    chunks = per_rank // CH

    def run4():
        for _ in range(chunks):
            lib.fingerprint_batch_dev(d4.data_ptr(), n4, st4, CH, p1, f4.data_ptr())
    ms = timed(run4, 1)
    T4 = (n4 - WIN) // HOP + 1
    k4 = profiled(lambda: lib.fingerprint_batch_dev(d4.data_ptr(), n4, st4, CH, p1, f4.data_ptr()))
    # host leg: the same chunk count from pinned host memory through sonar_fingerprint_batch_f64 (H2D + D2H inside)
    hostbuf = torch.empty((64, st4), dtype=torch.float64).pin_memory()
    hostbuf.copy_(d4[:64].cpu())
    hv = hostbuf.numpy()
    streams = [hv[i, :n4] for i in range(64)]
    bufs = lib.alloc_batch_outputs([n4] * 64, p1)
    lib.fingerprint_batch(streams, p1, buffers=bufs)
    barrier()
    t0 = time.perf_counter()
    reps4 = 4
    for _ in range(reps4):
        lib.fingerprint_batch(streams, p1, buffers=bufs)
    barrier()
    e_ms = reduce_max(1e3 * (time.perf_counter() - t0) / reps4)
    out["c4"] = {"workload": "batch fingerprinting of 4,096 x 60 s 44.1 kHz streams, sharded over the ranks (no collective)",
                 "streams_total": 4096, "streams_per_rank": per_rank, "chunk_streams": CH,
                 "resident": {"value": 4096 * 60.0 / (ms / 1e3), "unit": UNIT, "ms_whole_job": ms,
                              "note": "every rank fingerprints its 4096/N streams, 256 resident streams per chunk"},
                 "e2e_sample": {"value": world * 64 * 60.0 / (e_ms / 1e3), "unit": UNIT, "ms_per_64_streams": e_ms,
                                "h2d_bytes": int(64 * n4 * 8), "note": "64 streams per rank per call from pinned host memory "
                                "through sonar_fingerprint_batch_f64 (PCIe bound); the whole job moves 86.7 GB"},
                 "roofline": _roof_from_profile(k4, CH * T4, ALGO_BYTES_PER_FRAME, peak)}
    del d4, f4, hostbuf

    # ---- C5: 1,024 x 10-min pairs, +-60 s, DTW r=50, sharded by pair over the ranks; 8 resident pairs per chunk ----
    n5 = 600 * sr
    st5 = (n5 + 1) & ~1
    P5 = 32
    pairs_rank = 1024 // world
    d5 = torch.empty((2 * P5, st5), dtype=torch.float64, device="cuda")
    true_lags = []
    # the 32 resident pairs are windows of ONE long realisation of the C2 process per rank (envelope x noise): the CDN
    # copy starts at a_i, the source at a_i + offset_i, each CDN copy gets its own additive noise (built on the device:
    # generating 32 independent 10-min pairs with numpy took a minute of host time per rank)
    span = n5 + int(180 * sr)
    base5 = torch.from_numpy(synth.envelope_noise(span, sr, seed=300 + rank)).cuda()
    gen = torch.Generator(device="cuda").manual_seed(500 + rank)
    for i in range(P5):
        rng = np.random.default_rng(200 + rank * P5 + i)
        off = int(round(float(rng.uniform(-55.0, 55.0)) * sr))
        a_i = int(60 * sr) + int(rng.integers(0, 60 * sr))
        d5[2 * i, :n5] = base5[a_i + off: a_i + off + n5]
        d5[2 * i + 1, :n5] = base5[a_i: a_i + n5] + 0.02 * torch.randn(n5, dtype=torch.float64, device="cuda", generator=gen)
        true_lags.append(off / HOP)
    del base5
    bufs5 = lib.alloc_pair_outputs(P5, n5, p1, MAX_LAG_S, features=False, corr=False)
    res5 = [None]

    def run5():
        for _ in range(pairs_rank // P5):
            res5[0] = lib.align_pairs_dev(d5.data_ptr(), n5, st5, P5, p1, MAX_LAG_S, DTW_BAND, buffers=bufs5)
    ms = timed(run5, 1)
    lags5 = [int(x["xcorr"].peak_lag) for x in res5[0]]
    out["c5"] = {"workload": "batched alignment of 1,024 source/CDN pairs (10 min each, +-60 s lag, banded DTW r=50), "
                             "sharded by pair over the ranks (no collective)",
                 "pairs_total": 1024, "pairs_per_rank": pairs_rank, "chunk_pairs": P5,
                 "resident": {"alignments_per_s": 1024 / (ms / 1e3), "value": 1024 * 1200.0 / (ms / 1e3), "unit": UNIT,
                              "ms_whole_job": ms},
                 "detected_lags_frames": lags5, "true_offsets_frames": [round(v, 1) for v in true_lags],
                 "lags_within_one_frame_of_truth": bool(all(abs(a - b) <= 1.0 for a, b in zip(lags5, true_lags)))}
    del d5
    return out


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------

def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    p = pkg()
    capi, synth = p.capi, p.synth
    numa = None if args.no_numa else bind_to_gpu_numa(local_rank)  # before any pinned allocation (first touch)
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = capi.SonarLib(init=False)
    ids = (C.c_int * 1)(local_rank)
    lib._chk(lib.lib.sonar_init(1, ids, C.byref(lib.ctx)))
    assert lib.backend == "cuda-sm100a"
    ext = torch.cuda.ExternalStream(lib.stream(), device=torch.device("cuda", local_rank))

    P, seconds = args.pairs, float(args.seconds)
    n = int(round(seconds * SR))
    stride = (n + 1) & ~1
    NS = 2 * P
    prm = lib.default_params(algo_sample_rate=SR, call_sample_rate=SR)
    sz = lib.fp_sizes(prm, n)
    L = lib.fp_dev_layout(prm, n)
    T, Te = sz.n_frames, sz.n_energy_frames
    max_lag = int(MAX_LAG_S * SR) // HOP
    dtw_len = Te - max_lag

    # ---- synthetic inputs: pinned host PCM (e2e leg) and an HBM-resident copy (value leg) ----
    host = torch.empty((NS, stride), dtype=torch.float64).pin_memory()
    hv = host.numpy()
    for i in range(P):
        q, r = make_pair(synth, seconds, rank * P + i)
        hv[2 * i, :n], hv[2 * i + 1, :n] = q, r
    pcm_dev = host.to("cuda", non_blocking=False)
    torch.cuda.synchronize()

    # The public call for this workload is the chained pair pipeline (include/sonar.h: sonar_align_pairs_*):
    # fingerprint both streams -> "corr_energy" NCC -> trim by the detected lag -> banded DTW, per pair, with
    # nothing returning to the host in between.  Result buffers are caller-owned and reused every step.
    bufs_dev = lib.alloc_pair_outputs(P, n, prm, MAX_LAG_S, features=False, corr=False)
    # e2e returns what GenerateFingerprint + ExtractAlignmentFeatures return: every feature array, the alignment scalars
    # and the DTW path.  The correlation curve is internal to the reference (stats.CorrelationResult, never part of
    # AlignmentFeatures), so it is not requested.
    bufs_e2e = lib.alloc_pair_outputs(P, n, prm, MAX_LAG_S, features=True, corr=False)
    q_list = [hv[2 * i, :n] for i in range(P)]
    r_list = [hv[2 * i + 1, :n] for i in range(P)]

    def step_resident():  # PCM already in HBM; per-pair results (lag summary + DTW path) return to the host
        res = lib.align_pairs_dev(pcm_dev.data_ptr(), n, stride, P, prm, MAX_LAG_S, DTW_BAND, buffers=bufs_dev)
        return [x["xcorr"].peak_lag for x in res], res

    def step_e2e():  # host PCM in, every feature array + alignment scalars + DTW path out
        res = lib.align_pairs(q_list, r_list, prm, MAX_LAG_S, DTW_BAND, buffers=bufs_e2e)
        return [x["xcorr"].peak_lag for x in res], res, res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        lib.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- value leg: device-resident, CUDA events on the library's stream ----
    for _ in range(args.warmup):
        lags, paths = step_resident()
    lib.profile_read()
    barrier()
    sampler = ClockSampler(local_rank) if ((rank == 0 or world > 1) and not args.no_clocks) else None
    l0 = lib.kernel_launches()
    lib.profile_enable(False)  # the headline is timed WITHOUT the per-kernel event pairs
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(ext)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lags, paths = step_resident()
    ev1.record(ext)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    dev_ms = ev0.elapsed_time(ev1)
    launches = lib.kernel_launches() - l0
    clocks = sampler.stop() if sampler else None
    # per-kernel table and the roofline kernel's launch duration: the same steps once more, now with every launch
    # bracketed by a CUDA-event pair on its own stream (library side, sonar_profile_*), outside the headline timing
    kern, prof_steps = {}, 0
    if not args.no_profile:
        prof_steps = max(1, min(args.steps, 5))
        lib.profile_read()
        lib.profile_enable(True)
        for _ in range(prof_steps):
            step_resident()
        barrier()
        lib.profile_enable(False)
        kern = lib.profile_read()
    step_ms = reduce_max(dev_ms / args.steps)
    # N > 1: every rank's own step time, SM clock and kernel table (the weak-scaling loss at N = 8 had
    # no attribution -- there is no data-path collective, so a slower rank shows up here as a clock or a kernel)
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "ms_per_step": dev_ms / args.steps,
                "sm_mhz": (clocks or {}).get("sm_mhz"), "reasons": (clocks or {}).get("reasons"),
                "kernels_ms": {k: round(v[0] / max(prof_steps, 1), 3)
                               for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])[:6]}}
        allr = [None] * world
        dist.all_gather_object(allr, mine)
        per_rank = allr
    audio_s = world * NS * seconds
    value = audio_s / (step_ms / 1e3)

    # ---- extra leg: GenerateFingerprint alone on the same resident streams (the first half of the BASELINE metric,
    #      "audio-sec/s fingerprinted", without the alignment that `value` includes) ----
    feat_dev = torch.empty(NS * L.total, dtype=torch.float64, device="cuda")
    for _ in range(2):
        lib.fingerprint_batch_dev(pcm_dev.data_ptr(), n, stride, NS, prm, feat_dev.data_ptr())
    barrier()
    fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fe0.record(ext)
    for _ in range(3):
        lib.fingerprint_batch_dev(pcm_dev.data_ptr(), n, stride, NS, prm, feat_dev.data_ptr())
    fe1.record(ext)
    barrier()
    fp_ms = reduce_max(fe0.elapsed_time(fe1) / 3)
    # the same leg once more with the per-kernel event pairs on: here no alignment branch runs beside the fingerprint
    # kernels, so this is the STFT kernel's duration when it has the GPU to itself (roofline.alone)
    kern_fp = {}
    if not args.no_profile:
        lib.profile_read()
        lib.profile_enable(True)
        for _ in range(3):
            lib.fingerprint_batch_dev(pcm_dev.data_ptr(), n, stride, NS, prm, feat_dev.data_ptr())
        barrier()
        lib.profile_enable(False)
        kern_fp = lib.profile_read()
    exact_spectral, exact_pitch = lib.exact_counts()
    del feat_dev

    # ---- e2e leg: host buffers through the public C ABI (wall clock: the calls block the host) ----
    n_e2e = max(1, min(args.steps, 3))
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        lags_e, paths_e, fps = step_e2e()
    barrier()
    e2e_ms = reduce_max(1e3 * (time.perf_counter() - t0) / n_e2e)
    assert lags_e == lags, "host-pointer and device-resident legs disagree on the detected lags"
    h2d = NS * n * 8
    d2h = sum(a.nbytes for a in fps[0]["query"].arrays.values()) * NS + sum(len(pp["path_query"]) * 16 for pp in paths_e)

    # ---- extra legs: the SAME call from memory a Go caller actually holds.  `pageable`: ordinary heap
    #      arrays (a Go []float64): the runtime stages every copy through its own bounce buffer.  `registered`: the same
    #      arrays page-locked in place once with sonar_host_register (what go/sonargpu does for long-lived buffers), the
    #      registration itself outside the timed region.
    mem_legs = None
    if not args.no_s16:
        heap = np.empty((NS, n), dtype=np.float64)
        heap[:] = hv[:, :n]
        qh = [heap[2 * i] for i in range(P)]
        rh = [heap[2 * i + 1] for i in range(P)]

        def timed_heap():
            lib.align_pairs(qh, rh, prm, MAX_LAG_S, DTW_BAND, buffers=bufs_e2e)
            barrier()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                res = lib.align_pairs(qh, rh, prm, MAX_LAG_S, DTW_BAND, buffers=bufs_e2e)
            barrier()
            return reduce_max(1e3 * (time.perf_counter() - t0) / n_e2e), [x["xcorr"].peak_lag for x in res]

        pg_ms, pg_lags = timed_heap()
        t0 = time.perf_counter()
        lib.host_register(heap)
        reg_cost_ms = 1e3 * (time.perf_counter() - t0)
        try:
            rg_ms, rg_lags = timed_heap()
        finally:
            lib.host_unregister(heap)
        mem_legs = {"pageable": {"value": audio_s / (pg_ms / 1e3), "unit": UNIT, "ms_per_step": pg_ms,
                                 "lags_equal": bool(pg_lags == lags)},
                    "registered": {"value": audio_s / (rg_ms / 1e3), "unit": UNIT, "ms_per_step": rg_ms,
                                   "lags_equal": bool(rg_lags == lags), "one_time_register_ms": reg_cost_ms,
                                   "bytes_registered": int(heap.nbytes)},
                    "note": "e2e (the headline) reads pinned memory from sonar_host_alloc / torch pin_memory; these two "
                            "legs read a plain heap array, as a Go []float64 is, without and with sonar_host_register"}
        del heap, qh, rh

    # ---- extra leg: the same step with int16 PCM (what the decoder holds before the reference widens it to float64;
    #      sonar_align_pairs_pcm).  A quarter of the bytes cross PCIe; the samples are the float64 ones
    #      quantised to 16 bits, so this is reported beside `e2e`, never instead of it.
    s16 = None
    if not args.no_s16:
        host16 = torch.empty((NS, n), dtype=torch.int16).pin_memory()
        h16 = host16.numpy()
        for i in range(NS):
            h16[i] = np.clip(np.rint(hv[i, :n] * 8192.0), -32768, 32767).astype(np.int16)
        q16 = [h16[2 * i] for i in range(P)]
        r16 = [h16[2 * i + 1] for i in range(P)]

        def step_s16():
            res = lib.align_pairs_pcm(q16, r16, prm, MAX_LAG_S, DTW_BAND, buffers=bufs_e2e)
            return [x["xcorr"].peak_lag for x in res]

        step_s16()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            lags16 = step_s16()
        barrier()
        s16_ms = reduce_max(1e3 * (time.perf_counter() - t0) / n_e2e)
        s16 = {"value": audio_s / (s16_ms / 1e3), "unit": UNIT, "ms_per_step": s16_ms,
               "h2d_bytes_per_step": int(NS * n * 2), "d2h_bytes_per_step": int(d2h),
               "lags_equal_f64_leg": bool(lags16 == lags),
               "note": "int16 host PCM through sonar_align_pairs_pcm (widened to float64 on the device)"}

    # ---- N > 1 only: ONE long correlation (+-60 s) split by lag range over the ranks INSIDE the library: every rank
    #      z-scores, evaluates its lags in reference order, one ncclAllGather of the curve shards on the library's stream,
    #      peak analysis on every rank (csrc/nccl_shard.cu).  Reported beside the main metric, not inside it.
    lag_sharded = None
    if world > 1:
        dev = torch.device("cuda", local_rank)
        p.sharding.nccl_setup(lib, device=dev)
        rng = np.random.default_rng(7)  # identical on every rank: both sequences are replicated
        sweep = []
        for minutes, ml in ((2.5, max_lag), (5, max_lag), (10, max_lag), (20, max_lag), (40, max_lag), (20, 100000),
                            (40, 200000)):
            tn = (int(minutes * 60 * SR) - WIN) // HOP + 1
            base = np.convolve(rng.standard_normal(tn + 6000), np.ones(32) / 32, mode="same") + 1.0
            qa, rb = base[3000:3000 + tn].copy(), base[3000 - 2345:3000 - 2345 + tn] + 0.01 * rng.standard_normal(tn)
            lib.xcorr_lag_sharded(qa, rb, ml)  # warm-up
            barrier()
            t0 = time.perf_counter()
            reps = 5
            for _ in range(reps):
                sh_summ, _ = lib.xcorr_lag_sharded(qa, rb, ml)
            barrier()
            sh_ms = reduce_max(1e3 * (time.perf_counter() - t0) / reps)
            row = None
            if rank == 0:
                lib.xcorr(qa, rb, ml, want_corr=True)
                t0 = time.perf_counter()
                for _ in range(reps):
                    _, whole = lib.xcorr(qa, rb, ml, want_corr=True)
                exact_ms = 1e3 * (time.perf_counter() - t0) / reps
                t0 = time.perf_counter()
                for _ in range(reps):
                    _, scr = lib.xcorr(qa, rb, ml, want_corr=False)
                scr_ms = 1e3 * (time.perf_counter() - t0) / reps
                row = {"minutes": minutes, "frames": int(tn), "lags": 2 * min(ml, tn - 1) + 1, "sharded_ms": sh_ms,
                       "single_gpu_exact_curve_ms": exact_ms, "single_gpu_screened_ms": scr_ms,
                       "peak_lag": int(sh_summ.peak_lag),
                       "matches_unsharded": bool(sh_summ.peak_lag == whole.peak_lag and
                                                 sh_summ.peak_correlation == whole.peak_correlation and
                                                 sh_summ.second_peak == whole.second_peak and sh_summ.snr == whole.snr)}
            sweep.append(row)
        if rank == 0:
            ten = [r for r in sweep if r["minutes"] == 10][0]
            wide = [r for r in sweep if r["lags"] > 2 * max_lag + 1]
            cross = [(r["minutes"], r["lags"]) for r in sweep if r["sharded_ms"] < 0.9 * r["single_gpu_exact_curve_ms"]]
            lag_sharded = {"frames": ten["frames"], "lags": ten["lags"], "ranks": world, "ms": ten["sharded_ms"],
                           "single_gpu_ms": ten["single_gpu_exact_curve_ms"],
                           "single_gpu_screened_ms": ten["single_gpu_screened_ms"], "peak_lag": ten["peak_lag"],
                           "matches_unsharded": bool(all(r["matches_unsharded"] for r in sweep)),
                           "collectives": "one ncclAllGather of the curve shards (%d doubles per rank) on the library stream"
                                          % (-(-ten["lags"] // world)),
                           "sweep": sweep,
                           "configurations_where_sharding_wins_by_10pct": cross,
                           "wide_lag_speedup": [round(r["single_gpu_exact_curve_ms"] / r["sharded_ms"], 2) for r in wide],
                           "note": "bit-exact reference order makes BOTH phases dependent-add chains of the sequence "
                                   "length n (z-score: 2 n adds on one thread; every lag: its own sum over i), which no "
                                   "split of the lag range shortens: at +-60 s (20,671 lags = 65 CTAs) one GPU already "
                                   "runs every lag concurrently and the time is the chain.  Sharding pays once the lags "
                                   "exceed what one GPU runs concurrently (the two wide-lag rows); the screened "
                                   "single-GPU form is what the pair pipeline uses"}
        lib.nccl_shutdown()
    # ---- outside every timed region: the detected lags of the first pairs against the oracle.  The e2e leg returned
    #      the short-time energies (bit-exact with the oracle's: tests/test_gpu_fullsize.py); the oracle's own
    #      time-domain per-lag NCC over them must find the same lag index.  Pair 0 additionally runs the whole oracle
    #      pipeline from the PCM in the cpu_baseline leg below.
    lag_check = None
    if rank == 0 and not args.no_cpu:
        ora = capi.SonarLib(os.path.join(ROOT, "oracle", "libsonar_oracle.so"))
        idx = list(range(min(4, P)))
        ora_lags = []
        for i in idx:
            ea, eb = fps[i]["query"].short_time_energy, fps[i]["reference"].short_time_energy
            _, xs, _ = ora.align_xcorr(ea, eb, max_lag, HOP, SR)
            ora_lags.append(int(xs.peak_lag))
        lag_check = {"pairs": idx, "oracle_lags_frames": ora_lags, "gpu_lags_frames": [int(lags[i]) for i in idx],
                     "equal": bool(ora_lags == [int(lags[i]) for i in idx]),
                     "method": "oracle time-domain NCC (correlation.go:373-449 restated) over the returned energies"}
        assert lag_check["equal"], f"detected lags differ from the oracle: {lag_check}"
    legs = None
    if not args.no_legs:
        legs = extra_legs(lib, capi, synth, torch, ext, barrier, reduce_max, rank, world, measured_peak()[0])
    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms, k_n = kern.get("stft_features_kernel", (0.0, 0))
        roof = None
        if k_n:
            frames_per_launch = NS * T * prof_steps / k_n  # the library may split a step into chunks of pairs
            avg_ms = k_ms / k_n
            achieved = frames_per_launch * ALGO_BYTES_PER_FRAME / (avg_ms / 1e3) / 1e9
            tr = ncu_traffic()
            roof = {"kernel": "stft_features_kernel (fused framed STFT + MFCC + spectral descriptors: transform + scan kernel "
                              "pair, one timing pair around both)", "bound": "hbm",
                    "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    # ncu capture of the same kernel, scaled from its frames per launch to this run's
                    "traffic": (tr["dram_bytes_per_launch"] / tr["frames_per_launch"] * frames_per_launch) if tr else None,
                    "traffic_source": tr.get("source") if tr else None, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": frames_per_launch * ALGO_BYTES_PER_FRAME,
                    "avg_launch_ms": avg_ms, "launches_timed": k_n,
                    "timing": f"library-side CUDA-event pair around every launch on its own stream, over {prof_steps} "
                              "profiled step(s) run right after the timed region (the headline is timed without them); in "
                              "the step the alignment branch of the previous kernels runs beside this kernel on another "
                              "stream and shares the SMs with it"}
            a_ms, a_n = kern_fp.get("stft_features_kernel", (0.0, 0))
            if a_n:
                a_ach = frames_per_launch * ALGO_BYTES_PER_FRAME / (a_ms / a_n / 1e3) / 1e9
                roof["alone"] = {"achieved": a_ach, "frac": a_ach / peak, "avg_launch_ms": a_ms / a_n, "launches_timed": a_n,
                                 "note": "the same launch in the fingerprint-only leg (no alignment branch beside it)"}
        total_k = sum(v[0] for v in kern.values()) or 1.0
        shares = {k: {"ms_per_step": v[0] / max(prof_steps, 1), "launches_per_step": v[1] / max(prof_steps, 1),
                      "share": v[0] / total_k} for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])}
        cpu = cpu_baseline_single(capi, synth) if (world == 1 and not args.no_cpu) else None
        if cpu is not None and rank * P == 0:  # pair 0 of rank 0 is the pair the cpu_baseline leg ran end to end
            assert cpu["detected_lag_frames"] == int(lags[0]), (cpu["detected_lag_frames"], lags[0])
            lag_check["pair0_full_oracle_pipeline_equal"] = True
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(P, seconds, T, max_lag, world),
            "parallelism": f"pair-sharded x{world} (one process per GPU, no data-path collective)",
            "alignments_per_s": world * P / (step_ms / 1e3),
            "wall_ms_per_step": wall_ms / args.steps,
            "e2e": {"value": audio_s / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "timing": "wall clock around blocking C-ABI calls",
                    "alignments_per_s": world * P / (e2e_ms / 1e3)},
            "e2e_s16_ingest": s16,
            "fingerprint_only": {"value": audio_s / (fp_ms / 1e3), "unit": UNIT, "ms_per_step": fp_ms,
                                 "note": "GenerateFingerprint of the same resident streams without the alignment",
                                 "kernels_ms": {k: v[0] / max(v[1], 1) * (v[1] / 3.0) for k, v in
                                                sorted(kern_fp.items(), key=lambda kv: -kv[1][0])[:8]},
                                 "frames_reevaluated_in_float64": {"stft": exact_spectral, "of": int(NS * T),
                                                                   "yin": exact_pitch, "of_pitch_frames": int(NS * sz.n_pitch_frames)}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "kernels": shares,
            "cpu_baseline": cpu,
            "detected_lags_frames": lags,
            "lags_checked_vs_oracle": lag_check,
            "lag_sharded": lag_sharded,
            "e2e_host_memory": mem_legs,
            "numa_binding": numa,
            "per_rank": per_rank,
            "legs": legs,
        }
        print(json.dumps(line), file=_OUT, flush=True)
    lib.close()
    if world > 1:
        dist.destroy_process_group()


def _quiet_stdout():
    """stdout must carry exactly ONE JSON line; libraries (NCCL's version banner) write to fd 1 directly.
    Point fd 1 at stderr for the whole run and hand back a file object on the original stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def bind_to_gpu_numa(local_rank):
    """Pins this process (and the threads the library starts from it) to the CPUs of the NUMA node its GPU hangs off, so
    that the pinned staging buffers are first-touched on that node and the H2D copies do not cross the socket
    interconnect (all ranks sat on node 0).  Returns {"node": n, "cpus": k} or None when the topology
    is not exposed (single-socket box, container without sysfs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=32, help="source/CDN pairs per GPU per step")
    ap.add_argument("--seconds", type=float, default=300.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-s16", action="store_true", help="skip the int16-ingest extra leg")
    ap.add_argument("--no-clocks", action="store_true", help="do not sample nvidia-smi clocks (diagnostic)")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernels with CUDA events (diagnostic)")
    ap.add_argument("--no-legs", action="store_true", help="skip the extra C1 / C3 / C4 / C5 legs")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node (diagnostic)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    global _OUT
    _OUT = _quiet_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
