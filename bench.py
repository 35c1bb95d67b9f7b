#!/usr/bin/env python
"""Benchmark of the MFCC + modulation-spectrum hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): per GPU a batch of 1024 x 10 s 16 kHz mono
clips, 25 ms / 10 ms frames, 512-point FFT, 40 mel bands, 13 MFCC + delta +
MFCC-change curve + MFCC modulation spectrum.  One step = one pass of the whole
path over the batch.  ``value`` is audio-seconds per second with the PCM already
resident in HBM; ``e2e`` is the same work through the host-buffer C-ABI call
(pinned host PCM in, host features out, copies inside the timed region).

``--impl reference`` times the reference path's CPU arithmetic (the numpy/scipy
oracle -- librosa itself is not installable here) on all host
cores for the same metric and config, on a bounded sample of the workload.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 16000
SECONDS = 10.0
N_SAMPLES = int(SR * SECONDS)
CLIPS = 1024
PARAMS = dict(tStep=0.01, winLen=0.025, n_fft=512, n_mels=40, n_mfcc=13)
METRIC = "audio-seconds/sec (MFCC+modulation spectrum)"
UNIT = "audio-s/s"
WORKLOAD = (
    "BASELINE configs[1]: 1024 x 10 s 16 kHz clips per GPU; win 400 / hop 160 / n_fft 512, 40 mel, "
    "13 MFCC + delta + MFCC-change (zero-phase Butterworth, gradient, norm) + modulation spectrum (1 s windows, 0.5 s hop)"
)


# --------------------------------------------------------------------------- CPU arm


def _cpu_init():
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(1)
    except Exception:
        pass


_CPU_BATCH = None


def _cpu_one(i):
    import oracle

    f = oracle.mfcc_features(_CPU_BATCH[i], SR, **PARAMS)
    return float(f["totChange"][0])


def cpu_throughput(n_clips: int, repeats: int = 1, warmup: int = 0):
    """audio-s/s of the oracle over ``n_clips`` synthetic clips on all host cores."""
    global _CPU_BATCH
    import multiprocessing as mp

    from modulation_mfcc_b200.synth import synth_batch

    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    _CPU_BATCH = synth_batch(0, n_clips, N_SAMPLES, SR)
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        chunk = max(1, n_clips // (cores * 4))
        for r in range(warmup + repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_one, range(n_clips), chunksize=chunk)
            dt = time.perf_counter() - t0
            if r >= warmup:
                times.append(dt)
    _CPU_BATCH = None
    return [n_clips * SECONDS / t for t in times], times, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cores = os.cpu_count() or 1
    n_clips = CLIPS  # the full batch of one step (about 20 core-seconds of numpy/scipy work per step)
    vals, times, cores = cpu_throughput(n_clips, repeats=args.steps, warmup=args.warmup)
    total = sum(times)
    value = n_clips * SECONDS * len(times) / total
    sample = f"all {n_clips} clips of the workload per step (oracle.mfcc_features per clip, fork pool over all cores, BLAS threads = 1)"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "clips_per_step": n_clips, "note": "CPU arm: the numpy/scipy restatement of the reference path (the reference itself cannot be installed offline: DESIGN.md section 2)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


# --------------------------------------------------------------------------- clocks


class ClockSampler(threading.Thread):
    """Polls NVML for SM clock and throttle reasons while the timed region runs."""

    def __init__(self, index: int, period: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------- GPU arm


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import modulation_mfcc_b200 as mm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    lib = mm.lib()
    fx = mm.FeatureExtractor(SR, device=local, flags=int(os.environ.get("MMF_FLAGS", "0")), **PARAMS)
    plan, prm = fx.plan, fx.prm
    pcm = mm.synth_batch_device(CLIPS, N_SAMPLES, SR, seed=1234 + rank, device=dev)
    T = plan.num_frames(N_SAMPLES)
    Lw, Hw, nfft, bins = fx.modspec_geometry(T)
    # the only collective (north_star): ONE final gather of the per-clip feature of every step, issued
    # after the last step and inside the timed region.  Each step parks its totChange in a slice of
    # `local_feats`; nothing communicates while the persistent kernels own the SMs.
    local_feats = gathered = None
    if world > 1:
        n_keep = max(args.steps, args.warmup, 3)
        local_feats = torch.empty((n_keep, CLIPS, T), device=dev, dtype=torch.float64)
        gathered = torch.empty((world, n_keep, CLIPS, T), device=dev, dtype=torch.float64)

    k1_events = []

    def step(record: bool):
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        lm, cmax = plan.logmel(pcm)  # fused frame/window/rFFT/|X|^2/mel/log kernel (+ clip-max init)
        if record:
            e1.record()
            k1_events.append((e0, e1))
        res = plan.change_from_logmel(lm, cmax, prm, clamp_in_place=False)  # clamp+DCT+delta, IIR, derivative+norm, IIR
        mag, band = plan.modspec(res["mfcc"], Lw, Hw, nfft, bins)
        if world > 1:
            local_feats[step.count % local_feats.shape[0]].copy_(res["totChange"])
            step.count += 1
        return res, mag, band

    step.count = 0

    def final_gather(n_steps):
        if world > 1:
            n = min(n_steps, local_feats.shape[0])
            dist.all_gather_into_tensor(gathered[:, :n].contiguous() if n < local_feats.shape[0] else gathered,
                                        local_feats[:n].contiguous() if n < local_feats.shape[0] else local_feats)

    for _ in range(max(args.warmup, 3)):
        step(False)
    final_gather(max(args.warmup, 3))
    step.count = 0
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    lib.mmf_launch_count(1)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(True)
    final_gather(args.steps)  # the gather of all steps' features lands before the clock stops
    ev1.record()
    torch.cuda.synchronize()
    barrier()
    launches = int(lib.mmf_launch_count(0))
    ms_total = ev0.elapsed_time(ev1)
    k1_ms = [a.elapsed_time(b) for a, b in k1_events]
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * CLIPS * SECONDS / (ms_step * 1e-3)

    # ---- end to end through the host-buffer C-ABI call (pinned host PCM in, host features out)
    pcm_host_t = torch.empty((CLIPS, N_SAMPLES), dtype=torch.float32, pin_memory=True)
    pcm_host_t.copy_(pcm)
    torch.cuda.synchronize()
    pcm_host = pcm_host_t.numpy()
    want = ("totChange", "mfcc", "delta", "modspec", "band_energy")
    n_win = 1 + (T - Lw) // Hw
    shapes = {
        "totChange": ((CLIPS, T), torch.float64),
        "mfcc": ((CLIPS, 13, T), torch.float32),
        "delta": ((CLIPS, 13, T), torch.float32),
        "modspec": ((CLIPS, 13, n_win, nfft // 2 + 1), torch.float32),
        "band_energy": ((CLIPS, n_win, len(bins)), torch.float32),
    }
    pinned = {k: torch.empty(s, dtype=d, pin_memory=True) for k, (s, d) in shapes.items()}
    out = {k: v.numpy() for k, v in pinned.items()}
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        fx.host_call(pcm_host, want=want, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fx.host_call(pcm_host, want=want, out=out)  # synchronous: returns when the results are in host memory
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    barrier()
    clocks = sampler.stop()
    e2e_value = world * CLIPS * SECONDS * e2e_steps / e2e_s
    # extra (not the contract's `e2e`): the same call fed with 16-bit PCM as it sits in a WAV file --
    # half the H2D bytes, scaled to float32 on the device; inputs are the float batch quantised to int16
    pcm16_t = torch.empty((CLIPS, N_SAMPLES), dtype=torch.int16, pin_memory=True)
    pcm16_t.copy_(torch.clamp(torch.round(pcm * 32768.0), -32768, 32767).to(torch.int16))
    torch.cuda.synchronize()
    pcm16_host = pcm16_t.numpy()
    for _ in range(2):
        fx.host_call(pcm16_host, want=want, out=out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fx.host_call(pcm16_host, want=want, out=out)
    e16_s = time.perf_counter() - t0
    te = torch.tensor([e16_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e16_s = float(te.item())
    barrier()
    e2e16_value = world * CLIPS * SECONDS * e2e_steps / e16_s
    h2d = CLIPS * N_SAMPLES * 4
    d2h = sum(int(np.prod(s)) * (8 if d == torch.float64 else 4) for s, d in shapes.values())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # dominant kernel: fused STFT/mel.  Algorithmic bytes per launch = PCM in + log-mel out.
    alg_bytes = CLIPS * (4 * N_SAMPLES + 4 * PARAMS["n_mels"] * T)
    k1 = statistics.mean(k1_ms)
    achieved = alg_bytes / (k1 * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "stft_mel_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu:
        n_cpu = CLIPS  # the whole batch of one step: ~20 core-seconds of numpy/scipy work
        vals, times, cores = cpu_throughput(n_cpu, repeats=1, warmup=0)
        cpu = {
            "value": vals[0],
            "unit": UNIT,
            "cores": cores,
            "kind": "port",
            "sample": f"{n_cpu} clips of the same workload (oracle.mfcc_features per clip, fork pool over all {cores} cores, BLAS threads = 1), {times[0]:.2f} s wall",
        }

    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD,
            "clips_per_gpu": CLIPS,
            "frames_per_clip": T,
            "l2": "inputs are 655 MB per step per GPU, larger than the 126 MB L2 (no flush needed)",
            "collective": "one final all_gather_into_tensor of every step's totChange, inside the timed region" if world > 1 else "none",
            "fp64_stages": "zero-phase Butterworth and derivative/norm run in f64; the modulation spectrum is a tcgen05 GEMM (fp16 operand pairs, f32 accumulate)",
        },
        "clocks": clocks,
        "e2e": {
            "value": e2e_value,
            "unit": UNIT,
            "h2d_bytes_per_step": h2d,
            "d2h_bytes_per_step": d2h,
            "steps": e2e_steps,
            "call": "mmf_features_host (one C-ABI call per step, pinned host buffers)",
        },
        "e2e_pcm16": {
            "value": e2e16_value,
            "unit": UNIT,
            "h2d_bytes_per_step": CLIPS * N_SAMPLES * 2,
            "d2h_bytes_per_step": d2h,
            "steps": e2e_steps,
            "call": "mmf_features_host_pcm16 (int16 WAV samples in, scaled on the device; extra, not the headline)",
        },
        "gpu_launches": launches,
        "roofline": {
            "kernel": "stft_mel_kernel<512, c2> (fused frame/window/rFFT/|X|^2/mel/log)",
            "bound": "hbm",
            "achieved": achieved,
            "peak": hbm_peak,
            "unit": "GB/s",
            "frac": achieved / hbm_peak,
            "traffic": traffic,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes,
            "kernel_ms": k1,
            "kernel_share_of_step": k1 / ms_step,
        },
        "cpu_baseline": cpu,
    }
    if world > 1:
        dist.destroy_process_group()
    return line


class _StdoutToStderr:
    """Everything written to fd 1 while the bench runs (NCCL's version banner, library chatter)
    goes to stderr, so that stdout carries exactly one line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    with _StdoutToStderr():
        line = run_reference(args) if args.impl == "reference" else run_b200(args)
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
