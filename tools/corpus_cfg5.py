#!/usr/bin/env python
"""BASELINE configs[4]: a 100k-clip synthetic 16 kHz corpus sharded across the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/corpus_cfg5.py [--clips 100000] [--batch 1024]

Clips are a contiguous block partition over ranks (modulation_mfcc_b200.shard); every rank runs
the whole path on its shard in batches (PCM synthesised on the device, batch by batch, outside
the timed region) and keeps the per-clip MFCC-change curve; the only collective is the final
gather of those curves.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import modulation_mfcc_b200 as mm

SR, SECONDS = 16000, 10.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=100_000)
    ap.add_argument("--batch", type=int, default=1024)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = mm.shard_range(a.clips, rank, world)
    n = int(SR * SECONDS)
    fx = mm.FeatureExtractor(SR, device=local, tStep=0.01, winLen=0.025, n_fft=512, n_mels=40, n_mfcc=13)
    T = fx.plan.num_frames(n)
    tot = torch.empty((hi - lo, T), device=dev, dtype=torch.float64)
    band = None
    compute_ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    warm = mm.synth_batch_device(min(a.batch, hi - lo), n, SR, seed=1, device=dev)
    fx(warm, want_logmel=False)  # plan workspace, kernel attributes, allocator pools
    del warm
    torch.cuda.synchronize()
    for b0 in range(lo, hi, a.batch):
        nb = min(a.batch, hi - b0)
        pcm = mm.synth_batch_device(nb, n, SR, seed=1234 + b0, device=dev)  # not timed: stands in for the loader
        e0.record()
        res = fx(pcm, want_logmel=False)
        tot[b0 - lo : b0 - lo + nb] = res["totChange"]
        e1.record()
        torch.cuda.synchronize()
        compute_ms += e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    full = mm.gather_features(tot, a.clips)
    torch.cuda.synchronize()
    gather_s = time.perf_counter() - t0
    t = torch.tensor([compute_ms, gather_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        compute_ms, gather_ms = float(t[0]), float(t[1])
        ok = bool(torch.isfinite(full).all()) and tuple(full.shape) == (a.clips, T)
        print(json.dumps({
            "workload": f"cfg5: {a.clips} x 10 s 16 kHz clips, block-partitioned over {world} GPU(s), batches of {a.batch}",
            "n_gpus": world, "clips_per_rank_max": (a.clips + world - 1) // world,
            "compute_ms_max_over_ranks": compute_ms, "final_gather_ms": gather_ms,
            "audio_s_per_s_compute": a.clips * SECONDS / (compute_ms * 1e-3),
            "audio_s_per_s_with_gather": a.clips * SECONDS / ((compute_ms + gather_ms) * 1e-3),
            "gathered_shape": list(full.shape), "gathered_finite": ok,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
