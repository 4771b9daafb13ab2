#!/usr/bin/env python
"""One large frame over the GPUs of one box by row bands (BASELINE.json configs[3]); launch with torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      scripts/run_bands.py --width 4096 --height 3072 --disp 384 [--check] [--reps 3]

Prints, on rank 0, the time of the banded frame (max over ranks, CUDA-synchronised wall clock around the whole banded call)
and, with --check, whether the map equals the one a single GPU computes (rank 0 runs that as well; needs the memory of the
whole frame on one GPU: 65 GB at 4096 x 3072 x 384)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sister_b200  # noqa: E402
from sister_b200.bands import (EngineBandWorker, as_uint16, compute_banded, connect_row_mailboxes, disconnect_row_mailboxes,  # noqa: E402
                               gather_band_rows)
from sister_b200.synth import make_rig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", dest="w", type=int, default=4096)
ap.add_argument("--height", dest="h", type=int, default=3072)
ap.add_argument("--disp", dest="d", type=int, default=384)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--check", action="store_true")
ap.add_argument("--no-stream-rows", action="store_true", help="hand the row sweeps' states over when a band has finished (sister_band_vertical)")
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
views = make_rig(a.w, a.h, a.d, seed=1234, channels=1)
hp = a.h + 2 * a.d
times = []
with sister_b200.Engine(a.w, a.h, a.d, n_slots=1, device=local) as eng:
    worker = EngineBandWorker(eng, views, a.d, rank, world, mode=a.mode)
    streamed = world > 1 and not a.no_stream_rows and connect_row_mailboxes(worker, world, rank)
    for rep in range(a.reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rows = compute_banded(worker, world, rank) if world > 1 else None
        if world == 1:
            from sister_b200.bands import run_bands_in_process
            rows = run_bands_in_process([worker])[0]
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rep > 0:
            times.append(float(dt.item()) * 1e3)
    full = gather_band_rows(rows, a.d, a.h, hp, dst=0) if world > 1 else rows
    ok = None
    single_ms = None
    if a.check and rank == 0:
        t0 = time.perf_counter()
        want = eng.compute(views, a.d, mode_mask=1 << a.mode)[a.mode]
        single_ms = (time.perf_counter() - t0) * 1e3
        ok = bool((as_uint16(full) == want).all())
        print("bands == single GPU:", ok)
    if getattr(worker, "stream_rows", False):
        disconnect_row_mailboxes(worker)
    if rank == 0:
        print(json.dumps({"what": "one frame by row bands", "row_sweeps_streamed": bool(streamed), "shape": [a.w, a.h, a.d], "n_gpus": world, "band_ms": times,
                          "single_gpu_host_call_ms": single_ms, "equal": ok}))
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok in (None, True) else 1)
