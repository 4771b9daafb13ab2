#!/usr/bin/env python
"""bench.py -- throughput of the 5-view disparity hot path (BASELINE.json metric: frames/s and Gcost-evals/s).

  python bench.py --gpus N --steps K --warmup W            our sm_100a path (one process per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    the reference's own CPU implementation (oracle/_ref)
  python bench.py --workload {c2,c1,c3,c4-bands,sweep}     the other shapes BASELINE.json names (default c2)

A "step" is one pass of the hot path over one batch of `--rigs-per-step` synthetic rigs per GPU; a frame is one rig ->
its multiview disparity map (doMultiStereo mode 0). Rigs are independent, so ranks shard by frame
with no data-path collective (weak scaling); for N > 1 the disparity maps of the previous step are gathered to rank 0 over
NCCL while the next step computes.

  value   frames/s with the rigs already resident in HBM (sister_submit_device), device-timed with CUDA events on the
          library's own streams (sister_region_begin/end), max over ranks.
  e2e     frames/s through the public host-buffer API (sister_compute_batch): every step copies its BGR inputs
          host->device from pinned staging and reads its disparity maps back, inside the timed region.

Workloads (BASELINE.json configs): c2 = configs[1], 1280x960 D=192 (the headline, default); c1 = configs[0], 640x480 D=192;
c3 = configs[2], a fixed batch of 256 c2 rigs split over the ranks (strong scaling); c4-bands = configs[3], one
4096x3072 D=384 frame by row bands over the ranks (one GPU: the whole frame); sweep = configs[4], 1280x960 at D = 128..512.
The default line carries the other shapes as `extras` (N = 1) and the band run as `bands_c4` (N > 1), each bounded.

PyTorch is used only for torch.distributed / NCCL plumbing and for the device tensors that hold inputs and outputs.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from sister_b200.synth import cost_evals, make_rig  # noqa: E402

SHAPES = {"c1": (640, 480, 192), "c2": (1280, 960, 192), "c3": (1280, 960, 192), "c4-bands": (4096, 3072, 384)}
SWEEP_D = (128, 256, 384, 512)


def workload_name(w, h, d, what="multiview map (mode 0)"):
    tag = {(1280, 960, 192): " (BASELINE.json configs[1])", (640, 480, 192): " (BASELINE.json configs[0])",
           (4096, 3072, 384): " (BASELINE.json configs[3])"}.get((w, h, d), "")
    return f"synthetic 5-view rig {w}x{h}, max disparity {d}{tag}, {what}"


def config_of(w, h, d):
    """The workload description both arms print verbatim (the driver compares the two dicts)."""
    return {"workload": workload_name(w, h, d), "width": w, "height": h, "disp_count": d, "mode": "multiview (doMultiStereo mode 0)",
            "evals_per_frame": cost_evals(w, h, d), "evals_per_frame_nominal": cost_evals(w, h, d, padded=False),
            "padded_cells_per_frame": (w + 2 * d) * (h + 2 * d) * d,
            "l2": "per-frame working set (fused volume + 4 pair volumes, 2.1 GB at 1280x960x192) exceeds the 126 MB L2; no flush needed"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def source_hash():
    """sha256 over the CUDA sources: ties profiles/*_traffic.json (an ncu capture) to the build it was taken from."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "sister_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            h.update(name.encode())
            h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()[:16]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


# ---------------------------------------------------------------------------------------------- the reference arm

def reference_objects():
    import oracle
    try:
        return oracle.Ref(), oracle.Oracle(), "reference"
    except FileNotFoundError:
        return None, oracle.Oracle(), "port"


def reference_mode0(ref, orc, pads, d):
    """One doMultiStereo(mode 0) on the padded frame (hpp:152-295) via the reference's own functions; returns the padded
    float disparity."""
    if ref is not None:
        return ref.multistereo_taps(pads, d, mode=0, want_volumes=False)["disp"]
    return orc.multistereo(pads, d, 0, want_volumes=False)["disp"]


def run_reference(args):
    """The reference's own CPU path (oracle/_ref = unmodified sources compiled in place), timed on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w, h, d = SHAPES.get(args.workload, SHAPES["c2"])
    ref, orc, kind = reference_objects()
    views = make_rig(w, h, d, seed=1234, channels=3)
    pads = [orc.pad_replicate(orc.grey_bgr(v), d) for v in views]  # staging is <1% of the reference's time
    for _ in range(args.warmup):
        reference_mode0(ref, orc, pads, d)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        reference_mode0(ref, orc, pads, d)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    cores = 4 if ref is not None else 1  # census.cpp:117: 4 OpenMP sections in hammingCost, 1 thread elsewhere
    line = {
        "impl": "reference", "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "gpus_used": 0, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "gcost_evals_per_s": fps * cost_evals(w, h, d) / 1e9, "gcost_evals_nominal_per_s": fps * cost_evals(w, h, d, padded=False) / 1e9,
        "config": config_of(w, h, d),
        "run": {"rigs_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "host_cores": os.cpu_count(),
                         "sample": "1 rig per step, one doMultiStereo(mode 0) on the padded frame (hpp:152-295) via the L2 functions"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


def cpu_baseline_leg(w, h, d, gpu_map, budget_s: float = 25.0):
    """Bounded CPU sample on the GPU box's host cores: the reference (or the oracle port) on ONE rig of the workload (seed
    1234), and -- since the reference map is there anyway -- the parity of the GPU's map of the same rig."""
    ref, orc, kind = reference_objects()
    views = make_rig(w, h, d, seed=1234, channels=3)
    pads = [orc.pad_replicate(orc.grey_bgr(v), d) for v in views]
    t0 = time.perf_counter()
    n = 0
    disp = None
    while True:
        disp = reference_mode0(ref, orc, pads, d)
        n += 1
        if time.perf_counter() - t0 > budget_s * 0.5 or n >= 3:
            break
    dt = time.perf_counter() - t0
    parity = None
    if gpu_map is not None:
        parity = bool((orc.encode_crop(disp, d) == gpu_map).all())
    return {"value": n / dt, "unit": "frames/s", "cores": 4 if ref is not None else 1, "kind": kind, "host_cores": os.cpu_count(),
            "sample": f"{n} rig(s) of the same workload, doMultiStereo mode 0 on the padded frame, {dt:.1f} s wall",
            "gcost_evals_per_s": n / dt * cost_evals(w, h, d) / 1e9, "parity_checked": parity}


# The contract is ONE JSON line on stdout. Libraries write there too (NCCL prints its version banner to fd 1 under
# torchrun), so fd 1 is pointed at stderr for the whole run and the line goes to the original stdout.
_REAL_STDOUT = None


def claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


# ---------------------------------------------------------------------------------------------- our arm

class Dist:
    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (sister_b200 has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def device_rate(ctx: Dist, eng, sb, w, h, d, rigs_t, out_t, B, S, steps, warmup, mode_mask, gather_total=None):
    """Device-resident, device-timed rate of `steps` steps of B rigs on this rank (rigs / outputs are torch tensors).
    gather_total: with N > 1, the maps of step k - 1 travel to rank 0 (NCCL, asynchronous) while step k computes; out_t is
    then double-buffered [2, B, H, W]. Returns (ms over the timed region, max over ranks; launches)."""
    torch = ctx.torch
    from sister_b200.sharding import gather_maps
    nmaps = bin(mode_mask).count("1")

    def submit(buf):
        for k in range(B):
            ptrs = [0, 0, 0]
            j = 0
            for m in range(3):
                if (mode_mask >> m) & 1:
                    ptrs[m] = buf[k, j].data_ptr()
                    j += 1
            eng.submit_device(k % S, rigs_t[k].data_ptr(), w, h, 3, d, mode_mask, ptrs)

    pending = [None]

    def step(i):
        buf = out_t[i & 1]
        submit(buf)
        if gather_total is not None:
            if pending[0] is not None:
                pending[0]()            # the previous step's maps: complete on the device since the sync below
            eng.sync()                  # this step's kernels (the gather above ran beside them)
            pending[0] = gather_maps(buf[:, 0], gather_total, dst=0, async_op=True)

    for i in range(warmup):
        step(i)
    if pending[0] is not None:
        pending[0](); pending[0] = None
    ctx.barrier()
    launches0 = eng.launch_count()
    t_wall0 = time.perf_counter()
    eng.region_begin()
    for i in range(steps):
        step(i)
    ms = eng.region_end()
    if pending[0] is not None:
        pending[0](); pending[0] = None
        torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    ctx.barrier()
    # with the NCCL gather in the loop the device bracket misses the collective: use the larger of the two clocks
    if gather_total is not None:
        ms = max(ms, t_wall * 1e3)
    return ctx.max(ms), eng.launch_count() - launches0, nmaps


def e2e_rate(ctx: Dist, eng, sb, rigs_bgr, w, h, d, B, steps, mode_mask):
    """Host buffers through sister_compute_batch: H2D of every view and D2H of every map inside the timed region."""
    nm = [(mode_mask >> m) & 1 for m in range(3)]
    outs = [[np.zeros((h, w), np.uint16) if nm[m] else None for m in range(3)] for _ in range(B)]
    pinned = eng.host_array((B, 5, h, w, 3), np.uint8)   # what a capture pipeline hands over: page-locked
    for k in range(B):
        for v in range(5):
            pinned[k, v] = rigs_bgr[k][v]
    rigs = [[pinned[k, v] for v in range(5)] for k in range(B)]
    for _ in range(2):
        eng.compute_batch(rigs, d, mode_mask, outs=outs)
    ctx.barrier()
    t0 = time.perf_counter()
    checksum = 0
    first = next(m for m in range(3) if nm[m])
    for _ in range(steps):
        eng.compute_batch(rigs, d, mode_mask, outs=outs)
        checksum = int(outs[0][first][::97, ::89].sum())  # the step's result is read on the host
    dt = ctx.max(time.perf_counter() - t0)
    return dt, outs, checksum, B * 5 * w * h * 3, B * sum(nm) * w * h * 2


def single_rig_stages(eng, sb, rig_ptr, out_ptr, w, h, d, n=4, mode_mask=1):
    """Per-stage CUDA-event times of one rig alone on the GPU (median of n after one warm-up); out_ptr: room for the maps of
    the requested modes, back to back."""
    eng.sync()
    eng.set_profiling(True)
    acc = {k: [] for k in sb.STAGE_NAMES}
    ptrs, j = [0, 0, 0], 0
    for m in range(3):
        if (mode_mask >> m) & 1:
            ptrs[m] = out_ptr + j * w * h * 2
            j += 1
    for _ in range(n + 1):
        eng.submit_device(0, rig_ptr, w, h, 3, d, mode_mask, ptrs)
        eng.sync(0)
        for k, v in eng.stage_ms(0).items():
            acc[k].append(v)
    launches = eng.stage_launches(0)
    eng.set_profiling(False)
    return {k: statistics.median(v[1:]) for k, v in acc.items()}, launches


def roofline_of(agg_ms, cells, launches, note, traffic_key="aggregation_dram_bytes_per_frame"):
    peak, peak_src = load_peaks()
    algo = 8 * cells  # read C twice, write S once, read S once, per padded cell, uint16 volumes
    achieved = algo / (agg_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes of the same launch group from a committed ncu --set full capture (per frame, like `achieved`)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if tj.get("source_hash") == source_hash():
            traffic = tj[traffic_key]
            traffic_src = "profiles/r02_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, same sources)"
        else:
            traffic_src = "profiles/r02_traffic.json is from other kernel sources (hash mismatch): not quoted"
    except Exception:
        pass
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src, "traffic_gbs": (traffic / (agg_ms * 1e-3) / 1e9) if traffic else None,
            "kernel": "aggregation passes (SGM, sgm.cpp:26-455): k_sgm_sweeps + k_sgm_final", "algorithmic_bytes_per_launch_group": algo,
            "launches_in_group": launches, "duration_ms": agg_ms, "peak_source": peak_src, "how": note}


def run_frames(ctx: Dist, args, w, h, d, extras_ok=True):
    """The frame-sharded workloads (c1, c2, c3): returns the JSON line."""
    import sister_b200 as sb
    torch = ctx.torch
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    strong = args.workload == "c3"
    B = args.rigs_per_step if not strong else max(1, 256 // world)
    S = min(args.slots, B)
    eng = sb.Engine(w, h, d, n_slots=S, device=ctx.local)
    cfg = config_of(w, h, d)
    evals, cells = cfg["evals_per_frame"], cfg["padded_cells_per_frame"]
    # ---- synthetic rigs: the job is world * B rigs, sharded by frame (sister_b200/sharding.py); seeds follow the global
    # rig index; c3 cycles 16 distinct rigs per rank to bound the generation time
    from sister_b200.sharding import frame_shard
    g0, g1 = frame_shard(world * B, world, rank)
    distinct = min(B, 16)
    rigs_bgr = [make_rig(w, h, d, seed=1234 + g0 + k, channels=3) for k in range(distinct)]
    base_t = [torch.from_numpy(np.stack(r)).to(dev) for r in rigs_bgr]          # [5, H, W, 3] uint8 each
    rigs_t = [base_t[k % distinct] for k in range(B)]
    out_t = torch.zeros((2, B, 1, h, w), dtype=torch.int16, device=dev)         # uint16 bit patterns (NCCL has no u16)
    torch.cuda.synchronize()
    sampler = ClockSampler(ctx.local)
    sampler.start()
    ms, launches, _ = device_rate(ctx, eng, sb, w, h, d, rigs_t, out_t, B, S, args.steps, max(args.warmup, 3), sb.MODE_MULTIVIEW,
                                  gather_total=world * B if world > 1 else None)
    clocks = sampler.stop()
    fps = world * B * args.steps / (ms * 1e-3)
    line = {
        "metric": "frames_per_s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "u16", "data": "synthetic",
        "gcost_evals_per_s": fps * evals / 1e9, "gcost_evals_nominal_per_s": fps * cfg["evals_per_frame_nominal"] / 1e9,
        "config": cfg,
        "run": {"workload_key": args.workload, "rigs_per_step_per_gpu": B, "rigs_in_flight_per_gpu": S,
                "sharding": "by frame, no data-path collective" + ("; maps of step k-1 gathered to rank 0 over NCCL while step k computes" if world > 1 else "")},
        "clocks": clocks, "gpu_launches": int(launches),
    }
    # ---- e2e: host buffers through the public API, copies inside the timed region
    gpu_map_seed1234 = None
    if not args.no_e2e:
        Be = min(B, 16)
        dt, outs, checksum, h2d, d2h = e2e_rate(ctx, eng, sb, [rigs_bgr[k % distinct] for k in range(Be)], w, h, d, Be, args.steps, sb.MODE_MULTIVIEW)
        same = bool((torch.from_numpy(outs[0][0].astype(np.int32)).to(dev) == (out_t[(args.steps - 1) & 1, 0, 0].to(torch.int32) & 0xFFFF)).all().item())
        line["e2e"] = {"value": world * Be * args.steps / dt, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "rigs_per_step_per_gpu": Be, "input": "BGR uint8 (cv::imread layout), page-locked host memory", "checksum": checksum,
                       "matches_device_path": same}
        if rank == 0:
            gpu_map_seed1234 = outs[0][0].copy()
    # ---- roofline of the dominant kernel group (aggregation passes), one rig in flight, CUDA events per stage
    stages, stage_launches = single_rig_stages(eng, sb, rigs_t[0].data_ptr(), out_t[0, 0].data_ptr(), w, h, d)
    line["roofline"] = roofline_of(stages["aggregate"], cells, stage_launches["aggregate"],
                                   "CUDA events on the slot stream around the aggregation kernels, 1 rig in flight, median of 4; crop-only aggregation (the product path)")
    line["stage_ms_single_rig"] = stages
    line["single_rig_latency_ms"] = sum(stages.values())
    if rank == 0 and world == 1 and extras_ok and not args.no_extras:
        line["extras"] = extras(ctx, args, eng, sb, rigs_bgr, rigs_t, w, h, d, cells)
    eng.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"] = cpu_baseline_leg(w, h, d, gpu_map_seed1234)
        except Exception as ex:  # the baseline is informative; never lose the GPU line to it
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "unavailable", "sample": repr(ex)}
    return line


def extras(ctx: Dist, args, eng, sb, rigs_bgr, rigs_t, w, h, d, cells):
    """N = 1 only: the other things the contract names, each bounded (whole block: well under two minutes)."""
    torch = ctx.torch
    dev = ctx.dev
    t_start = time.perf_counter()
    ex = {}
    B, S = len(rigs_t), eng.n_slots
    # (1) the call a drop-in user makes: all three maps (hpp:77-89), device-resident and end to end
    out3 = torch.zeros((2, B, 3, h, w), dtype=torch.int16, device=dev)
    ms, _, _ = device_rate(ctx, eng, sb, w, h, d, rigs_t, out3, B, S, max(2, args.steps // 2), 2, sb.MODE_ALL)
    calls = B * max(2, args.steps // 2) / (ms * 1e-3)
    Be = min(B, 8)
    dt, _, _, h2d, d2h = e2e_rate(ctx, eng, sb, [rigs_bgr[k % len(rigs_bgr)] for k in range(Be)], w, h, d, Be, max(2, args.steps // 2), sb.MODE_ALL)
    st3, _ = single_rig_stages(eng, sb, rigs_t[0].data_ptr(), out3[0, 0].data_ptr(), w, h, d, n=3, mode_mask=sb.MODE_ALL)
    ex["api_call_3maps"] = {"what": "compute_disparities as the reference class runs it: multiview + horizontal + vertical maps (hpp:26-119)",
                            "calls_per_s": calls, "e2e_calls_per_s": Be * max(2, args.steps // 2) / dt, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                            "single_call_latency_ms": sum(st3.values()), "stage_ms_single_call": st3}
    # (2) the whole padded frame (what a raw_disp caller gets; the product path aggregates the crop only)
    eng.set_full_frame(True)
    stf, lf = single_rig_stages(eng, sb, rigs_t[0].data_ptr(), out3[0, 0, 0].data_ptr(), w, h, d, n=3)
    eng.set_full_frame(False)
    ex["full_frame"] = {"aggregate_ms": stf["aggregate"], "roofline_frac": 8 * cells / (stf["aggregate"] * 1e-3) / 1e9 / load_peaks()[0],
                        "note": "aggregation of the whole padded frame (sister_set_full_frame): every cell of the 8 B/cell model is touched"}
    del out3
    # (3) the other named shapes: configs[0], the D sweep of configs[4], configs[3] on one GPU
    def one_shape(ww, hh, dd, n=3, slots=2, B2=4):
        if time.perf_counter() - t_start > 150:
            return {"skipped": "time budget of the extras"}
        views = make_rig(ww, hh, dd, seed=1234, channels=1)
        with sb.Engine(ww, hh, dd, n_slots=slots, device=ctx.local) as e2:
            rt = torch.from_numpy(np.stack(views)).to(dev)
            ot = torch.zeros((max(B2, 1), hh, ww), dtype=torch.int16, device=dev)
            for _ in range(2):
                e2.submit_device(0, rt.data_ptr(), ww, hh, 1, dd, sb.MODE_MULTIVIEW, [ot[0].data_ptr(), 0, 0]); e2.sync(0)
            e2.set_profiling(True)
            lat = []
            for _ in range(n):
                e2.submit_device(0, rt.data_ptr(), ww, hh, 1, dd, sb.MODE_MULTIVIEW, [ot[0].data_ptr(), 0, 0]); e2.sync(0)
                lat.append(e2.stage_ms(0))
            e2.set_profiling(False)
            med = {k: statistics.median(x[k] for x in lat) for k in lat[0]}
            res = {"single_rig_latency_ms": sum(med.values()), "aggregate_ms": med["aggregate"],
                   "roofline_frac": 8 * (ww + 2 * dd) * (hh + 2 * dd) * dd / (med["aggregate"] * 1e-3) / 1e9 / load_peaks()[0]}
            if B2 > 1:
                e2.region_begin()
                reps = 3
                for _ in range(reps):
                    for k in range(B2):
                        e2.submit_device(k % slots, rt.data_ptr(), ww, hh, 1, dd, sb.MODE_MULTIVIEW, [ot[k].data_ptr(), 0, 0])
                ms2 = e2.region_end()
                res["frames_per_s"] = reps * B2 / (ms2 * 1e-3)
                res["gcost_evals_per_s"] = res["frames_per_s"] * cost_evals(ww, hh, dd) / 1e9
        return res
    ex["c1_640x480_d192"] = one_shape(640, 480, 192, B2=8, slots=4)
    ex["sweep_1280x960"] = {f"d{dd}": one_shape(1280, 960, dd, B2=4, slots=2) for dd in SWEEP_D}
    ex["c4_4096x3072_d384_one_gpu"] = one_shape(4096, 3072, 384, n=2, slots=1, B2=1)
    ex["seconds"] = time.perf_counter() - t_start
    return ex


def run_bands(ctx: Dist, args, reps=2, check=True):
    """BASELINE.json configs[3]: ONE 4096x3072 D=384 frame by row bands over the ranks (sister_b200/bands.py): the column /
    diagonal state crosses the band borders exactly, over NCCL. Returns the dict for the line."""
    import sister_b200 as sb
    from sister_b200.bands import (EngineBandWorker, as_uint16, compute_banded, connect_row_mailboxes, disconnect_row_mailboxes, gather_band_rows,
                                   run_bands_in_process)
    torch = ctx.torch
    w, h, d = SHAPES["c4-bands"]
    views = make_rig(w, h, d, seed=1234, channels=1)
    hp = h + 2 * d
    times = []
    with sb.Engine(w, h, d, n_slots=1, device=ctx.local) as eng:
        worker = EngineBandWorker(eng, views, d, ctx.rank, ctx.world, mode=0)
        streamed = False
        if ctx.world > 1:
            streamed = connect_row_mailboxes(worker, ctx.world, ctx.rank)  # the row sweeps of all bands run together, streamed over NVLink
        rows = None
        for rep in range(reps + 1):
            ctx.barrier()
            t0 = time.perf_counter()
            rows = compute_banded(worker, ctx.world, ctx.rank) if ctx.world > 1 else run_bands_in_process([worker])[0]
            torch.cuda.synchronize()
            dt = ctx.max(time.perf_counter() - t0)
            if rep > 0:
                times.append(dt * 1e3)
        full = gather_band_rows(rows, d, h, hp, dst=0) if ctx.world > 1 else rows
        timeline = None
        if ctx.world > 1:  # one more frame with the device drained after every phase: where the time of a banded frame goes
            trace = []
            ctx.barrier()
            compute_banded(worker, ctx.world, ctx.rank, trace=trace)
            mine = [[label, round(t * 1e3, 2)] for label, t in trace]
            allt = [None] * ctx.world
            ctx.dist.all_gather_object(allt, mine)
            timeline = {f"rank{r}": t for r, t in enumerate(allt)}
            if streamed:
                disconnect_row_mailboxes(worker)
        equal, single_ms = None, None
        if check and ctx.rank == 0:
            eng.compute(views, d, mode_mask=1)
            t0 = time.perf_counter()
            want = eng.compute(views, d, mode_mask=1)[0]
            single_ms = (time.perf_counter() - t0) * 1e3
            equal = bool((as_uint16(full) == want).all())
    ms = statistics.median(times)
    return {"what": "one frame by row bands, exact state hand-over between the bands (row sweeps: mailboxes in peer memory; column sweeps: NCCL send/recv)", "shape": [w, h, d], "n_gpus": ctx.world,
            "ms_per_frame": ms, "frames_per_s": 1e3 / ms, "gcost_evals_per_s": cost_evals(w, h, d) / ms / 1e6,
            "single_gpu_host_call_ms": single_ms, "equal_to_single_gpu": equal,
            "wta_rows_shared": ctx.world > 1, "row_sweeps_streamed": streamed, "timeline_ms_end_of_phase": timeline}


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c1", "c2", "c3", "c4-bands", "sweep"])
    ap.add_argument("--rigs-per-step", type=int, default=16, help="rigs per GPU per step (the end-to-end call drains its pipeline once per step)")
    ap.add_argument("--slots", type=int, default=4, help="rigs in flight per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the other named shapes / the band run that ride along with the default line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import sister_b200
    ctx = Dist()
    if not os.path.exists(sister_b200.library_path()):
        sister_b200.build_library()

    if args.workload == "c4-bands":
        b = run_bands(ctx, args)
        w, h, d = SHAPES["c4-bands"]
        line = {"metric": "frames_per_s", "value": b["frames_per_s"], "unit": "frames/s", "n_gpus": ctx.world, "steps": 2, "warmup": 1,
                "ms_per_step": b["ms_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u16",
                "data": "synthetic", "gcost_evals_per_s": b["gcost_evals_per_s"], "config": config_of(w, h, d), "run": {"workload_key": "c4-bands"}, "bands": b}
    elif args.workload == "sweep":
        line = None
        res = {}
        for dd in SWEEP_D:
            args.workload = "c2"
            sub = run_frames(ctx, argparse.Namespace(**{**vars(args), "no_cpu_baseline": True, "no_extras": True, "rigs_per_step": 8}), 1280, 960, dd, extras_ok=False)
            res[f"d{dd}"] = {k: sub[k] for k in ("value", "gcost_evals_per_s", "ms_per_step", "single_rig_latency_ms")} | {"e2e": sub.get("e2e", {}).get("value"),
                                                                                                                             "roofline_frac": sub["roofline"]["frac"]}
            line = sub
        line["run"]["workload_key"] = "sweep"
        line["sweep"] = res
    else:
        w, h, d = SHAPES[args.workload]
        line = run_frames(ctx, args, w, h, d)
        if ctx.world > 1 and not args.no_extras and args.workload == "c2":
            try:
                line_b = run_bands(ctx, args, reps=2, check=True)
            except Exception as ex:  # the headline line must survive the side measurement
                line_b = {"error": repr(ex)}
            line["bands_c4"] = line_b
    if ctx.rank == 0:
        emit(line)
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
