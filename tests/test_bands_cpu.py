"""Row-band sharding (sister_b200/bands.py) on CPU: the schedule for every world size, the band decomposition of the
aggregation against the whole-frame spec (tests/sgm_spec.py, itself proven equal to the reference recurrence), and the
real thing over world_size-2 and -3 gloo processes with the numpy band worker."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from band_spec import NumpyBandWorker, StreamedNumpyBandWorker  # noqa: E402
from sgm_spec import sgm_decomposed  # noqa: E402

from sister_b200.bands import (band_program, band_rows, compute_banded, crop_rows_of_band, gather_band_rows, run_bands_in_process,  # noqa: E402
                               simulate_programs)

D, H, W = 8, 14, 12
HP, WP = H + 2 * D, W + 2 * D


def fused(seed=5):
    rng = np.random.default_rng(seed)
    C = rng.integers(0, 253, (HP, WP, D), dtype=np.int64)
    C[rng.random((HP, WP, D)) < 0.2] = 0
    return C


def whole_frame_map(C):
    S = sgm_decomposed(C).astype(np.int64)
    out = np.zeros((H, W), np.uint16)
    for ii in range(H):
        for jj in range(W):
            i, j = ii + D, jj + D
            out[ii, jj] = min(int(S[i, j, : min(j, D - 1) + 1].argmin()) * 255, 65535)
    return out


def test_schedule_is_deadlock_free_and_overlaps_the_two_passes():
    for world in range(1, 17):
        steps = simulate_programs(world)
        # two wavefronts of `world` computes each, overlapped: about world slots of (compute + transfer), not 2 * world
        assert steps <= 2 * world + 2
        for r in range(world):
            prog = band_program(world, r)
            assert [o for o in prog if o[0] == "compute"] in ([("compute", 0, -1), ("compute", 1, -1)], [("compute", 1, -1), ("compute", 0, -1)])


def test_bands_partition_the_padded_rows():
    for hp in (9, 64, 1344, 3840):
        for world in (1, 2, 3, 8):
            spans = [band_rows(hp, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == hp
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert crop_rows_of_band(8, 14, 0, 8) == (0, 0)
    assert crop_rows_of_band(8, 14, 4, 12) == (0, 4)
    assert crop_rows_of_band(8, 14, 20, 30) == (12, 14)


@pytest.mark.parametrize("world", [1, 2, 3, 5])
def test_band_decomposition_equals_whole_frame(world):
    C = fused()
    workers = [NumpyBandWorker(C, D, H, W, *band_rows(HP, world, r)) for r in range(world)]
    rows = run_bands_in_process(workers)
    got = np.concatenate([r.numpy().view(np.uint16) for r in rows], axis=0)
    assert (got == whole_frame_map(C)).all()


@pytest.mark.parametrize("world", [2, 3, 5])
def test_streamed_schedule_in_process_equals_whole_frame(world):
    """run_bands_in_process with share_match + stream_rows workers: the WTA shares are gathered, the row sweeps run producers
    first (pass 0 top to bottom, pass 1 bottom to top), then the column sweeps alone go through the two wavefronts."""
    C = fused()
    boxes = {}
    workers = [StreamedNumpyBandWorker(C, D, H, W, *band_rows(HP, world, r), r, world, boxes) for r in range(world)]
    rows = run_bands_in_process(workers)
    got = np.concatenate([r.numpy().view(np.uint16) for r in rows], axis=0)
    assert (got == whole_frame_map(C)).all()
    for w in workers:
        assert w.calls[:4] == ["share", "rest", "rows1", "rows2"] and sorted(w.calls[4:]) == ["columns0", "columns1"]


def _streamed_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = StreamedNumpyBandWorker(fused(), D, H, W, *band_rows(HP, world, rank), rank, world)
        rows = compute_banded(w, world, rank)
        assert w.calls[:3] == ["share", "rest", "rows3"] and sorted(w.calls[3:]) == ["columns0", "columns1"]
        full = gather_band_rows(rows, D, H, HP, dst=0)
        if rank == 0:
            q.put(full.numpy().view(np.uint16).copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_with_the_streamed_schedule(world):
    """compute_banded over gloo with share_match + stream_rows: all_gather_into_tensor of the shares, rows() once per rank, the
    column states through band_program's blocking sends and receives."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_streamed_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert (got == whole_frame_map(fused())).all()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = NumpyBandWorker(fused(), D, H, W, *band_rows(HP, world, rank))
        rows = compute_banded(w, world, rank)
        full = gather_band_rows(rows, D, H, HP, dst=0)
        if rank == 0:
            q.put(full.numpy().view(np.uint16).copy())
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_reproduce_the_whole_frame_map(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert (got == whole_frame_map(fused())).all()
