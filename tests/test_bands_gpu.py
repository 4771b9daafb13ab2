"""Row bands on the GPU (sister_band_*, sister_b200/bands.py): G bands of one frame run on ONE GPU (one slot per band,
states handed over in process) must reproduce sister_compute bit for bit -- every chain cut at every band border and
continued from the stored state; and, when the box has 2+ GPUs, the same over NCCL with one process per GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

from sister_b200.synth import make_rig

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("share_match", [True, False])
@pytest.mark.parametrize("w,h,D,world,mode", [(96, 64, 32, 2, 0), (100, 76, 40, 3, 1), (52, 88, 136, 5, 2), (64, 48, 16, 7, 0), (72, 60, 264, 2, 0),
                                              (640, 480, 192, 4, 0)])
def test_bands_on_one_gpu_equal_the_single_band_map(w, h, D, world, mode, share_match):
    """share_match: every band matches 1/G of each view's rows and the packed maps are gathered (sister_band_submit_share /
    _rest); otherwise every band matches the whole frame (sister_band_submit)."""
    import sister_b200
    from sister_b200.bands import EngineBandWorker, as_uint16, run_bands_in_process

    views = make_rig(w, h, D, seed=31 + world, channels=3 if w < 600 else 1)
    with sister_b200.Engine(w, h, D, n_slots=world) as eng:
        want = eng.compute(views, D, mode_mask=1 << mode)[mode]
        workers = [EngineBandWorker(eng, views, D, r, world, mode=mode, slot=r, share_match=share_match) for r in range(world)]
        rows = run_bands_in_process(workers)
        got = np.concatenate([as_uint16(r) for r in rows], axis=0)
    assert got.shape == want.shape
    assert (got == want).all(), f"{(got != want).sum()} pixels differ"


@pytest.mark.parametrize("w,h,D,world", [(96, 64, 32, 2), (96, 64, 32, 5), (64, 48, 16, 7), (52, 88, 136, 3), (640, 480, 192, 4)])
def test_row_sweeps_streamed_between_the_bands(w, h, D, world):
    """sister_band_rows / sister_band_columns: the row sweeps of all bands run at the same time, each reading the rider states
    from a mailbox the band before writes while it runs (tagged words, as between two blocks of a sweep); two frames in a row
    (the tag changes), all three modes of one shape."""
    import sister_b200
    from sister_b200.bands import EngineBandWorker, as_uint16, connect_row_mailboxes_in_process, run_bands_in_process

    views = make_rig(w, h, D, seed=91 + world, channels=1)
    views2 = make_rig(w, h, D, seed=92 + world, channels=1, kind="plane")
    with sister_b200.Engine(w, h, D, n_slots=world) as eng:
        for mode in ((0, 1, 2) if w < 100 else (0,)):
            for vv in (views, views2):
                want = eng.compute(vv, D, mode_mask=1 << mode)[mode]
                workers = [EngineBandWorker(eng, vv, D, r, world, mode=mode, slot=r) for r in range(world)]
                connect_row_mailboxes_in_process(workers)
                for frame in range(2):
                    rows = run_bands_in_process(workers)
                    got = np.concatenate([as_uint16(r) for r in rows], axis=0)
                    assert (got == want).all(), f"mode {mode}, frame {frame}: {(got != want).sum()} pixels differ"
                for wk in workers:
                    for m in wk.mbox:
                        eng.dev_free(m)


@pytest.mark.parametrize("w,h,D,world", [(96, 64, 32, 3), (128, 96, 64, 4), (52, 88, 136, 2)])
def test_band_contexts_hold_a_band_of_the_volumes_only(w, h, D, world):
    """sister_create_band: the fused and pair volumes of a slot hold max_band_rows rows of the padded frame; G such slots
    together hold one frame's worth, and the banded map is still the single-GPU map. Full-frame calls are refused."""
    import sister_b200
    from sister_b200.bands import EngineBandWorker, as_uint16, band_rows, run_bands_in_process

    views = make_rig(w, h, D, seed=77 + world, channels=1)
    hp = h + 2 * D
    rows_max = max(b - a for a, b in (band_rows(hp, world, r) for r in range(world)))
    with sister_b200.Engine(w, h, D, n_slots=1) as eng:
        want = eng.compute(views, D, mode_mask=1)[0]
    with sister_b200.Engine(w, h, D, n_slots=world, max_band_rows=rows_max) as eng:
        workers = [EngineBandWorker(eng, views, D, r, world, mode=0, slot=r) for r in range(world)]
        rows = run_bands_in_process(workers)
        got = np.concatenate([as_uint16(r) for r in rows], axis=0)
        assert (got == want).all(), f"{(got != want).sum()} pixels differ"
        with pytest.raises(sister_b200.SisterError) as e:
            eng.compute(views, D, mode_mask=1)
        assert e.value.code == -3
    with pytest.raises(sister_b200.SisterError):
        with sister_b200.Engine(w, h, D, n_slots=1, max_band_rows=rows_max - 1) as eng:
            EngineBandWorker(eng, views, D, 0, world, mode=0, slot=0).submit()


def test_bands_over_nccl_when_the_box_has_two_gpus():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU: the NCCL leg runs under gpurun --gpus 2 (scripts/run_bands.py)")
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "scripts", "run_bands.py"), "--width", "320", "--height", "240", "--disp", "64", "--check"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "bands == single GPU: True" in r.stdout
