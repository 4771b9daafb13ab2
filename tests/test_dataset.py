"""Dataset walker and evaluation (sister_b200/dataset.py) on a tiny synthetic tree in the
layout of the reference README (README.md:35-61) and of the raw tree python/extract_dataset.py converts from."""
import os

import numpy as np
import pytest

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
cv2 = pytest.importorskip("cv2")

from sister_b200.dataset import (SisterDataset, VIEW_ORDER, decode_disparity, depth_to_disparity, disparity_to_depth, evaluate_depth,  # noqa: E402
                                 evaluate_rig, export_raw_tree, run_dataset)
from sister_b200.synth import make_rig  # noqa: E402

W, H, D, FOCAL = 96, 64, 32, 800.0
PLANE = int(round(D / 3.0))  # synth.disparity_field(kind="plane") rounds D / 3


def write_rig(folder, views):
    os.makedirs(folder, exist_ok=True)
    for name, v in zip(VIEW_ORDER, views):
        assert cv2.imwrite(os.path.join(folder, name + ".png"), v)


def build_tree(root):
    """two objects, one distance each, two baselines; the plane rigs have disparity PLANE = round(D / 3) everywhere"""
    rigs = {}
    for k, obj in enumerate(("washer", "hexa_screw")):
        for b, base in enumerate(("025mm", "050mm")):
            views = make_rig(W, H, D, seed=50 + 2 * k + b, kind="plane", noise=0, channels=3)
            write_rig(os.path.join(root, obj, "10cm", base), views)
            rigs[(obj, base)] = views
        depth = np.full((H, W), FOCAL * 0.025 / PLANE, np.float32)
        depth[:4] = 0  # invalid ground truth rows
        assert cv2.imwrite(os.path.join(root, obj, "10cm", "gt_depth.exr"), depth)
    return rigs


def test_walk_load_and_units(tmp_path):
    rigs = build_tree(str(tmp_path))
    ds = SisterDataset(tmp_path)
    assert len(ds) == 4 and ds.object_names == ["hexa_screw", "washer"]
    r = next(ds.rigs("washer", baseline="050mm"))
    assert abs(r.baseline_m - 0.050) < 1e-9 and abs(r.distance_m - 0.10) < 1e-9
    views = ds.load_views(r)
    for got, want in zip(views, rigs[("washer", "050mm")]):
        assert got.dtype == np.uint8 and got.shape == (H, W, 3) and (got == want).all()
    gt = ds.load_gt_depth(r)
    assert gt.dtype == np.float32 and gt.shape == (H, W) and (gt[:4] == 0).all()


def test_disparity_depth_round_trip_and_metrics():
    disp = np.array([[0, 255, 2550, 65535]], np.uint16)
    d = decode_disparity(disp)
    assert np.allclose(d, [[0, 1, 10, 257]])
    z = disparity_to_depth(d, FOCAL, 0.05)
    assert z[0, 0] == 0 and np.isclose(z[0, 2], FOCAL * 0.05 / 10)
    assert np.allclose(depth_to_disparity(z, FOCAL, 0.05)[0, 1:], d[0, 1:])
    gt = np.array([[1.0, 4.0, 0.0, 2.0]], np.float32)
    pred = np.array([[1.001, 0.0, 3.0, 2.004]], np.float32)
    m = evaluate_depth(pred, gt, bad_thresholds_m=(0.002,))
    assert m["gt_pixels"] == 3 and m["evaluated_pixels"] == 2 and np.isclose(m["completeness"], 2 / 3)
    assert np.isclose(m["mae_m"], 0.0025, atol=1e-6) and np.isclose(m["bad_2mm"], 0.5)


def test_run_dataset_with_a_stand_in_compute(tmp_path):
    build_tree(str(tmp_path))
    ds = SisterDataset(tmp_path)

    def compute(batch, disp_count):
        assert disp_count == D and all(len(r) == 5 for r in batch)
        m = np.full((H, W), PLANE * 255, np.uint16)
        return [[m, m, None] for _ in batch]

    rep = run_dataset(ds, compute, D, FOCAL, batch=3)
    assert len(rep) == 4
    for e in rep:
        assert set(e["metrics"]) == {"multiview", "horizontal"}
        exact = e["baseline"] == "025mm"  # the ground truth was written for the 25 mm baseline
        assert (e["metrics"]["multiview"]["mae_m"] < 1e-6) == exact


def test_export_raw_tree(tmp_path):
    scenes, gts, out = tmp_path / "scenes", tmp_path / "gt", tmp_path / "out"
    views = make_rig(W, H, D, seed=9, kind="plane", noise=0, channels=3)
    for level, base in (("1", "010"), ("2", "100")):
        folder = scenes / "arduino" / f"scene_{level}_{base}"
        os.makedirs(folder)
        for name, v in zip(VIEW_ORDER, views):
            cv2.imwrite(str(folder / f"00000_{name}.png"), v)
    os.makedirs(scenes / "no_gt_object" / "scene_1_010")
    os.makedirs(gts)
    for level in ("1", "2"):
        cv2.imwrite(str(gts / f"arduino_{level}.exr"), np.full((H, W), 0.05 * int(level), np.float32))
    assert export_raw_tree(scenes, gts, out) == 2
    ds = SisterDataset(out)
    assert [(r.object_name, r.distance, r.baseline) for r in ds.rigs()] == [("arduino", "10cm", "100mm"), ("arduino", "5cm", "010mm")]
    assert np.allclose(ds.load_gt_depth(next(ds.rigs(distance="5cm"))), 0.05)


@pytest.mark.gpu
def test_engine_over_the_dataset_recovers_the_plane(tmp_path):
    import sister_b200
    build_tree(str(tmp_path))
    ds = SisterDataset(tmp_path)
    with sister_b200.Engine(W, H, D, n_slots=2) as eng:
        rep = run_dataset(ds, lambda views, dc: eng.compute_batch(views, dc, mode_mask=sister_b200.MODE_ALL), D, FOCAL,
                          rigs=list(ds.rigs(baseline="025mm")))
    assert len(rep) == 2
    for e in rep:
        m = e["metrics"]["multiview"]
        assert m["completeness"] > 0.8 and m["median_m"] < 1e-6, m
