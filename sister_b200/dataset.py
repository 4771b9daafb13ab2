"""The data formats either side of the path: the SiSter dataset tree and the evaluation
of disparity maps against its ground-truth depth.

Layout of the published dataset (reference README.md:35-61):

    <object> / <camera distance, e.g. 10cm> / <baseline, e.g. 025mm> / {center,left,top,right,bottom}.png
    <object> / <camera distance> / gt_depth.exr          float32 metres, registered to the centre view

and of the raw acquisition tree the reference's python/extract_dataset.py converts from (extract_dataset.py:9-96):

    <scenes>/<object>/<prefix>_<level>_<baseline>/00000_<direction>.png,   <gt>/<object>_<level>.exr
    level 0 / 1 / 2  =  1cm / 5cm / 10cm                                      (extract_dataset.py:10)

`SisterDataset` walks the published layout, `export_raw_tree` performs the raw -> published conversion
(extract_dataset.py:148-183), `evaluate_rig` turns the uint16 maps of compute_disparities (disparity * 255, hpp:116-118)
into depth and compares with the ground truth. Host-side tooling: numpy + OpenCV's Python module for PNG / EXR only.
"""
from __future__ import annotations

import os
import shutil
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np

VIEW_ORDER = ("center", "right", "top", "left", "bottom")  # constructor order, hpp:22
LEVEL_TO_DISTANCE = {"0": "1cm", "1": "5cm", "2": "10cm"}   # extract_dataset.py:10


def _cv2():
    os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")  # must be set before cv2 touches an .exr
    import cv2

    return cv2


@dataclass(frozen=True)
class RigRef:
    """One 5-view rig of the dataset."""
    object_name: str
    distance: str   # folder name, e.g. "10cm"
    baseline: str   # folder name, e.g. "025mm"
    folder: Path

    @property
    def baseline_m(self) -> float:
        return float("".join(c for c in self.baseline if c.isdigit() or c == ".")) * 1e-3

    @property
    def distance_m(self) -> float:
        return float("".join(c for c in self.distance if c.isdigit() or c == ".")) * 1e-2


class SisterDataset:
    """Walker over the published layout. `rigs()` yields every folder that holds all five views."""

    def __init__(self, root):
        self.root = Path(root)
        if not self.root.is_dir():
            raise FileNotFoundError(f"dataset root {root} does not exist")
        self._rigs: List[RigRef] = []
        for obj in sorted(p for p in self.root.iterdir() if p.is_dir()):
            for dist in sorted(p for p in obj.iterdir() if p.is_dir()):
                for base in sorted(p for p in dist.iterdir() if p.is_dir()):
                    if all((base / f"{v}.png").is_file() for v in VIEW_ORDER):
                        self._rigs.append(RigRef(obj.name, dist.name, base.name, base))

    def rigs(self, object_name: Optional[str] = None, distance: Optional[str] = None, baseline: Optional[str] = None) -> Iterator[RigRef]:
        for r in self._rigs:
            if (object_name is None or r.object_name == object_name) and (distance is None or r.distance == distance) and \
                    (baseline is None or r.baseline == baseline):
                yield r

    def __len__(self):
        return len(self._rigs)

    @property
    def object_names(self) -> List[str]:
        return sorted({r.object_name for r in self._rigs})

    def load_views(self, rig: RigRef) -> List[np.ndarray]:
        """The five views as H x W x 3 BGR uint8 in constructor order, exactly what cv::imread gives the reference
        (compute_disp.cpp:19-23)."""
        cv2 = _cv2()
        views = []
        for v in VIEW_ORDER:
            img = cv2.imread(str(rig.folder / f"{v}.png"), cv2.IMREAD_COLOR)
            if img is None:
                raise IOError(f"cannot read {rig.folder / (v + '.png')}")
            views.append(img)
        if len({v.shape for v in views}) != 1:
            raise ValueError(f"views of {rig.folder} differ in size")
        return views

    def gt_path(self, rig: RigRef) -> Path:
        return self.root / rig.object_name / rig.distance / "gt_depth.exr"

    def load_gt_depth(self, rig: RigRef) -> Optional[np.ndarray]:
        """Ground-truth depth in metres (float32, H x W), or None when the distance folder has no gt_depth.exr."""
        p = self.gt_path(rig)
        if not p.is_file():
            return None
        cv2 = _cv2()
        d = cv2.imread(str(p), cv2.IMREAD_ANYCOLOR | cv2.IMREAD_ANYDEPTH)  # README.md:57-61
        if d is None:
            raise IOError(f"cannot read {p} (OpenCV built without OpenEXR?)")
        if d.ndim == 3:
            d = d[:, :, 0]
        return d.astype(np.float32)


def export_raw_tree(scenes_folder, gt_folder, output_folder) -> int:
    """Raw acquisition tree -> published layout (what extract_dataset.py's `export` command does, :148-183). Returns the
    number of rigs written."""
    scenes, gts, out = Path(scenes_folder), Path(gt_folder), Path(output_folder)
    gt_of: Dict[str, Dict[str, Path]] = {}
    for f in sorted(p for p in gts.iterdir() if p.is_file()):
        name, _, level = f.stem.rpartition("_")
        gt_of.setdefault(name, {})[level] = f
    n = 0
    for obj in sorted(p for p in scenes.iterdir() if p.is_dir()):
        if obj.name not in gt_of:
            continue  # the reference exports only objects that have ground truth (extract_dataset.py:27-30)
        for sub in sorted(p for p in obj.iterdir() if p.is_dir()):
            parts = sub.name.split("_")
            if len(parts) != 3:
                continue
            _, level, baseline = parts
            if level not in gt_of[obj.name] or level not in LEVEL_TO_DISTANCE:
                continue
            dist_dir = out / obj.name / LEVEL_TO_DISTANCE[level]
            dst = dist_dir / f"{baseline}mm"
            dst.mkdir(parents=True, exist_ok=True)
            for png in sorted(sub.glob("*.png")):
                shutil.copy(png, dst / png.name.replace("00000_", ""))
            shutil.copy(gt_of[obj.name][level], dist_dir / "gt_depth.exr")
            n += 1
    return n


# ---------------------------------------------------------------------------------------------------------- evaluation

def decode_disparity(map_u16: np.ndarray) -> np.ndarray:
    """uint16 map of compute_disparities -> disparity in pixels (hpp:116-118 stores disparity * 255, saturated)."""
    return map_u16.astype(np.float32) / 255.0


def disparity_to_depth(disp_px: np.ndarray, focal_px: float, baseline_m: float) -> np.ndarray:
    """Pinhole triangulation Z = f B / d; 0 where the disparity is 0 (the matcher's "no estimate", hpp:203)."""
    z = np.zeros_like(disp_px, dtype=np.float32)
    ok = disp_px > 0
    z[ok] = focal_px * baseline_m / disp_px[ok]
    return z


def depth_to_disparity(depth_m: np.ndarray, focal_px: float, baseline_m: float) -> np.ndarray:
    d = np.zeros_like(depth_m, dtype=np.float32)
    ok = depth_m > 0
    d[ok] = focal_px * baseline_m / depth_m[ok]
    return d


def evaluate_depth(pred_m: np.ndarray, gt_m: np.ndarray, bad_thresholds_m: Sequence[float] = (0.001, 0.002, 0.005)) -> Dict[str, float]:
    """Error statistics over pixels where both maps are valid (> 0 and finite)."""
    gt_ok = np.isfinite(gt_m) & (gt_m > 0)
    both = gt_ok & np.isfinite(pred_m) & (pred_m > 0)
    out = {"gt_pixels": int(gt_ok.sum()), "evaluated_pixels": int(both.sum()),
           "completeness": float(both.sum() / max(int(gt_ok.sum()), 1))}
    if not both.any():
        return out
    err = np.abs(pred_m[both].astype(np.float64) - gt_m[both].astype(np.float64))
    out.update(mae_m=float(err.mean()), rmse_m=float(np.sqrt((err ** 2).mean())), median_m=float(np.median(err)))
    for t in bad_thresholds_m:
        out[f"bad_{t * 1e3:g}mm"] = float((err > t).mean())
    return out


def evaluate_rig(maps_u16: Sequence[Optional[np.ndarray]], gt_depth_m: np.ndarray, focal_px: float, baseline_m: float) -> Dict[str, Dict[str, float]]:
    """maps_u16: [multiview, horizontal, vertical] as compute_disparities returns them (entries may be None)."""
    res = {}
    for name, m in zip(("multiview", "horizontal", "vertical"), maps_u16):
        if m is not None:
            res[name] = evaluate_depth(disparity_to_depth(decode_disparity(m), focal_px, baseline_m), gt_depth_m)
    return res


def run_dataset(dataset: SisterDataset, compute, disp_count: int, focal_px: float, rigs: Optional[Sequence[RigRef]] = None,
                batch: int = 8) -> List[Dict]:
    """Run `compute(list of 5-view rigs, disp_count) -> list of [mv, horiz, vert]` (Engine.compute_batch with mode_mask 7,
    or any stand-in) over the dataset in batches of equally sized rigs and evaluate against the ground truth."""
    todo = list(rigs) if rigs is not None else list(dataset.rigs())
    report = []
    i = 0
    while i < len(todo):
        chunk, views = [], []
        while i < len(todo) and len(chunk) < batch:
            v = dataset.load_views(todo[i])
            if views and v[0].shape != views[0][0].shape:
                break  # a batch shares one shape
            chunk.append(todo[i])
            views.append(v)
            i += 1
        outs = compute(views, disp_count)
        for rig, maps in zip(chunk, outs):
            gt = dataset.load_gt_depth(rig)
            entry = {"object": rig.object_name, "distance": rig.distance, "baseline": rig.baseline}
            if gt is not None:
                entry["metrics"] = evaluate_rig(maps, gt, focal_px, rig.baseline_m)
            report.append(entry)
    return report


def main(argv=None):
    import argparse
    import json

    ap = argparse.ArgumentParser(description="Run the 5-view path over a SiSter dataset tree and evaluate against gt_depth.exr")
    ap.add_argument("root")
    ap.add_argument("--disp", type=int, default=192, help="dispCount (hpp:26)")
    ap.add_argument("--focal-px", type=float, required=True, help="focal length of the centre camera in pixels")
    ap.add_argument("--object", default=None)
    ap.add_argument("--distance", default=None)
    ap.add_argument("--baseline", default=None)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    import sister_b200

    ds = SisterDataset(a.root)
    rigs = list(ds.rigs(a.object, a.distance, a.baseline))
    if not rigs:
        raise SystemExit("no rig matches")
    h, w = ds.load_views(rigs[0])[0].shape[:2]
    with sister_b200.Engine(w, h, a.disp, n_slots=4, device=a.device) as eng:
        report = run_dataset(ds, lambda views, D: eng.compute_batch(views, D, mode_mask=sister_b200.MODE_ALL), a.disp, a.focal_px, rigs)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
