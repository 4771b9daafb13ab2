#!/usr/bin/env python
"""Per-stage CUDA-event times of one rig of BASELINE.json configs[1] (median of N), for quick experiments."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sister_b200  # noqa: E402
from sister_b200.synth import make_rig  # noqa: E402

W, H, D = (int(x) for x in os.environ.get("SISTER_SHAPE", "1280,960,192").split(","))
views = make_rig(W, H, D, seed=1234, channels=3)
with sister_b200.Engine(W, H, D, n_slots=1) as eng:
    rig = eng.upload_rig(views)
    out = eng.dev_alloc(W * H * 2)
    eng.set_profiling(True)
    acc = {k: [] for k in sister_b200.STAGE_NAMES}
    for it in range(7):
        eng.submit_device(0, rig, W, H, 3, D, sister_b200.MODE_MULTIVIEW, [out, 0, 0])
        eng.sync(0)
        if it >= 2:
            for k, v in eng.stage_ms(0).items():
                acc[k].append(v)
    print("nw", os.environ.get("SISTER_DEBUG_SWEEP_NW", "auto"), {k: round(statistics.median(v), 4) for k, v in acc.items()})
