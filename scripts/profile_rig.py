#!/usr/bin/env python
"""Run a few rigs of BASELINE.json configs[1] through the device-resident path; the command ncu wraps (B200_PROFILING.md).

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_rig.py
  ncu --set full --clock-control none --import-source on -k regex:k_sgm -s 7 -c 2 -o gpurun_out/prof python scripts/profile_rig.py
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sister_b200  # noqa: E402
from sister_b200.synth import make_rig  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rigs", type=int, default=2)
ap.add_argument("--w", type=int, default=1280)
ap.add_argument("--h", type=int, default=960)
ap.add_argument("--d", type=int, default=192)
ap.add_argument("--mode", type=int, default=sister_b200.MODE_MULTIVIEW)
a = ap.parse_args()
views = make_rig(a.w, a.h, a.d, seed=1234, channels=3)
with sister_b200.Engine(a.w, a.h, a.d, n_slots=1) as eng:
    rig = eng.upload_rig(views)
    out = eng.dev_alloc(a.w * a.h * 2 * 3)
    outs = [out + k * a.w * a.h * 2 if (a.mode >> k) & 1 else 0 for k in range(3)]
    for _ in range(a.rigs):
        eng.submit_device(0, rig, a.w, a.h, 3, a.d, a.mode, outs)
        eng.sync(0)
    res = np.zeros((a.h, a.w), np.uint16)
    eng.dev_download(out, res)
    print("checksum", int(res[::97, ::89].sum()), "launches", eng.launch_count())
