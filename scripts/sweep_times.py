#!/usr/bin/env python
"""Measurement aid (library built with -DSISTER_DEBUG_HOOKS): per-block start / end times of the last k_sgm_sweeps launch."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sister_b200  # noqa: E402
from sister_b200.synth import make_rig  # noqa: E402

W, H, D = (int(x) for x in os.environ.get("SISTER_SHAPE", "1280,960,192").split(","))
views = make_rig(W, H, D, seed=1234, channels=3)
with sister_b200.Engine(W, H, D, n_slots=1) as eng:
    rig = eng.upload_rig(views)
    out = eng.dev_alloc(W * H * 2)
    for it in range(3):
        eng.submit_device(0, rig, W, H, 3, D, sister_b200.MODE_MULTIVIEW, [out, 0, 0])
        eng.sync(0)
    t = np.zeros((2048, 4), np.uint64)
    rc = eng.lib.sister_debug_sweep_times(t.ctypes.data_as(C.c_void_p), C.c_size_t(t.nbytes))
    assert rc == 0
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
print("blocks", len(t), "kernel span us", (t[:, 1].max() - t0) / 1e3)
for s in range(4):
    m = (t[:, 3] >> np.uint64(16)) == s
    ts = t[m]
    ts = ts[np.argsort(ts[:, 3])]
    st = (ts[:, 0] - t0) / 1e3
    en = (ts[:, 1] - t0) / 1e3
    print(f"sweep {s}: blocks {len(ts)}")
    for k in range(0, len(ts), max(1, len(ts) // 12)):
        print(f"   b={k:3d} sm={int(ts[k,2]):3d} start {st[k]:8.1f} end {en[k]:8.1f} dur {en[k]-st[k]:8.1f}")
    print(f"   last  start {st[-1]:8.1f} end {en[-1]:8.1f} dur {en[-1]-st[-1]:8.1f}")
