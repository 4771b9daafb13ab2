#!/usr/bin/env python
"""Debug-hook build only: run one rig and print the median kernel's per-phase clocks (block 0)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sister_b200  # noqa: E402
from sister_b200.synth import make_rig  # noqa: E402

W, H, D = 1280, 960, 192
views = make_rig(W, H, D, seed=1234, channels=3)
with sister_b200.Engine(W, H, D, n_slots=1) as eng:
    maps = eng.compute(views, D, sister_b200.MODE_MULTIVIEW)
    k = np.zeros((32, 8), np.int64)
    eng.lib.sister_debug_median_clk(k.ctypes.data_as(C.c_void_p))
    rows = H + 2 * D
    print("cycles per row: stage | wait_group+raw_row | wait left | wait right | bnd+median | sts+arrive | stg | loop")
    for c in (0, 1, 6, 11, 12):
        print(c, np.round(k[c] / rows, 1), round(k[c].sum() / rows, 1))
