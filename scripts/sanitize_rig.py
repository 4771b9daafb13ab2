#!/usr/bin/env python
"""One small rig through every kernel of the path -- the 5-view call (3 modes, crop-only and whole frame), the batch
pipeline, the two-view path, an SGM stage call, a 3-band run on one GPU (states handed over at the end of a band, then streamed) and a second, wider rig (the chunked median kernel) --
for compute-sanitizer
(scripts/sanitize.sh runs it under memcheck, racecheck, initcheck and synccheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sister_b200  # noqa: E402
from sister_b200.synth import make_rig  # noqa: E402

w, h, D = 96, 64, 32
views = make_rig(w, h, D, seed=42, channels=3, colour=True)
with sister_b200.Engine(w, h, D, n_slots=3) as eng:
    a = eng.compute(views, D)
    b, raw = eng.compute(views, D, want_raw=True)
    assert all((x == y).all() for x, y in zip(a, b))
    batch = eng.compute_batch([views] * 4, D, mode_mask=1)
    assert all((r[0] == a[0]).all() for r in batch)
    g = [v[:, :, 1].copy() for v in views]
    eng.stereo(g[0], g[1], D)
    vol = np.random.default_rng(1).integers(0, 253, (16, 24, 40), dtype=np.uint8)
    eng.test_sgm(vol)
    if "--bands" in sys.argv:
        from sister_b200.bands import EngineBandWorker, as_uint16, run_bands_in_process
        workers = [EngineBandWorker(eng, views, D, r, 3, mode=0, slot=r) for r in range(3)]
        rows = run_bands_in_process(workers)
        got = np.concatenate([as_uint16(r) for r in rows], axis=0)
        assert (got == a[0]).all()
        # the same with the row sweeps streamed between the bands (tagged mailboxes) and the column sweeps on their own stream
        from sister_b200.bands import connect_row_mailboxes_in_process
        connect_row_mailboxes_in_process(workers)
        for _ in range(2):
            rows = run_bands_in_process(workers)
            got = np.concatenate([as_uint16(r) for r in rows], axis=0)
            assert (got == a[0]).all()
# a frame wide enough (264 x 256 padded) for the chunked median kernel, small D
w2, h2, D2 = 248, 240, 8
views2 = make_rig(w2, h2, D2, seed=43, channels=1)
with sister_b200.Engine(w2, h2, D2, n_slots=1) as eng:
    c = eng.compute(views2, D2, mode_mask=1)
    d = eng.compute(views2, D2, mode_mask=1)
    assert (c[0] == d[0]).all()
print("sanitize_rig ok")
