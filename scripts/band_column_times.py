#!/usr/bin/env python
"""Measurement aid: the eight bands of a config-4 frame in ONE process on one GPU (band-sized volumes), with the wall time of
every band's column sweep of pass 0 taken alone (everything else drained) -- what one hop of the column wavefront costs."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sister_b200  # noqa: E402
from sister_b200.bands import EngineBandWorker, as_uint16, band_rows, connect_row_mailboxes_in_process  # noqa: E402
from sister_b200.synth import make_rig  # noqa: E402

W, H, D = (int(x) for x in os.environ.get("SISTER_SHAPE", "4096,3072,384").split(","))
G = int(os.environ.get("SISTER_BANDS", "8"))
views = make_rig(W, H, D, seed=1234, channels=1)
hp = H + 2 * D
rows_max = max(b - a for a, b in (band_rows(hp, G, r) for r in range(G)))
with sister_b200.Engine(W, H, D, n_slots=G, max_band_rows=rows_max) as eng:
    workers = [EngineBandWorker(eng, views, D, r, G, mode=0, slot=r) for r in range(G)]
    connect_row_mailboxes_in_process(workers)
    for rep in range(2):
        gathered = torch.cat([w.submit_share(r, G) for r, w in enumerate(workers)])
        for w in workers:
            w.submit_rest(gathered, G)
        for w in workers:            # (one GPU cannot hold the row kernels of eight config-4 bands at once: one after the other)
            w.rows(passes=1)
            w.drain()
        for w in reversed(workers):
            w.rows(passes=2)
            w.drain()
        state, times = None, []
        for r, w in enumerate(workers):
            t0 = time.perf_counter()
            state = w.columns(0, state, r + 1 < G)
            w.drain()
            times.append((time.perf_counter() - t0) * 1e3)
        state = None
        for r in range(G - 1, -1, -1):
            state = workers[r].columns(1, state, r > 0)
        outs = [w.finish() for w in workers]
        print("column sweep of pass 0, ms per band:", [round(t, 2) for t in times])
    got = np.concatenate([as_uint16(r) for r in outs], axis=0)
    print("checksum", int(got[::97, ::89].sum()))
