"""The two-view path (doStereo, hpp:122-150).
CPU: the plain-C oracle composed in doStereo's order reproduces the fixtures the compiled reference produced
(tests/golden/stereo_pairs.npz, scripts/gen_golden.py --stereo-only), and, where the reference objects are present, the
reference itself. GPU: sister_stereo against the fixtures, and SGM on uint8 volumes holding the 255 marker."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_views, grey_of
from sister_b200.synth import make_rig

G = np.load(os.path.join(GOLDEN, "stereo_pairs.npz"))
N_PAIRS = sum(1 for k in G.files if k.startswith("shape_"))


def pair(k):
    w, h, D, seed = (int(x) for x in G[f"shape_{k}"])
    views = make_rig(w, h, D, seed=seed, kind="smooth", channels=1)
    return views[0], views[1], D


@pytest.mark.parametrize("k", range(N_PAIRS))
def test_oracle_two_view_path_reproduces_reference_fixtures(oracle_lib, k):
    c, r, D = pair(k)
    L, R = oracle_lib.do_stereo(c, r, D)
    assert (L.astype(np.int16) == G[f"left_{k}"]).all()
    assert (R.astype(np.int16) == G[f"right_{k}"]).all()


def test_oracle_sgm_with_invalid_cost_marker(oracle_lib):
    for k in range(3):
        assert (oracle_lib.sgm(G[f"sgm255_in_{k}"].astype(np.uint16)) == G[f"sgm255_out_{k}"]).all()


def test_reference_two_view_path_matches_fixture(ref_lib):
    c, r, D = pair(0)
    L, R = ref_lib.do_stereo(c, r, D)
    assert (L.astype(np.int16) == G["left_0"]).all() and (R.astype(np.int16) == G["right_0"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("k", range(N_PAIRS))
def test_gpu_two_view_path_reproduces_reference_fixtures(k):
    import sister_b200
    c, r, D = pair(k)
    h, w = c.shape
    with sister_b200.Engine(w, h, D, n_slots=1) as eng:
        L, R = eng.stereo(c, r, D)
        assert (L.astype(np.int16) == G[f"left_{k}"]).all(), f"{(L.astype(np.int16) != G[f'left_{k}']).sum()} px differ (left)"
        assert (R.astype(np.int16) == G[f"right_{k}"]).all(), f"{(R.astype(np.int16) != G[f'right_{k}']).sum()} px differ (right)"


@pytest.mark.gpu
def test_gpu_sgm_with_invalid_cost_marker():
    import sister_b200
    with sister_b200.Engine(64, 64, 192, n_slots=1) as eng:
        for k in range(3):
            s, _ = eng.test_sgm(G[f"sgm255_in_{k}"])
            assert (s == G[f"sgm255_out_{k}"]).all(), f"KAT {k}: {(s != G[f'sgm255_out_{k}']).sum()} cells differ"


@pytest.mark.gpu
def test_gpu_two_view_then_five_view_on_one_context(oracle_lib):
    import sister_b200
    g = np.load(os.path.join(GOLDEN, "rig_64x48_d16.npz"))
    views = golden_views(g)
    c, r, D = pair(3)
    with sister_b200.Engine(96, 112, 16, n_slots=1) as eng:
        a = eng.compute(views, 16)
        L, _ = eng.stereo(c, r, D)
        b = eng.compute(views, 16)
        assert (L.astype(np.int16) == G["left_3"]).all()
        for m, key in enumerate(("disp_mv", "disp_h", "disp_v")):
            assert (a[m] == g[key]).all() and (b[m] == g[key]).all()
