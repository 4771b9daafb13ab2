"""Property test (hypothesis): random W, H (multiples of 4), D (multiples of 8), random seeds, grey / replicated / true
colour input -- the CUDA path through the C ABI equals the CPU oracle bit for bit: the three encoded maps and the raw
padded disparity of every mode (SURVEY.md section 4's test plan)."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402

from sister_b200.synth import make_rig  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import sister_b200
    with sister_b200.Engine(96, 80, 72, n_slots=1) as eng:
        yield eng


@settings(max_examples=25, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(w4=st.integers(4, 24), h4=st.integers(4, 20), d8=st.integers(1, 9), seed=st.integers(0, 10 ** 6),
       kind=st.sampled_from(["smooth", "plane"]), chan=st.sampled_from(["grey", "bgr", "colour"]), modes=st.integers(1, 7))
def test_random_shapes_equal_the_oracle(engine, oracle_lib, w4, h4, d8, seed, kind, chan, modes):
    w, h, D = 4 * w4, 4 * h4, 8 * d8
    views = make_rig(w, h, D, seed=seed, kind=kind, channels=1 if chan == "grey" else 3, colour=chan == "colour")
    outs, raw = engine.compute(views, D, mode_mask=modes, want_raw=True)
    crop = engine.compute(views, D, mode_mask=modes)  # the product path: crop-only aggregation
    ref, ref_raw = oracle_lib.compute_disparities(views, D, mode_mask=modes, want_raw=True)
    for m in range(3):
        if not (modes >> m) & 1:
            assert outs[m] is None
            continue
        assert (raw[m] == ref_raw[m]).all(), f"{w}x{h} D={D} mode {m}: {(raw[m] != ref_raw[m]).sum()} padded pixels differ"
        assert (outs[m] == ref[m]).all() and (crop[m] == ref[m]).all()


@pytest.fixture(scope="module")
def wide_engine():
    import sister_b200
    with sister_b200.Engine(640, 400, 16, n_slots=1) as eng:
        eng.set_test_taps(True)
        yield eng


@settings(max_examples=20, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(w4=st.integers(56, 160), h4=st.integers(56, 100), d8=st.integers(1, 2), seed=st.integers(0, 10 ** 6),
       kind=st.sampled_from(["smooth", "plane"]))
def test_wide_frames_masks_equal_the_oracle(wide_engine, oracle_lib, w4, h4, d8, seed, kind):
    """Frames wide enough (padded sides 240 .. 672) for the chunked median kernel, whatever the number of chunks and the fill
    of the last one; sides that are not a multiple of 8 take the barrier kernel. The per-view median + LRC maps and the masks
    against the oracle (small D keeps the oracle fast)."""
    w, h, D = 4 * w4, 4 * h4, 8 * d8
    views = make_rig(w, h, D, seed=seed, kind=kind, channels=1)
    wp, hp = w + 2 * D, h + 2 * D
    pads = [oracle_lib.pad_replicate(v, D) for v in views]
    t = oracle_lib.multistereo(pads, D, 0)
    out = wide_engine.compute(views, D, mode_mask=1)[0]
    lr = wide_engine.fetch("lr_final", (4, wp * hp), np.int16)
    masks = wide_engine.fetch("masks", (4, hp, wp), np.uint8)
    for v in range(4):
        assert (lr[v] == t["lr"][v]).all(), f"{w}x{h} D={D}: median + LRC, view {v}: {(lr[v] != t['lr'][v]).sum()} px"
    assert (masks == t["masks"]).all()
    ref = oracle_lib.compute_disparities(views, D, mode_mask=1)
    assert (out == ref[0]).all()
