"""The oracle against the reference ITSELF (oracle/_ref, the unmodified sources compiled in place), on seeded
random inputs; plus the fake-OpenCV shim's image ops against cv2. CPU only; skipped where _ref is absent."""
import numpy as np
import pytest

from sister_b200.synth import make_rig


@pytest.mark.parametrize("h,w,D", [(40, 52, 16), (36, 88, 32), (48, 40, 8)])
def test_stages_match_reference(oracle_lib, ref_lib, h, w, D):
    rng = np.random.default_rng(h * 1000 + w)
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    b = rng.integers(0, 256, (h, w), dtype=np.uint8)
    assert (ref_lib.census(a) == oracle_lib.census(a)).all()
    va = ref_lib.ad_census(a, b, D)
    assert (va == oracle_lib.ad_census(a, b, D)).all()
    rv = rng.integers(0, 70, (h, w, D), dtype=np.uint16)
    for vol in (va, rv):
        Lr, Rr = ref_lib.wta(vol)
        Lo, Ro = oracle_lib.wta(vol)
        assert (Lr == Lo).all() and (Rr == Ro).all()
    mr = ref_lib.median_inplace(Lr)
    assert (mr == oracle_lib.median_inplace(Lr)).all()
    assert (ref_lib.lrcheck(mr, Rr) == oracle_lib.lrcheck(mr, Rr)).all()
    cv = rng.integers(0, 1021, (h, w, D), dtype=np.uint16)
    cv[rng.random((h, w, D)) < 0.05] = 255
    cv[0][rng.random((w, D)) < 0.3] = 255
    assert (ref_lib.sgm(cv) == oracle_lib.sgm(cv)).all()


@pytest.mark.parametrize("w,h,D,kind", [(64, 48, 16, "smooth"), (40, 56, 8, "plane")])
def test_header_end_to_end(oracle_lib, ref_lib, w, h, D, kind):
    views = make_rig(w, h, D, seed=77, kind=kind, channels=3)
    mv, hz, vt = ref_lib.compute_disparities(views, D)
    outs = oracle_lib.compute_disparities(views, D)
    assert (mv == outs[0]).all() and (hz == outs[1]).all() and (vt == outs[2]).all()
    pads = [oracle_lib.pad_replicate(oracle_lib.grey_bgr(v), D) for v in views]
    for mode in range(3):
        tr, to = ref_lib.multistereo_taps(pads, D, mode), oracle_lib.multistereo(pads, D, mode)
        for key in ("masks", "fused", "sum", "disp"):
            assert (tr[key] == to[key]).all(), (mode, key)


def test_image_ops_match_cv2(oracle_lib):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    g = oracle_lib.grey_bgr(bgr)
    assert (g == cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)).all()
    assert (oracle_lib.pad_replicate(g, 7) == cv2.copyMakeBorder(g, 7, 7, 7, 7, cv2.BORDER_REPLICATE)).all()
    assert (oracle_lib.orient(g, 180) == cv2.flip(g, 1)).all()
    assert (oracle_lib.orient(g, 90) == cv2.flip(cv2.transpose(g), -1)).all()
    assert (oracle_lib.orient(g, 270) == cv2.flip(cv2.transpose(g), 0)).all()
