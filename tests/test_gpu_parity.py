"""Parity of the CUDA path (through the C ABI) against the oracle and the committed golden vectors.
Everything here is bit-exact: the whole path is integer (the reference has no sub-pixel step, postprocess.cpp:141)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_rigs, golden_views, grey_of
import sister_b200
from sister_b200.synth import make_rig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import sister_b200
    if not os.path.exists(sister_b200.library_path()):
        sister_b200.build_library()
    eng = sister_b200.Engine(256, 256, 96, n_slots=3)
    eng.set_test_taps(True)
    yield eng
    eng.close()


@pytest.mark.parametrize("name", golden_rigs())
def test_golden_rigs(engine, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    w, h, D = int(g["w"]), int(g["h"]), int(g["D"])
    views = golden_views(g)
    outs, raw = engine.compute(views, D, want_raw=True)
    for m, key in enumerate(("disp_mv", "disp_h", "disp_v")):
        assert (raw[m] == g[f"raw_disp_m{m}"]).all(), f"raw disparity differs, mode {m}: {(raw[m] != g[f'raw_disp_m{m}']).sum()} px"
        assert (outs[m] == g[key]).all(), key
    # grey input == grey replicated to BGR; single mode leaves the other outputs untouched
    outs1 = engine.compute([grey_of(v) for v in views], D, mode_mask=1)
    assert (outs1[0] == g["disp_mv"]).all() and outs1[1] is None and outs1[2] is None
    outs2 = engine.compute(views, D, mode_mask=6)
    assert outs2[0] is None and (outs2[1] == g["disp_h"]).all() and (outs2[2] == g["disp_v"]).all()


@pytest.mark.parametrize("w,h,D,kind,seed,colour", [(96, 64, 32, "smooth", 11, False), (40, 56, 8, "plane", 12, False), (72, 60, 24, "smooth", 13, False),
                                                        (48, 40, 96, "smooth", 14, False),  # D = 32, 96: lane-interleaved cells (4 chains per warp)
                                                        (88, 72, 40, "smooth", 15, True), (56, 44, 16, "smooth", 16, True)])  # B != G != R: hpp:29-33
def test_every_stage_against_oracle(engine, oracle_lib, w, h, D, kind, seed, colour):
    views = make_rig(w, h, D, seed=seed, kind=kind, channels=3, colour=colour)
    if colour:
        assert all((v[:, :, 0] != v[:, :, 2]).mean() > 0.5 for v in views)
    wp, hp = w + 2 * D, h + 2 * D
    px = wp * hp
    pads = [oracle_lib.pad_replicate(oracle_lib.grey_bgr(v), D) for v in views]
    rots = (0, 180, 90, 270)
    side = (1, 3, 2, 4)
    for mode in (2, 1, 0):
        engine.compute(views, D, mode_mask=1 << mode)
        t = oracle_lib.multistereo(pads, D, mode)
        fused = engine.fetch("fused", (hp, wp, D), np.uint8)
        assert (fused == t["fused"]).all(), f"fused volume, mode {mode}"
        ssum = engine.fetch("sum", (hp, wp, D), np.uint16)
        assert (ssum == t["sum"]).all(), f"aggregated volume, mode {mode}: {(ssum != t['sum']).sum()} cells"
        raw = engine.fetch("raw_disp", (3, hp, wp), np.int16)
        assert (raw[mode] == t["disp"]).all()
    # after mode 0 all four views were matched: compare the per-view products
    ori = engine.fetch("oriented", (8, px), np.uint8)
    cen = engine.fetch("census", (8, px), np.uint64)
    wl = engine.fetch("wta_l", (4, px), np.int16)
    wr = engine.fetch("wta_r", (4, px), np.int16)
    lr = engine.fetch("lr_final", (4, px), np.int16)
    masks = engine.fetch("masks", (4, hp, wp), np.uint8)
    for v in range(4):
        c_img = oracle_lib.orient(pads[0], rots[v])
        s_img = oracle_lib.orient(pads[side[v]], rots[v])
        assert (ori[2 * v] == c_img.ravel()).all() and (ori[2 * v + 1] == s_img.ravel()).all(), f"oriented images, view {v}"
        c1, c2 = oracle_lib.census(c_img), oracle_lib.census(s_img)
        assert (cen[2 * v] == c1.ravel()).all() and (cen[2 * v + 1] == c2.ravel()).all(), f"census, view {v}"
        L, R = oracle_lib.wta(oracle_lib.cost_volume(c1, c2, D))
        assert (wl[v] == L.ravel()).all(), f"WTA-left, view {v}"
        assert (wr[v] == R.ravel()).all(), f"WTA-right, view {v}"
        assert (lr[v] == t["lr"][v]).all(), f"median + LRC, view {v}"
    assert (masks == t["masks"]).all()


@pytest.mark.parametrize("w,h,D,seed", [(240, 240, 8, 31), (248, 240, 8, 32), (376, 248, 8, 33), (1000, 264, 16, 34), (252, 244, 8, 35), (2080, 496, 8, 36)])
def test_median_kernels_by_frame_shape(oracle_lib, w, h, D, seed):
    """The recursive median (postprocess.cpp:31-67) has two kernels: warps chained by neighbour hand-over (128 columns each) for
    padded frames of 256 .. 2048 with sides a multiple of 8, one block barrier per row otherwise. Shapes: exactly two chunks,
    a last chunk of two lanes (264 = 2 * 128 + 8), different chunk counts for the two orientations, nine chunks, one
    shape (268 x 260: multiples of 4, not of 8) that takes the barrier kernel, and one wider than 2048 (2096 x 512: eight pixels per
    lane, a last chunk of six lanes)."""
    views = make_rig(w, h, D, seed=seed, channels=1)
    wp, hp = w + 2 * D, h + 2 * D
    pads = [oracle_lib.pad_replicate(v, D) for v in views]
    t = oracle_lib.multistereo(pads, D, 0)
    with sister_b200.Engine(w, h, D, n_slots=1) as eng:
        eng.set_test_taps(True)
        eng.compute(views, D, mode_mask=1)
        lr = eng.fetch("lr_final", (4, wp * hp), np.int16)
        masks = eng.fetch("masks", (4, hp, wp), np.uint8)
    for v in range(4):
        assert (lr[v] == t["lr"][v]).all(), f"median + LRC, view {v}: {(lr[v] != t['lr'][v]).sum()} px"
    assert (masks == t["masks"]).all()


def test_one_fuse_pass_for_several_modes(engine, oracle_lib):
    """A call that asks for several modes fuses once: the horizontal, vertical and multiview volumes leave one pass
    (hpp:262-276: C_mv = C_h + C_v). The volume of the last mode run and every map must be those of the per-mode passes."""
    views = make_rig(88, 72, 40, seed=17, kind="smooth", channels=3, colour=True)
    D, hp, wp = 40, 72 + 80, 88 + 80
    pads = [oracle_lib.pad_replicate(oracle_lib.grey_bgr(v), D) for v in views]
    for mask, last in ((7, 2), (3, 1), (5, 2), (6, 2)):
        outs = engine.compute(views, D, mode_mask=mask)
        t = oracle_lib.multistereo(pads, D, last)
        fused = engine.fetch("fused", (hp, wp, D), np.uint8)
        assert (fused == t["fused"]).all(), f"fused volume of mode {last} from the shared pass (mode_mask {mask})"
        for m in range(3):
            if (mask >> m) & 1:
                single = engine.compute(views, D, mode_mask=1 << m)[m]
                assert (outs[m] == single).all(), f"mode {m} of mode_mask {mask}"


def test_sgm_known_answers(engine, oracle_lib):
    k = np.load(os.path.join(GOLDEN, "stage_kats.npz"))
    for i in range(3):
        s, disp = engine.test_sgm(k[f"sgm8_in_{i}"])
        assert (s == k[f"sgm8_out_{i}"]).all(), f"KAT {i}"
    rng = np.random.default_rng(99)
    for (h, w, D) in [(40, 52, 16), (36, 88, 32), (24, 40, 8), (30, 44, 72), (20, 36, 200), (16, 24, 136), (12, 20, 512)]:
        vol = rng.integers(0, 253, (h, w, D), dtype=np.uint8)
        vol[rng.random((h, w, D)) < 0.2] = 0
        s, disp = engine.test_sgm(vol)
        ref = oracle_lib.sgm(vol.astype(np.uint16))
        assert (s == ref).all(), f"SGM {h}x{w}x{D}: {(s != ref).sum()} cells differ"
        L, _ = oracle_lib.wta(ref)
        assert (disp == L).all(), f"final WTA {h}x{w}x{D}"


def test_shape_preconditions_are_reported(engine):
    import sister_b200
    views = make_rig(64, 48, 16, channels=1)
    for bad_d in (12, 20):
        with pytest.raises(sister_b200.SisterError) as e:
            engine.compute(views, bad_d)
        assert e.value.code == -2
    with pytest.raises(sister_b200.SisterError) as e:
        engine.compute(make_rig(62, 48, 16, channels=1), 16)  # (w + 2D) % 4 != 0 (postprocess.cpp:18)
    assert e.value.code == -2
    with pytest.raises(sister_b200.SisterError) as e:
        engine.compute(make_rig(512, 48, 16, channels=1), 16)
    assert e.value.code == -3


def test_batch_equals_single_and_is_deterministic(engine):
    rigs = [make_rig(96, 64, 32, seed=100 + k, channels=1) for k in range(7)]
    single = [engine.compute(r, 32, mode_mask=1)[0] for r in rigs]
    for _ in range(2):
        batch = engine.compute_batch(rigs, 32, mode_mask=1)
        for a, b in zip(single, batch):
            assert (a == b[0]).all()


def test_reference_class_mirror(engine):
    import sister_b200
    g = np.load(os.path.join(GOLDEN, "rig_64x48_d16.npz"))
    views = golden_views(g)
    s = sister_b200.SisterMultiviewDisparities(*views, engine=engine)
    mv, hz, vt = s.compute_disparities(16)
    assert (mv == g["disp_mv"]).all() and (hz == g["disp_h"]).all() and (vt == g["disp_v"]).all()


def test_config1_against_oracle(oracle_lib):
    """BASELINE.json configs[0]: 640x480, D = 192, the multiview map, against the CPU oracle (a few seconds)."""
    import sister_b200
    views = make_rig(640, 480, 192, seed=1234, channels=1)
    with sister_b200.Engine(640, 480, 192, n_slots=1) as eng:
        outs, raw = eng.compute(views, 192, mode_mask=1, want_raw=True)
    ref, ref_raw = oracle_lib.compute_disparities(views, 192, mode_mask=1, want_raw=True)
    assert (raw[0] == ref_raw[0]).all(), f"{(raw[0] != ref_raw[0]).sum()} padded pixels differ"
    assert (outs[0] == ref[0]).all()


@pytest.mark.parametrize("name", golden_rigs())
def test_crop_only_aggregation_reproduces_golden_maps(name):
    """The product path aggregates for the crop Rect(D, D, W, H) only (sister_set_full_frame): the three maps must be
    the reference's, and asking for the raw padded map must switch the whole frame back on for that call."""
    import sister_b200
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    w, h, D = int(g["w"]), int(g["h"]), int(g["D"])
    views = golden_views(g)
    with sister_b200.Engine(w, h, D, n_slots=2) as eng:  # no taps: crop only
        outs = eng.compute(views, D)
        assert (outs[0] == g["disp_mv"]).all() and (outs[1] == g["disp_h"]).all() and (outs[2] == g["disp_v"]).all()
        batch = eng.compute_batch([views, views, views], D, mode_mask=5)
        for r in batch:
            assert (r[0] == g["disp_mv"]).all() and (r[2] == g["disp_v"]).all()
        outs2, raw = eng.compute(views, D, want_raw=True)
        for m in range(3):
            assert (raw[m] == g[f"raw_disp_m{m}"]).all()
        for m in range(3):
            assert (outs2[m] == outs[m]).all()


@pytest.mark.parametrize("w,h,D,seed", [(100, 76, 40, 5), (52, 88, 136, 6), (1280, 960, 192, 8), (640, 480, 128, 9), (320, 240, 256, 10)])
def test_crop_only_equals_full_frame(w, h, D, seed):
    """Same maps with and without the crop restriction, at shapes where chains are cut short on every side; includes
    BASELINE.json configs[1] (1280 x 960, D = 192)."""
    import sister_b200
    views = make_rig(w, h, D, seed=seed, channels=1)
    with sister_b200.Engine(w, h, D, n_slots=1) as eng:
        crop = eng.compute(views, D, mode_mask=3 if max(w, h) > 1000 else 7)
        eng.set_full_frame(True)
        full = eng.compute(views, D, mode_mask=3 if max(w, h) > 1000 else 7)
    for a, b in zip(crop, full):
        if a is not None:
            assert (a == b).all(), f"{(a != b).sum()} pixels differ"


def test_full_size_properties():
    """BASELINE.json configs[1] shape (1280x960, D = 192): properties that need no CPU oracle."""
    import sister_b200
    D = 192
    plane = make_rig(1280, 960, D, seed=7, kind="plane", noise=0, channels=1)
    with sister_b200.Engine(1280, 960, D, n_slots=2) as eng:
        a = eng.compute(plane, D, mode_mask=7)
        b = eng.compute(plane, D, mode_mask=7)
        for m in range(3):
            assert (a[m] == b[m]).all(), "not deterministic"
            d = a[m] // 255
            assert (a[m] % 255 == 0).all()
            inner = d[32:-32, 32:-32]
            assert (inner == D // 3).mean() > 0.9, f"fronto-parallel plane not recovered in mode {m}"
        batch = eng.compute_batch([plane, plane, plane], D, mode_mask=1)
        for r in batch:
            assert (r[0] == a[0]).all()


@pytest.mark.parametrize("w,h,D,seed,modes", [(64, 48, 384, 21, 1), (32, 16, 512, 22, 1), (40, 32, 264, 23, 7),
                                                  (48, 32, 320, 24, 1)])  # D = 320: lane-interleaved cells, 10 registers per lane
def test_large_disparity_ranges_against_oracle(oracle_lib, w, h, D, seed, modes):
    """D beyond the reference's own limit (postprocess.cpp:193 overflows at D >= 272): the widened oracle is the check
    (BASELINE.json configs[3]/[4] use D = 384 and 512)."""
    import sister_b200
    views = make_rig(w, h, D, seed=seed, kind="smooth", channels=1)
    with sister_b200.Engine(w, h, D, n_slots=1) as eng:
        outs, raw = eng.compute(views, D, mode_mask=modes, want_raw=True)
    ref, ref_raw = oracle_lib.compute_disparities(views, D, mode_mask=modes, want_raw=True)  # scalar: ~1 minute in total
    for m in range(3):
        if not (modes >> m) & 1:
            continue
        assert (raw[m] == ref_raw[m]).all(), f"mode {m}: {(raw[m] != ref_raw[m]).sum()} padded pixels differ"
        assert (outs[m] == ref[m]).all()


@pytest.mark.parametrize("D", [128, 256, 512])
def test_baseline_distance_sweep_shapes(D):
    """BASELINE.json configs[4]: 1280x960 rigs at D = 128..512 (up to 2.3e9 cells, beyond int32 indexing): properties
    that need no CPU oracle -- determinism, and a fronto-parallel plane at disparity D/3 is recovered."""
    import sister_b200
    plane = make_rig(1280, 960, D, seed=5, kind="plane", noise=0, channels=1)
    with sister_b200.Engine(1280, 960, D, n_slots=1) as eng:
        a = eng.compute(plane, D, mode_mask=1)[0]
        b = eng.compute(plane, D, mode_mask=1)[0]
    assert (a == b).all()
    # the generator rounds D / 3 to the nearest pixel; disp * 255 saturates only from disp >= 258 (hpp:116-118)
    assert (a[32:-32, 32:-32] // 255 == int(round(D / 3.0))).mean() > 0.9


def test_config4_large_frame_single_gpu():
    """BASELINE.json configs[3] shape: 4096x3072, D = 384 (7.2e9 cells, 64.6 GB of volumes) on one B200."""
    import sister_b200
    D = 384
    plane = make_rig(4096, 3072, D, seed=9, kind="plane", noise=0, channels=1)
    with sister_b200.Engine(4096, 3072, D, n_slots=1) as eng:
        a = eng.compute(plane, D, mode_mask=1)[0]
        b = eng.compute(plane, D, mode_mask=1)[0]
    assert (a == b).all()
    inner = a[64:-64, 64:-64] // 255
    assert (inner == int(round(D / 3.0))).mean() > 0.9


def test_pinned_and_strided_inputs(engine):
    """Views in page-locked memory are copied to the device directly, pageable or row-padded ones are staged: same maps."""
    g = np.load(os.path.join(GOLDEN, "rig_96x64_d32.npz"))
    w, h, D = int(g["w"]), int(g["h"]), int(g["D"])
    views = golden_views(g)
    pinned = engine.host_array((5, h, w, 3), np.uint8)
    for k in range(5):
        pinned[k] = views[k]
    a = engine.compute([pinned[k] for k in range(5)], D, mode_mask=1)[0]
    assert (a == g["disp_mv"]).all()
    batch = engine.compute_batch([[pinned[k] for k in range(5)]] * 3, D, mode_mask=1)
    assert all((b[0] == g["disp_mv"]).all() for b in batch)
    # row-padded views (cv::Mat::step > w * channels): regions of interest of larger images, pageable and page-locked
    for big in (np.full((5, h + 9, w + 23, 3), 77, np.uint8), engine.host_array((5, h + 9, w + 23, 3), np.uint8)):
        big[:] = 77
        roi = [big[k, 4:4 + h, 11:11 + w] for k in range(5)]
        for k in range(5):
            roi[k][:] = views[k]
        assert roi[0].strides[0] == (w + 23) * 3 and not roi[0].flags["C_CONTIGUOUS"]
        c = engine.compute(roi, D, mode_mask=1)[0]
        assert (c == g["disp_mv"]).all(), "row-padded BGR views"
        batch = engine.compute_batch([roi, roi], D, mode_mask=1)
        assert all((b[0] == g["disp_mv"]).all() for b in batch)
    bigg = np.zeros((5, h + 2, w + 40), np.uint8)
    roig = [bigg[k, 1:1 + h, 8:8 + w] for k in range(5)]
    for k in range(5):
        roig[k][:] = grey_of(views[k])
    assert (engine.compute(roig, D, mode_mask=1)[0] == g["disp_mv"]).all(), "row-padded grey views"


def test_config2_against_oracle(oracle_lib):
    """BASELINE.json configs[1], the headline shape: 1280x960, D = 192, the multiview map (doMultiStereo mode 0,
    hpp:152-295) -- raw padded disparity and encoded map against the reference itself when oracle/_ref is present (it travels
    with the repo), else against the plain-C port."""
    import oracle
    import sister_b200
    W, H, D = 1280, 960, 192
    views = make_rig(W, H, D, seed=1234, channels=1)
    with sister_b200.Engine(W, H, D, n_slots=1) as eng:
        outs, raw = eng.compute(views, D, mode_mask=1, want_raw=True)
        crop_only = eng.compute(views, D, mode_mask=1)[0]
    pads = [oracle_lib.pad_replicate(v, D) for v in views]
    if os.path.exists(oracle.REF_SO):
        ref_raw = oracle.Ref().multistereo_taps(pads, D, mode=0, want_volumes=False)["disp"]
    else:
        ref_raw = oracle_lib.multistereo(pads, D, 0, want_volumes=False)["disp"]
    assert (raw[0] == ref_raw).all(), f"{(raw[0] != ref_raw).sum()} padded pixels differ"
    ref_map = oracle_lib.encode_crop(ref_raw, D)
    assert (outs[0] == ref_map).all() and (crop_only == ref_map).all()


def test_output_saturation_above_258(oracle_lib):
    """hpp:116-118: the uint16 map is disparity * 255 with saturation, so every disparity >= 258 reads 65535. A rig whose
    true disparity is 270 at D = 288 (beyond the reference's own D < 272 limit, postprocess.cpp:193: the widened port is
    the check) reaches it in the horizontal map."""
    import sister_b200
    w, h, D = 320, 16, 288
    views = make_rig(w, h, D, seed=31, kind=270, channels=1)
    with sister_b200.Engine(w, h, D, n_slots=1) as eng:
        outs, raw = eng.compute(views, D, mode_mask=2, want_raw=True)
    ref, ref_raw = oracle_lib.compute_disparities(views, D, mode_mask=2, want_raw=True)
    assert (raw[1] == ref_raw[1]).all()
    assert (outs[1] == ref[1]).all()
    crop = raw[1][D:D + h, D:D + w]
    assert (crop >= 258).sum() > 500 and (outs[1][crop >= 258] == 65535).all() and (outs[1][crop < 258] < 65535).all()


def test_two_contexts_on_two_devices_in_one_process():
    """One context per GPU inside ONE process: the shared-memory opt-ins of the kernels are per
    device, so the second device must work like the first. Skips on a one-GPU box."""
    import sister_b200
    g = np.load(os.path.join(GOLDEN, "rig_128x96_d64.npz"))
    views = golden_views(g)
    w, h, D = int(g["w"]), int(g["h"]), int(g["D"])
    try:
        eng1 = sister_b200.Engine(1280, 960, 192, n_slots=1, device=1)
    except sister_b200.SisterError as e:
        if e.code == -5:
            pytest.skip("one GPU")
        raise
    with eng1, sister_b200.Engine(1280, 960, 192, n_slots=1, device=0) as eng0:
        for eng in (eng0, eng1, eng0):
            outs = eng.compute(views, D)
            assert (outs[0] == g["disp_mv"]).all() and (outs[1] == g["disp_h"]).all() and (outs[2] == g["disp_v"]).all()
        # the headline shape needs the large opt-ins (k_fuse 115 KB) on both devices
        plane = make_rig(1280, 960, 192, seed=7, kind="plane", noise=0, channels=1)
        a = eng0.compute(plane, 192, mode_mask=1)[0]
        b = eng1.compute(plane, 192, mode_mask=1)[0]
        assert (a == b).all()


def test_status_word_survives_queued_submits(engine):
    """Several device submits queued on one slot before sister_sync share the slot's status word: it accumulates, and a
    clean run leaves it clean (sister_sync returns OK and the next host submit is unaffected)."""
    g = np.load(os.path.join(GOLDEN, "rig_64x48_d16.npz"))
    views = golden_views(g)
    rig = engine.upload_rig(views)
    out = engine.dev_alloc(3 * 64 * 48 * 2)
    for _ in range(3):
        engine.submit_device(1, rig, 64, 48, 3, 16, 7, [out, out + 64 * 48 * 2, out + 2 * 64 * 48 * 2])
    engine.sync(1)
    got = np.zeros((3, 48, 64), np.uint16)
    engine.dev_download(out, got)
    assert (got[0] == g["disp_mv"]).all() and (got[1] == g["disp_h"]).all() and (got[2] == g["disp_v"]).all()
    engine.dev_free(rig)
    engine.dev_free(out)
