#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libsister_ref.so).

Run in the authoring container (where /root/reference exists):  python scripts/gen_golden.py
The reference ships no golden vectors; these fixtures pin the oracle and the CUDA path to
outputs of the reference itself. Inputs are regenerated from seeds (sister_b200/synth.py), so only outputs,
compact intermediates and sha256 digests of the big volumes are stored.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from sister_b200.synth import make_rig  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, W, H, D, seed, kind, colour)
RIGS = [
    ("rig_64x48_d16", 64, 48, 16, 1234, "smooth", False),
    ("rig_40x56_d8", 40, 56, 8, 1235, "plane", False),       # portrait
    ("rig_96x64_d32", 96, 64, 32, 1236, "smooth", False),
    ("rig_160x120_d32", 160, 120, 32, 1237, "smooth", False),
    ("rig_128x96_d64", 128, 96, 64, 1238, "smooth", False),
    ("rig_72x60_d24", 72, 60, 24, 1239, "plane", False),      # D not a multiple of 16/32
    # true colour (B != G != R): the fixed-point BGR2GRAY of hpp:29-33 on what cv::imread hands over (compute_disp.cpp:19-23)
    ("rig_96x64_d32_colour", 96, 64, 32, 1240, "smooth", True),
    ("rig_80x72_d24_colour", 80, 72, 24, 1241, "smooth", True),
    # padded 264 x 256: wide enough for the chunked median kernel (two full chunks + a last chunk of two lanes one way, exactly
    # two chunks the other way); small D keeps the fixture small
    ("rig_248x240_d8", 248, 240, 8, 1242, "smooth", False),
]


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    oracle.build(ref=True)
    ref = oracle.Ref()
    orc = oracle.Oracle()
    os.makedirs(OUT, exist_ok=True)
    for name, w, h, D, seed, kind, colour in RIGS:
        views = make_rig(w, h, D, seed=seed, kind=kind, channels=3, colour=colour)
        mv, hz, vt = ref.compute_disparities(views, D)
        pads = [orc.pad_replicate(orc.grey_bgr(v), D) for v in views]
        rec = dict(w=w, h=h, D=D, seed=seed, kind=str(kind), colour=int(colour), disp_mv=mv, disp_h=hz, disp_v=vt,
                   input_sha=np.array([sha(v) for v in views]))
        for mode in range(3):
            t = ref.multistereo_taps(pads, D, mode)
            assert (orc.encode_crop(t["disp"], D) == (mv, hz, vt)[mode]).all()
            rec[f"raw_disp_m{mode}"] = t["disp"].astype(np.int16)
            rec[f"fused_sha_m{mode}"] = np.array(sha(t["fused"]))
            rec[f"sum_sha_m{mode}"] = np.array(sha(t["sum"]))
            if mode == 0:
                rec["masks"] = np.packbits(t["masks"], axis=None)
                rec["lr"] = t["lr"].astype(np.int16)  # 4 x (hp*wp): left maps after median + LRC, view frames
        # census of the padded centre view (right-view orientation) incl. the bit-63 carry
        rec["census_center_sha"] = np.array(sha(ref.census(pads[0])))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **rec)
        print(name, "ok", os.path.getsize(os.path.join(OUT, name + ".npz")), "bytes")

    # stage KATs on random data (catch the quirks that never reach the disparity)
    rng = np.random.default_rng(20261017)
    kat = {}
    for k, (h, w, D) in enumerate([(20, 24, 16), (18, 28, 8), (24, 20, 24)]):
        vol = rng.integers(0, 1021, (h, w, D), dtype=np.uint16)
        vol[rng.random((h, w, D)) < 0.05] = 255
        vol[0][rng.random((w, D)) < 0.3] = 255
        vol[-1][rng.random((w, D)) < 0.3] = 255
        kat[f"sgm_in_{k}"] = vol
        kat[f"sgm_out_{k}"] = ref.sgm(vol)
        vol8 = rng.integers(0, 253, (h, w, D), dtype=np.uint16)
        kat[f"sgm8_in_{k}"] = vol8.astype(np.uint8)
        kat[f"sgm8_out_{k}"] = ref.sgm(vol8)
        a = rng.integers(0, 256, (h, w), dtype=np.uint8)
        b = rng.integers(0, 256, (h, w), dtype=np.uint8)
        kat[f"img_a_{k}"] = a
        kat[f"img_b_{k}"] = b
        kat[f"census_a_{k}"] = ref.census(a)
        cv = ref.ad_census(a, b, D)
        kat[f"cost_{k}"] = cv.astype(np.uint8)
        L, R = ref.wta(cv)
        kat[f"wtaL_{k}"] = L.astype(np.int16)
        kat[f"wtaR_{k}"] = R.astype(np.int16)
        Lm, Rm = ref.median_inplace(L), ref.median_inplace(R)
        kat[f"medL_{k}"] = Lm.astype(np.int16)
        kat[f"medR_{k}"] = Rm.astype(np.int16)
        kat[f"lrc_{k}"] = ref.lrcheck(Lm, Rm, 5).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "stage_kats.npz"), **kat)
    print("stage_kats ok", os.path.getsize(os.path.join(OUT, "stage_kats.npz")), "bytes")


# (W, H, D, seed): two-view pairs (center, right) for doStereo (hpp:122-150), frame = image, W % 4 == H % 4 == 0
STEREO = [(96, 64, 32, 2001), (72, 60, 24, 2002), (128, 96, 64, 2003), (64, 80, 8, 2004)]


def stereo():
    """tests/golden/stereo_pairs.npz: the reference's two-view path composed from its own functions in doStereo's order
    (oracle.Ref.do_stereo), plus SGM known answers on uint8 volumes that contain the 255 marker (the first-line
    255 -> 0 substitution of sgm.cpp:109,123,146, which the fused volumes of the 5-view path never exercise)."""
    oracle.build(ref=True)
    ref = oracle.Ref()
    rec = {}
    for k, (w, h, D, seed) in enumerate(STEREO):
        views = make_rig(w, h, D, seed=seed, kind="smooth", channels=1)
        L, R = ref.do_stereo(views[0], views[1], D)
        rec[f"shape_{k}"] = np.array([w, h, D, seed])
        rec[f"left_{k}"] = L.astype(np.int16)
        rec[f"right_{k}"] = R.astype(np.int16)
    rng = np.random.default_rng(20261018)
    for k, (h, w, D) in enumerate([(20, 24, 16), (14, 36, 40), (24, 20, 192)]):
        vol = rng.integers(0, 253, (h, w, D), dtype=np.uint16)
        vol[rng.random((h, w, D)) < 0.15] = 255
        vol[0][rng.random((w, D)) < 0.4] = 255
        vol[-1][rng.random((w, D)) < 0.4] = 255
        rec[f"sgm255_in_{k}"] = vol.astype(np.uint8)
        rec[f"sgm255_out_{k}"] = ref.sgm(vol)
    np.savez_compressed(os.path.join(OUT, "stereo_pairs.npz"), **rec)
    print("stereo_pairs ok", os.path.getsize(os.path.join(OUT, "stereo_pairs.npz")), "bytes")


if __name__ == "__main__":
    if "--stereo-only" in sys.argv:
        stereo()
    else:
        main()
        stereo()
