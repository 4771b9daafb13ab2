"""TEST INFRASTRUCTURE ONLY: ctypes loaders for the two CPU checkers.

* ``Ref``    -- oracle/_ref/libsister_ref.so: the UNMODIFIED reference (CVLAB-Unibo/sister) compiled in place
               by oracle/Makefile against the fake OpenCV shim. Present in this container (built from
               /root/reference) and shipped to the GPU box as a prebuilt, git-ignored artefact.
* ``Oracle`` -- oracle/_build/libsister_oracle.so: our plain-C restatement (oracle/sister_oracle.c).

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this package. The product (``sister_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsister_ref.so")
ORACLE_SO = os.path.join(HERE, "_build", "libsister_oracle.so")

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_u64p = C.POINTER(C.c_uint64)
_f32p = C.POINTER(C.c_float)


def build(ref: bool = True, quiet: bool = True) -> None:
    """Compile the oracle (always) and, when the reference tree is present, oracle/_ref."""
    target = ["all"] if ref else ["oracle"]
    subprocess.run(["make", "-C", HERE, "-f", os.path.join(HERE, "Makefile")] + target,
                   check=True, stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a: np.ndarray | None, typ):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(typ)


def _views_array(views):
    arr = (_u8p * 5)()
    keep = []
    for k, v in enumerate(views):
        v = np.ascontiguousarray(v, dtype=np.uint8)
        keep.append(v)
        arr[k] = v.ctypes.data_as(_u8p)
    return arr, keep


class Ref:
    """The reference itself (hpp:22-119 end to end, plus the L2 free functions as stage taps)."""

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(path)
        L = self.lib
        L.ref_compute_disparities.restype = C.c_int
        L.ref_compute_disparities.argtypes = [C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, _u16p, _u16p, _u16p, C.c_int]
        L.ref_census.argtypes = [_u8p, C.c_int, C.c_int, _u64p]
        L.ref_ad_census.argtypes = [_u8p, _u8p, C.c_int, C.c_int, C.c_int, _u16p]
        L.ref_wta.argtypes = [_u16p, C.c_int, C.c_int, C.c_int, _f32p, _f32p]
        L.ref_median_inplace.argtypes = [_f32p, C.c_int, C.c_int]
        L.ref_lrcheck.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.ref_sgm.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u16p, _u16p]
        L.ref_multistereo_taps.restype = C.c_int
        L.ref_multistereo_taps.argtypes = [C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u16p, _u16p, _f32p, _f32p]

    def compute_disparities(self, views_bgr, disp_count: int, quiet: bool = True):
        """views_bgr: 5 arrays H x W x 3 uint8 (center, right, top, left, bottom) -> (mv, horiz, vert) uint16 H x W."""
        h, w = views_bgr[0].shape[:2]
        arr, keep = _views_array(views_bgr)
        outs = [np.zeros((h, w), np.uint16) for _ in range(3)]
        rc = self.lib.ref_compute_disparities(arr, w, h, disp_count, *[_ptr(o, _u16p) for o in outs], int(quiet))
        if rc != 0:
            raise RuntimeError(f"ref_compute_disparities failed rc={rc}")
        return tuple(outs)

    def census(self, img):
        h, w = img.shape
        out = np.zeros((h, w), np.uint64)
        self.lib.ref_census(_ptr(np.ascontiguousarray(img), _u8p), w, h, _ptr(out, _u64p))
        return out

    def ad_census(self, im1, im2, D):
        h, w = im1.shape
        out = np.zeros((h, w, D), np.uint16)
        self.lib.ref_ad_census(_ptr(np.ascontiguousarray(im1), _u8p), _ptr(np.ascontiguousarray(im2), _u8p), h, w, D, _ptr(out, _u16p))
        return out

    def wta(self, vol):
        h, w, D = vol.shape
        L = np.zeros((h, w), np.float32)
        R = np.zeros((h, w), np.float32)
        self.lib.ref_wta(_ptr(np.ascontiguousarray(vol), _u16p), w, h, D, _ptr(L, _f32p), _ptr(R, _f32p))
        return L, R

    def median_inplace(self, img):
        out = np.ascontiguousarray(img, dtype=np.float32).copy()
        h, w = out.shape
        self.lib.ref_median_inplace(_ptr(out, _f32p), w, h)
        return out

    def lrcheck(self, L, R, thr=5):
        out = np.ascontiguousarray(L, dtype=np.float32).copy()
        h, w = out.shape
        self.lib.ref_lrcheck(_ptr(out, _f32p), _ptr(np.ascontiguousarray(R, dtype=np.float32), _f32p), w, h, thr)
        return out

    def sgm(self, vol, img=None):
        h, w, D = vol.shape
        if img is None:
            img = np.zeros((h, w), np.uint8)
        out = np.zeros((h, w, D), np.uint16)
        self.lib.ref_sgm(_ptr(np.ascontiguousarray(img), _u8p), h, w, D, _ptr(np.ascontiguousarray(vol), _u16p), _ptr(out, _u16p))
        return out

    def do_stereo(self, center, side, D):
        """The reference's two-view path, composed from its own functions in the order of doStereo (hpp:122-150):
        ad_census (hpp:132), sgm (hpp:135), WTA left / right on the aggregated volume (hpp:137-138), in-place median on
        both maps (hpp:139-140), doLRCheck with threshold 5 (hpp:143). Returns (left, right) float32 maps."""
        S = self.sgm(self.ad_census(center, side, D))
        L, R = self.wta(S)
        L, R = self.median_inplace(L), self.median_inplace(R)
        return self.lrcheck(L, R, 5), R

    def multistereo_taps(self, views_padded, D, mode=0, want_volumes=True):
        hp, wp = views_padded[0].shape
        arr, keep = _views_array(views_padded)
        masks = np.zeros((4, hp, wp), np.uint8)
        fused = np.zeros((hp, wp, D), np.uint16) if want_volumes else None
        ssum = np.zeros((hp, wp, D), np.uint16) if want_volumes else None
        disp = np.zeros((hp, wp), np.float32)
        lr = np.zeros((4, hp * wp), np.float32)
        rc = self.lib.ref_multistereo_taps(arr, wp, hp, D, mode, _ptr(masks, _u8p), _ptr(fused, _u16p), _ptr(ssum, _u16p), _ptr(disp, _f32p), _ptr(lr, _f32p))
        if rc != 0:
            raise RuntimeError(f"ref_multistereo_taps rc={rc}")
        return dict(masks=masks, fused=fused, sum=ssum, disp=disp, lr=lr)


class Oracle:
    """The plain-C restatement (oracle/sister_oracle.c)."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = C.CDLL(path)
        L = self.lib
        L.so_grey_bgr.argtypes = [_u8p, C.c_int, C.c_int, C.c_size_t, _u8p]
        L.so_pad_replicate.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.so_orient.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, _u8p]
        L.so_encode_crop.argtypes = [_f32p, C.c_int, C.c_int, C.c_int, _u16p]
        L.so_census.argtypes = [_u8p, C.c_int, C.c_int, _u64p]
        L.so_cost_volume.argtypes = [_u64p, _u64p, C.c_int, C.c_int, C.c_int, _u16p]
        L.so_wta_left.argtypes = [_u16p, C.c_int, C.c_int, C.c_int, _f32p]
        L.so_wta_right.argtypes = [_u16p, C.c_int, C.c_int, C.c_int, _f32p]
        L.so_median_inplace.argtypes = [_f32p, C.c_int, C.c_int]
        L.so_lrcheck.argtypes = [_f32p, _f32p, C.c_int, C.c_int, C.c_int]
        L.so_sgm.restype = C.c_int
        L.so_sgm.argtypes = [_u16p, C.c_int, C.c_int, C.c_int, _u16p]
        L.so_multistereo.restype = C.c_int
        L.so_multistereo.argtypes = [C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, C.c_int, _u8p, _u16p, _u16p, _f32p, _f32p]
        L.so_compute_disparities.restype = C.c_int
        L.so_compute_disparities.argtypes = [C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_uint,
                                             C.POINTER(_u16p), _f32p]

    def grey_bgr(self, bgr):
        h, w = bgr.shape[:2]
        out = np.zeros((h, w), np.uint8)
        b = np.ascontiguousarray(bgr)
        self.lib.so_grey_bgr(_ptr(b, _u8p), w, h, w * 3, _ptr(out, _u8p))
        return out

    def pad_replicate(self, g, D):
        h, w = g.shape
        out = np.zeros((h + 2 * D, w + 2 * D), np.uint8)
        self.lib.so_pad_replicate(_ptr(np.ascontiguousarray(g), _u8p), w, h, D, _ptr(out, _u8p))
        return out

    def orient(self, x, rot):
        h, w = x.shape
        out = np.zeros((h, w) if rot in (0, 180) else (w, h), np.uint8)
        self.lib.so_orient(_ptr(np.ascontiguousarray(x), _u8p), w, h, rot, _ptr(out, _u8p))
        return out

    def encode_crop(self, disp, D):
        hp, wp = disp.shape
        out = np.zeros((hp - 2 * D, wp - 2 * D), np.uint16)
        self.lib.so_encode_crop(_ptr(np.ascontiguousarray(disp, dtype=np.float32), _f32p), wp, hp, D, _ptr(out, _u16p))
        return out

    def census(self, img):
        h, w = img.shape
        out = np.zeros((h, w), np.uint64)
        self.lib.so_census(_ptr(np.ascontiguousarray(img), _u8p), w, h, _ptr(out, _u64p))
        return out

    def cost_volume(self, c1, c2, D):
        h, w = c1.shape
        out = np.zeros((h, w, D), np.uint16)
        self.lib.so_cost_volume(_ptr(np.ascontiguousarray(c1), _u64p), _ptr(np.ascontiguousarray(c2), _u64p), h, w, D, _ptr(out, _u16p))
        return out

    def ad_census(self, im1, im2, D):
        return self.cost_volume(self.census(im1), self.census(im2), D)

    def wta(self, vol):
        h, w, D = vol.shape
        v = np.ascontiguousarray(vol)
        L = np.zeros((h, w), np.float32)
        R = np.zeros((h, w), np.float32)
        self.lib.so_wta_left(_ptr(v, _u16p), w, h, D, _ptr(L, _f32p))
        self.lib.so_wta_right(_ptr(v, _u16p), w, h, D, _ptr(R, _f32p))
        return L, R

    def median_inplace(self, img):
        out = np.ascontiguousarray(img, dtype=np.float32).copy()
        h, w = out.shape
        self.lib.so_median_inplace(_ptr(out, _f32p), w, h)
        return out

    def lrcheck(self, L, R, thr=5):
        out = np.ascontiguousarray(L, dtype=np.float32).copy()
        h, w = out.shape
        self.lib.so_lrcheck(_ptr(out, _f32p), _ptr(np.ascontiguousarray(R, dtype=np.float32), _f32p), w, h, thr)
        return out

    def sgm(self, vol):
        h, w, D = vol.shape
        out = np.zeros((h, w, D), np.uint16)
        rc = self.lib.so_sgm(_ptr(np.ascontiguousarray(vol), _u16p), h, w, D, _ptr(out, _u16p))
        if rc != 0:
            raise MemoryError("so_sgm")
        return out

    def do_stereo(self, center, side, D):
        """The reference's two-view path, composed from its own functions in the order of doStereo (hpp:122-150):
        ad_census (hpp:132), sgm (hpp:135), WTA left / right on the aggregated volume (hpp:137-138), in-place median on
        both maps (hpp:139-140), doLRCheck with threshold 5 (hpp:143). Returns (left, right) float32 maps."""
        S = self.sgm(self.ad_census(center, side, D))
        L, R = self.wta(S)
        L, R = self.median_inplace(L), self.median_inplace(R)
        return self.lrcheck(L, R, 5), R

    def multistereo(self, views_padded, D, mode=0, want_volumes=True):
        hp, wp = views_padded[0].shape
        arr, keep = _views_array(views_padded)
        masks = np.zeros((4, hp, wp), np.uint8)
        fused = np.zeros((hp, wp, D), np.uint16) if want_volumes else None
        ssum = np.zeros((hp, wp, D), np.uint16) if want_volumes else None
        disp = np.zeros((hp, wp), np.float32)
        lr = np.zeros((4, hp * wp), np.float32)
        rc = self.lib.so_multistereo(arr, wp, hp, D, mode, _ptr(masks, _u8p), _ptr(fused, _u16p), _ptr(ssum, _u16p), _ptr(disp, _f32p), _ptr(lr, _f32p))
        if rc != 0:
            raise RuntimeError(f"so_multistereo rc={rc}")
        return dict(masks=masks, fused=fused, sum=ssum, disp=disp, lr=lr)

    def compute_disparities(self, views, disp_count: int, mode_mask: int = 7, want_raw: bool = False):
        """views: 5 arrays H x W x 3 (BGR) or H x W (grey). Returns list of 3 uint16 maps (None if bit clear)."""
        v0 = views[0]
        h, w = v0.shape[:2]
        ch = 3 if v0.ndim == 3 else 1
        arr, keep = _views_array(views)
        outs = [np.zeros((h, w), np.uint16) if (mode_mask >> k) & 1 else None for k in range(3)]
        oarr = (_u16p * 3)()
        for k in range(3):
            oarr[k] = _ptr(outs[k], _u16p) if outs[k] is not None else None
        raw = np.zeros((3, h + 2 * disp_count, w + 2 * disp_count), np.float32) if want_raw else None
        rc = self.lib.so_compute_disparities(arr, w, h, ch, w * ch, disp_count, mode_mask, oarr, _ptr(raw, _f32p))
        if rc != 0:
            raise RuntimeError(f"so_compute_disparities rc={rc}")
        return (outs, raw) if want_raw else outs
