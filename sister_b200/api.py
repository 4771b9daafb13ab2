"""ctypes binding of include/sister_b200.h plus the Python mirror of the reference helper class."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsister_b200.so")

MODE_MULTIVIEW, MODE_HORIZONTAL, MODE_VERTICAL, MODE_ALL = 1, 2, 4, 7
STAGE_NAMES = ("prep", "census", "match", "mask", "fuse", "aggregate", "select")
TAPS = dict(oriented=0, census=1, wta_l=2, wta_r=3, lr_final=4, masks=5, fused=6, sum=7, raw_disp=8)

_u8p = C.POINTER(C.c_uint8)
_u16p = C.POINTER(C.c_uint16)
_i16p = C.POINTER(C.c_int16)

# every symbol include/sister_b200.h declares (tests/test_abi.py checks the library exports all of them)
ABI_SYMBOLS = (
    "sister_create", "sister_destroy", "sister_compute", "sister_compute_batch", "sister_submit", "sister_wait",
    "sister_submit_device", "sister_sync", "sister_dev_alloc", "sister_dev_free", "sister_dev_memset", "sister_ipc_export", "sister_ipc_open", "sister_ipc_close", "sister_host_alloc", "sister_host_free", "sister_dev_upload",
    "sister_dev_download", "sister_set_profiling", "sister_region_begin", "sister_region_end", "sister_get_stage_ms", "sister_get_stage_launches",
    "sister_get_launch_count", "sister_debug_fetch", "sister_set_test_taps", "sister_set_full_frame",
    "sister_stereo", "sister_create_band", "sister_band_state_bytes", "sister_band_submit", "sister_band_share_bytes", "sister_band_submit_share", "sister_band_submit_rest", "sister_band_rows", "sister_band_columns", "sister_band_columns_wait", "sister_band_vertical", "sister_band_finish", "sister_test_sgm", "sister_strerror", "sister_last_error",
    "sister_version",
)


class SisterError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"sister_b200 error {code}: {text}")
        self.code = code


def library_path() -> str:
    return _LIB_PATH


def build_library(verbose: bool = False) -> str:
    """Compile sister_b200/csrc/*.cu for sm_100a into the in-tree shared library (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def load_library():
    """Load the CUDA library. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise FileNotFoundError(f"{_LIB_PATH} not built: run `make -C sister_b200/csrc` (or __graft_entry__.build()); "
                                "sister_b200 has no CPU fallback")
    L = C.CDLL(_LIB_PATH)
    vp = C.c_void_p
    L.sister_create.restype = C.c_int
    L.sister_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.sister_create_band.restype = C.c_int
    L.sister_create_band.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.sister_destroy.argtypes = [vp]
    L.sister_compute.restype = C.c_int
    L.sister_compute.argtypes = [vp, C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_uint,
                                 C.POINTER(_u16p), _i16p]
    L.sister_compute_batch.restype = C.c_int
    L.sister_compute_batch.argtypes = [vp, C.c_int, C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                       C.c_uint, C.POINTER(_u16p)]
    L.sister_submit.restype = C.c_int
    L.sister_submit.argtypes = [vp, C.c_int, C.POINTER(_u8p), C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_uint]
    L.sister_wait.restype = C.c_int
    L.sister_wait.argtypes = [vp, C.c_int, C.POINTER(_u16p), _i16p]
    L.sister_submit_device.restype = C.c_int
    L.sister_submit_device.argtypes = [vp, C.c_int, C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]
    L.sister_sync.restype = C.c_int
    L.sister_sync.argtypes = [vp, C.c_int]
    L.sister_dev_alloc.restype = C.c_int
    L.sister_dev_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.sister_dev_free.restype = C.c_int
    L.sister_dev_free.argtypes = [vp, vp]
    L.sister_dev_memset.restype = C.c_int
    L.sister_dev_memset.argtypes = [vp, vp, C.c_int, C.c_size_t]
    L.sister_ipc_export.restype = C.c_int
    L.sister_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.sister_ipc_open.restype = C.c_int
    L.sister_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.sister_ipc_close.restype = C.c_int
    L.sister_ipc_close.argtypes = [vp, vp]
    L.sister_host_alloc.restype = C.c_int
    L.sister_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
    L.sister_host_free.restype = C.c_int
    L.sister_host_free.argtypes = [vp, vp]
    L.sister_dev_upload.restype = C.c_int
    L.sister_dev_upload.argtypes = [vp, vp, vp, C.c_size_t]
    L.sister_dev_download.restype = C.c_int
    L.sister_dev_download.argtypes = [vp, vp, vp, C.c_size_t]
    L.sister_set_profiling.argtypes = [vp, C.c_int]
    L.sister_region_begin.restype = C.c_int
    L.sister_region_begin.argtypes = [vp]
    L.sister_region_end.restype = C.c_int
    L.sister_region_end.argtypes = [vp, C.POINTER(C.c_float)]
    L.sister_get_stage_ms.restype = C.c_int
    L.sister_get_stage_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.c_int]
    L.sister_get_stage_launches.restype = C.c_int
    L.sister_get_stage_launches.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.c_int]
    L.sister_get_launch_count.restype = C.c_uint64
    L.sister_get_launch_count.argtypes = [vp]
    L.sister_debug_fetch.restype = C.c_int
    L.sister_debug_fetch.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t]
    L.sister_set_test_taps.restype = C.c_int
    L.sister_set_test_taps.argtypes = [vp, C.c_int]
    L.sister_set_full_frame.restype = C.c_int
    L.sister_set_full_frame.argtypes = [vp, C.c_int]
    L.sister_stereo.restype = C.c_int
    L.sister_stereo.argtypes = [vp, _u8p, _u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.sister_band_state_bytes.restype = C.c_size_t
    L.sister_band_state_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
    L.sister_band_submit.restype = C.c_int
    L.sister_band_submit.argtypes = [vp, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.sister_band_share_bytes.restype = C.c_size_t
    L.sister_band_share_bytes.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
    L.sister_band_submit_share.restype = C.c_int
    L.sister_band_submit_share.argtypes = [vp, C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, vp]
    L.sister_band_submit_rest.restype = C.c_int
    L.sister_band_submit_rest.argtypes = [vp, C.c_int, vp, C.c_int]
    L.sister_band_rows.restype = C.c_int
    L.sister_band_rows.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_uint]
    L.sister_band_columns.restype = C.c_int
    L.sister_band_columns.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.sister_band_columns_wait.restype = C.c_int
    L.sister_band_columns_wait.argtypes = [vp, C.c_int]
    L.sister_band_vertical.restype = C.c_int
    L.sister_band_vertical.argtypes = [vp, C.c_int, C.c_int, vp, vp]
    L.sister_band_finish.restype = C.c_int
    L.sister_band_finish.argtypes = [vp, C.c_int, vp]
    L.sister_test_sgm.restype = C.c_int
    L.sister_test_sgm.argtypes = [vp, _u8p, C.c_int, C.c_int, C.c_int, _u16p, _i16p]
    L.sister_strerror.restype = C.c_char_p
    L.sister_strerror.argtypes = [C.c_int]
    L.sister_last_error.restype = C.c_char_p
    L.sister_last_error.argtypes = [vp]
    L.sister_version.restype = C.c_int
    _lib = L
    return L


def _views_ptrs(views):
    keep = [np.ascontiguousarray(v, dtype=np.uint8) for v in views]
    arr = (_u8p * len(keep))(*[k.ctypes.data_as(_u8p) for k in keep])
    return arr, keep


def _views_ptrs_strided(views):
    """Pointers + the common row stride (cv::Mat::step) of views whose rows are dense but may be padded -- an ROI of a
    larger image, what `Mat(rect)` hands the reference class. Falls back to dense copies when the views do not share
    one such stride."""
    vs = [np.asarray(v) for v in views]
    ch = 3 if vs[0].ndim == 3 else 1
    inner = (ch, 1) if ch == 3 else (1,)
    step = vs[0].strides[0]
    if all(v.dtype == np.uint8 and v.strides[1:] == inner and v.strides[0] == step and step >= v.shape[1] * ch for v in vs):
        arr = (_u8p * len(vs))(*[C.cast(v.ctypes.data, _u8p) for v in vs])
        return arr, vs, step
    arr, keep = _views_ptrs(vs)
    return arr, keep, vs[0].shape[1] * ch


class Engine:
    """One context on one GPU: ``n_slots`` rigs in flight, sized for rigs up to max_w x max_h x max_disp."""

    def __init__(self, max_w: int, max_h: int, max_disp: int, n_slots: int = 1, device: int = 0, max_band_rows: int = 0):
        """max_band_rows > 0: a context for row bands only (sister_create_band): the volumes hold that many rows of the
        padded frame instead of all of them."""
        self.lib = load_library()
        self.ctx = C.c_void_p()
        self.n_slots = n_slots
        if max_band_rows > 0:
            self._chk(self.lib.sister_create_band(C.byref(self.ctx), device, max_w, max_h, max_disp, n_slots, max_band_rows), create=True)
        else:
            self._chk(self.lib.sister_create(C.byref(self.ctx), device, max_w, max_h, max_disp, n_slots), create=True)
        self._dev_allocs = []
        self._host_allocs = []

    # -- plumbing
    def _chk(self, rc: int, create: bool = False):
        if rc != 0:
            detail = self.lib.sister_strerror(rc).decode()
            if not create and self.ctx:
                detail += ": " + self.lib.sister_last_error(self.ctx).decode()
            raise SisterError(rc, detail)

    def close(self):
        if getattr(self, "ctx", None):
            for p in getattr(self, "_host_allocs", []):
                self.lib.sister_host_free(self.ctx, C.c_void_p(p))
            self._host_allocs = []
            self.lib.sister_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def _shape(views):
        v0 = views[0]
        h, w = v0.shape[:2]
        ch = 3 if v0.ndim == 3 else 1
        return w, h, ch

    # -- the hot path
    def compute(self, views, disp_count: int, mode_mask: int = MODE_ALL, want_raw: bool = False):
        """views: center, right, top, left, bottom (H x W x 3 BGR or H x W grey, uint8). Returns [mv, horiz, vert]."""
        w, h, ch = self._shape(views)
        arr, keep, step = _views_ptrs_strided(views)
        outs = [np.zeros((h, w), np.uint16) if (mode_mask >> k) & 1 else None for k in range(3)]
        oarr = (_u16p * 3)(*[o.ctypes.data_as(_u16p) if o is not None else None for o in outs])
        raw = np.zeros((3, h + 2 * disp_count, w + 2 * disp_count), np.int16) if want_raw else None
        self._chk(self.lib.sister_compute(self.ctx, arr, w, h, ch, step, disp_count, mode_mask, oarr,
                                          raw.ctypes.data_as(_i16p) if want_raw else None))
        return (outs, raw) if want_raw else outs

    def compute_batch(self, rigs, disp_count: int, mode_mask: int = MODE_MULTIVIEW, outs=None):
        """rigs: list of 5-view lists, one shape. Returns list of [mv, horiz, vert] per rig."""
        n = len(rigs)
        w, h, ch = self._shape(rigs[0])
        flat = [v for rig in rigs for v in rig]
        arr, keep, step = _views_ptrs_strided(flat)
        if outs is None:
            outs = [[np.zeros((h, w), np.uint16) if (mode_mask >> k) & 1 else None for k in range(3)] for _ in range(n)]
        oarr = (_u16p * (3 * n))(*[o.ctypes.data_as(_u16p) if o is not None else None for rig in outs for o in rig])
        self._chk(self.lib.sister_compute_batch(self.ctx, n, arr, w, h, ch, step, disp_count, mode_mask, oarr))
        return outs

    def submit(self, slot: int, views, disp_count: int, mode_mask: int = MODE_ALL):
        w, h, ch = self._shape(views)
        arr, keep = _views_ptrs(views)
        self._chk(self.lib.sister_submit(self.ctx, slot, arr, w, h, ch, w * ch, disp_count, mode_mask))
        self._last = (w, h, disp_count, mode_mask)

    def wait(self, slot: int, shape, mode_mask: int = MODE_ALL):
        h, w = shape
        outs = [np.zeros((h, w), np.uint16) if (mode_mask >> k) & 1 else None for k in range(3)]
        oarr = (_u16p * 3)(*[o.ctypes.data_as(_u16p) if o is not None else None for o in outs])
        self._chk(self.lib.sister_wait(self.ctx, slot, oarr, None))
        return outs

    # -- memory shared with the other processes of the box (row bands)
    def dev_memset(self, ptr: int, value: int, nbytes: int):
        self._chk(self.lib.sister_dev_memset(self.ctx, C.c_void_p(ptr), value, nbytes))

    def ipc_export(self, ptr: int) -> bytes:
        buf = C.create_string_buffer(64)
        self._chk(self.lib.sister_ipc_export(self.ctx, C.c_void_p(ptr), buf))
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._chk(self.lib.sister_ipc_open(self.ctx, C.create_string_buffer(handle, 64), C.byref(p)))
        return p.value

    def ipc_close(self, ptr: int):
        self._chk(self.lib.sister_ipc_close(self.ctx, C.c_void_p(ptr)))

    # -- device-resident path (bench `value`)
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        self._chk(self.lib.sister_dev_alloc(self.ctx, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, ptr: int):
        self._chk(self.lib.sister_dev_free(self.ctx, C.c_void_p(ptr)))

    def host_array(self, shape, dtype=np.uint8) -> np.ndarray:
        """A numpy array in page-locked host memory (sister_host_alloc); views passed from it skip the staging memcpy.
        The memory belongs to the engine and is released by close()."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._chk(self.lib.sister_host_alloc(self.ctx, nbytes, C.byref(p)))
        self._host_allocs.append(p.value)
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def dev_upload(self, ptr: int, a: np.ndarray):
        a = np.ascontiguousarray(a)
        self._chk(self.lib.sister_dev_upload(self.ctx, C.c_void_p(ptr), a.ctypes.data_as(C.c_void_p), a.nbytes))

    def dev_download(self, ptr: int, a: np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        self._chk(self.lib.sister_dev_download(self.ctx, a.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), a.nbytes))

    def upload_rig(self, views) -> int:
        """Copy a rig to one dense device buffer (5 views back to back); returns the device pointer."""
        stack = np.ascontiguousarray(np.stack([np.ascontiguousarray(v, dtype=np.uint8) for v in views]))
        p = self.dev_alloc(stack.nbytes)
        self.dev_upload(p, stack)
        return p

    def submit_device(self, slot: int, rig_ptr: int, w: int, h: int, channels: int, disp_count: int,
                      mode_mask: int, out_ptrs):
        vb = w * h * channels
        varr = (C.c_void_p * 5)(*[C.c_void_p(rig_ptr + k * vb) for k in range(5)])
        oarr = (C.c_void_p * 3)(*[C.c_void_p(p) if p else None for p in out_ptrs])
        self._chk(self.lib.sister_submit_device(self.ctx, slot, varr, w, h, channels, disp_count, mode_mask, oarr))

    def sync(self, slot: int = -1):
        self._chk(self.lib.sister_sync(self.ctx, slot))

    def stereo(self, center, side, disp_count: int):
        """The reference's two-view path (doStereo, hpp:122-150) on grey uint8 images: returns (left, right) float32 maps,
        left = LR-checked disparity of `center` (-10 = rejected), right = median-filtered disparity of `side`."""
        c = np.ascontiguousarray(center, dtype=np.uint8)
        r = np.ascontiguousarray(side, dtype=np.uint8)
        h, w = c.shape
        outL = np.zeros((h, w), np.float32)
        outR = np.zeros((h, w), np.float32)
        fp = C.POINTER(C.c_float)
        self._chk(self.lib.sister_stereo(self.ctx, c.ctypes.data_as(_u8p), r.ctypes.data_as(_u8p), w, h, w, disp_count,
                                         outL.ctypes.data_as(fp), outR.ctypes.data_as(fp)))
        return outL, outR

    # -- row bands: one frame over several GPUs (sister_b200/bands.py drives these)
    def band_state_bytes(self, w: int, h: int, disp_count: int) -> int:
        return int(self.lib.sister_band_state_bytes(w, h, disp_count))

    def band_submit(self, slot: int, rig_ptr: int, w: int, h: int, channels: int, disp_count: int, mode: int, row0: int, row1: int):
        vb = w * h * channels
        varr = (C.c_void_p * 5)(*[C.c_void_p(rig_ptr + k * vb) for k in range(5)])
        self._chk(self.lib.sister_band_submit(self.ctx, slot, varr, w, h, channels, disp_count, mode, row0, row1))

    def band_share_bytes(self, w: int, h: int, disp_count: int, n_shares: int) -> int:
        return int(self.lib.sister_band_share_bytes(w, h, disp_count, n_shares))

    def band_submit_share(self, slot: int, rig_ptr: int, w: int, h: int, channels: int, disp_count: int, mode: int, row0: int, row1: int,
                          share: int, n_shares: int, share_out_ptr: int):
        """Staging, census and this share of the raw-cost WTA rows (packed into share_out_ptr); band_submit_rest continues."""
        vb = w * h * channels
        varr = (C.c_void_p * 5)(*[C.c_void_p(rig_ptr + k * vb) for k in range(5)])
        self._chk(self.lib.sister_band_submit_share(self.ctx, slot, varr, w, h, channels, disp_count, mode, row0, row1, share, n_shares,
                                                    C.c_void_p(share_out_ptr)))

    def band_submit_rest(self, slot: int, shares_ptr: int, n_shares: int):
        self._chk(self.lib.sister_band_submit_rest(self.ctx, slot, C.c_void_p(shares_ptr), n_shares))

    def band_rows(self, slot: int, passes: int, in0: int, out0: int, in1: int, out1: int, tag: int):
        """Row sweeps of the band with their rider states streamed to / from the neighbouring bands' mailboxes (0 = none)."""
        p = lambda x: C.c_void_p(x) if x else None  # noqa: E731
        self._chk(self.lib.sister_band_rows(self.ctx, slot, passes, p(in0), p(out0), p(in1), p(out1), tag))

    def band_columns(self, slot: int, pass_: int, state_in_ptr: int, state_out_ptr: int):
        self._chk(self.lib.sister_band_columns(self.ctx, slot, pass_, C.c_void_p(state_in_ptr) if state_in_ptr else None,
                                               C.c_void_p(state_out_ptr) if state_out_ptr else None))

    def band_columns_wait(self, slot: int):
        self._chk(self.lib.sister_band_columns_wait(self.ctx, slot))

    def band_vertical(self, slot: int, pass_: int, state_in_ptr: int, state_out_ptr: int):
        self._chk(self.lib.sister_band_vertical(self.ctx, slot, pass_, C.c_void_p(state_in_ptr) if state_in_ptr else None,
                                                C.c_void_p(state_out_ptr) if state_out_ptr else None))

    def band_finish(self, slot: int, out_ptr: int):
        self._chk(self.lib.sister_band_finish(self.ctx, slot, C.c_void_p(out_ptr)))

    # -- measurement / taps
    def set_profiling(self, on: bool):
        self.lib.sister_set_profiling(self.ctx, int(on))

    def set_test_taps(self, on: bool):
        """Keep the aggregated volume of later submits for fetch("sum", ...) (tests only; 2 * cells bytes per slot)."""
        self._chk(self.lib.sister_set_test_taps(self.ctx, int(on)))

    def set_full_frame(self, on: bool):
        """Aggregate the whole padded frame on later submits instead of the crop the caller sees (needed for raw_disp)."""
        self._chk(self.lib.sister_set_full_frame(self.ctx, int(on)))

    def region_begin(self):
        self._chk(self.lib.sister_region_begin(self.ctx))

    def region_end(self) -> float:
        ms = C.c_float()
        self._chk(self.lib.sister_region_end(self.ctx, C.byref(ms)))
        return float(ms.value)

    def stage_ms(self, slot: int = 0):
        ms = (C.c_float * len(STAGE_NAMES))()
        self._chk(self.lib.sister_get_stage_ms(self.ctx, slot, ms, len(STAGE_NAMES)))
        return dict(zip(STAGE_NAMES, [float(x) for x in ms]))

    def stage_launches(self, slot: int = 0):
        n = (C.c_int * len(STAGE_NAMES))()
        self._chk(self.lib.sister_get_stage_launches(self.ctx, slot, n, len(STAGE_NAMES)))
        return dict(zip(STAGE_NAMES, [int(x) for x in n]))

    def launch_count(self) -> int:
        return int(self.lib.sister_get_launch_count(self.ctx))

    def fetch(self, what: str, shape, dtype, slot: int = 0) -> np.ndarray:
        out = np.zeros(shape, dtype)
        self._chk(self.lib.sister_debug_fetch(self.ctx, slot, TAPS[what], out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def test_sgm(self, fused: np.ndarray):
        """fused: uint8 [h][w][D] -> (sum uint16 [h][w][D], disp int16 [h][w])."""
        h, w, D = fused.shape
        f = np.ascontiguousarray(fused, dtype=np.uint8)
        s = np.zeros((h, w, D), np.uint16)
        disp = np.zeros((h, w), np.int16)
        self._chk(self.lib.sister_test_sgm(self.ctx, f.ctypes.data_as(_u8p), w, h, D, s.ctypes.data_as(_u16p), disp.ctypes.data_as(_i16p)))
        return s, disp


class SisterMultiviewDisparities:
    """Mirror of the reference helper class (hpp:18-26): construct with the five views, call compute_disparities.

    >>> s = SisterMultiviewDisparities(center, right, top, left, bottom)
    >>> disp_multiview, disp_horizontal, disp_vertical = s.compute_disparities(192)

    Inputs are H x W x 3 uint8 BGR arrays as cv2.imread returns them (compute_disp.cpp:19-23); outputs are H x W uint16
    holding disparity * 255 (hpp:116-118). Unlike the reference this raises SisterError instead of aborting when the
    implicit shape preconditions are violated, and prints nothing.
    """

    def __init__(self, center, right, top, left, bottom, engine: Engine | None = None, device: int = 0):
        self.views = [center, right, top, left, bottom]
        shapes = {np.asarray(v).shape for v in self.views}
        if len(shapes) != 1:
            raise ValueError("the five views must have identical shapes")
        self._engine = engine
        self._device = device

    def compute_disparities(self, dispCount: int):
        h, w = np.asarray(self.views[0]).shape[:2]
        eng = self._engine or Engine(w, h, dispCount, n_slots=1, device=self._device)
        try:
            outs = eng.compute(self.views, dispCount, MODE_ALL)
        finally:
            if self._engine is None:
                eng.close()
        return outs[0], outs[1], outs[2]
