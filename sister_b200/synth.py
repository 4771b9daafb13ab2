"""Deterministic synthetic 5-view rigs.

A rig is the set of images a SiSter robot captures by moving one camera to the centre / right / top / left /
bottom positions of a cross (reference README.md:35-47). Here the five views are rendered from one texture ``T``
and a disparity field ``d`` with the matcher's conventions (reference census.cpp:70-84 + hpp:56-70):

    center(y, x) = T(y, x)      right(y, x) = T(y, x + d)     left(y, x) = T(y, x - d)
    top(y, x)    = T(y - d, x)  bottom(y, x) = T(y + d, x)

so that the centre pixel (i, j) at disparity d matches right(i, j - d), left(i, j + d), top(i + d, j) and
bottom(i - d, j). numpy only, no GPU; the generator is a splitmix64 hash so that fixtures do not depend on numpy's
Generator streams.
"""
from __future__ import annotations

import numpy as np

VIEW_NAMES = ("center", "right", "top", "left", "bottom")


def splitmix64(seed: int, n: int) -> np.ndarray:
    """n 64-bit values of the splitmix64 sequence started at ``seed`` (vectorised)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _box5(a: np.ndarray) -> np.ndarray:
    """5x5 box blur with edge replication, integer arithmetic."""
    p = np.pad(a.astype(np.uint32), 2, mode="edge")
    cs = np.cumsum(np.cumsum(p, axis=0), axis=1)
    cs = np.pad(cs, ((1, 0), (1, 0)))
    h, w = a.shape
    s = cs[5:5 + h, 5:5 + w] - cs[0:h, 5:5 + w] - cs[5:5 + h, 0:w] + cs[0:h, 0:w]
    return ((s + 12) // 25).astype(np.uint8)


def texture(h: int, w: int, seed: int) -> np.ndarray:
    """Blurred noise with flat 128-valued patches (ties / low-confidence areas for the LRC masks)."""
    raw = (splitmix64(seed, h * w) >> np.uint64(56)).astype(np.uint8).reshape(h, w)
    t = _box5(raw)
    # stretch contrast back after the blur
    t = np.clip((t.astype(np.int32) - 128) * 4 + 128, 0, 255).astype(np.uint8)
    yy, xx = np.mgrid[0:h, 0:w]
    flat = ((yy // 37) + (xx // 53)) % 3 == 0
    flat &= ((yy // 37) + 2 * (xx // 53)) % 5 == 0
    t[flat] = 128
    return t


def disparity_field(h: int, w: int, disp_count: int, kind: str = "smooth") -> np.ndarray:
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    if kind == "plane":
        d = np.full((h, w), disp_count / 3.0)
    elif kind == "smooth":
        d = disp_count * (0.30 + 0.18 * np.sin(0.013 * yy) * np.cos(0.017 * xx))
        checker = ((yy // 90).astype(np.int64) + (xx // 120).astype(np.int64)) % 2 == 1
        d = np.where(checker, d * 0.6, d)
    elif isinstance(kind, (int, float)):
        d = np.full((h, w), float(kind))
    else:
        raise ValueError(kind)
    return np.clip(np.rint(d), 0, disp_count - 1).astype(np.int64)


def make_rig(w: int, h: int, disp_count: int, seed: int = 1234, kind="smooth", noise: int = 3, channels: int = 1, colour: bool = False):
    """Return the 5 views (center, right, top, left, bottom) as uint8 arrays H x W (or H x W x 3).

    channels = 3 replicates grey to B = G = R unless ``colour`` is set: then the three channels of every view get their own
    gain, offset and noise stream (B != G != R almost everywhere), so that the fixed-point BGR2GRAY of hpp:29-33 is
    actually exercised and a swapped or mis-weighted channel changes the result."""
    m = disp_count  # margin so that every shifted lookup stays inside the texture
    T = texture(h + 2 * m, w + 2 * m, seed)
    d = disparity_field(h, w, disp_count, kind)
    yy, xx = np.mgrid[0:h, 0:w]
    yy = yy + m
    xx = xx + m
    views = [T[yy, xx], T[yy, xx + d], T[yy - d, xx], T[yy, xx - d], T[yy + d, xx]]
    out = []
    for k, v in enumerate(views):
        if noise:
            n = (splitmix64(seed * 7919 + 13 * (k + 1), h * w) % np.uint64(2 * noise + 1)).astype(np.int32).reshape(h, w) - noise
            v = np.clip(v.astype(np.int32) + n, 0, 255).astype(np.uint8)
        v = np.ascontiguousarray(v)
        if channels == 3 and colour:
            chans = []
            for c, (gain_num, offs) in enumerate(((3, 40), (4, 0), (5, -30))):  # B, G, R: gain / 4, offset
                nc = (splitmix64(seed * 104729 + 31 * (k + 1) + 7 * (c + 1), h * w) % np.uint64(9)).astype(np.int32).reshape(h, w) - 4
                chans.append(np.clip(v.astype(np.int32) * gain_num // 4 + offs + nc, 0, 255).astype(np.uint8))
            v = np.ascontiguousarray(np.stack(chans, axis=2))
        elif channels == 3:
            v = np.ascontiguousarray(np.repeat(v[:, :, None], 3, axis=2))
        out.append(v)
    return out


def cost_evals(w: int, h: int, disp_count: int, views: int = 4, padded: bool = True) -> int:
    """Cost evaluations of one doMultiStereo (px * disp * view); padded = what hammingCost fills (census.cpp:63-88)."""
    if padded:
        return views * (w + 2 * disp_count) * (h + 2 * disp_count) * disp_count
    return views * w * h * disp_count
