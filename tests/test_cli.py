"""tools/compute_disp.cpp (the counterpart of the reference's sample executable, compute_disp.cpp) builds, fails loudly
without a GPU, and on the GPU box writes the golden maps and the same pictures cv2 derives from them."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_views, grey_of
from sister_b200.synth import VIEW_NAMES, make_rig

import sister_b200

EXE = os.path.join(ROOT, "tests", "cpp", "_build", "compute_disp")


@pytest.fixture(scope="module")
def cli():
    sister_b200.build_library()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.dirname(sister_b200.library_path())
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", os.path.join(ROOT, "tools", "compute_disp.cpp"), "-o", EXE,
                    "-L", libdir, "-lsister_b200", f"-Wl,-rpath,{libdir}"], check=True)
    return EXE


def write_rig(folder, views):
    for name, v in zip(VIEW_NAMES, views):
        with open(os.path.join(folder, name + ".ppm"), "wb") as f:
            f.write(b"P6\n%d %d\n255\n" % (v.shape[1], v.shape[0]))
            f.write(np.ascontiguousarray(v[:, :, ::-1]).tobytes())  # files are RGB, the views are BGR


def read_pnm(path):
    with open(path, "rb") as f:
        magic = f.readline().strip()
        w, h = map(int, f.readline().split())
        maxv = int(f.readline())
        data = f.read()
    if magic == b"P5" and maxv == 65535:
        return np.frombuffer(data, ">u2").reshape(h, w).astype(np.uint16)
    assert magic == b"P6"
    return np.frombuffer(data, np.uint8).reshape(h, w, 3)[:, :, ::-1]  # -> BGR


def test_cli_builds_and_fails_loudly_without_a_gpu(cli, tmp_path):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    write_rig(str(tmp_path), make_rig(64, 48, 16, seed=1, channels=3))
    p = subprocess.run([cli, str(tmp_path) + "/", "16"], capture_output=True, text=True)
    assert p.returncode == 1 and "sm_100" in p.stderr
    p = subprocess.run([cli, str(tmp_path) + "/nothing/", "16"], capture_output=True, text=True)
    assert p.returncode == 1 and "cannot read" in p.stderr


@pytest.mark.gpu
def test_cli_outputs(cli, tmp_path):
    cv2 = pytest.importorskip("cv2")
    g = np.load(os.path.join(GOLDEN, "rig_128x96_d64.npz"))
    w, h, D = int(g["w"]), int(g["h"]), int(g["D"])
    views = golden_views(g)
    folder = str(tmp_path) + "/"
    write_rig(folder, views)
    p = subprocess.run([cli, folder, str(D)], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    for name, key in (("disp_multiview", "disp_mv"), ("disp_horizontal", "disp_h"), ("disp_vertical", "disp_v")):
        m = read_pnm(folder + name + ".pgm")
        assert (m == g[key]).all(), name
        want = cv2.applyColorMap(cv2.normalize(m, None, 0, 255, cv2.NORM_MINMAX, cv2.CV_8UC1), cv2.COLORMAP_MAGMA)
        assert (read_pnm(folder + name + ".ppm") == want).all(), name + " colour map"
        if name == "disp_multiview":
            blend = cv2.addWeighted(views[0], 0.1, want, 0.9, 0.0)
            assert (read_pnm(folder + "blended.ppm") == blend).all()
