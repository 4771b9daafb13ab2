"""The header-only C++ drop-in (include/sister/SisterMultiviewDisparities.hpp) compiles against a cv::Mat, links the
C-ABI library and -- on the GPU box -- reproduces the golden maps through the reference's own calling sequence."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_views, grey_of
from sister_b200.synth import make_rig

import sister_b200

EXE = os.path.join(ROOT, "tests", "cpp", "_build", "shim_driver")


@pytest.fixture(scope="module")
def driver():
    sister_b200.build_library()
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    libdir = os.path.dirname(sister_b200.library_path())
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "oracle", "fake_cv"), "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "shim_driver.cpp"), "-o", EXE, "-L", libdir, "-lsister_b200",
                    f"-Wl,-rpath,{libdir}"], check=True)
    return EXE


def run_driver(exe, tmp_path, name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    w, h, D = int(g["w"]), int(g["h"]), int(g["D"])
    views = golden_views(g)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    np.stack(views).astype(np.uint8).tofile(fin)
    p = subprocess.run([exe, fin, fout, str(w), str(h), str(D)], capture_output=True, text=True)
    return g, (w, h), p, fout


def test_shim_compiles_and_fails_loudly_without_a_gpu(driver, tmp_path):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    _, _, p, _ = run_driver(driver, tmp_path, "rig_64x48_d16")
    assert p.returncode == 3 and "sm_100" in p.stderr, (p.returncode, p.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["rig_64x48_d16", "rig_128x96_d64"])
def test_shim_reproduces_golden_maps(driver, tmp_path, name):
    g, (w, h), p, fout = run_driver(driver, tmp_path, name)
    assert p.returncode == 0, p.stderr
    maps = np.fromfile(fout, np.uint16).reshape(3, h, w)
    assert (maps[0] == g["disp_mv"]).all() and (maps[1] == g["disp_h"]).all() and (maps[2] == g["disp_v"]).all()
