import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.build(ref=os.path.isdir("/root/reference"))
    return oracle.Oracle()


@pytest.fixture(scope="session")
def ref_lib():
    import oracle
    if not os.path.exists(oracle.REF_SO):
        if not os.path.isdir("/root/reference"):
            pytest.skip("oracle/_ref not built and no reference tree")
        oracle.build(ref=True)
    return oracle.Ref()


def golden_rigs():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith("rig_") and f.endswith(".npz"))


def golden_views(g, channels=3):
    """The rig a golden record was generated from (scripts/gen_golden.py): seeds, not pixels, are stored."""
    from sister_b200.synth import make_rig
    colour = bool(int(g["colour"])) if "colour" in g.files else False
    return make_rig(int(g["w"]), int(g["h"]), int(g["D"]), seed=int(g["seed"]), kind=str(g["kind"]), channels=channels, colour=colour)


def grey_of(view):
    """OpenCV-4 BGR2GRAY (hpp:29-33), numpy: what a caller who converts first would pass as 1 channel."""
    import numpy as np
    if view.ndim == 2:
        return view
    v = view.astype(np.int64)
    return ((3735 * v[:, :, 0] + 19235 * v[:, :, 1] + 9798 * v[:, :, 2] + 16384) >> 15).astype(np.uint8)
