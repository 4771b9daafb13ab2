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
