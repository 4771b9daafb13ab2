"""The path-volume decomposition the CUDA aggregation uses (tests/sgm_spec.py) equals the reference recurrence
(oracle so_sgm, pinned against the compiled reference in test_oracle_ref.py) bit for bit. CPU only."""
import numpy as np
import pytest

from sgm_spec import combine, path_volumes, sgm_decomposed


@pytest.mark.parametrize("h,w,D,seed", [(12, 20, 16, 1), (20, 12, 8, 2), (9, 33, 24, 3), (31, 7, 16, 4), (16, 16, 40, 5)])
def test_decomposition_equals_reference_recurrence(oracle_lib, h, w, D, seed):
    rng = np.random.default_rng(seed)
    vol = rng.integers(0, 253, (h, w, D), dtype=np.uint16)
    vol[rng.random((h, w, D)) < 0.25] = 0
    if seed % 2:
        vol[:, :, : D // 2] = np.minimum(vol[:, :, : D // 2], 40)  # smooth-ish region: exercises the P1 branches
    ref = oracle_lib.sgm(vol)
    got = sgm_decomposed(vol)
    assert (got == ref).all(), f"{(got != ref).sum()} cells differ"


def test_path_terms_fit_a_byte_and_pairs_too(oracle_lib):
    rng = np.random.default_rng(7)
    vol = rng.integers(0, 253, (14, 18, 16), dtype=np.uint16)
    Q = path_volumes(vol)
    inner = Q[:, 1:-1]
    assert inner.max() <= 100  # P2 bounds every penalty term off the first lines (sgm.cpp:282-297)
    assert (Q[1:4, 0] == 0).all() and (Q[5:8, -1] == 0).all()  # r1..r3 do not contribute on a first line
