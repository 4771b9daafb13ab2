"""The path-volume decomposition the CUDA aggregation uses (tests/sgm_spec.py) equals the reference recurrence
(oracle so_sgm, pinned against the compiled reference in test_oracle_ref.py) bit for bit. CPU only."""
import numpy as np
import pytest

from sgm_spec import combine, path_volumes, sgm_decomposed, sgm_paired, sweep_volumes


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


@pytest.mark.parametrize("h,w,D,seed", [(12, 20, 16, 1), (20, 12, 8, 2), (9, 33, 24, 3), (31, 7, 16, 4)])
def test_paired_sweeps_equal_reference_recurrence(oracle_lib, h, w, D, seed):
    """The four two-path sweeps of the CUDA aggregation (rider state handed from chain to chain between steps) reproduce the
    reference volume, and each pair volume is the sum of its two independent-chain path volumes."""
    rng = np.random.default_rng(seed)
    vol = rng.integers(0, 253, (h, w, D), dtype=np.uint16)
    vol[rng.random((h, w, D)) < 0.25] = 0
    if seed % 2:
        vol[rng.random((h, w, D)) < 0.1] = 255  # raw volumes of the two-view path carry the invalid marker (census.cpp:76)
        vol[0][rng.random((w, D)) < 0.4] = 255
        vol[-1][rng.random((w, D)) < 0.4] = 255
    assert (sgm_paired(vol) == oracle_lib.sgm(vol)).all()
    if not seed % 2:
        Q = path_volumes(vol).astype(np.int64)
        V = sweep_volumes(vol)
        for s in range(4):
            assert (V[s] == Q[2 * s] + Q[2 * s + 1]).all()
            assert V[s][1:-1].max() <= 200


def test_paired_sweeps_restricted_to_the_crop():
    """Crop-only aggregation: chains outside the region run their rider only or not at all; inside the region the pair
    volumes are those of the whole frame."""
    rng = np.random.default_rng(11)
    h, w, D = 22, 26, 8
    vol = rng.integers(0, 253, (h, w, D), dtype=np.uint16)
    roi = (5, 17, 6, 21)
    full = sweep_volumes(vol)
    crop = sweep_volumes(vol, roi)
    assert (crop[:, 5:17, 6:21] == full[:, 5:17, 6:21]).all()
    outside = crop.copy()
    outside[:, 5:17, 6:21] = 0
    assert not outside.any()
