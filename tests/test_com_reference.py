"""CoM reference tables without CasADi against the fixture made by importing the reference's own
`compute_knot` / `built_the_*` (tests/golden/make_walk_inputs.py)."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def walk():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "walk_inputs.npz")))


def test_tables_match_the_reference_sampling(pkg, walk):
    from cmpc_b200.com_reference import references_from_knots
    ref = references_from_knots(walk["knot_x"], walk["knot_y"], walk["seq_x"], walk["seq_y"])
    for k in ("pos_x", "vel_x", "acc_x", "pos_y", "vel_y", "acc_y", "pos_z", "vel_z", "acc_z"):
        assert len(ref[k]) == len(walk["ref_" + k]), k                      # x / z: 1971 ticks, y: 2000 (kept quirk)
        assert np.abs(ref[k] - walk["ref_" + k]).max() <= 1e-11, k


def test_spline_system_is_the_references(pkg, walk):
    """Row count 4 n - 1, full row rank, knots interpolated, C2 joints, and the minimum-norm property."""
    from cmpc_b200.com_reference import quintic_coefficients, spline_system
    x = walk["knot_x"]
    n = len(x)
    A, b = spline_system(x)
    assert A.shape == (4 * n - 1, 6 * n) and np.linalg.matrix_rank(A) == 4 * n - 1
    p = quintic_coefficients(x)
    assert np.abs(A @ p - b).max() <= 1e-10
    c = p.reshape(n, 6)
    assert np.abs(c[:-1, 0] - x[:-1]).max() <= 1e-12 and np.abs(c[:-1].sum(axis=1) - x[1:]).max() <= 1e-10
    null = np.linalg.svd(A)[2][4 * n - 1:]                                   # any other solution is longer
    assert np.abs(null @ p).max() <= 1e-9


def test_sampling_quirks(pkg):
    from cmpc_b200.com_reference import sample_tables
    coeff = np.array([1.0, 2.0, 3.0, 0, 0, 0,  0.5, 0, 0, 0, 0, 1.0])
    pos, vel, acc = sample_tables([4, 6], coeff)
    assert len(pos) == 6
    assert pos[1] == 1.0 + 2.0 * 0.25 + 3.0 * 0.25 ** 2
    assert vel[1] == 2.0 + 2 * 3.0 * 0.25                                     # d/dtau, not divided by the 4 ticks
    assert acc[1] == 2 * 3.0 / 16.0                                           # divided by ticks squared
    assert pos[5] == 0.5 + 0.5 ** 5 and acc[5] == 20 * 0.5 ** 3 / 4.0
    with pytest.raises(ValueError):
        sample_tables([4, 4], coeff)
