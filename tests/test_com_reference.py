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


def test_compute_knot_restated_from_the_plan(pkg, walk):
    """`compute_knot` (functions.py:11-56) restated on the plan alone reproduces the knots / segment ticks the reference's own
    compute_knot produced with its foot trajectory generator (fixture made by importing the reference), and the whole
    `references` chain reproduces the tables."""
    from cmpc_b200.com_reference import compute_knot, references
    plan = [{"pos": walk["plan_pos"][j], "ang": walk["plan_ang"][j], "ss_duration": int(walk["plan_ss"][j]), "ds_duration": int(walk["plan_ds"][j]),
             "foot_id": "lfoot" if int(walk["plan_foot"][j]) == 0 else "rfoot"} for j in range(len(walk["plan_ss"]))]
    initial = {"lfoot": {"pos": walk["lfoot0"]}, "rfoot": {"pos": walk["rfoot0"]}}
    kx, ky, sx, sy = compute_knot(plan, initial)
    assert list(sx) == list(walk["seq_x"]) and list(sy) == list(walk["seq_y"])
    assert np.array_equal(np.array(kx), walk["knot_x"]) and np.array_equal(np.array(ky), walk["knot_y"])
    ref = references(plan, initial)
    for k in ("pos_x", "vel_y", "acc_x", "pos_z"):
        assert np.abs(ref[k] - walk["ref_" + k]).max() <= 1e-11, k


def test_feet_positions_restated(pkg, walk):
    """x / y of `generate_feet_trajectories_at_time` against the reference's precomputed foot tables (code/Debug/Pos {L,R}foot
    pre trj reproduce these bit-exactly): the contact tables hold the planned poses in stance."""
    from cmpc_b200.com_reference import feet_xy_at
    plan = [{"pos": walk["plan_pos"][j], "ang": walk["plan_ang"][j], "ss_duration": int(walk["plan_ss"][j]), "ds_duration": int(walk["plan_ds"][j]),
             "foot_id": "lfoot" if int(walk["plan_foot"][j]) == 0 else "rfoot"} for j in range(len(walk["plan_ss"]))]
    initial = {"lfoot": {"pos": walk["lfoot0"]}, "rfoot": {"pos": walk["rfoot0"]}}
    f = feet_xy_at(plan, initial, 50)
    assert f["lfoot"] == tuple(walk["lfoot0"][3:5]) and f["rfoot"] == tuple(walk["rfoot0"][3:5])
    f = feet_xy_at(plan, initial, 285)                               # double support of step 1: support plan[1], landed plan[2]
    sup = plan[1]["foot_id"]; sw = "lfoot" if sup == "rfoot" else "rfoot"
    assert f[sup] == tuple(plan[1]["pos"][0:2]) and f[sw] == tuple(plan[2]["pos"][0:2])
    f = feet_xy_at(plan, initial, 335)                               # mid swing of step 2: half way between plan[1] and plan[3]
    sup = plan[2]["foot_id"]; sw = "lfoot" if sup == "rfoot" else "rfoot"
    assert abs(f[sw][0] - 0.5 * (plan[1]["pos"][0] + plan[3]["pos"][0])) < 1e-12
