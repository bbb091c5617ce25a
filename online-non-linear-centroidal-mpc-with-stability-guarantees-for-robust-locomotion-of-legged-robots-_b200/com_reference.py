"""CoM reference tables without CasADi: the part of `functions.references`
(`code/functions.py:60-124`) that follows `compute_knot` -- quintic spline coefficients through the knots and the
position / velocity / acceleration tables the MPC samples every tick (MPC file :565-570).

Restated behaviour (file:line of the reference):
  * `quintic_spline` (`functions.py:129-157`) poses an equality-only feasibility problem (objective 0) in the 6 n
    polynomial coefficients of n segments and hands it to IPOPT from p = 0.  The constraints are linear, full row rank
    and fewer than the unknowns (4 n - 1 rows): position at both ends of the first n - 1 segments, zero velocity at the
    start of the first and of the last segment, velocity and acceleration continuity across the n - 1 joints, zero
    initial acceleration.  One Newton step of an interior-point method from the origin on such a system lands on the
    MINIMUM-NORM solution, which is what is computed here (ASSUMPTION: nothing in the reference tree pins the solver's
    answer; stated in DESIGN.md).
  * `built_the_reference / _velocity / _acceleration` (`functions.py:196-248`): segment i spans the ticks
    [sequence[i-1], sequence[i]) (the first from 0), tau = tick offset / segment length.  Quirks kept: the velocity is
    the derivative with respect to tau, NOT divided by the segment length (:222); the acceleration IS divided by the
    segment length squared (:243); the tables stop at sequence[-1], so x/z and y have different lengths.
  * z reference: constant 0.72 with zero velocity / acceleration, as long as the x table (:97-99).
  * `compute_knot` (`functions.py:11-56`): the knots are mid-points of the two feet (x) and 0.6 x the lateral position
    of the NEXT support foot (y), sampled one tick after every landing (ticks first_time_knot + ss + 1 + k (ss + ds)), with
    two leading knots from the initial stance; the segment end ticks are those ticks for x and ds - 1 ticks later for y.
    The foot positions it reads come from `FootTrajectoryGenerator.generate_feet_trajectories_at_time`
    (`foot_trajectory_generator.py:12-116`); the x / y components of that function are restated in `feet_xy_at` (initial
    pose during step 0, planned poses in double support, the cubic swing profile in single support), so the generator is
    not needed to make references for another walk.
"""
from __future__ import annotations

import numpy as np

COM_HEIGHT_REF = 0.72        # functions.py:97


def _step_index_at(plan, time):
    """`FootstepPlanner.get_step_index_at_time` (footstep_planner_vertices.py:82-88)."""
    t = 0
    for i, step in enumerate(plan):
        t += step["ss_duration"] + step["ds_duration"]
        if t > time:
            return i
    return None


def feet_xy_at(plan, initial, time):
    """x, y of both feet at tick `time`: {'lfoot': (x, y), 'rfoot': (x, y)} -- the `['pos'][3:5]` entries of
    `generate_feet_trajectories_at_time` (foot_trajectory_generator.py:12-116)."""
    idx = _step_index_at(plan, time)
    if idx == 0:                                                     # :21-35 initial poses during the first step
        return {"lfoot": tuple(np.asarray(initial["lfoot"]["pos"], float)[3:5]), "rfoot": tuple(np.asarray(initial["rfoot"]["pos"], float)[3:5])}
    start = sum(s["ss_duration"] + s["ds_duration"] for s in plan[:idx])
    tin = time - start
    support = plan[idx]["foot_id"]
    swing = "lfoot" if support == "rfoot" else "rfoot"
    T = plan[idx]["ss_duration"]
    if tin >= T:                                                     # :38-59 double support: planned poses
        return {support: tuple(np.asarray(plan[idx]["pos"], float)[0:2]), swing: tuple(np.asarray(plan[idx + 1]["pos"], float)[0:2])}
    a, b = np.asarray(plan[idx - 1]["pos"], float), np.asarray(plan[idx + 1]["pos"], float)      # :62-76 cubic swing profile
    sft = -2.0 / T ** 3 * tin ** 3 + 3.0 / T ** 2 * tin ** 2
    sw = a + (b - a) * sft
    return {support: tuple(np.asarray(plan[idx]["pos"], float)[0:2]), swing: (sw[0], sw[1])}


def compute_knot(plan, initial):
    """knot_x, knot_y, sequence_x, sequence_y of `functions.compute_knot` (functions.py:11-56), from the plan alone."""
    ss, ds = plan[2]["ss_duration"], plan[2]["ds_duration"]          # :17-18
    f0 = feet_xy_at(plan, initial, 0)
    knot_x = [(f0["lfoot"][0] + f0["rfoot"][0]) / 2, (f0["lfoot"][0] + f0["rfoot"][0]) / 2]      # :20, :23
    knot_y = [(f0["lfoot"][1] + f0["rfoot"][1]) / 2, f0[plan[1]["foot_id"]][1] * 0.6]            # :21, :24-25
    scale = ss + ds
    first = int(2 * scale)                                           # :28-29
    seq_x, seq_y = [first], [first]
    first_contact = first + ss + 1                                   # :33
    for i in range(first, len(plan) * scale - 1):                    # :35
        if (i - first_contact) % scale == 0:
            f = feet_xy_at(plan, initial, i)
            knot_x.append((f["lfoot"][0] + f["rfoot"][0]) / 2)
            seq_x.append(i)
            contact = plan[_step_index_at(plan, i) + 1]["foot_id"]   # :41-45
            knot_y.append(f[contact][1] * 0.6)
            seq_y.append(i + ds - 1)                                 # :49
    return knot_x, knot_y, seq_x, seq_y


def references(plan, initial):
    """`functions.references` (functions.py:58-124) end to end: knots from the plan, minimum-norm quintic splines, tables."""
    return references_from_knots(*compute_knot(plan, initial))


def spline_system(knots):
    """Constraint matrix and right-hand side of `quintic_spline` (rows in the reference's order)."""
    x = np.asarray(knots, float).ravel()
    n = len(x)
    m = 4 * n - 1
    A = np.zeros((m, 6 * n))
    b = np.zeros(m)
    seg = 6 * np.arange(n - 1)
    r = 2 * np.arange(n - 1)
    A[r, seg] = 1.0; b[r] = x[:-1]                                   # p_i(0) = x_i
    for j in range(6):
        A[r + 1, seg + j] = 1.0                                      # p_i(1) = x_{i+1}
    b[r + 1] = x[1:]
    k = 2 * (n - 1)
    A[k, 1] = 1.0                                                    # p_0'(0) = 0
    A[k + 1, 6 * (n - 1) + 1] = 1.0                                  # p_{n-1}'(0) = 0
    k += 2
    rows = k + np.arange(n - 1)
    for j in range(1, 6):
        A[rows, seg + j] = float(j)                                  # p_i'(1) = p_{i+1}'(0)
    A[rows, seg + 6 + 1] -= 1.0
    k += n - 1
    A[k, 2] = 2.0                                                    # p_0''(0) = 0
    k += 1
    rows = k + np.arange(n - 1)
    for j, cf in ((2, 2.0), (3, 6.0), (4, 12.0), (5, 20.0)):
        A[rows, seg + j] = cf                                        # p_i''(1) = p_{i+1}''(0)
    A[rows, seg + 6 + 2] -= 2.0
    return A, b


def quintic_coefficients(knots):
    """Minimum-norm coefficients (6 n,) of the spline through `knots`: p = A' (A A')^-1 b."""
    A, b = spline_system(knots)
    p = A.T @ np.linalg.solve(A @ A.T, b)
    if np.max(np.abs(A @ p - b)) > 1e-9 * max(1.0, np.max(np.abs(b))):
        raise np.linalg.LinAlgError("spline constraints are not satisfied (rank-deficient knot system)")
    return p


def sample_tables(sequence, coeff):
    """pos, vel, acc tables (length sequence[-1]) from the segment end ticks and the spline coefficients."""
    seq = np.asarray(sequence, int).ravel()
    c = np.asarray(coeff, float).ravel()
    starts = np.concatenate([[0], seq[:-1]])
    length = seq - starts
    if np.any(length <= 0):
        raise ValueError("segment end ticks must increase")
    seg = np.repeat(np.arange(len(seq)), length)                     # segment of every tick
    tau = (np.arange(seq[-1]) - starts[seg]) / length[seg]
    a = c[: 6 * len(seq)].reshape(len(seq), 6)[seg]                  # coefficients of every tick's segment
    t2, t3, t4, t5 = tau ** 2, tau ** 3, tau ** 4, tau ** 5
    pos = a[:, 0] + a[:, 1] * tau + a[:, 2] * t2 + a[:, 3] * t3 + a[:, 4] * t4 + a[:, 5] * t5
    vel = a[:, 1] + 2 * a[:, 2] * tau + 3 * a[:, 3] * t2 + 4 * a[:, 4] * t3 + 5 * a[:, 5] * t4          # per tau, :222
    acc = (2 * a[:, 2] + 6 * a[:, 3] * tau + 12 * a[:, 4] * t2 + 20 * a[:, 5] * t3) / (length[seg] ** 2)  # :243
    return pos, vel, acc


def references_from_knots(knot_x, knot_y, sequence_x, sequence_y):
    """The dict `functions.references` returns (`pos_x` ... `acc_z`), from the output of `compute_knot`."""
    ref = {}
    for ax, knots, seq in (("x", knot_x, sequence_x), ("y", knot_y, sequence_y)):
        pos, vel, acc = sample_tables(seq, quintic_coefficients(knots))
        ref["pos_" + ax], ref["vel_" + ax], ref["acc_" + ax] = pos, vel, acc
    n = len(ref["pos_x"])
    ref["pos_z"], ref["vel_z"], ref["acc_z"] = np.full(n, COM_HEIGHT_REF), np.zeros(n), np.zeros(n)
    return ref
