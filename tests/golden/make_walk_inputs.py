"""Generate tests/golden/walk_inputs.npz by importing the reference's own pure-numpy planner.

Run in the development container only (needs /root/reference; the GPU box has no copy):

    python tests/golden/make_walk_inputs.py

What is taken from the reference *by importing it* (stub `casadi`/`matplotlib` modules, because
`utils.py:1` / `functions.py:3-9` import them at module level):
  * `FootstepPlanner` (`code/footstep_planner_vertices.py:6`): plan + 2000x6 contact reference tables,
  * `FootTrajectoryGenerator` (`code/foot_trajectory_generator.py`) -- only for `compute_knot`,
  * `compute_knot`, `built_the_reference/velocity/acceleration` (`code/functions.py:11,196-248`).
What is restated: `quintic_spline` (`functions.py:129-157`) is an equality-only feasibility NLP
(`f = 0`) solved by IPOPT from p = 0; one Newton step from 0 on a full-row-rank linear system lands
on the minimum-norm solution, so `numpy.linalg.lstsq` is used (ASSUMPTION, nothing pins it).
The measured whole-body angular momentum of the reference's own N=10 walk (`original_code/cuhw.txt`,
1962 rows; `original_code/plot.py:14-18` reads it) is stored as `hw_meas`: it is the only recorded
closed-loop signal in the reference tree and is replayed as the "measured" h_w of the surrogate plant.
Inputs that need DART (initial foot poses, robot mass) are the literals recovered from
`code/Debug/contact_trj_from_centroidal_MPC` line 1 and the URDF mass sum.
"""
import contextlib
import io
import os
import sys
import types

import numpy as np

REF = "/root/reference/code"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "walk_inputs.npz")

LFOOT0 = np.array([0.0, -1.3870180197464853e-16, -4.639952657811354e-18,
                   1.0310923973763693e-17, 0.10163857612916291, -1.3877787807814457e-17])
RFOOT0 = np.array([0.0, -1.3870180197464853e-16, 4.639952657811354e-18,
                   1.0310923973763693e-17, -0.10163857612916291, -1.3877787807814457e-17])
MASS = 40.05487735
VREF = [(0.15, 0., 0)] * 11 + [(0.13, 0, 0)] * 4 + [(0.1, 0., 0)] * 2 + [(0., 0, 0)] * 3  # simulation.py:97


def quintic_spline_minnorm(x):
    """Linear system of functions.py:135-151, minimum-norm solution."""
    n = len(x)
    rows, rhs = [], []

    def row(pairs, b):
        r = np.zeros(6 * n)
        for j, c in pairs:
            r[j] += c
        rows.append(r)
        rhs.append(b)

    for i in range(n - 1):
        row([(6 * i, 1.0)], x[i])
        row([(6 * i + j, 1.0) for j in range(6)], x[i + 1])
    row([(1, 1.0)], 0.0)
    row([(6 * (n - 1) + 1, 1.0)], 0.0)
    for i in range(n - 1):
        row([(6 * i + j, float(j)) for j in range(1, 6)] + [(6 * (i + 1) + 1, -1.0)], 0.0)
    row([(2, 2.0)], 0.0)
    for i in range(n - 1):
        row([(6 * i + 2, 2.0), (6 * i + 3, 6.0), (6 * i + 4, 12.0), (6 * i + 5, 20.0), (6 * (i + 1) + 2, -2.0)], 0.0)
    A, b = np.array(rows), np.array(rhs)
    p = np.linalg.lstsq(A, b, rcond=None)[0]
    assert np.max(np.abs(A @ p - b)) < 1e-10
    return p.reshape(-1, 1)


def main():
    for name in ["casadi", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits",
                 "mpl_toolkits.mplot3d", "dartpy"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["mpl_toolkits.mplot3d"].Axes3D = object
    sys.path.insert(0, REF)
    params = {'g': 9.81, 'h': 0.72, 'foot_size': 0.1, 'step_height': 0.02, 'world_time_step': 0.01,
              'ss_duration': 70, 'ds_duration': 30, 'first_swing': 'rfoot', 'µ': 0.5, 'N': 10,
              'mass': MASS, 'update_contact': 'YES', 'mpc_rate': 1}
    params['eta'] = np.sqrt(params['g'] / params['h'])
    initial = {'lfoot': {'pos': LFOOT0.copy(), 'vel': np.zeros(6), 'acc': np.zeros(6)},
               'rfoot': {'pos': RFOOT0.copy(), 'vel': np.zeros(6), 'acc': np.zeros(6)}}
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink):
        import footstep_planner_vertices as fpv
        import foot_trajectory_generator as ftg
        import functions as fn
        planner = fpv.FootstepPlanner(VREF, initial['lfoot']['pos'], initial['rfoot']['pos'], params)
        gen = ftg.FootTrajectoryGenerator(initial, planner, params)
        fn.plot_spline = lambda knot: None
        knot_x, knot_y, seq_x, seq_y = fn.compute_knot(gen, planner)
        co_x, co_y = quintic_spline_minnorm(knot_x), quintic_spline_minnorm(knot_y)
        ref = {}
        for ax, seq, co in (("x", seq_x, co_x), ("y", seq_y, co_y)):
            ref["pos_" + ax] = np.concatenate(fn.built_the_reference(seq, co))
            ref["vel_" + ax] = np.concatenate(fn.built_the_velocity(seq, co))
            ref["acc_" + ax] = np.concatenate(fn.built_the_acceleration(seq, co))
        nz = len(ref["pos_x"])
        ref["pos_z"] = np.full(nz, 0.72)          # functions.py:97-99
        ref["vel_z"] = np.zeros(nz)
        ref["acc_z"] = np.zeros(nz)
    plan_pos = np.array([s['pos'] for s in planner.plan])
    plan_ang = np.array([s['ang'] for s in planner.plan])
    plan_ss = np.array([s['ss_duration'] for s in planner.plan])
    plan_ds = np.array([s['ds_duration'] for s in planner.plan])
    plan_foot = np.array([0 if s['foot_id'] == 'lfoot' else 1 for s in planner.plan])
    hw_meas = np.loadtxt("/root/reference/original_code/cuhw.txt")
    np.savez_compressed(
        OUT, hw_meas=hw_meas, plan_pos=plan_pos, plan_ang=plan_ang, plan_ss=plan_ss, plan_ds=plan_ds, plan_foot=plan_foot,
        contact_left=planner.position_contacts_ref['contact_left'],
        contact_right=planner.position_contacts_ref['contact_right'],
        knot_x=np.array(knot_x), knot_y=np.array(knot_y), seq_x=np.array(seq_x), seq_y=np.array(seq_y),
        lfoot0=LFOOT0, rfoot0=RFOOT0, mass=MASS,
        **{"ref_" + k: np.asarray(v, float).ravel() for k, v in ref.items()})
    print("wrote", OUT, "len x/y/z", len(ref["pos_x"]), len(ref["pos_y"]), len(ref["pos_z"]))
    print("plan x", plan_pos[:, 0])
    print("knot_x", np.round(knot_x, 4))
    print("seq_x", seq_x)


if __name__ == "__main__":
    main()
