"""ORACLE (test infrastructure) -- symbolic restatement of the reference NLP.

This file restates, symbol by symbol, the CasADi `Opti` problem that the reference builds in
`code/centroidal_mpc_vertices.py:126-353` (+ `centroidal_dynamic`, :371-461).  CasADi is not
installable offline, so the role CasADi plays in the reference (expression graph + automatic
derivatives) is played here by sympy: one *stage block* i of the NLP is written once in
symbols, differentiated symbolically (gradient, Jacobians, Hessian of the block Lagrangian) and
turned into numpy callables (`lambdify`) or C code (`oracle/gen_c.py`).

Nothing in the product (`*_b200/`) imports this file.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg may.  **Parity status: unpinned** -- the reference ships no golden
vectors for the solve and CasADi/IPOPT cannot run here.

Block i (i = 0..N-1) touches v = (x_i[20], u_i[32], x_{i+1}[20], u_{i+1}[32]) and carries

  d_i  : x_{i+1} - x_i - delta*f(x_i, ref_i, gamma_i, u_i) = 0              (MPC file :187-190)
  g_i  : the inequality rows the reference adds for stage i, as g <= 0       (:218-271)
  l_i  : the summand of the cost for stage i (+ the force-rate term i)       (:311-351)

State / input layout (MPC file :149-166, :382-403):
  x = [p(3) v(3) h(3) theta(3) psi_l p_l(3) psi_r p_r(3)]
  u = [f_l1..f_l4 (12) f_r1..f_r4 (12) v_l(3) v_r(3) w_l w_r]
"""
from __future__ import annotations

import sympy as sp

NX, NU = 20, 32
NV = NX + NU + NX + NU  # block variables
MU = 0.5                # MPC file :41
FOOT_L, FOOT_W = 0.25, 0.13  # MPC file :51-52
# foot polygon in the foot frame, MPC file :55-60
FOOT_POLY = [(FOOT_L / 2, FOOT_W / 2, 0.0), (FOOT_L / 2, -FOOT_W / 2, 0.0),
             (-FOOT_L / 2, -FOOT_W / 2, 0.0), (-FOOT_L / 2, FOOT_W / 2, 0.0)]
PZ_MAX = 0.76           # MPC file :230
BOX = (0.01, 0.005, 0.00005)  # MPC file :259-271

# parameter vector of one block (order matters: used by the numeric drivers)
PARAM_NAMES = (
    ["ref%d" % j for j in range(9)] +            # com_ref[:, i]  (pos, vel, acc)         :172
    ["plr%d" % j for j in range(3)] +            # pos_contact_l_ref[:, i]                :174
    ["prr%d" % j for j in range(3)] +            # pos_contact_r_ref[:, i]                :175
    ["alr", "arr"] +                             # ang_contact_{l,r}_ref[i]               :176-177
    ["gl", "gr", "gln", "grn"] +                 # gamma_l[i], gamma_r[i], gamma_*[i+1]   :179-181
    ["wz", "wrate", "hw_on"] +                   # w_z[i] (:301-305), rate weight (:339-351), row :224 on/off
    ["mass", "k1", "k2", "delta", "grav", "eps_reg"]
)
NP = len(PARAM_NAMES)
# inequality rows of a block: Lyapunov, hw, pz, 32 friction, 8 unilateral, 12 foot box
NG = 1 + 1 + 1 + 32 + 8 + 12


def _cross(a, b):
    return sp.Matrix([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def centroidal_dynamic(x, ref, gl, gr, u, P):
    """f(x, ref, gamma_l, gamma_r, u) of MPC file :371-461 (returns a 20-vector)."""
    mass, k1, grav = P["mass"], P["k1"], P["grav"]
    p, v, th = x[0:3, 0], x[3:6, 0], x[9:12, 0]
    psi_l, p_l, psi_r, p_r = x[12], x[13:16, 0], x[16], x[17:20, 0]
    fl = [u[3 * k:3 * k + 3, 0] for k in range(4)]
    fr = [u[12 + 3 * k:12 + 3 * k + 3, 0] for k in range(4)]
    v_l, v_r, w_l, w_r = u[24:27, 0], u[27:30, 0], u[30], u[31]
    Vl = (fl[0] + fl[1] + fl[2] + fl[3]) * gl            # :405
    Vr = (fr[0] + fr[1] + fr[2] + fr[3]) * gr            # :406
    z1 = p - ref[0:3, 0]                                 # :408
    z2 = k1 * z1 + (v - ref[3:6, 0])                     # :409

    def verts(pos, yaw):                                 # :413-429
        c, s = sp.cos(yaw), sp.sin(yaw)
        Rz = sp.Matrix([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        return [Rz * sp.Matrix(ck) + pos for ck in FOOT_POLY]

    vl, vr = verts(p_l, psi_l), verts(p_r, psi_r)
    tl = sp.zeros(3, 1)
    tr = sp.zeros(3, 1)
    for k in range(4):                                   # :445-446
        tl += _cross(vl[k] - p, fl[k])
        tr += _cross(vr[k] - p, fr[k])
    tl, tr = gl * tl, gr * tr
    gvec = sp.Matrix([0, 0, -grav])
    dcom = v                                             # :452
    ddcom = gvec + (Vl + Vr + th * 0) / mass             # :453
    dhw = tl + tr                                        # :454
    return sp.Matrix.vstack(dcom, ddcom, dhw, z2 / mass,               # :459
                            sp.Matrix([(1 - gl) * w_l]), (1 - gl) * v_l,  # :455,:457
                            sp.Matrix([(1 - gr) * w_r]), (1 - gr) * v_r)  # :456,:458


def build_block():
    """Return the symbols and expressions of one stage block."""
    V = sp.Matrix(sp.symbols("v0:%d" % NV, real=True))
    Pv = sp.Matrix(sp.symbols(" ".join(PARAM_NAMES), real=True))
    P = {n: Pv[j] for j, n in enumerate(PARAM_NAMES)}
    xi, ui = V[0:NX, 0], V[NX:NX + NU, 0]
    xn, un = V[NX + NU:NX + NU + NX, 0], V[NX + NU + NX:NV, 0]
    ref = Pv[0:9, 0]
    plr, prr = Pv[9:12, 0], Pv[12:15, 0]
    alr, arr = P["alr"], P["arr"]
    gl, gr, gln, grn = P["gl"], P["gr"], P["gln"], P["grn"]
    mass, k1, k2, delta, grav = P["mass"], P["k1"], P["k2"], P["delta"], P["grav"]

    # --- dynamics defect, MPC file :188-190
    d = xn - xi - delta * centroidal_dynamic(xi, ref, gl, gr, ui, P)

    fl = [ui[3 * k:3 * k + 3, 0] for k in range(4)]
    fr = [ui[12 + 3 * k:12 + 3 * k + 3, 0] for k in range(4)]
    fln = [un[3 * k:3 * k + 3, 0] for k in range(4)]
    frn = [un[12 + 3 * k:12 + 3 * k + 3, 0] for k in range(4)]

    # --- Lyapunov row, MPC file :202-220
    z1 = xn[0:3, 0] - ref[0:3, 0]
    z2 = k1 * z1 + (xn[3:6, 0] - ref[3:6, 0])
    gvec = sp.Matrix([0, 0, -grav])
    u_n = -(k1 + k2) * z2 + k1 * k1 * z1 - gvec + ref[6:9, 0] - xi[9:12, 0] / mass   # :207-208
    Vl = (fl[0] + fl[1] + fl[2] + fl[3]) * gl / mass                                # :214
    Vr = (fr[0] + fr[1] + fr[2] + fr[3]) * gr / mass                                # :215
    lyap = (-(z1.T * (k1 * z1))[0] - (z2.T * (k2 * z2))[0] + (z1.T * z2)[0]
            + (z2.T * ((Vl + Vr) - u_n))[0])                                        # :219-220
    g = [lyap]
    # --- angular momentum row (only stage 0), MPC file :223-224 ;  hw_1'hw_1 - hw_0'hw_0 <= 0
    hw_row = P["hw_on"] * ((xn[6:9, 0].T * xn[6:9, 0])[0] - (xi[6:9, 0].T * xi[6:9, 0])[0])
    g.append(hw_row)
    # --- CoM height, :230
    g.append(xi[2] - PZ_MAX)
    # --- friction pyramid, :44-47, :236-244  (A f) * gamma <= 0
    A = sp.Matrix([[1, 0, -MU], [-1, 0, -MU], [0, 1, -MU], [0, -1, -MU]])
    for k in range(4):
        r = A * fl[k] * gl
        g += [r[j] for j in range(4)]
    for k in range(4):
        r = A * fr[k] * gr
        g += [r[j] for j in range(4)]
    # --- unilateral rows, :246-254   f_z * gamma >= 0
    for k in range(4):
        g.append(-fl[k][2] * gl)
    for k in range(4):
        g.append(-fr[k][2] * gr)
    # --- foot placement box, :258-271
    for j in range(3):
        e = (xn[13 + j] - plr[j]) * gln
        g += [e - BOX[j], -e - BOX[j]]
    for j in range(3):
        e = (xn[17 + j] - prr[j]) * grn
        g += [e - BOX[j], -e - BOX[j]]
    g = sp.Matrix(g)
    assert g.shape[0] == NG

    # --- cost summand i, :311-337
    def sumsqr(m):
        return sum(e * e for e in m)

    avg_l = sp.Rational(1, 4) * Vl * gl * mass        # :278
    avg_r = sp.Rational(1, 4) * Vr * gr * mass        # :279
    cost = 1000 * sumsqr(xi[6:9, 0])
    cost += (xn[0] - ref[0]) ** 2 + (xn[1] - ref[1]) ** 2 + P["wz"] * (xn[2] - ref[2]) ** 2
    cost += 1000 * sumsqr((xn[13:16, 0] - plr) * gln) + 1000 * sumsqr((xn[17:20, 0] - prr) * grn)
    cost += 1000 * ((xn[12] - alr) * gln) ** 2 + 1000 * ((xn[16] - arr) * grn) ** 2
    for k in range(4):
        cost += 10 * sumsqr(avg_l - fl[k]) * gl + 10 * sumsqr(avg_r - fr[k]) * gr
        cost += 10 * sumsqr(fl[k]) * (1 - gl) + 10 * sumsqr(fr[k]) * (1 - gr)
    # --- force-rate term i (:343-351); wrate = weight_f_rate for i < N-1, else 0
    for k in range(4):
        cost += P["wrate"] * ((fln[k][2] - fl[k][2]) ** 2 * gl + (frn[k][2] - fr[k][2]) ** 2 * gr)
    # --- NOT in the reference: Tikhonov term on the cost-free foot inputs u[24:32] (v_l, v_r, w_l,
    #     w_r have no cost and, for a stance foot, no effect -- MPC file :455-458).  IPOPT hides the
    #     resulting singular directions with inertia-correction; here eps_reg*|u_foot|^2 picks the
    #     minimum-norm member.  eps_reg = 0 gives the literal reference cost.
    cost += P["eps_reg"] * sumsqr(ui[24:32, 0])
    return V, Pv, d, g, cost


class BlockFunctions:
    """numpy callables of one stage block (built once, ~10 s of sympy work)."""

    def __init__(self):
        V, Pv, d, g, cost = build_block()
        self.V, self.Pv = V, Pv
        Y = sp.Matrix(sp.symbols("y0:%d" % NX, real=True))
        L = sp.Matrix(sp.symbols("l0:%d" % NG, real=True))
        args = [list(V), list(Pv)]
        grad = sp.Matrix([cost]).jacobian(V).T
        Jd = d.jacobian(V)
        Jg = g.jacobian(V)
        self.jd_idx = [(r, c) for r in range(NX) for c in range(NV) if Jd[r, c] != 0]
        self.jg_idx = [(r, c) for r in range(NG) for c in range(NV) if Jg[r, c] != 0]
        lag = cost + (Y.T * d)[0] + (L.T * g)[0]
        glag = sp.Matrix([lag]).jacobian(V)
        H = glag.jacobian(V)
        self.h_idx = [(r, c) for r in range(NV) for c in range(NV) if H[r, c] != 0]
        mods = "numpy"
        self.f_cost = sp.lambdify(args, cost, mods, cse=True)
        self.f_grad = sp.lambdify(args, list(grad), mods, cse=True)
        self.f_d = sp.lambdify(args, list(d), mods, cse=True)
        self.f_g = sp.lambdify(args, list(g), mods, cse=True)
        self.f_jd = sp.lambdify(args, [Jd[r, c] for r, c in self.jd_idx], mods, cse=True)
        self.f_jg = sp.lambdify(args, [Jg[r, c] for r, c in self.jg_idx], mods, cse=True)
        self.f_h = sp.lambdify(args + [list(Y), list(L)], [H[r, c] for r, c in self.h_idx], mods, cse=True)
        # symbolic objects kept for the C generator
        self.sym = dict(V=V, Pv=Pv, Y=Y, L=L, d=d, g=g, cost=cost, grad=grad, Jd=Jd, Jg=Jg, H=H)


_CACHE = None


def block_functions() -> BlockFunctions:
    global _CACHE
    if _CACHE is None:
        _CACHE = BlockFunctions()
    return _CACHE
