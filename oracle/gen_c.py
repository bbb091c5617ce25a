"""ORACLE (test infrastructure) -- C code generator for the stage functions of the C oracle (oracle/ipm_c.c).

The reference evaluates its NLP through CasADi's SX virtual machine with automatic derivatives
(`expand=True`, `code/centroidal_mpc_vertices.py:127`).  Here the same role is played by sympy: the stage
functions are written in symbols (re-using the literal restatement of `oracle/spec.py`), rewritten in
stage-wise form by SYMBOLIC SUBSTITUTION of the linear part of the dynamics (so the Lyapunov and
angular-momentum rows, :193-224, depend on (x_i, u_i) only), differentiated symbolically and printed as C.
Nothing is shared with the hand-derived derivatives of the CUDA path.

    python -m oracle.gen_c          # writes oracle/_gen/stage_gen.c  (committed, ~1 min of sympy work)

Stage variable z = [u(32) ; x(28)], x = [p v h theta psi_l p_l psi_r p_r | q(8)] with q = previous vertex f_z
(the force-rate term :343-351 becomes a stage cost).
"""
from __future__ import annotations

import os

import sympy as sp

from .spec import BOX, MU, PZ_MAX, centroidal_dynamic

NXA, NU, NZ, NG = 28, 32, 60, 55
PNAMES = (["ref%d" % j for j in range(9)] + ["refp%d" % j for j in range(9)] + ["frp%d" % j for j in range(8)] +
          ["gl", "gr", "glp", "grp", "wz", "has_u", "has_track", "hw_on", "w_rate", "mass", "k1", "delta", "grav", "eps_reg"])
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_gen", "stage_gen.c")


def build():
    Z = sp.Matrix(sp.symbols("z0:%d" % NZ, real=True))
    Pv = sp.Matrix(sp.symbols(" ".join(PNAMES), real=True))
    P = {n: Pv[j] for j, n in enumerate(PNAMES)}
    u, x = Z[0:NU, 0], Z[NU:NZ, 0]
    xs = x[0:20, 0]
    q = x[20:28, 0]
    ref, refp, frp = Pv[0:9, 0], Pv[9:18, 0], Pv[18:26, 0]
    gl, gr, glp, grp = P["gl"], P["gr"], P["glp"], P["grp"]
    mass, k1, delta, grav = P["mass"], P["k1"], P["delta"], P["grav"]
    # dynamics (MPC file :188-190 with f of :371-461) + q+ = f_z
    f = centroidal_dynamic(xs, ref, gl, gr, u, {"mass": mass, "k1": k1, "grav": grav})
    phi = sp.Matrix.vstack(xs + delta * f, sp.Matrix([u[3 * v + 2] for v in range(8)]))
    fl = [u[3 * k:3 * k + 3, 0] for k in range(4)]
    fr = [u[12 + 3 * k:12 + 3 * k + 3, 0] for k in range(4)]
    # Lyapunov row (:202-220) with x_{i+1}[0:6] := phi[0:6]  (k2 cancels; written with k2 = 0)
    z1 = phi[0:3, 0] - ref[0:3, 0]
    z2 = k1 * z1 + (phi[3:6, 0] - ref[3:6, 0])
    gvec = sp.Matrix([0, 0, -grav])
    u_n = -k1 * z2 + k1 * k1 * z1 - gvec + ref[6:9, 0] - xs[9:12, 0] / mass
    Vl = (fl[0] + fl[1] + fl[2] + fl[3]) * gl / mass
    Vr = (fr[0] + fr[1] + fr[2] + fr[3]) * gr / mass
    lyap = (-(z1.T * (k1 * z1))[0] + (z1.T * z2)[0] + (z2.T * ((Vl + Vr) - u_n))[0])
    g = [lyap]
    g.append(P["hw_on"] * ((phi[6:9, 0].T * phi[6:9, 0])[0] - (xs[6:9, 0].T * xs[6:9, 0])[0]))      # :224
    g.append(xs[2] - PZ_MAX)                                                                      # :230
    A = sp.Matrix([[1, 0, -MU], [-1, 0, -MU], [0, 1, -MU], [0, -1, -MU]])                         # :44-47
    for k in range(4):
        g += list(A * fl[k])
    for k in range(4):
        g += list(A * fr[k])
    g += [-fl[k][2] for k in range(4)] + [-fr[k][2] for k in range(4)]                            # :246-254
    for e in range(2):                                                                            # :258-271 (on x_i, ref col i-1)
        for j in range(3):
            err = xs[(13 if e == 0 else 17) + j] - frp[3 * e + j]
            g += [err - BOX[j], -err - BOX[j]]
    g = sp.Matrix(g)
    assert g.shape[0] == NG

    def sumsqr(m):
        return sum(e * e for e in m)

    # cost: tracking of x_i with reference column i-1 (:313-319) ...
    wz = P["wz"]
    track = ((xs[0] - refp[0]) ** 2 + (xs[1] - refp[1]) ** 2 + wz * (xs[2] - refp[2]) ** 2
             + 1000 * gl * sumsqr(xs[13:16, 0] - frp[0:3, 0]) + 1000 * gr * sumsqr(xs[17:20, 0] - frp[3:6, 0])
             + 1000 * gl * (xs[12] - frp[6]) ** 2 + 1000 * gr * (xs[16] - frp[7]) ** 2)
    # ... and the input terms of stage i (:312, :320-335), the rate term i-1 (:343-351), the Tikhonov term
    avg_l = (fl[0] + fl[1] + fl[2] + fl[3]) / 4
    avg_r = (fr[0] + fr[1] + fr[2] + fr[3]) / 4
    inp = 1000 * sumsqr(xs[6:9, 0])
    for k in range(4):
        inp += 10 * gl * sumsqr(avg_l - fl[k]) + 10 * gr * sumsqr(avg_r - fr[k])
        inp += 10 * (1 - gl) * sumsqr(fl[k]) + 10 * (1 - gr) * sumsqr(fr[k])
        inp += P["has_track"] * P["w_rate"] * (glp * (fl[k][2] - q[k]) ** 2 + grp * (fr[k][2] - q[4 + k]) ** 2)
    inp += P["eps_reg"] * sumsqr(u[24:32, 0])
    cost = P["has_track"] * track + P["has_u"] * inp
    return Z, Pv, phi, g, cost


def main():
    Z, Pv, phi, g, cost = build()
    Y = sp.Matrix(sp.symbols("y0:%d" % NXA, real=True))
    L = sp.Matrix(sp.symbols("l0:%d" % NG, real=True))
    grad = sp.Matrix([cost]).jacobian(Z).T
    Jphi = phi.jacobian(Z)
    Jg = g.jacobian(Z)
    lag = cost + (Y.T * phi)[0] + (L.T * g)[0]
    H = sp.Matrix([lag]).jacobian(Z).jacobian(Z)
    jphi_idx = [(r, c) for r in range(NXA) for c in range(NZ) if Jphi[r, c] != 0]
    jg_idx = [(r, c) for r in range(NG) for c in range(NZ) if Jg[r, c] != 0]
    h_idx = [(r, c) for r in range(NZ) for c in range(r + 1) if H[r, c] != 0]
    outs = ([cost] + list(grad) + list(phi) + [Jphi[r, c] for r, c in jphi_idx] + list(g) +
            [Jg[r, c] for r, c in jg_idx] + [H[r, c] for r, c in h_idx])
    repl, red = sp.cse(outs, symbols=sp.numbered_symbols("t"), optimizations="basic")
    names = (["*cost"] + ["grad[%d]" % j for j in range(NZ)] + ["phi[%d]" % j for j in range(NXA)] +
             ["jphi[%d]" % j for j in range(len(jphi_idx))] + ["g[%d]" % j for j in range(NG)] +
             ["jg[%d]" % j for j in range(len(jg_idx))] + ["hess[%d]" % j for j in range(len(h_idx))])
    sub = {Z[j]: sp.Symbol("z[%d]" % j) for j in range(NZ)}
    sub.update({Pv[j]: sp.Symbol("p[%d]" % j) for j in range(len(PNAMES))})
    sub.update({Y[j]: sp.Symbol("y[%d]" % j) for j in range(NXA)})
    sub.update({L[j]: sp.Symbol("lam[%d]" % j) for j in range(NG)})
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as fh:
        fh.write("/* GENERATED by oracle/gen_c.py (sympy %s) -- do not edit.  Stage functions of the C oracle. */\n" % sp.__version__)
        fh.write("#include <math.h>\n")
        fh.write("#define SG_NP %d\n#define SG_NJPHI %d\n#define SG_NJG %d\n#define SG_NH %d\n" % (len(PNAMES), len(jphi_idx), len(jg_idx), len(h_idx)))
        for nm, idx in (("jphi", jphi_idx), ("jg", jg_idx), ("h", h_idx)):
            fh.write("const int sg_%s_r[] = {%s};\n" % (nm, ",".join(str(r) for r, _ in idx)))
            fh.write("const int sg_%s_c[] = {%s};\n" % (nm, ",".join(str(c) for _, c in idx)))
        fh.write("/* parameter order: %s */\n" % " ".join(PNAMES))
        fh.write("void sg_stage(const double* z, const double* p, const double* y, const double* lam, double* cost, double* grad,\n"
                 "              double* phi, double* jphi, double* g, double* jg, double* hess) {\n")
        for s_, e in repl:
            fh.write("  const double %s = %s;\n" % (s_, sp.ccode(e.xreplace(sub))))
        for nm, e in zip(names, red):
            fh.write("  %s = %s;\n" % (nm, sp.ccode(sp.sympify(e).xreplace(sub))))
        fh.write("}\n")
    print("wrote", OUT, "jphi", len(jphi_idx), "jg", len(jg_idx), "hess", len(h_idx), "cse temporaries", len(repl))


if __name__ == "__main__":
    main()
