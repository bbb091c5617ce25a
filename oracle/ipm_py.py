"""ORACLE (test infrastructure) -- pure numpy/scipy interior-point solve of the reference NLP.

Restates what `self.opt.solve()` does in the reference (`code/centroidal_mpc_vertices.py:606`):
CasADi hands the NLP of :185-353 to IPOPT (options at :127-128).  IPOPT's source is not in the
reference tree and not installable offline (pip line `README.md:20`, version unpinned), so its
*published* algorithm is restated: Waechter & Biegler, "On the implementation of an
interior-point filter line-search algorithm for large-scale nonlinear programming", Math. Prog.
106 (2006) -- slack reformulation g(w)+s=0, primal-dual Newton steps on the barrier KKT system
(eq. 11/13), fraction-to-boundary rule (15), filter line search (18-20), monotone barrier
update (7), Hessian regularisation when the step is not a descent direction (sec. 3.1; the
inertia test is replaced by the curvature test of the inertia-free variant because scipy's
sparse LU reports no inertia), bound relaxation `bound_relax_factor` = 1e-8 (sec. 3.5).

Two settings:  oracle-T (tol 1e-10: the parity target) and oracle-R (tol 1e-3, IPOPT defaults
otherwise: what the reference actually runs).  Linear algebra here is a *generic sparse LU* of
the whole KKT matrix, deliberately unlike the stage-wise Riccati recursion of the CUDA path.

Parity status: UNPINNED (no CasADi/IPOPT here, no golden vectors in the reference).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sps
import scipy.linalg as sla
import scipy.sparse.linalg as spla

from .spec import NG, NP, NU, NV, NX, PARAM_NAMES, block_functions

NS = NX + NU  # stride of one stage in w
_PI = {n: j for j, n in enumerate(PARAM_NAMES)}


@dataclass
class Problem:
    """Parameters of one NLP instance (what the reference feeds with opt.set_value, :511-600)."""
    N: int
    x0: np.ndarray            # (20,)
    com_ref: np.ndarray       # (9, N)
    pl_ref: np.ndarray        # (3, N)
    pr_ref: np.ndarray        # (3, N)
    al_ref: np.ndarray        # (N,)
    ar_ref: np.ndarray        # (N,)
    gl: np.ndarray            # (N+1,)
    gr: np.ndarray            # (N+1,)
    mass: float = 40.05487735
    k1: float = 4.0
    k2: float = 0.1
    delta: float = 0.01
    grav: float = 9.81
    w_rate: float = 1.0       # weight_f_rate (:339-341)
    eps_reg: float = 1e-9

    def block_params(self, i: int) -> np.ndarray:
        P = np.zeros(NP)
        P[0:9] = self.com_ref[:, i]
        P[9:12] = self.pl_ref[:, i]
        P[12:15] = self.pr_ref[:, i]
        P[_PI["alr"]] = self.al_ref[i]
        P[_PI["arr"]] = self.ar_ref[i]
        P[_PI["gl"]], P[_PI["gr"]] = self.gl[i], self.gr[i]
        P[_PI["gln"]], P[_PI["grn"]] = self.gl[i + 1], self.gr[i + 1]
        P[_PI["wz"]] = 1000.0 * math.exp(-i) + 1000.0      # :301-305
        P[_PI["wrate"]] = self.w_rate if i < self.N - 1 else 0.0
        P[_PI["hw_on"]] = 1.0 if i == 0 else 0.0
        P[_PI["mass"]], P[_PI["k1"]], P[_PI["k2"]] = self.mass, self.k1, self.k2
        P[_PI["delta"]], P[_PI["grav"]], P[_PI["eps_reg"]] = self.delta, self.grav, self.eps_reg
        return P


@dataclass
class Options:
    tol: float = 1e-10
    max_iter: int = 1000
    mu_init: float = 0.1
    mu_min: float = None
    bound_relax: float = 1e-8
    constr_viol_tol: float = 1e-4
    compl_inf_tol: float = 1e-4
    dual_inf_tol: float = 1.0
    kappa_eps: float = 10.0
    kappa_mu: float = 0.2
    theta_mu: float = 1.5
    tau_min: float = 0.99
    bound_push: float = 1e-2
    verbose: bool = False
    linesearch: str = "relaxed"  # "filter" | "relaxed" | "none"
    inertia: bool = True          # exact inertia from a dense LDL' (slow, robust) vs curvature test on a sparse LU

    @staticmethod
    def oracle_T():
        return Options(tol=1e-8, mu_min=1e-9, constr_viol_tol=1e-8, compl_inf_tol=1e-8, dual_inf_tol=1e-6)

    @staticmethod
    def oracle_R():
        """IPOPT settings of the reference: tol=1e-3 (:128), everything else default."""
        return Options(tol=1e-3)


@dataclass
class Result:
    w: np.ndarray
    cost: float
    viol: float
    iters: int
    status: int               # 0 converged, 1 max_iter, 2 line-search failure
    y: np.ndarray = None
    lam: np.ndarray = None
    s: np.ndarray = None
    kkt: float = 0.0
    nfact: int = 0
    log: list = field(default_factory=list)

    def X(self, N):
        return np.stack([self.w[NS * i:NS * i + NX] for i in range(N + 1)], axis=1)

    def U(self, N):
        return np.stack([self.w[NS * i + NX:NS * i + NS] for i in range(N)], axis=1)


class NLP:
    """Numeric assembly of f, c, g and their derivatives from the symbolic stage block.

    All N stage blocks are evaluated in one vectorised call of each lambdified function
    (every block symbol is fed an array of length N)."""

    def __init__(self, prob: Problem):
        self.p = prob
        self.bf = block_functions()
        N = prob.N
        self.N = N
        self.n = NS * N + NX
        self.m_eq = NX * (N + 1)
        self.Pm = np.stack([prob.block_params(i) for i in range(N)], axis=1)     # (NP, N)
        self.Pl = list(self.Pm)
        keep = np.ones((N, NG), bool)
        for i in range(N):
            k = keep[i]
            if i > 0:
                k[1] = False
            gl, gr, gln, grn = prob.gl[i], prob.gr[i], prob.gl[i + 1], prob.gr[i + 1]
            if gl == 0:
                k[3:19] = False
                k[35:39] = False
            if gr == 0:
                k[19:35] = False
                k[39:43] = False
            if gln == 0:
                k[43:49] = False
            if grn == 0:
                k[49:55] = False
        self.keep = keep
        self.m_in = int(keep.sum())
        self.rowmap = -np.ones((N, NG), int)
        self.rowmap[keep] = np.arange(self.m_in)
        # column map of block i -> global index (u_N does not exist -> -1)
        cm = np.arange(NV)[None, :] + NS * np.arange(N)[:, None]
        cm[cm >= self.n] = -1
        self.colmap = cm                                                    # (N, NV)
        self.cvalid = cm >= 0
        bf = self.bf
        jd_r = np.array([rc[0] for rc in bf.jd_idx]); jd_c = np.array([rc[1] for rc in bf.jd_idx])
        jg_r = np.array([rc[0] for rc in bf.jg_idx]); jg_c = np.array([rc[1] for rc in bf.jg_idx])
        h_r = np.array([rc[0] for rc in bf.h_idx]); h_c = np.array([rc[1] for rc in bf.h_idx])
        self._jd = (NX * (np.arange(N)[:, None] + 1) + jd_r[None, :], cm[:, jd_c])
        self._jg = (self.rowmap[:, jg_r], cm[:, jg_c])
        self._h = (cm[:, h_r], cm[:, h_c])

    def _vl(self, w):
        v = np.zeros((self.N, NV))
        v[self.cvalid] = w[self.colmap[self.cvalid]]
        return list(v.T)

    def _arr(self, out):
        return np.stack([np.broadcast_to(np.asarray(e, float), (self.N,)) for e in out], axis=1)   # (N, k)

    def cost(self, w):
        return float(np.sum(self.bf.f_cost(self._vl(w), self.Pl)))

    def cost_reference(self, w):
        """Cost exactly as the reference defines it (no eps_reg term)."""
        c = self.cost(w)
        for i in range(self.N):
            c -= self.p.eps_reg * float(np.sum(w[NS * i + NX + 24:NS * i + NS] ** 2))
        return c

    def grad(self, w):
        gi = self._arr(self.bf.f_grad(self._vl(w), self.Pl))
        gvec = np.zeros(self.n)
        np.add.at(gvec, self.colmap[self.cvalid], gi[self.cvalid])
        return gvec

    def eq(self, w):
        c = np.zeros(self.m_eq)
        c[0:NX] = w[0:NX] - self.p.x0
        c[NX:] = self._arr(self.bf.f_d(self._vl(w), self.Pl)).ravel()
        return c

    def ineq(self, w):
        gi = self._arr(self.bf.f_g(self._vl(w), self.Pl))
        return gi[self.keep]

    def _sparse(self, vals, rc, shape, extra=None):
        r, c = rc
        ok = (r >= 0) & (c >= 0)
        rr, cc, vv = r[ok], c[ok], vals[ok]
        if extra is not None:
            rr = np.concatenate([extra[0], rr]); cc = np.concatenate([extra[1], cc]); vv = np.concatenate([extra[2], vv])
        return sps.csr_matrix((vv, (rr, cc)), shape=shape)

    def jac_eq(self, w):
        jv = self._arr(self.bf.f_jd(self._vl(w), self.Pl))
        return self._sparse(jv, self._jd, (self.m_eq, self.n), (np.arange(NX), np.arange(NX), np.ones(NX)))

    def jac_in(self, w):
        jv = self._arr(self.bf.f_jg(self._vl(w), self.Pl))
        return self._sparse(jv, self._jg, (self.m_in, self.n))

    def hess(self, w, y, lam):
        li = np.zeros((self.N, NG))
        li[self.keep] = lam
        yi = y[NX:].reshape(self.N, NX)
        hv = self._arr(self.bf.f_h(self._vl(w), self.Pl, list(yi.T), list(li.T)))
        return self._sparse(hv, self._h, (self.n, self.n))

    def violation(self, w):
        """max unscaled violation of the reference's constraints (no relaxation)."""
        return float(max(np.max(np.abs(self.eq(w))), np.max(np.maximum(self.ineq(w), 0.0), initial=0.0)))


def solve(prob: Problem, w0: np.ndarray = None, opts: Options = None) -> Result:
    opts = opts or Options.oracle_T()
    nlp = NLP(prob)
    n, me, mi = nlp.n, nlp.m_eq, nlp.m_in
    w = np.zeros(n) if w0 is None else np.array(w0, float).copy()
    relax = opts.bound_relax
    mu = opts.mu_init
    mu_floor = opts.tol / 10.0 if opts.mu_min is None else opts.mu_min
    g = nlp.ineq(w) - relax
    s = np.maximum(-g, opts.bound_push)           # slack push (IPOPT sec. 3.6)
    lam = np.ones(mi)                              # bound_mult_init_val = 1
    y = np.zeros(me)
    filt = []
    log = []
    nfact = 0
    dw_last = 0.0
    status = 1
    it = 0
    theta0 = None

    def err(mu_t, gradL, c, r_g, s, lam, y):
        s_max = 100.0
        sd = max(s_max, (np.abs(y).sum() + np.abs(lam).sum()) / (me + mi)) / s_max
        sc = max(s_max, np.abs(lam).sum() / max(mi, 1)) / s_max
        e_d = np.max(np.abs(gradL)) / sd
        e_p = max(np.max(np.abs(c)), np.max(np.abs(r_g)))
        e_c = np.max(np.abs(s * lam - mu_t)) / sc
        return max(e_d, e_p, e_c), (e_d, e_p, e_c)

    for it in range(opts.max_iter + 1):
        f = nlp.cost(w)
        gf = nlp.grad(w)
        c = nlp.eq(w)
        g = nlp.ineq(w) - relax
        Jc = nlp.jac_eq(w)
        Jg = nlp.jac_in(w)
        r_g = g + s
        gradL = gf + Jc.T @ y + Jg.T @ lam
        E0, parts0 = err(0.0, gradL, c, r_g, s, lam, y)
        viol_u = max(np.max(np.abs(c)), np.max(np.maximum(g + relax, 0)))
        log.append((it, f, parts0[0], parts0[1], parts0[2], mu))
        if opts.verbose:
            print("it %3d f %.10e dual %.2e prim %.2e compl %.2e mu %.1e dw %.1e al %.2e" % (it, f, *parts0, mu, dw_last, getattr(solve, "_al", 0)))
        Ef, partsf = err(mu_floor, gradL, c, r_g, s, lam, y)
        if mu <= mu_floor and Ef <= opts.tol and partsf[1] <= opts.constr_viol_tol and partsf[0] <= opts.dual_inf_tol \
                and partsf[2] <= opts.compl_inf_tol:
            status = 0
            break
        if it == opts.max_iter:
            break
        # barrier update (eq. 7)
        while True:
            Emu, _ = err(mu, gradL, c, r_g, s, lam, y)
            if Emu <= opts.kappa_eps * mu and mu > mu_floor:
                mu = max(mu_floor, min(opts.kappa_mu * mu, mu ** opts.theta_mu))
                filt = []
            else:
                break
        tau = max(opts.tau_min, 1.0 - mu)
        # reduced KKT:  (W + Jg' Sigma Jg + dw I) dw + Jc' dy = -(gf + Jc' y + Jg' (lam + Sigma r_g - mu/s + lam... ))
        Sig = lam / s
        W = nlp.hess(w, y, lam)
        # elimination:  ds = -r_g - Jg dw ;  dlam = -lam + mu/s - Sig*ds
        rhs_w = -(gf + Jc.T @ y + Jg.T @ (mu / s + Sig * r_g))
        Hbar = (W + Jg.T @ sps.diags(Sig) @ Jg).tocsc()
        dw_reg = 0.0
        attempt = 0
        while True:
            K = sps.bmat([[Hbar + dw_reg * sps.identity(n), Jc.T], [Jc, None]], format="csc")
            ok = False
            if opts.inertia:
                # inertia of the KKT matrix from a dense Bunch-Kaufman LDL' (IPOPT sec. 3.1 asks for
                # exactly n positive and m_eq negative eigenvalues)
                Kd = K.toarray()
                lu_, d_, perm_ = sla.ldl(Kd, lower=True)
                nneg = 0
                k_ = 0
                nd = d_.shape[0]
                while k_ < nd:
                    if k_ + 1 < nd and d_[k_ + 1, k_] != 0.0:
                        ev = np.linalg.eigvalsh(d_[k_:k_ + 2, k_:k_ + 2]); nneg += int((ev < 0).sum()); k_ += 2
                    else:
                        nneg += int(d_[k_, k_] < 0); k_ += 1
                nfact += 1
                if nneg == me:
                    sol = np.linalg.solve(Kd, np.concatenate([rhs_w, -c]))
                    # one step of iterative refinement
                    res_ = np.concatenate([rhs_w, -c]) - Kd @ sol
                    sol = sol + np.linalg.solve(Kd, res_)
                    ok = bool(np.all(np.isfinite(sol)))
                    if ok:
                        dw = sol[:n]; dy = sol[n:]
                        break
            else:
                try:
                    lu = spla.splu(K)
                    nfact += 1
                    sol = lu.solve(np.concatenate([rhs_w, -c]))
                    ok = bool(np.all(np.isfinite(sol)))
                except RuntimeError:
                    ok = False
                if ok:
                    dw = sol[:n]
                    dy = sol[n:]
                    curv = dw @ (Hbar @ dw) + dw_reg * (dw @ dw)
                    if curv >= 1e-12 * (dw @ dw) or (dw @ dw) == 0.0:
                        break
            # regularise (sec. 3.1): first 1e-4 (or last/3), then x100 / x8
            if dw_reg == 0.0:
                dw_reg = 1e-4 if dw_last == 0.0 else max(1e-20, dw_last / 3.0)
            else:
                dw_reg *= 100.0 if dw_last == 0.0 else 8.0
            attempt += 1
            if dw_reg > 1e40:
                return Result(w, f, viol_u, it, 3, y, lam, s, E0, nfact, log)
        if dw_reg > 0:
            dw_last = dw_reg
        ds = -r_g - Jg @ dw
        dlam = -lam + mu / s - Sig * ds
        # fraction to the boundary (15)
        def amax(v, dv):
            neg = dv < 0
            return min(1.0, float(np.min(-tau * v[neg] / dv[neg]))) if neg.any() else 1.0
        a_p, a_d = amax(s, ds), amax(lam, dlam)
        # filter line search on (theta, phi)
        theta = np.abs(c).sum() + np.abs(r_g).sum()
        if theta0 is None:
            theta0 = theta
            theta_max = 1e4 * max(1.0, theta0)
            theta_min = 1e-4 * max(1.0, theta0)
        phi = f - mu * np.log(s).sum()
        dphi = gf @ dw - mu * (ds / s).sum()
        alpha = a_p
        accepted = False
        g_th, g_ph, eta = 1e-5, 1e-5, 1e-4
        for ls in range({"filter": 40, "relaxed": 3}.get(opts.linesearch, 0)):
            wt = w + alpha * dw
            st = s + alpha * ds
            ct = nlp.eq(wt)
            gt = nlp.ineq(wt) - relax
            th_t = np.abs(ct).sum() + np.abs(gt + st).sum()
            ph_t = nlp.cost(wt) - mu * np.log(st).sum()
            okf = th_t <= theta_max and all((th_t < (1 - g_th) * tf) or (ph_t < pf - g_ph * tf) for tf, pf in filt)
            if okf:
                sw = theta <= theta_min and dphi < 0 and alpha * (-dphi) ** 2.3 > theta ** 1.1
                if sw:
                    if ph_t <= phi + eta * alpha * dphi + 10 * np.finfo(float).eps * abs(phi):
                        accepted = True
                        break
                else:
                    if th_t <= (1 - g_th) * theta or ph_t <= phi - g_ph * theta:
                        accepted = True
                        filt.append(((1 - g_th) * theta, phi - g_ph * theta))
                        break
            alpha *= 0.5
        if opts.linesearch != "filter" and not accepted:
            alpha = a_p
            wt = w + alpha * dw
            st = s + alpha * ds
            solve._fail = 0
        elif not accepted:
            # no restoration phase here: fall back to the full fraction-to-boundary step once the
            # filter is reset; give up if that happens twice in a row
            if getattr(solve, "_fail", 0) >= 3:
                status = 2
                solve._fail = 0
                break
            solve._fail = getattr(solve, "_fail", 0) + 1
            filt = []
            alpha = a_p * 1e-2
            wt = w + alpha * dw
            st = s + alpha * ds
        else:
            solve._fail = 0
        w, s = wt, st
        solve._al = alpha
        y = y + alpha * dy
        lam = lam + a_d * dlam
        # keep Sigma within [mu/(k s), k mu / s]  (eq. 16), k = 1e10
        lam = np.clip(lam, mu / (1e10 * s), 1e10 * mu / s)
    fref = nlp.cost_reference(w)
    return Result(w, fref, nlp.violation(w), it, status, y, lam, s, E0, nfact, log)


def unpack_problem(N, x0, com_ref, foot_ref, gamma, mass, k1, eps_reg=1e-9, w_rate=1.0) -> Problem:
    """instance-major arrays of the C ABI (include/cmpc.h) -> Problem"""
    com_ref = np.asarray(com_ref, float).reshape(N, 9)
    foot_ref = np.asarray(foot_ref, float).reshape(N, 8)
    gamma = np.asarray(gamma, float).reshape(N + 1, 2)
    return Problem(N=N, x0=np.asarray(x0, float), com_ref=com_ref.T.copy(), pl_ref=foot_ref[:, 0:3].T.copy(),
                   pr_ref=foot_ref[:, 3:6].T.copy(), al_ref=foot_ref[:, 6].copy(), ar_ref=foot_ref[:, 7].copy(),
                   gl=gamma[:, 0].copy(), gr=gamma[:, 1].copy(), mass=float(mass), k1=float(k1), eps_reg=eps_reg, w_rate=w_rate)


def neutral_start(prob: Problem) -> np.ndarray:
    """x_i = x0, vertex f_z = m g / #contact vertices: the cold start used for all oracle-T golden vectors."""
    N = prob.N
    w = np.zeros(NS * N + NX)
    for i in range(N + 1):
        w[NS * i:NS * i + NX] = prob.x0
    for i in range(N):
        n = prob.gl[i] + prob.gr[i]
        for v in range(8):
            ge = prob.gl[i] if v < 4 else prob.gr[i]
            w[NS * i + NX + 3 * v + 2] = ge * prob.mass * prob.grav / (4 * max(n, 1))
    return w


def solve_packed(N, x0, com_ref, foot_ref, gamma, mass, k1, opts: Options = None):
    prob = unpack_problem(N, x0, com_ref, foot_ref, gamma, mass, k1)
    r = solve(prob, neutral_start(prob), opts or Options.oracle_T())
    return {"status": r.status, "iters": r.iters, "cost": r.cost, "viol": r.viol, "x1": r.X(N)[:, 1], "u0": r.U(N)[:, 0]}
