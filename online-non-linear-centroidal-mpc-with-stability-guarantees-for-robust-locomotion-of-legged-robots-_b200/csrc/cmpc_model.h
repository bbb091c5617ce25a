// cmpc_model.h -- the centroidal NLP of the reference, stage by stage, with hand-derived analytic
// derivatives (no autodiff tape).  Host/device: compiled by nvcc for sm_100a (product) and by g++
// for the host-side simulator under tests/ (debugging aid only, never shipped).
//
// What is restated (reference = code/centroidal_mpc_vertices.py):
//   dynamics           :371-461, Euler step :187-190
//   Lyapunov rows      :193-220 (k2 cancels algebraically)
//   angular-mom. row   :223-224
//   CoM height         :229-230
//   friction/unilateral:44-47, :235-254
//   foot-placement box :258-271
//   cost               :275-353
//
// Formulation used by the solver (same optimum as the reference's literal NLP):
//   * multiple shooting with stage variables x_i (28) and u_i (32);
//   * state augmented with q = previous vertex f_z (8) so the force-rate term (:343-351) is a stage
//     cost:  x = [p v h theta psi_l p_l psi_r p_r | q(8)],  q_{i+1} = f_z(u_i);
//   * rows that the reference writes on x_{i+1} (Lyapunov, angular momentum) are written on
//     (x_i, u_i) through the *linear* part of the dynamics (p+ = p + d v, v+ = v + d (g + F/m));
//   * tracking cost / box rows on x_{i+1} with reference column i are carried by stage i+1.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CMPC_HD __host__ __device__ inline
#else
#define CMPC_HD inline
#endif

namespace cmpc {

constexpr int NXP = 20;          // physical states (reference layout, MPC file :164-166)
constexpr int NQ = 8;            // augmented previous f_z
constexpr int NX = NXP + NQ;     // 28
constexpr int NU = 32;           // MPC file :149-152
constexpr int NZ = NU + NX;      // 60, stage ordering [u ; x]
constexpr int NR = 56;           // inequality row slots per stage
constexpr int NMAX = 64;         // max horizon supported by the shared-memory tables

// row slots
constexpr int R_LYAP = 0, R_HW = 1, R_PZ = 2, R_FRIC = 3, R_UNI = 35, R_BOX = 43;

// x offsets
constexpr int IP = 0, IV = 3, IH = 6, ITH = 9, IPSL = 12, IPL = 13, IPSR = 16, IPR = 17, IQ = 20;

struct Config {
  int N;
  double delta, grav, mu_fric;      // :11, :18, :41
  double hl, hw;                    // foot half length / width (:51-60)
  double w_h, w_xy, w_zc, w_foot, w_sym, w_swing, w_rate, eps_reg;   // :301-351
  double pz_max, box[3];            // :230, :259-271
  double relax;                     // IPOPT bound_relax_factor (1e-8)
  // interior-point options
  double mu_init, mu_final, tol, kappa_eps, kappa_mu, theta_mu, tau_min, bound_push;
  double mu_warm;                   // initial barrier for full warm starts
  double warm_push, warm_comp;      // full warm start: slack floor, and cap on s*lam/mu_warm (0 = none)
  int max_iter, ls_max;
  int stall_window;                 // iterations without halving the barrier-problem error before an attempt is abandoned (0 = off)
  int stall_final;                  // the same at the final barrier value
  int jam_window;                   // consecutive steps shorter than 0.1 before a WARM attempt is abandoned (0 = off)
  int warm_stall_window;            // stall window of a WARM attempt above the final barrier value (0 = stall_window)
  int crawl_window; double crawl_alpha;   // consecutive steps shorter than crawl_alpha, whatever the residual does (0 = off)
};

CMPC_HD Config default_config(int N) {
  Config c;
  c.N = N; c.delta = 0.01; c.grav = 9.81; c.mu_fric = 0.5;
  c.hl = 0.125; c.hw = 0.065;
  c.w_h = 1000.0; c.w_xy = 1.0; c.w_zc = 2000.0; c.w_foot = 1000.0; c.w_sym = 10.0; c.w_swing = 10.0;
  c.w_rate = 1.0; c.eps_reg = 1e-5;
  c.pz_max = 0.76; c.box[0] = 0.01; c.box[1] = 0.005; c.box[2] = 0.00005;
  c.relax = 1e-8;
  c.mu_init = 0.1; c.mu_final = 1e-9; c.tol = 1e-8; c.kappa_eps = 10.0; c.kappa_mu = 0.2;
  c.theta_mu = 1.5; c.tau_min = 0.99; c.bound_push = 1e-2; c.mu_warm = 1e-3; c.warm_push = 3e-5; c.warm_comp = 0.0;
  // (long horizons: off.  Measured at N = 60, 16 384 warm recorded-walk instances on one B200, converged fraction / solves
  // per second: no rule, cap 5 N iterations 99.66 % / 2954;  end-game rule with window N, cap 5 N: 99.22 % / 3513;  same, cap
  // 10 N / 3: 98.98 % / 4171;  both rules as at N = 20: 98.27 % / 3983 -- the rules abort attempts that would still converge)
  c.stall_window = N > 20 ? 0 : 60; c.stall_final = N > 20 ? 0 : 20;
  c.jam_window = 6;
  // (short horizons: a warm attempt that has not halved its error in 25 iterations at one barrier value restarts cold.  All 1926
  // recorded N = 10 ticks, warm, CPU build: longest solve 90 -> 55 iterations -- the warm attempts of ticks 473 / 772 / 1474 / 1574 sit at
  // step lengths of 0.7 for the 60 iterations of the general window --, mean 10.69 -> 10.52, same KKT points.  At N = 20 windows of
  // 25 / 30 lengthen the longest solve (54 -> 58 / 62) and 40 changes nothing: off)
  c.warm_stall_window = N <= 12 ? 25 : 0;
  c.crawl_window = N > 20 ? 0 : 12; c.crawl_alpha = 0.05;    // (long horizons: off -- at N = 60 warm attempts crawl and still beat a cold start: 24.8 -> 28.1 iterations with the rule)
  // iteration cap PER ATTEMPT.  N <= 20: 70 -- an instance that fails every attempt is a serial chain of 3-4 caps on one CTA and sets the
  // duration of a cold batch (4096 perturbed states, CPU build: same 4090 converged with 100 and 70, longest chain 300 -> 210 iterations,
  // mean unchanged; 50 loses three instances); long horizons (several contact switches inside) need more than 100 from cold
  c.max_iter = N > 20 ? 5 * N : 70; c.ls_max = 3;
  return c;
}

// One NLP instance: what the reference feeds through opt.set_value (:511-600).
struct Instance {
  const double* x0;        // [20]
  const double* com_ref;   // [N][9]   pos, vel, acc of column i
  const double* foot_ref;  // [N][8]   p_l(3) p_r(3) psi_l psi_r of column i
  const double* gamma;     // [N+1][2] gamma_l, gamma_r
  double mass, k1;
};

CMPC_HD double wz_of(const Config& c, int j) { return (c.w_zc - 0.5 * c.w_zc) * exp(-(double)j) + 0.5 * c.w_zc; }  // :301-305

// foot polygon corner k in the foot frame (:55-60)
CMPC_HD void corner(const Config& c, int k, double& cx, double& cy) {
  cx = (k < 2) ? c.hl : -c.hl;
  cy = (k == 0 || k == 3) ? c.hw : -c.hw;
}

// lever arms r_ek = R(psi_e) c_k + p_e - p  and their yaw derivative R'(psi_e) c_k
struct Arms { double r[8][3]; double dr[8][2]; double rc[8][2]; };

CMPC_HD void lever_arms(const Config& c, const double* x, Arms& a) {
  for (int e = 0; e < 2; ++e) {
    const double psi = x[e ? IPSR : IPSL];
    const double* pe = x + (e ? IPR : IPL);
    const double cs = cos(psi), sn = sin(psi);
    for (int k = 0; k < 4; ++k) {
      double cx, cy; corner(c, k, cx, cy);
      const double rx = cs * cx - sn * cy, ry = sn * cx + cs * cy;
      const int v = 4 * e + k;
      a.rc[v][0] = rx; a.rc[v][1] = ry;
      a.r[v][0] = rx + pe[0] - x[0]; a.r[v][1] = ry + pe[1] - x[1]; a.r[v][2] = pe[2] - x[2];
      a.dr[v][0] = -sn * cx - cs * cy; a.dr[v][1] = cs * cx - sn * cy;
    }
  }
}

// phi_i(x,u): Euler step of the centroidal dynamics (:187-190, :371-461) plus q+ = f_z.
CMPC_HD void dyn_step(const Config& c, const Instance& in, int i, const double* x, const double* u, double* xn) {
  const double gl = in.gamma[2 * i], gr = in.gamma[2 * i + 1];
  const double d = c.delta, m = in.mass;
  const double* ref = in.com_ref + 9 * i;
  Arms a; lever_arms(c, x, a);
  double F[3] = {0, 0, 0}, T[3] = {0, 0, 0};
  for (int v = 0; v < 8; ++v) {
    const double ge = (v < 4) ? gl : gr;
    const double* f = u + 3 * v;
    F[0] += ge * f[0]; F[1] += ge * f[1]; F[2] += ge * f[2];
    T[0] += ge * (a.r[v][1] * f[2] - a.r[v][2] * f[1]);
    T[1] += ge * (a.r[v][2] * f[0] - a.r[v][0] * f[2]);
    T[2] += ge * (a.r[v][0] * f[1] - a.r[v][1] * f[0]);
  }
  for (int j = 0; j < 3; ++j) {
    xn[IP + j] = x[IP + j] + d * x[IV + j];
    xn[IV + j] = x[IV + j] + d * ((j == 2 ? -c.grav : 0.0) + F[j] / m);
    xn[IH + j] = x[IH + j] + d * T[j];
    xn[ITH + j] = x[ITH + j] + (d / m) * (in.k1 * (x[IP + j] - ref[j]) + x[IV + j] - ref[3 + j]);
    xn[IPL + j] = x[IPL + j] + d * (1.0 - gl) * u[24 + j];
    xn[IPR + j] = x[IPR + j] + d * (1.0 - gr) * u[27 + j];
  }
  xn[IPSL] = x[IPSL] + d * (1.0 - gl) * u[30];
  xn[IPSR] = x[IPSR] + d * (1.0 - gr) * u[31];
  for (int v = 0; v < 8; ++v) xn[IQ + v] = u[3 * v + 2];
}

// tracking part of the cost carried by stage i >= 1 (reference column i-1, gamma[i]) (:313-319)
CMPC_HD double track_cost(const Config& c, const Instance& in, int i, const double* x) {
  const double* ref = in.com_ref + 9 * (i - 1);
  const double* fr = in.foot_ref + 8 * (i - 1);
  const double gl = in.gamma[2 * i], gr = in.gamma[2 * i + 1];
  double J = c.w_xy * ((x[0] - ref[0]) * (x[0] - ref[0]) + (x[1] - ref[1]) * (x[1] - ref[1]))
           + wz_of(c, i - 1) * (x[2] - ref[2]) * (x[2] - ref[2]);
  for (int j = 0; j < 3; ++j) {
    J += c.w_foot * gl * (x[IPL + j] - fr[j]) * (x[IPL + j] - fr[j]);
    J += c.w_foot * gr * (x[IPR + j] - fr[3 + j]) * (x[IPR + j] - fr[3 + j]);
  }
  J += c.w_foot * gl * (x[IPSL] - fr[6]) * (x[IPSL] - fr[6]);
  J += c.w_foot * gr * (x[IPSR] - fr[7]) * (x[IPSR] - fr[7]);
  return J;
}

// stage cost l_i(x_i, u_i): i < N input terms + h cost (+ rate term with q), i >= 1 tracking.
// `with_reg` = include the eps_reg Tikhonov term (not part of the reference cost).
CMPC_HD double stage_cost(const Config& c, const Instance& in, int i, const double* x, const double* u, bool with_reg) {
  double J = 0.0;
  if (i >= 1) J += track_cost(c, in, i, x);
  if (i >= c.N) return J;
  const double gl = in.gamma[2 * i], gr = in.gamma[2 * i + 1];
  J += c.w_h * (x[IH] * x[IH] + x[IH + 1] * x[IH + 1] + x[IH + 2] * x[IH + 2]);       // :312
  for (int e = 0; e < 2; ++e) {
    const double ge = e ? gr : gl;
    const double* f = u + 12 * e;
    double s2 = 0.0, mean[3] = {0, 0, 0};
    for (int k = 0; k < 4; ++k)
      for (int j = 0; j < 3; ++j) { s2 += f[3 * k + j] * f[3 * k + j]; mean[j] += 0.25 * f[3 * k + j]; }
    const double dev = s2 - 4.0 * (mean[0] * mean[0] + mean[1] * mean[1] + mean[2] * mean[2]);
    J += ge * c.w_sym * dev + (1.0 - ge) * c.w_swing * s2;                              // :320-335
    if (i >= 1) {                                                                       // :343-351, term i-1
      const double gp = in.gamma[2 * (i - 1) + e];
      for (int k = 0; k < 4; ++k) {
        const double dz = f[3 * k + 2] - x[IQ + 4 * e + k];
        J += c.w_rate * gp * dz * dz;
      }
    }
  }
  if (with_reg)
    for (int j = 24; j < 32; ++j) J += c.eps_reg * u[j] * u[j];
  return J;
}

// Lyapunov row written on (x_i, u_i) (:202-220 with p+, v+ substituted).  Returns the row value and,
// if G != nullptr, the gradient wrt xi = (p, v, theta, F) (12) and the 4x4 coefficient matrix C
// (Hessian = C (x) I_3 in xi-space).
CMPC_HD double lyapunov(const Config& c, const Instance& in, int i, const double* x, const double* F,
                        double* G, double* C) {
  const double d = c.delta, m = in.mass, k1 = in.k1;
  const double* ref = in.com_ref + 9 * i;
  double z1[3], z2[3], ae[3];
  for (int j = 0; j < 3; ++j) {
    const double grav = (j == 2) ? -c.grav : 0.0;
    const double vp = x[IV + j] + d * (grav + F[j] / m);
    z1[j] = x[IP + j] + d * x[IV + j] - ref[j];
    z2[j] = k1 * z1[j] + vp - ref[3 + j];
    ae[j] = F[j] / m + grav - ref[6 + j] + x[ITH + j] / m;
  }
  double q = 0.0;
  for (int j = 0; j < 3; ++j)
    q += -k1 * z1[j] * z1[j] + k1 * z2[j] * z2[j] + (1.0 - k1 * k1) * z1[j] * z2[j] + z2[j] * ae[j];
  if (G) {
    // T = d(z1, z2, ae)/d(p, v, theta, F)  (scalar blocks)
    const double T[3][4] = {{1.0, d, 0.0, 0.0}, {k1, k1 * d + 1.0, 0.0, d / m}, {0.0, 0.0, 1.0 / m, 1.0 / m}};
    const double Hz[3][3] = {{-2.0 * k1, 1.0 - k1 * k1, 0.0}, {1.0 - k1 * k1, 2.0 * k1, 1.0}, {0.0, 1.0, 0.0}};
    for (int j = 0; j < 3; ++j) {
      const double g1 = -2.0 * k1 * z1[j] + (1.0 - k1 * k1) * z2[j];
      const double g2 = 2.0 * k1 * z2[j] + (1.0 - k1 * k1) * z1[j] + ae[j];
      const double g3 = z2[j];
      for (int a = 0; a < 4; ++a) G[3 * a + j] = T[0][a] * g1 + T[1][a] * g2 + T[2][a] * g3;
    }
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) {
        double s = 0.0;
        for (int p = 0; p < 3; ++p)
          for (int r = 0; r < 3; ++r) s += T[p][a] * Hz[p][r] * T[r][b];
        C[4 * a + b] = s;
      }
  }
  return q;
}

// Curvature of the Lyapunov row in xi-space, C (x) I_3, C = T' Hz T: depends on (k1, delta, m) only -> once per instance.
CMPC_HD void lyapunov_consts(const Config& c, const Instance& in, double* C) {
  const double d = c.delta, m = in.mass, k1 = in.k1;
  const double T[3][4] = {{1.0, d, 0.0, 0.0}, {k1, k1 * d + 1.0, 0.0, d / m}, {0.0, 0.0, 1.0 / m, 1.0 / m}};
  const double Hz[3][3] = {{-2.0 * k1, 1.0 - k1 * k1, 0.0}, {1.0 - k1 * k1, 2.0 * k1, 1.0}, {0.0, 1.0, 0.0}};
  for (int a = 0; a < 4; ++a)
    for (int b = 0; b < 4; ++b) {
      double s = 0.0;
      for (int p = 0; p < 3; ++p)
        for (int r = 0; r < 3; ++r) s += T[p][a] * Hz[p][r] * T[r][b];
      C[4 * a + b] = s;
    }
}

// Lyapunov row value and gradient wrt xi = (p, v, theta, F) (12 values), no local arrays.
CMPC_HD double lyapunov_grad(const Config& c, const Instance& in, int i, const double* x, const double* F, double* G) {
  const double d = c.delta, m = in.mass, k1 = in.k1;
  const double* ref = in.com_ref + 9 * i;
  const double T[3][4] = {{1.0, d, 0.0, 0.0}, {k1, k1 * d + 1.0, 0.0, d / m}, {0.0, 0.0, 1.0 / m, 1.0 / m}};
  double q = 0.0;
  for (int j = 0; j < 3; ++j) {
    const double grav = (j == 2) ? -c.grav : 0.0;
    const double vp = x[IV + j] + d * (grav + F[j] / m);
    const double z1 = x[IP + j] + d * x[IV + j] - ref[j];
    const double z2 = k1 * z1 + vp - ref[3 + j];
    const double ae = F[j] / m + grav - ref[6 + j] + x[ITH + j] / m;
    q += -k1 * z1 * z1 + k1 * z2 * z2 + (1.0 - k1 * k1) * z1 * z2 + z2 * ae;
    const double g1 = -2.0 * k1 * z1 + (1.0 - k1 * k1) * z2;
    const double g2 = 2.0 * k1 * z2 + (1.0 - k1 * k1) * z1 + ae;
    const double g3 = z2;
    for (int a = 0; a < 4; ++a) G[3 * a + j] = T[0][a] * g1 + T[1][a] * g2 + T[2][a] * g3;
  }
  return q;
}

// All inequality rows of stage i as g <= 0 (unrelaxed).  xpred = phi_i(x,u) (needed for the hw row).
// Rows outside `mask` are left untouched.
CMPC_HD void stage_ineq(const Config& c, const Instance& in, int i, uint64_t mask, const double* x,
                        const double* u, const double* xpred, double* g) {
  if (i < c.N) {
    const double gl = in.gamma[2 * i], gr = in.gamma[2 * i + 1];
    if (mask & (1ull << R_LYAP)) {
      double F[3] = {0, 0, 0};
      for (int v = 0; v < 8; ++v) {
        const double ge = (v < 4) ? gl : gr;
        F[0] += ge * u[3 * v]; F[1] += ge * u[3 * v + 1]; F[2] += ge * u[3 * v + 2];
      }
      g[R_LYAP] = lyapunov(c, in, i, x, F, nullptr, nullptr);
    }
    if (mask & (1ull << R_HW))
      g[R_HW] = xpred[IH] * xpred[IH] + xpred[IH + 1] * xpred[IH + 1] + xpred[IH + 2] * xpred[IH + 2]
              - (x[IH] * x[IH] + x[IH + 1] * x[IH + 1] + x[IH + 2] * x[IH + 2]);
    for (int v = 0; v < 8; ++v) {
      if (!(mask & (1ull << (R_UNI + v)))) continue;
      const double* f = u + 3 * v;
      g[R_FRIC + 4 * v + 0] = f[0] - c.mu_fric * f[2];
      g[R_FRIC + 4 * v + 1] = -f[0] - c.mu_fric * f[2];
      g[R_FRIC + 4 * v + 2] = f[1] - c.mu_fric * f[2];
      g[R_FRIC + 4 * v + 3] = -f[1] - c.mu_fric * f[2];
      g[R_UNI + v] = -f[2];
    }
  }
  if (mask & (1ull << R_PZ)) g[R_PZ] = x[IP + 2] - c.pz_max;
  if (i >= 1) {
    const double* fr = in.foot_ref + 8 * (i - 1);
    for (int e = 0; e < 2; ++e) {
      if (!(mask & (1ull << (R_BOX + 6 * e)))) continue;
      for (int j = 0; j < 3; ++j) {
        const double err = x[(e ? IPR : IPL) + j] - fr[3 * e + j];
        g[R_BOX + 6 * e + 2 * j] = err - c.box[j];
        g[R_BOX + 6 * e + 2 * j + 1] = -err - c.box[j];
      }
    }
  }
}

// Active-row masks of all stages; rows whose value cannot depend on any decision variable
// (p_z of the fixed x_0; box rows of a foot that has been in stance since stage 0) are parametric:
// they are left out of the solve and only reported through `param_viol`.
CMPC_HD void build_masks(const Config& c, const Instance& in, uint64_t* mask, double* param_viol) {
  bool moved[2] = {false, false};
  double pv = in.x0[2] - c.pz_max;
  for (int i = 0; i <= c.N; ++i) {
    uint64_t mk = 0;
    const double gl = in.gamma[2 * i], gr = in.gamma[2 * i + 1];
    if (i < c.N) {
      mk |= 1ull << R_LYAP;
      if (i == 0) mk |= 1ull << R_HW;
      for (int v = 0; v < 8; ++v)
        if (((v < 4) ? gl : gr) > 0.5) mk |= (0xFull << (R_FRIC + 4 * v)) | (1ull << (R_UNI + v));
    }
    if (i >= 1 && i < c.N) mk |= 1ull << R_PZ;
    if (i >= 1) {
      for (int e = 0; e < 2; ++e) {
        if ((e ? gr : gl) < 0.5) continue;
        if (moved[e]) mk |= 0x3Full << (R_BOX + 6 * e);
        else {
          const double* fr = in.foot_ref + 8 * (i - 1);
          for (int j = 0; j < 3; ++j) {
            const double err = fabs(in.x0[(e ? IPR : IPL) + j] - fr[3 * e + j]) - c.box[j];
            pv = err > pv ? err : pv;
          }
        }
      }
    }
    mask[i] = mk;
    if (i < c.N) { if (gl < 0.5) moved[0] = true; if (gr < 0.5) moved[1] = true; }
  }
  *param_viol = pv;
}

// ---------------------------------------------------------------------------------------------
// Sparse [B A] of stage i: column j of the 28 x 60 matrix [d phi/du, d phi/dx] has at most 4
// non-zeros; the row pattern is structural.
CMPC_HD int ba_row(int j, int slot) {
  if (j < 24) {
    const int a = j % 3;
    if (slot == 0) return IV + a;
    if (slot == 1) return IH + (a + 1) % 3;
    if (slot == 2) return IH + (a + 2) % 3;
    return (a == 2) ? IQ + j / 3 : -1;
  }
  if (j < 32) {
    if (slot) return -1;
    if (j < 27) return IPL + (j - 24);
    if (j < 30) return IPR + (j - 27);
    return j == 30 ? IPSL : IPSR;
  }
  const int cidx = j - 32;
  if (cidx < 3) {
    if (slot == 0) return IP + cidx;
    if (slot == 1) return ITH + cidx;
    if (slot == 2) return IH + (cidx + 1) % 3;
    return IH + (cidx + 2) % 3;
  }
  if (cidx < 6) {
    const int a = cidx - 3;
    if (slot == 0) return IP + a;
    if (slot == 1) return IV + a;
    if (slot == 2) return ITH + a;
    return -1;
  }
  if (cidx < 12) return slot == 0 ? cidx : -1;
  if (cidx == IPSL || cidx == IPSR) return slot == 0 ? cidx : IH + slot - 1;
  if (cidx < 20) {
    const int a = (cidx < IPSR) ? cidx - IPL : cidx - IPR;
    if (slot == 0) return cidx;
    if (slot == 1) return IH + (a + 1) % 3;
    if (slot == 2) return IH + (a + 2) % 3;
    return -1;
  }
  return -1;
}

// values of the sparse [B A] (60 x 4), given arms and forces
CMPC_HD void ba_values(const Config& c, const Instance& in, int i, const double* u, const Arms& a, double* bav) {
  const double gl = in.gamma[2 * i], gr = in.gamma[2 * i + 1];
  const double d = c.delta, m = in.mass;
  double Fe[2][3] = {{0, 0, 0}, {0, 0, 0}}, dT[2][3] = {{0, 0, 0}, {0, 0, 0}};
  for (int v = 0; v < 8; ++v) {
    const int e = v / 4;
    const double* f = u + 3 * v;
    for (int j = 0; j < 3; ++j) Fe[e][j] += f[j];
    // (R' c_k) x f
    dT[e][0] += a.dr[v][1] * f[2];
    dT[e][1] += -a.dr[v][0] * f[2];
    dT[e][2] += a.dr[v][0] * f[1] - a.dr[v][1] * f[0];
  }
  for (int j = 0; j < NZ * 4; ++j) bav[j] = 0.0;
  for (int j = 0; j < 24; ++j) {
    const int v = j / 3, ax = j % 3;
    const double ge = (v < 4) ? gl : gr;
    bav[4 * j + 0] = d * ge / m;
    bav[4 * j + 1] = d * ge * a.r[v][(ax + 2) % 3];
    bav[4 * j + 2] = -d * ge * a.r[v][(ax + 1) % 3];
    bav[4 * j + 3] = (ax == 2) ? 1.0 : 0.0;
  }
  for (int j = 24; j < 27; ++j) bav[4 * j] = d * (1.0 - gl);
  for (int j = 27; j < 30; ++j) bav[4 * j] = d * (1.0 - gr);
  bav[4 * 30] = d * (1.0 - gl);
  bav[4 * 31] = d * (1.0 - gr);
  const double Ft[3] = {gl * Fe[0][0] + gr * Fe[1][0], gl * Fe[0][1] + gr * Fe[1][1], gl * Fe[0][2] + gr * Fe[1][2]};
  for (int ax = 0; ax < 3; ++ax) {
    double* col = bav + 4 * (32 + IP + ax);
    col[0] = 1.0; col[1] = d * in.k1 / m;
    col[2] = d * Ft[(ax + 2) % 3]; col[3] = -d * Ft[(ax + 1) % 3];
    col = bav + 4 * (32 + IV + ax);
    col[0] = d; col[1] = 1.0; col[2] = d / m;
    bav[4 * (32 + IH + ax)] = 1.0;
    bav[4 * (32 + ITH + ax)] = 1.0;
    for (int e = 0; e < 2; ++e) {
      const double ge = e ? gr : gl;
      col = bav + 4 * (32 + (e ? IPR : IPL) + ax);
      col[0] = 1.0;
      col[1] = -d * ge * Fe[e][(ax + 2) % 3];
      col[2] = d * ge * Fe[e][(ax + 1) % 3];
    }
  }
  for (int e = 0; e < 2; ++e) {
    const double ge = e ? gr : gl;
    double* col = bav + 4 * (32 + (e ? IPSR : IPSL));
    col[0] = 1.0;
    col[1] = d * ge * dT[e][0]; col[2] = d * ge * dT[e][1]; col[3] = d * ge * dT[e][2];
  }
}

}  // namespace cmpc
