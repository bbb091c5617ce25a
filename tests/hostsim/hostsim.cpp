// tests/hostsim/hostsim.cpp -- TEST AID ONLY (never loaded by the product).
// Compiles the product's solver core (csrc/cmpc_solver.h) with g++ and a serial execution policy so the
// algorithm can be debugged in a container without a GPU.  The product path is the CUDA build in
// csrc/cmpc_kernels.cu and fails loudly when that extension is missing; nothing here is a fallback.
#define CMPC_TRACE 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "cmpc_solver.h"

using namespace cmpc;

struct ParSerial {
  void* smem_ptr = nullptr;
  template <class T> T& smem() const { return *static_cast<T*>(smem_ptr); }
  void bind(void* p) { smem_ptr = p; }
  int tid() const { return 0; }
  int nt() const { return 1; }
  void sync() const {}
  int lane() const { return 0; }
  int warp() const { return 0; }
  int nwarps() const { return 1; }
  int lanes() const { return 1; }
  void sync_warp() const {}
  bool sync_and(bool p) const { return p; }
  bool any(bool p) const { return p; }
  double wmax(double v) const { return v; }
  double wmin(double v) const { return v; }
  double wsum(double v) const { return v; }
  void copy_async(double* dst, const double* src, int n) const { for (int t = 0; t < n; ++t) dst[t] = src[t]; }
  void commit_async() const {}
  void wait_async() const {}
  double shfl4(double v, int) const { return v; }
  double shfl4x(double v, int) const { return v; }
  void prefetch_l2(const double*, int) const {}
  static constexpr int TPT = 120, CPT = 16;
  static constexpr bool GAINS4 = false;
};

extern "C" {

void hostsim_trace(int on) { cmpc::cmpc_trace_on = on; }

int hostsim_work_doubles(int N) { return (int)work_doubles(N); }

// cfg_over: {eps_reg, relax, mu_init, mu_final, tol, max_iter, ls_max, w_rate, mu_warm, kappa_eps, kappa_mu, theta_mu, tau_min, warm_push, warm_comp, (8 unused), stall_window, stall_final} (NaN = keep default)
// Debug: stage-i Lagrangian gradient (60) and assembled stage block M (60x60, lower) at the iterate stored in
// `work` (X, U, Y, S, LAM as laid out by carve_work).  Used by tests to check the analytic Hessian by finite
// differences of the analytic gradient.
int hostsim_stage_debug(int N, const double* x0, const double* com_ref, const double* foot_ref, const double* gamma,
                        double mass, double k1, double* work, int i, double mu, double* grad_out, double* M_out) {
  Config c = default_config(N);
  Instance in{x0, com_ref, foot_ref, gamma, mass, k1};
  Work w = carve_work(work, N);
  Smem* sm = new Smem();
  ParSerial par;
  Solver<ParSerial> sol(c, in, w, *sm, par);
  double pv; build_masks(c, in, sm->mask, &pv);
  sol.setup();
  sol.mu = mu;
  double acc[8];
  stage_derivs(c, in, w, i, sm->mask[i], mu, acc);
  for (int j = 0; j < 60; ++j) grad_out[j] = cmpc_dbg_rd[j];
  for (int t = 0; t < NX * NX; ++t) sm->P[t] = 0.0;
  for (int t = 0; t < NX; ++t) sm->pv[t] = 0.0;
  for (int t = 0; t < Q_RG; ++t) sm->recb[0][t] = w.REC[(size_t)i * RECSZ + t];
  sol.assemble_stage(i, 0.0, sm->recb[0]);
  for (int r = 0; r < 60; ++r) for (int cc = 0; cc < 60; ++cc) {
    const int a = r < NU ? r : r + NW, b = cc < NU ? cc : cc + NW;
    M_out[r * 60 + cc] = sm->M[mi(a >= b ? a : b, a >= b ? b : a)];
  }
  delete sm;
  return 0;
}

// Debug: the CTA-wide eval pass against the thread-per-stage reference evaluation (stage_derivs) at the iterate stored in
// `work`.  out[0] = max relative difference over all record entries, out[1] = same over the per-stage statistics.
int hostsim_eval_compare(int N, const double* x0, const double* com_ref, const double* foot_ref, const double* gamma,
                         double mass, double k1, double* work, double* out) {
  Config c = default_config(N);
  Instance in{x0, com_ref, foot_ref, gamma, mass, k1};
  Work w = carve_work(work, N);
  Smem* sm = new Smem();
  ParSerial par;
  Solver<ParSerial> sol(c, in, w, *sm, par);
  double pv; build_masks(c, in, sm->mask, &pv);
  sol.setup();
  double ev[8];
  sol.eval(ev);
  std::vector<double> rec_new(w.REC, w.REC + (size_t)(N + 1) * RECSZ);
  std::vector<double> acc_new((N + 1) * 8);
  for (int i = 0; i <= N; ++i) for (int q = 0; q < 5; ++q) acc_new[i * 8 + q] = sm->acc[i][q];
  for (size_t t = 0; t < rec_new.size(); ++t) w.REC[t] = 0.0;
  double e_rec = 0.0, e_acc = 0.0;
  for (int i = 0; i <= N; ++i) {
    double acc[8];
    stage_derivs(c, in, w, i, sm->mask[i], 0.0, acc);
    for (int t = 0; t < RECSZ; ++t) {
      if (t >= 597 && t < 600) continue;                         // unused slots
      if (i == N && !((t >= Q_GC + 32 && t < Q_GC + 60) || (t >= Q_M1 + 32 && t < Q_M1 + 60) || (t >= Q_M2 + 32 && t < Q_M2 + 60) || (t >= Q_DIAG + 32 && t < Q_DIAG + 60))) continue;
      if ((t == Q_HRG || (t >= Q_HP && t < Q_HP + 3)) && !(sm->mask[i] & (1ull << R_HW))) continue;   // read at stage 0 only
      const double a = w.REC[(size_t)i * RECSZ + t], b = rec_new[(size_t)i * RECSZ + t];
      const double e = fabs(a - b) / (1.0 + fabs(a));
      if (e > e_rec) { e_rec = e; out[2] = i; out[3] = t; out[4] = a; out[5] = b; }
    }
    for (int q = 0; q < 5; ++q) {
      const double a = acc[q], b = acc_new[i * 8 + q];
      const double e = fabs(a - b) / (1.0 + fabs(a));
      if (e > e_acc) { e_acc = e; out[6] = i; out[7] = q; }
    }
  }
  out[0] = e_rec; out[1] = e_acc;
  // trial pass at a step of 0.37 along a pseudo-random direction against the per-stage evaluation
  {
    unsigned long long seed = 88172645463325252ull;
    auto rnd = [&]() { seed ^= seed << 13; seed ^= seed >> 7; seed ^= seed << 17; return (double)(seed % 2000001) / 1e6 - 1.0; };
    for (int t = 0; t < (N + 1) * NX; ++t) w.DX[t] = 1e-3 * rnd();
    for (int t = 0; t < N * NU; ++t) w.DU[t] = 1e-1 * rnd();
    for (int t = 0; t < (N + 1) * NR; ++t) w.DS[t] = 0.3 * w.S[t] * rnd();
    double tr[5], ref[5] = {0, 0, 0, 0, 0};
    sol.trial(0.37, tr);
    for (int i = 0; i <= N; ++i) {
      double a[8];
      stage_trial(c, in, w, i, sm->mask[i], 0.37, a);
      ref[0] += a[0]; ref[1] += a[1]; ref[2] += a[2]; ref[3] = a[3] > ref[3] ? a[3] : ref[3]; ref[4] += a[4];
    }
    double e = 0.0;
    for (int q = 0; q < 5; ++q) { const double v = fabs(tr[q] - ref[q]) / (1.0 + fabs(ref[q])); e = v > e ? v : e; }
    out[8] = e;
  }
  delete sm;
  return 0;
}

int hostsim_solve(int N, const double* x0, const double* com_ref, const double* foot_ref, const double* gamma,
                  double mass, double k1, const double* cfg_over, int warm, double* work, double* stats_out) {
  Config c = default_config(N);
  if (cfg_over) {
    if (cfg_over[0] == cfg_over[0]) c.eps_reg = cfg_over[0];
    if (cfg_over[1] == cfg_over[1]) c.relax = cfg_over[1];
    if (cfg_over[2] == cfg_over[2]) c.mu_init = cfg_over[2];
    if (cfg_over[3] == cfg_over[3]) c.mu_final = cfg_over[3];
    if (cfg_over[4] == cfg_over[4]) c.tol = cfg_over[4];
    if (cfg_over[5] == cfg_over[5]) c.max_iter = (int)cfg_over[5];
    if (cfg_over[6] == cfg_over[6]) c.ls_max = (int)cfg_over[6];
    if (cfg_over[7] == cfg_over[7]) c.w_rate = cfg_over[7];
    if (cfg_over[8] == cfg_over[8]) c.mu_warm = cfg_over[8];
    if (cfg_over[9] == cfg_over[9]) c.kappa_eps = cfg_over[9];
    if (cfg_over[10] == cfg_over[10]) c.kappa_mu = cfg_over[10];
    if (cfg_over[11] == cfg_over[11]) c.theta_mu = cfg_over[11];
    if (cfg_over[12] == cfg_over[12]) c.tau_min = cfg_over[12];
    if (cfg_over[13] == cfg_over[13]) c.warm_push = cfg_over[13];
    if (cfg_over[14] == cfg_over[14]) c.warm_comp = cfg_over[14];
    if (cfg_over[23] == cfg_over[23]) c.stall_window = (int)cfg_over[23];
    if (cfg_over[24] == cfg_over[24]) c.stall_final = (int)cfg_over[24];
    if (cfg_over[25] == cfg_over[25]) c.jam_window = (int)cfg_over[25];
    if (cfg_over[26] == cfg_over[26]) c.crawl_window = (int)cfg_over[26];
    if (cfg_over[27] == cfg_over[27]) c.crawl_alpha = cfg_over[27];
    if (cfg_over[28] == cfg_over[28]) c.warm_stall_window = (int)cfg_over[28];
  }
  Instance in{x0, com_ref, foot_ref, gamma, mass, k1};
  Work w = carve_work(work, N);
  Smem* sm = new Smem();
  ParSerial par;
  Solver<ParSerial> sol(c, in, w, *sm, par);
  Stats st;
  sol.run(warm, &st);
  stats_out[0] = st.cost; stats_out[1] = st.viol; stats_out[2] = st.kkt; stats_out[3] = st.mu;
  stats_out[4] = st.iters; stats_out[5] = st.status; stats_out[6] = st.nfact; stats_out[7] = st.nreg;
  delete sm;
  return st.status;
}
}
