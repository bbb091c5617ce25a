// tests/hostsim/hostsim.cpp -- TEST AID ONLY (never loaded by the product).
// Compiles the product's solver core (csrc/cmpc_solver.h) with g++ and a serial execution policy so the
// algorithm can be debugged in a container without a GPU.  The product path is the CUDA build in
// csrc/cmpc_kernels.cu and fails loudly when that extension is missing; nothing here is a fallback.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "cmpc_solver.h"

using namespace cmpc;

struct ParSerial {
  int tid() const { return 0; }
  int nt() const { return 1; }
  void sync() const {}
};

extern "C" {

int hostsim_work_doubles(int N) { return (int)work_doubles(N); }

// cfg_over: {eps_reg, relax, mu_init, mu_final, tol, max_iter, ls_max, w_rate, mu_warm} (NaN = keep default)
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
