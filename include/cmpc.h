/* cmpc.h -- C ABI of the B200 batched centroidal-MPC solver (libcmpc_b200.so).
 *
 * Drop-in boundary for the per-tick NLP solve of the reference
 * `code/centroidal_mpc_vertices.py` (and its `_payload` twin).  Each entry point replaces a group of
 * CasADi calls made by `centroidal_mpc.solve` (file:line relative to /root/reference/code):
 *
 *   cmpc_create            <- cs.Opti() / opt.solver('ipopt', ...) / the whole NLP build
 *                             (centroidal_mpc_vertices.py:126-353)
 *   cmpc_solve_device/host <- opt.set_value x (x0, gamma_l, gamma_r, com_ref, 4N foot refs)
 *                             (:511, :533-534, :590, :597-600), opt.solve() (:606),
 *                             sol.value(state[:,1]), sol.value(U[:,0]) (:614-616)
 *   cmpc_get_trajectory    <- sol.value(self.opti_state) / sol.value(self.U) (:617, :630-631)
 *   cmpc_set_warm          <- opt.set_initial(U, ...), opt.set_initial(state, ...) (:630-631)
 *   cmpc_reset_warm        <- a fresh Opti (first tick is solved from the solver's default guess)
 *   cmpc_assemble_device   <- the parameter assembly loops of `solve` (:482-600) + planner queries
 *                             (footstep_planner_vertices.py:82-103), batched on the device
 *
 * Batch-first: B independent instances per call, one CTA per instance, all FP64.  Instance-major
 * layouts (each instance contiguous).  No torch types; plain pointers and sizes.  Every function
 * returns 0 on success or a negative error code (never throws); cmpc_last_error() gives the text.
 * A handle is bound to one GPU and is not re-entrant.
 */
#ifndef CMPC_H
#define CMPC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMPC_NX 20   /* states  (MPC file :164-166) */
#define CMPC_NU 32   /* inputs  (MPC file :149-152) */

/* per-instance status codes */
enum {
  CMPC_CONVERGED = 0,
  CMPC_MAXITER = 1,
  CMPC_LINESEARCH = 2,
  CMPC_REGULARIZATION = 3,
  CMPC_INFEASIBLE_X0 = 4,   /* a row that depends on x0 only is violated (e.g. CoM z > 0.76, :230) */
  CMPC_NAN = 5,
  CMPC_STALL = 6            /* no progress within `stall_window` iterations at one barrier value (attempt abandoned) */
};

/* warm-start modes */
enum {
  CMPC_COLD = 0,          /* solver's own initial guess */
  CMPC_WARM_PRIMAL = 1,   /* states/inputs of the previous solve (what the reference does, :630-631) */
  CMPC_WARM_FULL = 2,     /* states, inputs, costates, slacks and multipliers of the previous solve */
  CMPC_WARM_SHIFTED = 3,  /* as FULL, moved one stage ahead first (consecutive ticks, mpc_rate * world_time_step = delta) */
  CMPC_WARM_AUTO = 4      /* per instance: SHIFTED while a landing (a foot's gamma going 0 -> 1) lies inside the horizon -- the
                             switch moves one stage per tick --, FULL otherwise (references move 1.5 mm per tick) */
};

typedef struct cmpc_config {
  int32_t N;              /* horizon, params['N'] (:10) */
  int32_t max_iter;       /* interior-point iteration cap */
  int32_t ls_max;         /* backtracking steps of the filter line search */
  int32_t threads;        /* threads per instance (CTA size): 128 */
  int32_t stall_window;   /* an attempt whose barrier-problem error has not halved in this many iterations is abandoned (0 = off) */
  int32_t stall_final;    /* the same at the final barrier value (a healthy end game takes 2-4 iterations) */
  int32_t jam_window;     /* a WARM attempt is abandoned (the instance restarts cold) after this many consecutive steps shorter than 0.1 (0 = off) */
  int32_t reserved0;
  double delta;           /* world_time_step * mpc_rate (:11) */
  double grav;            /* params['g'] (:18) */
  double mu_fric;         /* 0.5 (:41) */
  double foot_half_len;   /* 0.125 (:51) */
  double foot_half_wid;   /* 0.065 (:52) */
  double w_h, w_xy, w_zc, w_foot, w_sym, w_swing;   /* 1000, 1, 2000, 1000, 10, 10 (:301-335) */
  double w_rate;          /* 1, or 0 when mpc_rate == 10 (:339-341) */
  double eps_reg;         /* Tikhonov weight on the cost-free foot velocity / yaw-rate inputs */
  double pz_max;          /* 0.76 (:230) */
  double box[3];          /* 0.01, 0.005, 0.00005 (:259-271) */
  double relax;           /* inequality relaxation, IPOPT bound_relax_factor = 1e-8 */
  double mu_init, mu_final, mu_warm, tol;
  double kappa_eps, kappa_mu, theta_mu, tau_min, bound_push;
} cmpc_config;

typedef struct cmpc_handle cmpc_handle;

int cmpc_default_config(int32_t N, cmpc_config* cfg);
int cmpc_create(const cmpc_config* cfg, int32_t batch_capacity, int32_t device, cmpc_handle** out);
int cmpc_destroy(cmpc_handle* h);
const char* cmpc_last_error(void);
const char* cmpc_version(void);

/* Solve `batch` instances.  All pointers are DEVICE pointers valid on the handle's GPU.
 *   x0       [B][20]        opti_x0_param (:170)
 *   com_ref  [B][N][9]      opti_com_ref column i = (pos, vel, acc) (:172)
 *   foot_ref [B][N][8]      column i = (p_l(3), p_r(3), psi_l, psi_r) (:174-177)
 *   gamma    [B][N+1][2]    (gamma_l, gamma_r) (:179-181)
 *   mass,k1  [B]            params['mass'] (:15), k1 (:27; 7 in the payload variant)
 * Outputs (device): x1 [B][20] = state[:,1]; u0 [B][32] = U[:,0]; xN [B][20] = state[:,N];
 *   cost [B] (reference cost :311-351, without eps_reg term); viol [B] max unrelaxed violation of all
 *   rows and dynamics defects; status/iters [B].  Any output pointer may be NULL.
 * Asynchronous on `stream` (a cudaStream_t).  stream == NULL means the handle's own non-blocking stream, NOT the CUDA
 * default stream.  Every operation on a handle is ordered after the previous operation on that handle, whatever streams
 * the two were issued on (the handle records an event behind each operation and later operations wait for it).
 * The warm-start state lives in the handle.  An instance that does not converge is solved again inside the same launch
 * (it re-enters the work queue of the persistent CTAs): from the solver's cold start; then with a ten times larger initial
 * barrier value; then from another starting point, CoM states blended towards the reference along the horizon; `iters`
 * accumulates over the attempts. */
int cmpc_solve_device(cmpc_handle* h, int32_t batch, const double* x0, const double* com_ref,
                      const double* foot_ref, const double* gamma, const double* mass, const double* k1,
                      int32_t warm_mode, double* x1, double* u0, double* xN, double* cost, double* viol,
                      int32_t* status, int32_t* iters, void* stream);

/* Same call with HOST buffers: pinned staging, H2D, solve, D2H, synchronises before returning. */
int cmpc_solve_host(cmpc_handle* h, int32_t batch, const double* x0, const double* com_ref,
                    const double* foot_ref, const double* gamma, const double* mass, const double* k1,
                    int32_t warm_mode, double* x1, double* u0, double* xN, double* cost, double* viol,
                    int32_t* status, int32_t* iters);

/* As cmpc_solve_host, and the full primal trajectories of the first `traj_batch` instances (X [traj_batch][N+1][20],
 * U [traj_batch][N][32], host buffers) ride on the same device-to-host synchronisation: one call per control tick for the
 * drop-in class (x1, u0 and `x_collect`, :614-617). */
int cmpc_solve_host_traj(cmpc_handle* h, int32_t batch, const double* x0, const double* com_ref,
                         const double* foot_ref, const double* gamma, const double* mass, const double* k1,
                         int32_t warm_mode, double* x1, double* u0, double* xN, double* cost, double* viol,
                         int32_t* status, int32_t* iters, int32_t traj_batch, double* X, double* U);

/* Full primal trajectories of the last solve to HOST buffers: X [B][N+1][20], U [B][N][32]. */
int cmpc_get_trajectory(cmpc_handle* h, int32_t batch, double* X, double* U);
/* Primal warm start from HOST buffers (same layouts); used with CMPC_WARM_PRIMAL. */
int cmpc_set_warm(cmpc_handle* h, int32_t batch, const double* X, const double* U);
/* Snapshot / restore the warm-start state (states, inputs, costates, slacks, multipliers) of the first `batch`
 * instances inside the handle (device-to-device).  Lets a caller replay the same tick repeatedly (benchmarks,
 * what-if sweeps from one nominal solution).  Restore is asynchronous on `stream` (may be NULL). */
int cmpc_warm_save(cmpc_handle* h, int32_t batch);
int cmpc_warm_restore(cmpc_handle* h, int32_t batch, void* stream);
/* Forget the warm-start state: of all instances (mask == NULL) or of the instances b < n with mask[b] != 0 (host buffer).
 * A warm solve of an instance without warm-start state starts cold. */
int cmpc_reset_warm(cmpc_handle* h, const uint8_t* mask, int32_t n);
/* Sum over the last batch of (interior-point iterations, Riccati factorisations, regularisation retries),
 * device time of the last solve kernel in milliseconds (CUDA events) and kernels launched by it. */
int cmpc_last_stats(cmpc_handle* h, int64_t* iters, int64_t* nfact, int64_t* nreg, double* kernel_ms,
                    int32_t* launches);
/* Per-phase SM cycle counters of the last solve, summed over CTAs (zeros unless built with -DCMPC_PROFILE):
 * 11 values: eval, assemble, P[B A] products, factorisation, factor store, forward sweep, slack steps, trials, step,
 * whole solves, CTA lifetimes. */
int cmpc_phase_cycles(cmpc_handle* h, uint64_t* out11);
/* Device memory: bytes of iterate per instance (resident across ticks), bytes of scratch per resident CTA slot (Newton
 * step, derivative records, stage factors of the solve the slot is running), number of slots, shared memory per CTA. */
int cmpc_footprint(const cmpc_handle* h, size_t* iterate_bytes_per_instance, size_t* scratch_bytes_per_slot, int32_t* slots,
                   size_t* smem_bytes_per_cta);

/* ---- batched per-tick parameter assembly on the device ------------------------------------------
 * Tables of one walk, device pointers, built once by the caller from the planner / reference generator outputs:
 *   com_tab    [T_ref][9]   CoM_ref pos/vel/acc x,y,z per absolute tick (MPC file :64-74), cut to the shortest table
 *   foot_tab   [T_ref][8]   position_contacts_ref: p_l(3), p_r(3), yaw_l, yaw_r (:77-84)
 *   gamma_tab  [T_plan][2]  contact schedule per absolute tick from get_phase_at_time / plan[..]['foot_id'] (:515-534)
 *   step_index [T_plan]     get_step_index_at_time (footstep_planner_vertices.py:82-88) */
typedef struct cmpc_walk_tables {
  int32_t N, rate, first_swing_left, n_steps, T_ref, T_plan;
  const double* com_tab;
  const double* foot_tab;
  const double* gamma_tab;
  const int32_t* step_index;
} cmpc_walk_tables;

/* x0 / com_ref / foot_ref / gamma of `batch` robots (layouts of cmpc_solve_device), robot b at its own tick[b]: replaces
 * the O(N * steps) Python loops and 4N set_value calls of `solve` (:482-600).  Per robot inputs (device): measured CoM
 * position / velocity / angular momentum [B][3], theta_hat of the previous solution [B][3] (:485), measured foot yaws
 * [B][2], the robot's (step-adjusted) plan positions [B][n_steps][3].  err[b] (may be NULL): 0 ok, 1 the horizon runs
 * past the reference tables (the reference's IndexError, :567), 2 past the footstep plan; such robots are not written.
 * One gather kernel on `stream` (NULL = default stream) of GPU `device`. */
int cmpc_assemble_device(const cmpc_walk_tables* tb, int32_t device, int32_t batch, const int32_t* tick, const double* com_pos,
                         const double* com_vel, const double* hw, const double* theta, const double* yaw, const double* plan,
                         double* x0, double* com_ref, double* foot_ref, double* gamma, int32_t* err, void* stream);

/* ---- batched dense QP of the whole-body inverse-dynamics step -----------------------------------
 * Replaces `QPSolver.set_values / solve` (code/utils.py:40-92: CasADi Opti('conic') + OSQP) as used by
 * `InverseDynamics.get_joint_torques` (code/inverse_dynamics.py:30-135), for `batch` independent QPs of one size:
 *     min 1/2 x'H x + F'x   s.t.  A_eq x = b_eq,  A_in x <= b_in
 *   H [B][n][n] (symmetric), F [B][n], A_eq [B][m_eq][n], b_eq [B][m_eq], A_in [B][m_in][n], b_in [B][m_in], row-major.
 * n <= 96, m_eq <= 48, m_in <= 32 (the reference: n = 2 dofs + 12, m_eq = dofs, m_in = 16).  One CTA per QP, Mehrotra
 * predictor-corrector interior point on the quasi-definite KKT matrix in shared memory, FP64; H + 1e-9 I (the
 * reference's H is singular in the torque block, OSQP regularises likewise).  tol <= 0 -> 1e-9, max_iter <= 0 -> 60.
 * Outputs: x [B][n], status [B] (0 converged, 1 max_iter, 3 bad pivot, 5 nan), iters [B] (either may be NULL).
 * The reference returns zeros when OSQP fails (utils.py:85-92); here the caller decides from `status`. */
int cmpc_qp_solve_device(int32_t device, int32_t batch, int32_t n, int32_t m_eq, int32_t m_in, const double* H, const double* F,
                         const double* A_eq, const double* b_eq, const double* A_in, const double* b_in, double tol, int32_t max_iter,
                         double* x, int32_t* status, int32_t* iters, void* stream);
/* Same with HOST buffers (allocates device staging per call: a convenience path, not a hot path). */
int cmpc_qp_solve_host(int32_t device, int32_t batch, int32_t n, int32_t m_eq, int32_t m_in, const double* H, const double* F,
                       const double* A_eq, const double* b_eq, const double* A_in, const double* b_in, double tol, int32_t max_iter,
                       double* x, int32_t* status, int32_t* iters);

/* Peak-FP64 probe: runs a dependent-free DFMA loop on every SM and returns the measured TFLOP/s. */
int cmpc_measure_fp64_peak(int32_t device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* CMPC_H */
