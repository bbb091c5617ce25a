// cmpc_solver.h -- one NLP instance solved by one cooperative thread group (a CTA on the GPU).
//
// Replaces `self.opt.solve()` (code/centroidal_mpc_vertices.py:606: CasADi Opti -> IPOPT -> MUMPS) by a
// primal-dual interior-point Newton method whose linear algebra is a stage-wise Riccati recursion:
//
//   eval pass      thread-per-stage: residuals, analytic Jacobians, barrier terms -> per-stage record
//   backward pass  stage by stage, all threads: assemble the 60x60 stage KKT block in shared memory
//                  (cost + barrier + Lagrangian Hessian + [B A]' P [B A]), partial Cholesky of the
//                  32x32 input block, Schur complement = cost-to-go P_i; factors streamed to global
//   forward pass   du_i = -L^-T (l_m + L_S' dx_i), dx_{i+1} = A dx_i + B du_i + d_i, costates
//   line search    fraction-to-boundary + a short filter backtracking, evaluated thread-per-stage
//   convergence    per-stage partial results reduced by warp 0 with shuffles (__shfl_xor_sync), NaN detection by a warp
//                  vote, pivot tests by a barrier vote (__syncthreads_and)
//
// The code is written against an execution policy `Par` (tid, nt, sync, reductions) so the very same
// source runs as one CTA per instance on sm_100a and as a single serial "thread" under g++ in
// tests/hostsim (debug aid).  All arithmetic is FP64.
#pragma once
#include "cmpc_model.h"

#if defined(__CUDACC__)
#define CMPC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CMPC_HD_NOINLINE
#endif

#ifndef CMPC_PIPE_W
#define CMPC_PIPE_W 1      // rows of W = P [B A] whose loads are in flight together (per lane; measured: 1 is not slower than 4)
#endif
#ifndef CMPC_PIPE_M
#define CMPC_PIPE_M 4      // rows of M += [B A]' W whose loads are in flight together (per warp)
#endif
#ifndef CMPC_RSQ64H
#define CMPC_RSQ64H 1
#endif
#ifndef CMPC_SKIP_BLOCKS
#define CMPC_SKIP_BLOCKS 1
#endif

namespace cmpc {

constexpr int NW = 2;                       // multipliers kept as explicit unknowns per stage (Lyapunov row, angular-momentum row)
constexpr int NA = NU + NW;                 // 34: augmented input block [u ; w]
constexpr int XO = NA;                      // offset of x in the stage block [u ; w ; x]
constexpr int NZA = NA + NX;                // 62
constexpr int GR = NZA;                     // index of the gradient row
constexpr int MSZ = (NZA + 1) * (NZA + 2) / 2;   // stage block stored as a packed lower triangle (63 rows incl. the gradient row): 2016 doubles
CMPC_HD constexpr int mi(int r, int c) { return r * (r + 1) / 2 + c; }   // index of M(r, c), r >= c
constexpr int MROWS = NZA + 1;              // 62 variable rows + the gradient row
constexpr int TRI_A = NA * (NA + 1) / 2;    // 595
constexpr int TRI_X = NX * (NX + 1) / 2;    // 406
// per-stage factor record streamed to global memory
constexpr int KSZ = (NX + 1) * NA;          // gains K (28 x 34) + k (34)
constexpr int F_K = 0, F_P = F_K + KSZ, F_PV = F_P + NX * NX;   // K | k, cost-to-go P (full, symmetric: coalesced reads), p
constexpr int FACSZ = F_PV + NX;            // 1798 doubles
constexpr int FWDBUF = KSZ + NX;            // forward sweep staging of one stage: K | k | d
constexpr int NTILE = 15 * 16 / 2;          // 4 x 4 tiles of the padded 64 x 64 lower triangle right of tile column 0 (120)
constexpr int NLY = 33 + 3 * 66;               // entries of the stage block the Lyapunov row touches: 33 gradient + 198 same-axis pairs
constexpr int PSTR = 18;                    // doubles between tile rows of the pivot panel: 16-byte aligned (128-bit accesses), 8 distinct bank groups
struct alignas(16) Pair { double x, y; };   // two panel entries moved by one 128-bit shared-memory access
// per-stage derivative record written by the eval pass
constexpr int Q_GC = 0, Q_M1 = 60, Q_M2 = 120, Q_BA = 180, Q_D = 420, Q_DIAG = 448, Q_FRIC = 508,
              Q_LG = 556, Q_LC = 568, Q_LSIG = 584, Q_HP = 585, Q_HLAM = 588, Q_HSIG = 589, Q_YH = 590,
              Q_GAM = 593, Q_GAMP = 595, Q_DR = 600, Q_LRG = 616, Q_LLAM = 617, Q_HRG = 618,
              Q_RG = 620;                     // row residuals g - relax + s of the condensed rows (56 slots, by row)
constexpr int RECSZ = Q_RG + NR;            // 676
static_assert(Q_RG % 2 == 0 && Q_BA % 2 == 0 && Q_D % 2 == 0 && RECSZ % 2 == 0 && KSZ % 2 == 0 && FACSZ % 2 == 0 && NX % 2 == 0,
              "asynchronous copies move 16-byte granules: even offsets and lengths");

// optional phase timers (cycles, thread 0 of each CTA): -DCMPC_PROFILE
#if defined(CMPC_PROFILE) && defined(__CUDA_ARCH__)
#define CMPC_TIC(sm) long long tic_ = clock64()
#define CMPC_TOC(sm, k) do { if (threadIdx.x == 0) { long long now_ = clock64(); (sm).prof[k] += now_ - tic_; tic_ = now_; } } while (0)
#else
#define CMPC_TIC(sm) do {} while (0)
#define CMPC_TOC(sm, k) do {} while (0)
#endif
enum { PF_EVAL = 0, PF_ASM, PF_PBA, PF_CHOL, PF_STORE, PF_FWD, PF_SLACK, PF_TRIAL, PF_APPLY, PF_SOLVE, PF_CTA, PF_COUNT };   // PF_SOLVE: whole solves, PF_CTA: CTA lifetime

#ifdef CMPC_TRACE
static int cmpc_trace_on = 0;
static double cmpc_dbg_rd[64];
#endif

#if defined(__CUDA_ARCH__)
#define CMPC_SCHED_FENCE() asm volatile("" ::: "memory")
// the workspace and the instance data are global memory (lets the compiler emit LDG / STG instead of generic accesses)
#define CMPC_ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))
#else
#define CMPC_SCHED_FENCE() do {} while (0)
#define CMPC_ASSUME_GLOBAL(p) do {} while (0)
#endif


CMPC_HD int cmpc_popcount(uint64_t m) {
#if defined(__CUDA_ARCH__)
  return __popcll(m);
#else
  int n = 0;
  while (m) { n += (int)(m & 1ull); m >>= 1; }
  return n;
#endif
}

CMPC_HD double cmpc_rcp(double x) {
#if defined(__CUDA_ARCH__)
  return __drcp_rn(x);
#else
  return 1.0 / x;
#endif
}

// Branch-free reciprocal square root for pivots (1e-14 < x < 1e30): hardware seed (MUFU.RSQ64H, one instruction on the high
// word, relative error about 2^-22; the single-precision route costs two conversions, a range fix-up and the MUFU in sequence:
// six dependent instructions on the diagonal-tile chain everybody waits for), two Newton steps in double.
// Unlike rsqrt(double) it has no range check / slow-path branch, so two of them interleave in one instruction stream.
CMPC_HD double cmpc_rsqrt_nb(double x) {
#if defined(__CUDA_ARCH__)
#if CMPC_RSQ64H
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  double y = (double)rsqrtf((float)x);
#endif
  const double h = 0.5 * x;
  y = y * fma(-h, y * y, 1.5);
  y = y * fma(-h, y * y, 1.5);
  return y;
#else
  return 1.0 / sqrt(x);
#endif
}

CMPC_HD double cmpc_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}

enum Status { ST_CONVERGED = 0, ST_MAXITER = 1, ST_LINESEARCH = 2, ST_REGULARIZATION = 3, ST_INFEASIBLE_X0 = 4, ST_NAN = 5, ST_STALL = 6 };

// Global memory of one solve.  The ITERATE (states, inputs, costates, slacks, multipliers) belongs to the instance and
// stays device resident across ticks (warm starts); the Newton step, the per-stage derivative records and the stage
// factors are SCRATCH of the resident CTA slot that happens to run the solve (0.4 MB at N = 20: with one scratch block per
// slot instead of one per instance the working set of a launch is slots x 0.4 MB and mostly stays in L2).
struct Work {
  double *X, *U, *Y, *S, *LAM;        // iterate: (N+1)*28, N*32, (N+1)*28, (N+1)*56, (N+1)*56
  double *DX, *DU, *DS, *YN, *DW;     // Newton step (YN = full-step costates, DW = new multipliers of explicit rows)
  double *REC, *FAC;                  // (N+1)*RECSZ, N*FACSZ
};

CMPC_HD size_t iter_doubles(int N) { return (size_t)(N + 1) * NX * 2 + (size_t)N * NU + (size_t)(N + 1) * NR * 2; }
CMPC_HD size_t scratch_doubles(int N) {
  return (size_t)(N + 1) * NX * 2 + (size_t)N * NU + (size_t)N * NW + (size_t)(N + 1) * NR + (size_t)(N + 1) * RECSZ + (size_t)N * FACSZ;
}
CMPC_HD size_t work_doubles(int N) { return iter_doubles(N) + scratch_doubles(N); }

CMPC_HD Work carve_work(double* iter, double* scratch, int N) {
  Work w; double* p = iter;
  w.X = p; p += (N + 1) * NX;  w.U = p; p += N * NU;  w.Y = p; p += (N + 1) * NX;
  w.S = p; p += (N + 1) * NR;  w.LAM = p;
  p = scratch;
  w.DX = p; p += (N + 1) * NX; w.DU = p; p += N * NU; w.DS = p; p += (N + 1) * NR; w.YN = p; p += (N + 1) * NX; w.DW = p; p += N * NW;
  w.REC = p; p += (size_t)(N + 1) * RECSZ; w.FAC = p;
  return w;
}
CMPC_HD Work carve_work(double* base, int N) { return carve_work(base, base + iter_doubles(N), N); }   // one contiguous block (tests/hostsim)

// Shared-memory block of one instance.
struct alignas(16) Smem {
  double M[MSZ];             // stage KKT block [u ; w ; x] (+ gradient row 62), packed lower triangle
  double W[NX * NZ];         // P * [B A]
  double P[NX * NX];         // cost-to-go Hessian of stage i+1 (full symmetric)
  double pv[NX];             // cost-to-go gradient
  double tv[NX];             // p + P d
  double recb[2][Q_RG];      // derivative record of the current stage and, arriving asynchronously, of the next one to be assembled
  double zs[NA];
  double dpub[10];           // factored diagonal tile: reciprocal diagonal (4) and strict lower part (6)
  double rdiag[NA];          // reciprocal of the stored diagonal of L
  unsigned char pzf[16];     // per tile row of the current pivot panel: 1 = the tile is exactly zero
  double lyapC[16];          // curvature of the Lyapunov row in (p, v, theta, F) space (per instance)
  double red[2];
  double red5[5];            // block-wide reduction results (written by lane 0 of warp 0)
  uint64_t mask[NMAX + 1];
  double acc[NMAX + 1][5];   // per-stage partial results of the eval / slack-step / trial passes
  int flag;
  long long prof[PF_COUNT];
  short csr_ptr[NX + 1];      // structural pattern of [B A] by rows (gather form)
  unsigned char csr_idx[NZ * 4];
  signed char barow[NZ * 4];  // ba_row(j, q) table
  unsigned short baofs[NZ * 4];  // row offset ba_row * NZ into W of slot (j, q), 0 for an empty slot (whose [B A] value is 0)
  unsigned char barz[NZ * 4];    // ba_row with 0 for an empty slot
  unsigned short ly_m[NLY], ly_s[NLY];   // Lyapunov row scatter table: index into M, index of the coefficient in the record
  unsigned char ly_g[NLY];               // ... and the two gamma selectors (0: left, 1: right, 2: none), 4 bits each
  // options, instance and workspace pointers of the running solve.  The solver object itself lives in LOCAL memory on the GPU
  // (the interior-point loop is a function of its own), whose loads miss the thrashed L1 two times out of three: everything
  // the passes read per item comes from this copy instead (fixed shared-memory latency, address known at compile time)
  alignas(8) Config cfg;
  Instance inst;
  Work wk;
  // transaction barrier of the bulk asynchronous copies (GPU execution policy)
  alignas(8) unsigned long long mbar;
};

struct Stats { double cost, viol, kkt, mu; int iters, status, nfact, nreg; };

// ---------------------------------------------------------------------------------------------
// eval pass, one thread per stage: derivative record + KKT residual contributions.
// acc[i] = {prim_inf, dual_inf, max s*lam, min s*lam, sum |y|+|lam|, cost(with reg), theta, sum ln s}
CMPC_HD void stage_derivs(const Config& c, const Instance& in, const Work& w, int i, uint64_t mask, double mu,
                          double* acc) {
  const int N = c.N;
  const double* x = w.X + i * NX;
  const double* s = w.S + i * NR;
  const double* lam = w.LAM + i * NR;
  double* rec = w.REC + (size_t)i * RECSZ;
  double gc[NZ], m1[NZ], m2[NZ], gl_[NZ], diag[NZ];
  for (int j = 0; j < NZ; ++j) { gc[j] = 0; m1[j] = 0; m2[j] = 0; gl_[j] = 0; diag[j] = 0; }
  double prim = 0.0, smax = 0.0, smin = 1e300, lsum = 0.0, theta = 0.0, lns = 0.0;
  double g[NR];
  // helper: account one active row r with sparse Jacobian entries (idx, val) pairs
  auto row = [&](int r, double gval, const int* idx, const double* val, int nnz, bool explicit_row = false) {
    const double rg = gval - c.relax + s[r];
    const double sig = lam[r] / s[r];
    if (!explicit_row) rec[Q_RG + r] = rg;
    for (int t = 0; t < nnz; ++t) {
      if (!explicit_row) { m1[idx[t]] += val[t] / s[r]; m2[idx[t]] += sig * rg * val[t]; }
      gl_[idx[t]] += lam[r] * val[t];
    }
    const double ar = fabs(rg);
    prim = ar > prim ? ar : prim;                                 // (theta, sum ln s come from the trial pass)
    const double sl = s[r] * lam[r];
    smax = sl > smax ? sl : smax; smin = sl < smin ? sl : smin; lsum += lam[r];
    return sig;
  };

  if (i >= 1) {                                                 // tracking cost of stage i (ref col i-1)
    const double* ref = in.com_ref + 9 * (i - 1);
    const double* fr = in.foot_ref + 8 * (i - 1);
    const double gl = in.gamma[2 * i], gr = in.gamma[2 * i + 1];
    const double wz = wz_of(c, i - 1);
    gc[32 + 0] += 2.0 * c.w_xy * (x[0] - ref[0]); diag[32 + 0] += 2.0 * c.w_xy;
    gc[32 + 1] += 2.0 * c.w_xy * (x[1] - ref[1]); diag[32 + 1] += 2.0 * c.w_xy;
    gc[32 + 2] += 2.0 * wz * (x[2] - ref[2]);      diag[32 + 2] += 2.0 * wz;
    for (int j = 0; j < 3; ++j) {
      gc[32 + IPL + j] += 2.0 * c.w_foot * gl * (x[IPL + j] - fr[j]);     diag[32 + IPL + j] += 2.0 * c.w_foot * gl;
      gc[32 + IPR + j] += 2.0 * c.w_foot * gr * (x[IPR + j] - fr[3 + j]); diag[32 + IPR + j] += 2.0 * c.w_foot * gr;
    }
    gc[32 + IPSL] += 2.0 * c.w_foot * gl * (x[IPSL] - fr[6]); diag[32 + IPSL] += 2.0 * c.w_foot * gl;
    gc[32 + IPSR] += 2.0 * c.w_foot * gr * (x[IPSR] - fr[7]); diag[32 + IPSR] += 2.0 * c.w_foot * gr;
    // box rows
    for (int e = 0; e < 2; ++e) {
      if (!(mask & (1ull << (R_BOX + 6 * e)))) continue;
      for (int j = 0; j < 3; ++j) {
        const int xi = 32 + (e ? IPR : IPL) + j;
        const double err = x[(e ? IPR : IPL) + j] - fr[3 * e + j];
        const double one = 1.0, mone = -1.0;
        diag[xi] += row(R_BOX + 6 * e + 2 * j, err - c.box[j], &xi, &one, 1);
        diag[xi] += row(R_BOX + 6 * e + 2 * j + 1, -err - c.box[j], &xi, &mone, 1);
      }
    }
  }
  if (mask & (1ull << R_PZ)) {
    const int xi = 32 + IP + 2; const double one = 1.0;
    diag[xi] += row(R_PZ, x[IP + 2] - c.pz_max, &xi, &one, 1);
  }
  double cost = (i >= 1) ? track_cost(c, in, i, x) : 0.0;
  double dual = 0.0, ysum = 0.0;
  if (i < N) {
    const double* u = w.U + i * NU;
    const double gl = in.gamma[2 * i], gr = in.gamma[2 * i + 1];
    const double d = c.delta;
    cost = stage_cost(c, in, i, x, u, true);
    Arms a; lever_arms(c, x, a);
    double* bav = rec + Q_BA;
    ba_values(c, in, i, u, a, bav);
    // --- cost gradient / diagonal Hessian of the input terms
    for (int j = 0; j < 3; ++j) { gc[32 + IH + j] += 2.0 * c.w_h * x[IH + j]; diag[32 + IH + j] += 2.0 * c.w_h; }
    for (int e = 0; e < 2; ++e) {
      const double ge = e ? gr : gl;
      const double gp = (i >= 1) ? in.gamma[2 * (i - 1) + e] * c.w_rate : 0.0;
      const double* f = u + 12 * e;
      double mean[3] = {0, 0, 0};
      for (int k = 0; k < 4; ++k) for (int j = 0; j < 3; ++j) mean[j] += 0.25 * f[3 * k + j];
      for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 3; ++j) {
          const int ui = 12 * e + 3 * k + j;
          gc[ui] += ge * 2.0 * c.w_sym * (f[3 * k + j] - mean[j]) + (1.0 - ge) * 2.0 * c.w_swing * f[3 * k + j];
          diag[ui] += ge * 2.0 * c.w_sym * 0.75 + (1.0 - ge) * 2.0 * c.w_swing;
          if (j == 2) {
            const int qi = 32 + IQ + 4 * e + k;
            const double dz = f[3 * k + 2] - x[IQ + 4 * e + k];
            gc[ui] += 2.0 * gp * dz; gc[qi] -= 2.0 * gp * dz;
            diag[ui] += 2.0 * gp; diag[qi] += 2.0 * gp;
          }
        }
      rec[Q_GAM + e] = ge; rec[Q_GAMP + e] = gp;
    }
    for (int j = 24; j < 32; ++j) { gc[j] += 2.0 * c.eps_reg * u[j]; diag[j] += 2.0 * c.eps_reg; }
    // --- friction / unilateral rows
    for (int v = 0; v < 8; ++v) {
      double* blk = rec + Q_FRIC + 6 * v;
      for (int t = 0; t < 6; ++t) blk[t] = 0.0;
      if (!(mask & (1ull << (R_UNI + v)))) continue;
      const double* f = u + 3 * v;
      const double mf = c.mu_fric;
      int idx[2]; double val[2]; double sg[5];
      idx[0] = 3 * v; idx[1] = 3 * v + 2; val[1] = -mf;
      val[0] = 1.0;  sg[0] = row(R_FRIC + 4 * v + 0, f[0] - mf * f[2], idx, val, 2);
      val[0] = -1.0; sg[1] = row(R_FRIC + 4 * v + 1, -f[0] - mf * f[2], idx, val, 2);
      idx[0] = 3 * v + 1;
      val[0] = 1.0;  sg[2] = row(R_FRIC + 4 * v + 2, f[1] - mf * f[2], idx, val, 2);
      val[0] = -1.0; sg[3] = row(R_FRIC + 4 * v + 3, -f[1] - mf * f[2], idx, val, 2);
      idx[0] = 3 * v + 2; val[0] = -1.0;
      sg[4] = row(R_UNI + v, -f[2], idx, val, 1);
      blk[0] = sg[0] + sg[1];                 // xx
      blk[1] = 0.0;                           // xy
      blk[2] = -mf * (sg[0] - sg[1]);         // xz
      blk[3] = sg[2] + sg[3];                 // yy
      blk[4] = -mf * (sg[2] - sg[3]);         // yz
      blk[5] = mf * mf * (sg[0] + sg[1] + sg[2] + sg[3]) + sg[4];   // zz
      // (the diagonal entries travel with the other diagonal terms: the assembly then writes every entry once)
      diag[3 * v] += blk[0]; diag[3 * v + 1] += blk[3]; diag[3 * v + 2] += blk[5];
      blk[0] = 0.0; blk[3] = 0.0; blk[5] = 0.0;
    }
    // --- Lyapunov row
    {
      double F[3] = {0, 0, 0};
      for (int v = 0; v < 8; ++v) {
        const double ge = (v < 4) ? gl : gr;
        F[0] += ge * u[3 * v]; F[1] += ge * u[3 * v + 1]; F[2] += ge * u[3 * v + 2];
      }
      double G[12], C[16];
      const double q = lyapunov(c, in, i, x, F, G, C);
      int idx[33]; double val[33]; int n = 0;
      for (int v = 0; v < 8; ++v) {
        const double ge = (v < 4) ? gl : gr;
        if (ge < 0.5) continue;
        for (int j = 0; j < 3; ++j) { idx[n] = 3 * v + j; val[n] = G[9 + j]; ++n; }
      }
      for (int j = 0; j < 3; ++j) { idx[n] = 32 + IP + j; val[n] = G[j]; ++n; }
      for (int j = 0; j < 3; ++j) { idx[n] = 32 + IV + j; val[n] = G[3 + j]; ++n; }
      for (int j = 0; j < 3; ++j) { idx[n] = 32 + ITH + j; val[n] = G[6 + j]; ++n; }
      const double sig = row(R_LYAP, q, idx, val, n, true);
      rec[Q_LRG] = q - c.relax + s[R_LYAP]; rec[Q_LLAM] = lam[R_LYAP];
      for (int t = 0; t < 12; ++t) rec[Q_LG + t] = G[t];
      for (int t = 0; t < 16; ++t) rec[Q_LC + t] = lam[R_LYAP] * C[t];
      rec[Q_LSIG] = sig;
    }
    // --- dynamics defect and predicted state
    double xp[NX];
    dyn_step(c, in, i, x, u, xp);
    const double* xn = w.X + (i + 1) * NX;
    for (int j = 0; j < NX; ++j) {
      const double dj = xp[j] - xn[j];
      rec[Q_D + j] = dj;
      const double ad = fabs(dj);
      prim = ad > prim ? ad : prim; theta += ad;
    }
    // --- angular momentum row (stage 0 only): ||h+||^2 - ||h||^2 <= 0, function of u only (x_0 fixed)
    rec[Q_HLAM] = 0.0; rec[Q_HSIG] = 0.0;
    if (mask & (1ull << R_HW)) {
      const double ghw = xp[IH] * xp[IH] + xp[IH + 1] * xp[IH + 1] + xp[IH + 2] * xp[IH + 2]
                       - (x[IH] * x[IH] + x[IH + 1] * x[IH + 1] + x[IH + 2] * x[IH + 2]);
      int idx[24]; double val[24];
      for (int j = 0; j < 24; ++j) {
        // d h+/d u_j : rows IH+(a+1)%3 (slot 1) and IH+(a+2)%3 (slot 2)
        const int ax = j % 3;
        idx[j] = j;
        val[j] = 2.0 * (xp[IH + (ax + 1) % 3] * bav[4 * j + 1] + xp[IH + (ax + 2) % 3] * bav[4 * j + 2]);
      }
      const double sig = row(R_HW, ghw, idx, val, 24, true);
      rec[Q_HRG] = ghw - c.relax + s[R_HW];
      for (int j = 0; j < 3; ++j) rec[Q_HP + j] = xp[IH + j];
      rec[Q_HLAM] = lam[R_HW]; rec[Q_HSIG] = sig;
    }
    // --- Lagrangian curvature of the bilinear torque term: y_h' (r x f)
    const double* yn = w.Y + (i + 1) * NX;
    for (int j = 0; j < 3; ++j) rec[Q_YH + j] = d * yn[IH + j];
    for (int e = 0; e < 2; ++e) {
      const double ge = e ? gr : gl;
      double acc2 = 0.0;                  // -d*g * sum_k y . ((R c_k) x f_k)
      for (int k = 0; k < 4; ++k) {
        const int v = 4 * e + k; const double* f = u + 3 * v;
        const double cx = a.rc[v][0], cy = a.rc[v][1];
        acc2 += yn[IH] * (cy * f[2]) + yn[IH + 1] * (-cx * f[2]) + yn[IH + 2] * (cx * f[1] - cy * f[0]);
      }
      diag[32 + (e ? IPSR : IPSL)] += -d * ge * acc2;
      // store R'c_k for the (f, psi) mixed terms
      for (int k = 0; k < 4; ++k) { rec[Q_DR + 2 * (4 * e + k)] = a.dr[4 * e + k][0]; rec[Q_DR + 2 * (4 * e + k) + 1] = a.dr[4 * e + k][1]; }
    }
    // --- dual residual of stage i:  gc + J' lam + [B A]' y_{i+1} - [0; y_i]
    for (int j = 0; j < NZ; ++j) {
      double r = gc[j] + gl_[j];
      for (int t = 0; t < 4; ++t) { const int rr = ba_row(j, t); if (rr >= 0) r += bav[4 * j + t] * yn[rr]; }
      if (j >= 32) r -= w.Y[i * NX + (j - 32)];
      if (i == 0 && j >= 32) r = 0.0;                      // x_0 is fixed
      const double ar = fabs(r); dual = ar > dual ? ar : dual;
#ifdef CMPC_TRACE
      cmpc_dbg_rd[j] = r;
#endif
    }
  } else {
    for (int j = 32; j < NZ; ++j) {
      const double r = gc[j] + gl_[j] - w.Y[i * NX + (j - 32)];
      const double ar = fabs(r); dual = ar > dual ? ar : dual;
    }
  }
  for (int j = 0; j < NX; ++j) ysum += fabs(w.Y[i * NX + j]);
  for (int j = 0; j < NZ; ++j) { rec[Q_GC + j] = gc[j]; rec[Q_M1 + j] = m1[j]; rec[Q_M2 + j] = m2[j]; rec[Q_DIAG + j] = diag[j]; }
  acc[0] = prim; acc[1] = dual; acc[2] = smax; acc[3] = smin; acc[4] = ysum + lsum; acc[5] = cost; acc[6] = theta; acc[7] = lns;
  (void)g; (void)mu;
}

// trial evaluation for the line search, one thread per stage: theta, cost, sum ln s at
// (x + a dx, u + a du, s + a ds).  acc[i] = {theta, cost(with reg), sum ln s, max unrelaxed violation, cost(ref)}
CMPC_HD void stage_trial(const Config& c, const Instance& in, const Work& w, int i, uint64_t mask, double alpha, double* acc) {
  const int N = c.N;
  double x[NX], u[NU], xn[NX], xp[NX], g[NR];
  for (int j = 0; j < NX; ++j) x[j] = w.X[i * NX + j] + alpha * w.DX[i * NX + j];
  double theta = 0.0, lns = 0.0, viol = 0.0;
  if (i < N) {
    for (int j = 0; j < NU; ++j) u[j] = w.U[i * NU + j] + alpha * w.DU[i * NU + j];
    for (int j = 0; j < NX; ++j) xn[j] = w.X[(i + 1) * NX + j] + alpha * w.DX[(i + 1) * NX + j];
    dyn_step(c, in, i, x, u, xp);
    for (int j = 0; j < NX; ++j) { const double ad = fabs(xp[j] - xn[j]); theta += ad; viol = ad > viol ? ad : viol; }
  } else {
    for (int j = 0; j < NU; ++j) u[j] = 0.0;
    for (int j = 0; j < NX; ++j) xp[j] = x[j];
  }
  stage_ineq(c, in, i, mask, x, u, xp, g);
  for (int r = 0; r < NR; ++r) {
    if (!(mask & (1ull << r))) continue;
    const double st = w.S[i * NR + r] + alpha * w.DS[i * NR + r];
    theta += fabs(g[r] - c.relax + st);
    lns += log(st);
    viol = g[r] > viol ? g[r] : viol;
  }
  acc[0] = theta; acc[1] = stage_cost(c, in, i, x, u, true); acc[2] = lns; acc[3] = viol;
  acc[4] = stage_cost(c, in, i, x, u, false);
}

// Per-stage scratch of the eval pass (aliases the stage-block storage M | W | P, idle outside the backward sweep).
struct EvalScratch {
  double cs[2], sn[2];       // cos / sin of the two foot yaws
  double part[8][10];        // per vertex: f (3), r x f (3), (R'c) x f (3), y_h . ((R c) x f)
  double sum[17];            // Fe[2][3], T[3] (gamma-weighted), dT[2][3], psq[2]
  double F[3], xph[3];       // total contact force, predicted angular momentum
  double LG[12];             // gradient of the Lyapunov row wrt (p, v, theta, F)
  double sc[6];              // sigma, residual, multiplier of the Lyapunov row and of the angular-momentum row
  double st[11][5];          // per role: prim_inf, dual_inf, max s*lam, min s*lam, sum |y| + lam
};
constexpr int ECH = (int)((sizeof(double) * (MSZ + NX * NZ + NX * NX)) / sizeof(EvalScratch));   // stages per chunk (25)
static_assert(ECH >= 8, "eval scratch does not fit");

// Per-stage scratch of the trial pass (same storage as the eval scratch).
struct TrialScratch {
  double cs[2], sn[2];
  double part[8][7];         // per vertex: f (3), r x f (3), |f|^2
  double sum[11];            // Fe[2][3], T[3] (gamma-weighted), s2[2]
  double st[10][5];          // per role: theta, cost (with reg), sum ln s, max violation, cost (reference)
};
constexpr int TCH = (int)((sizeof(double) * (MSZ + NX * NZ + NX * NX)) / sizeof(TrialScratch));   // stages per chunk (37)
static_assert(TCH >= 8, "trial scratch does not fit");

// ---------------------------------------------------------------------------------------------
template <class Par>
struct Solver {
  // (the shared-memory block is not a member: every member function takes it from the execution policy, which on the
  // GPU derives it from the CTA's dynamic shared-memory symbol -- the compiler then knows the address space and emits
  // LDS / STS instead of generic loads and stores)
  const Config& c0; const Instance& in0; Work w0; Par& par;    // as given; the passes read the shared-memory copies C(), IN(), W()
  double mu, reg_last, mu_scale;
  int nfact, nreg;
  bool start_blend;                           // cold start variant of the last retry
  int tile_i[Par::TPT], tile_j[Par::TPT];     // this thread's 4 x 4 register tiles of the stage block (row, column; -1 = none)

  CMPC_HD Solver(const Config& c_, const Instance& in_, const Work& w_, Smem& sm_, Par& par_)
      : c0(c_), in0(in_), w0(w_), par(par_), mu(0), reg_last(0), mu_scale(1.0), nfact(0), nreg(0), start_blend(false) { par.bind(&sm_); }

  CMPC_HD const Config& C() const { return par.template smem<Smem>().cfg; }
  CMPC_HD const Instance& IN() const { return par.template smem<Smem>().inst; }
  CMPC_HD const Work& W() const { return par.template smem<Smem>().wk; }

  // workspace sections, known to be global memory
  CMPC_HD double* gX() const { double* p = W().X; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gU() const { double* p = W().U; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gY() const { double* p = W().Y; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gS() const { double* p = W().S; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gLAM() const { double* p = W().LAM; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gDX() const { double* p = W().DX; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gDU() const { double* p = W().DU; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gDS() const { double* p = W().DS; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gYN() const { double* p = W().YN; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gDW() const { double* p = W().DW; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gREC() const { double* p = W().REC; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD double* gFAC() const { double* p = W().FAC; CMPC_ASSUME_GLOBAL(p); return p; }

  // instance data, global memory as well
  CMPC_HD const double* i_x0() const { const double* p = IN().x0; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD const double* i_com_ref() const { const double* p = IN().com_ref; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD const double* i_foot_ref() const { const double* p = IN().foot_ref; CMPC_ASSUME_GLOBAL(p); return p; }
  CMPC_HD const double* i_gamma() const { const double* p = IN().gamma; CMPC_ASSUME_GLOBAL(p); return p; }

  CMPC_HD static int tri(int r, int cidx) { return r * (r + 1) / 2 + cidx; }

  // ---- initial point.  warm: 0 = cold (solver's own guess), 1 = primal (X, U given; slacks/duals reset as
  // IPOPT does), 2 = full (X, U, Y, S, LAM given).
  CMPC_HD void init_point(int warm) {
    Smem& sm = par.template smem<Smem>();
    const int N = C().N, tid = par.tid(), nt = par.nt();
    if (warm == 0) {
      for (int t = tid; t < (N + 1) * NX; t += nt) {
        const int i = t / NX, j = t % NX;
        double v = (j < NXP) ? i_x0()[j] : 0.0;
        // alternative start (last retry): CoM position / velocity blended from x0 towards the reference along the horizon
        if (start_blend && i >= 1 && j < 6) { const double a = (double)i / (double)N; v = (1.0 - a) * v + a * i_com_ref()[9 * (i - 1) + j]; }
        gX()[t] = v;
      }
      for (int t = tid; t < N * NU; t += nt) {
        const int i = t / NU, j = t % NU;
        const double gl = i_gamma()[2 * i], gr = i_gamma()[2 * i + 1];
        double v = 0.0;
        if (j < 24 && j % 3 == 2) {
          const double ge = (j < 12) ? gl : gr;
          v = ge * IN().mass * C().grav / (4.0 * (gl + gr > 0.5 ? gl + gr : 1.0));
        }
        gU()[t] = v;
      }
      par.sync();
      for (int t = tid; t < (N + 1) * NQ; t += nt) {
        const int i = t / NQ, v = t % NQ;
        gX()[i * NX + IQ + v] = (i >= 1) ? gU()[(i - 1) * NU + 3 * v + 2] : 0.0;
      }
    }
    if (warm == 4) {
      // automatic: shift while a landing lies inside the horizon (the switch moves one stage per tick, so the previous tick's
      // stage i + 1 is the better start for stage i: 20-27 instead of 26-36 iterations on those ticks), else keep the stage
      // alignment (references move 1.5 mm per tick: 5-6 instead of 8 iterations)
      bool sw = false;
      for (int i = 1; i <= N; ++i)                                   // a landing: a foot's gamma going 0 -> 1
        sw = sw || (i_gamma()[2 * i] > i_gamma()[2 * i - 2]) || (i_gamma()[2 * i + 1] > i_gamma()[2 * i - 1]);
      warm = sw ? 3 : 2;
    }
    if (warm == 3) {
      // MPC shift: the previous tick's stage i+1 becomes this tick's stage i (the last stage is repeated).  Staged
      // through the step buffers so the in-place move is race free.
      for (int t = tid; t < N * NX; t += nt) { gDX()[t] = gX()[t + NX]; gYN()[t] = gY()[t + NX]; }
      for (int t = tid; t < (N - 1) * NU; t += nt) gDU()[t] = gU()[t + NU];
      for (int t = tid; t < (N - 1) * NR; t += nt) {
        const int r = t % NR;
        // rows that exist only at stage 0 (angular momentum) keep their own history
        gDS()[t] = (r == R_HW && t < NR) ? gS()[t] : gS()[t + NR];
      }
      par.sync();
      for (int t = tid; t < N * NX; t += nt) { gX()[t] = gDX()[t]; gY()[t] = gYN()[t]; }
      for (int t = tid; t < (N - 1) * NU; t += nt) gU()[t] = gDU()[t];
      for (int t = tid; t < (N - 1) * NR; t += nt) gS()[t] = gDS()[t];
      par.sync();
      for (int t = tid; t < (N - 1) * NR; t += nt) { const int r = t % NR; gDS()[t] = (r == R_HW && t < NR) ? gLAM()[t] : gLAM()[t + NR]; }
      par.sync();
      for (int t = tid; t < (N - 1) * NR; t += nt) gLAM()[t] = gDS()[t];
      par.sync();
      warm = 2;
    }
    for (int t = tid; t < NX; t += nt) gX()[t] = (t < NXP) ? i_x0()[t] : 0.0;      // x_0 is data
    if (warm < 2) for (int t = tid; t < (N + 1) * NX; t += nt) gY()[t] = 0.0;
    par.sync();
    if (warm < 2) {
      mu = C().mu_init * mu_scale;
      for (int i = tid; i <= N; i += nt) {
        double x[NX], u[NU], xp[NX], g[NR];
        for (int j = 0; j < NX; ++j) x[j] = gX()[i * NX + j];
        if (i < N) { for (int j = 0; j < NU; ++j) u[j] = gU()[i * NU + j]; dyn_step(C(), IN(), i, x, u, xp); }
        else { for (int j = 0; j < NU; ++j) u[j] = 0.0; for (int j = 0; j < NX; ++j) xp[j] = x[j]; }
        stage_ineq(C(), IN(), i, sm.mask[i], x, u, xp, g);
        for (int r = 0; r < NR; ++r) {
          double sv = 1.0, lv = 0.0;
          if (sm.mask[i] & (1ull << r)) { sv = -(g[r] - C().relax); sv = sv > C().bound_push ? sv : C().bound_push; lv = 1.0; }
          gS()[i * NR + r] = sv; gLAM()[i * NR + r] = lv;
        }
      }
    } else {
      mu = C().mu_warm;
      // keep the previous slacks/multipliers but push them off the boundary: s >= sqrt(mu)*1e-2, lam = mu/s floor
      for (int t = tid; t < (N + 1) * NR; t += nt) {
        const int i = t / NR, r = t % NR;
        if (!(sm.mask[i] & (1ull << r))) { gS()[t] = 1.0; gLAM()[t] = 0.0; continue; }
        double sv = gS()[t], lv = gLAM()[t];
        if (!(sv > C().warm_push)) sv = C().warm_push;
        if (!(lv > mu / sv * 1e-3)) lv = mu / sv * 1e-3;
        if (C().warm_comp > 0.0 && lv > mu / sv * C().warm_comp) lv = mu / sv * C().warm_comp;
        gS()[t] = sv; gLAM()[t] = lv;
      }
    }
    par.sync();
  }

  // ---- eval pass + reductions.  out: prim, dual, smax, smin, scale sums (out[5..7] unused: the merit quantities come
  // from the trial pass).  CTA-wide: work items are (stage, role) pairs -- 8 vertex roles (friction rows, force columns),
  // a CoM role (p, v, h, theta) and a feet role (foot poses, foot inputs) -- in short barrier-separated phases, all
  // temporaries in registers or in the per-stage scratch: no per-thread arrays.
  CMPC_HD static void stat_row(double* st, double rg, double s, double lam) {
    const double ar = fabs(rg), sl = s * lam;
    st[0] = ar > st[0] ? ar : st[0]; st[2] = sl > st[2] ? sl : st[2]; st[3] = sl < st[3] ? sl : st[3]; st[4] += lam;
  }
  CMPC_HD static void stat_dual(double* st, double r) { const double ar = fabs(r); st[1] = ar > st[1] ? ar : st[1]; }

  CMPC_HD void eval(double* out) {
    Smem& sm = par.template smem<Smem>();
    const int N = C().N, tid = par.tid(), nt = par.nt();
    EvalScratch* es = reinterpret_cast<EvalScratch*>(sm.M);
    const double d = C().delta, m = IN().mass, k1 = IN().k1;
    for (int i0 = 0; i0 <= N; i0 += ECH) {
      const int ns = (N + 1 - i0) < ECH ? (N + 1 - i0) : ECH;
      // ---- P0: yaw sines / cosines; role statistics cleared
      for (int t = tid; t < ns * 2; t += nt) {
        const int il = t >> 1, e = t & 1, i = i0 + il;
        const double psi = gX()[i * NX + (e ? IPSR : IPSL)];
        es[il].cs[e] = cos(psi); es[il].sn[e] = sin(psi);
      }
      for (int t = tid; t < ns * 11; t += nt) {
        double* st = es[t / 11].st[t % 11];
        st[0] = 0.0; st[1] = 0.0; st[2] = 0.0; st[3] = 1e300; st[4] = 0.0;
      }
      par.sync();
      // ---- P1: per-vertex partial sums
      for (int t = tid; t < ns * 8; t += nt) {
        const int il = t >> 3, v = t & 7, i = i0 + il;
        if (i >= N) continue;
        const int e = v >> 2, k = v & 3;
        const double* x = gX() + i * NX; const double* u = gU() + i * NU; const double* yn = gY() + (i + 1) * NX;
        const double* pe = x + (e ? IPR : IPL);
        const double cs = es[il].cs[e], sn = es[il].sn[e];
        double cx, cy; corner(C(), k, cx, cy);
        const double rx = cs * cx - sn * cy, ry = sn * cx + cs * cy;
        const double r0 = rx + pe[0] - x[0], r1 = ry + pe[1] - x[1], r2 = pe[2] - x[2];
        const double dr0 = -sn * cx - cs * cy, dr1 = cs * cx - sn * cy;
        const double f0 = u[3 * v], f1 = u[3 * v + 1], f2 = u[3 * v + 2];
        double* p = es[il].part[v];
        p[0] = f0; p[1] = f1; p[2] = f2;
        p[3] = r1 * f2 - r2 * f1; p[4] = r2 * f0 - r0 * f2; p[5] = r0 * f1 - r1 * f0;
        p[6] = dr1 * f2; p[7] = -dr0 * f2; p[8] = dr0 * f1 - dr1 * f0;
        p[9] = yn[IH] * (ry * f2) + yn[IH + 1] * (-rx * f2) + yn[IH + 2] * (rx * f1 - ry * f0);
      }
      par.sync();
      // ---- P2: sums over the vertices
      for (int t = tid; t < ns * 17; t += nt) {
        const int il = t / 17, q = t - il * 17, i = i0 + il;
        if (i >= N) continue;
        const double gl = i_gamma()[2 * i], gr = i_gamma()[2 * i + 1];
        const double (*pt)[10] = es[il].part;
        double s = 0.0;
        if (q < 6) { const int e = q / 3, j = q - 3 * e; for (int k = 0; k < 4; ++k) s += pt[4 * e + k][j]; }
        else if (q < 9) { const int j = q - 6; for (int v = 0; v < 8; ++v) s += (v < 4 ? gl : gr) * pt[v][3 + j]; }
        else if (q < 15) { const int e = (q - 9) / 3, j = (q - 9) - 3 * e; for (int k = 0; k < 4; ++k) s += pt[4 * e + k][6 + j]; }
        else { const int e = q - 15; for (int k = 0; k < 4; ++k) s += pt[4 * e + k][9]; }
        es[il].sum[q] = s;
      }
      par.sync();
      // ---- P3: stage scalars: total force, predicted angular momentum, the two explicit rows
      for (int il = tid; il < ns; il += nt) {
        const int i = i0 + il;
        if (i >= N) continue;
        EvalScratch& E = es[il];
        const double* x = gX() + i * NX; const double* yn = gY() + (i + 1) * NX;
        const double* s = gS() + i * NR; const double* lam = gLAM() + i * NR;
        double* rec = gREC() + (size_t)i * RECSZ;
        const uint64_t mask = sm.mask[i];
        const double gl = i_gamma()[2 * i], gr = i_gamma()[2 * i + 1];
        for (int j = 0; j < 3; ++j) { E.F[j] = gl * E.sum[j] + gr * E.sum[3 + j]; E.xph[j] = x[IH + j] + d * E.sum[6 + j]; }
        const double q = lyapunov_grad(C(), IN(), i, x, E.F, E.LG);
        double* st = E.st[10];
        {
          const double sv = s[R_LYAP], lv = lam[R_LYAP], rg = q - C().relax + sv;
          stat_row(st, rg, sv, lv);
          E.sc[0] = lv / sv; E.sc[1] = rg; E.sc[2] = lv;
          rec[Q_LSIG] = lv / sv; rec[Q_LRG] = rg; rec[Q_LLAM] = lv;
          for (int t = 0; t < 16; ++t) rec[Q_LC + t] = lv * sm.lyapC[t];
          for (int t = 0; t < 12; ++t) rec[Q_LG + t] = E.LG[t];
        }
        E.sc[3] = 0.0; E.sc[4] = 0.0; E.sc[5] = 0.0;
        if (mask & (1ull << R_HW)) {
          const double ghw = E.xph[0] * E.xph[0] + E.xph[1] * E.xph[1] + E.xph[2] * E.xph[2]
                           - (x[IH] * x[IH] + x[IH + 1] * x[IH + 1] + x[IH + 2] * x[IH + 2]);
          const double sv = s[R_HW], lv = lam[R_HW], rg = ghw - C().relax + sv;
          stat_row(st, rg, sv, lv);
          E.sc[3] = lv / sv; E.sc[4] = rg; E.sc[5] = lv;
          rec[Q_HRG] = rg;
        }
        rec[Q_HSIG] = E.sc[3]; rec[Q_HLAM] = E.sc[5];
        for (int j = 0; j < 3; ++j) { rec[Q_HP + j] = E.xph[j]; rec[Q_YH + j] = d * yn[IH + j]; }
        for (int e = 0; e < 2; ++e) {
          rec[Q_GAM + e] = e ? gr : gl;
          rec[Q_GAMP + e] = (i >= 1) ? i_gamma()[2 * (i - 1) + e] * C().w_rate : 0.0;
        }
      }
      par.sync();
      // ---- P4: the roles (role-major item order: warps stay on one code path)
      for (int t = tid; t < ns * 10; t += nt) {
        const int role = t / ns, il = t - role * ns, i = i0 + il;
        const bool has_u = i < N;
        EvalScratch& E = es[il];
        const double* x = gX() + i * NX; const double* yi = gY() + i * NX;
        const double* u = gU() + (has_u ? i : 0) * NU;                 // (not read at the terminal stage)
        const double* yn = gY() + (has_u ? i + 1 : i) * NX;
        const double* xn = gX() + (has_u ? i + 1 : i) * NX;
        const double* s = gS() + i * NR; const double* lam = gLAM() + i * NR;
        double* rec = gREC() + (size_t)i * RECSZ;
        const uint64_t mask = sm.mask[i];
        const double gl = i_gamma()[2 * i], gr = i_gamma()[2 * i + 1];
        double* st = E.st[role];
        const double lamL = has_u ? E.sc[2] : 0.0, lamH = has_u ? E.sc[5] : 0.0;
        const bool x_free = i >= 1;                                   // x_0 is data: its dual residual is not a residual
        if (role < 8) {
          // ---------------- vertex v: three force inputs and the previous-f_z state q_v
          const int v = role, e = v >> 2, k = v & 3;
          const int qi = 32 + IQ + v;
          if (!has_u) {
            rec[Q_GC + qi] = 0.0; rec[Q_DIAG + qi] = 0.0; rec[Q_M1 + qi] = 0.0; rec[Q_M2 + qi] = 0.0;
            stat_dual(st, -yi[IQ + v]); st[4] += fabs(yi[IQ + v]);
            continue;
          }
          const double ge = e ? gr : gl;
          const double* pe = x + (e ? IPR : IPL);
          const double cs = E.cs[e], sn = E.sn[e];
          double cx, cy; corner(C(), k, cx, cy);
          const double rx = cs * cx - sn * cy, ry = sn * cx + cs * cy;
          const double rr[3] = {rx + pe[0] - x[0], ry + pe[1] - x[1], pe[2] - x[2]};
          rec[Q_DR + 2 * v] = -sn * cx - cs * cy; rec[Q_DR + 2 * v + 1] = cs * cx - sn * cy;
          const double fv[3] = {u[3 * v], u[3 * v + 1], u[3 * v + 2]};
          const double gp = (i >= 1) ? i_gamma()[2 * (i - 1) + e] * C().w_rate : 0.0;
          const double dz = fv[2] - x[IQ + v];
          double a_gc[3], a_dg[3], a_m1[3] = {0, 0, 0}, a_m2[3] = {0, 0, 0}, a_gl[3] = {0, 0, 0};
          for (int j = 0; j < 3; ++j) {
            const double mean = 0.25 * E.sum[3 * e + j];
            a_gc[j] = ge * 2.0 * C().w_sym * (fv[j] - mean) + (1.0 - ge) * 2.0 * C().w_swing * fv[j];
            a_dg[j] = ge * 2.0 * C().w_sym * 0.75 + (1.0 - ge) * 2.0 * C().w_swing;
          }
          a_gc[2] += 2.0 * gp * dz; a_dg[2] += 2.0 * gp;
          double blk[6] = {0, 0, 0, 0, 0, 0};
          if (mask & (1ull << (R_UNI + v))) {
            const double mf = C().mu_fric;
            const double gv[5] = {fv[0] - mf * fv[2], -fv[0] - mf * fv[2], fv[1] - mf * fv[2], -fv[1] - mf * fv[2], -fv[2]};
            double sg[5];
#pragma unroll
            for (int q = 0; q < 5; ++q) {
              const int r = (q < 4) ? R_FRIC + 4 * v + q : R_UNI + v;
              const double sv = s[r], lv = lam[r], inv = cmpc_rcp(sv), rg = gv[q] - C().relax + sv, sig = lv * inv;
              sg[q] = sig;
              stat_row(st, rg, sv, lv); rec[Q_RG + r] = rg;
              const int ax = (q < 2) ? 0 : (q < 4 ? 1 : 2);
              const double ja = (q == 4) ? -1.0 : ((q & 1) ? -1.0 : 1.0);
              a_m1[ax] += ja * inv; a_m2[ax] += ja * sig * rg; a_gl[ax] += ja * lv;
              if (q < 4) { a_m1[2] += -mf * inv; a_m2[2] += -mf * sig * rg; a_gl[2] += -mf * lv; }
            }
            blk[0] = sg[0] + sg[1]; blk[2] = -mf * (sg[0] - sg[1]);
            blk[3] = sg[2] + sg[3]; blk[4] = -mf * (sg[2] - sg[3]);
            blk[5] = mf * mf * (sg[0] + sg[1] + sg[2] + sg[3]) + sg[4];
            a_dg[0] += blk[0]; a_dg[1] += blk[3]; a_dg[2] += blk[5];        // diagonal entries travel with the other diagonal terms
            blk[0] = 0.0; blk[3] = 0.0; blk[5] = 0.0;
          }
#pragma unroll
          for (int q = 0; q < 6; ++q) rec[Q_FRIC + 6 * v + q] = blk[q];
          const bool hw = (mask >> R_HW) & 1ull;
#pragma unroll
          for (int ax = 0; ax < 3; ++ax) {
            const int ui = 3 * v + ax;
            const double c0 = d * ge / m, c1 = d * ge * rr[(ax + 2) % 3], c2 = -d * ge * rr[(ax + 1) % 3], c3 = (ax == 2) ? 1.0 : 0.0;
            double* col = rec + Q_BA + 4 * ui;
            col[0] = c0; col[1] = c1; col[2] = c2; col[3] = c3;
            double glv = a_gl[ax] + lamL * ge * E.LG[9 + ax];
            if (hw) glv += lamH * 2.0 * (E.xph[(ax + 1) % 3] * c1 + E.xph[(ax + 2) % 3] * c2);
            double r = a_gc[ax] + glv + c0 * yn[IV + ax] + c1 * yn[IH + (ax + 1) % 3] + c2 * yn[IH + (ax + 2) % 3];
            if (ax == 2) r += yn[IQ + v];
            stat_dual(st, r);
#ifdef CMPC_TRACE
            cmpc_dbg_rd[ui] = r;
#endif
            rec[Q_GC + ui] = a_gc[ax]; rec[Q_DIAG + ui] = a_dg[ax]; rec[Q_M1 + ui] = a_m1[ax]; rec[Q_M2 + ui] = a_m2[ax];
          }
          {
            double* col = rec + Q_BA + 4 * qi;
            col[0] = 0.0; col[1] = 0.0; col[2] = 0.0; col[3] = 0.0;
            const double gq = -2.0 * gp * dz;
            rec[Q_GC + qi] = gq; rec[Q_DIAG + qi] = 2.0 * gp; rec[Q_M1 + qi] = 0.0; rec[Q_M2 + qi] = 0.0;
            const double r = x_free ? gq - yi[IQ + v] : 0.0;
            stat_dual(st, r);
#ifdef CMPC_TRACE
            cmpc_dbg_rd[qi] = r;
#endif
            st[4] += fabs(yi[IQ + v]);
            // dynamics defect of q_v+ = f_z
            const double dj = fv[2] - xn[IQ + v];
            rec[Q_D + IQ + v] = dj;
            const double ad = fabs(dj); st[0] = ad > st[0] ? ad : st[0];
          }
        } else if (role == 8) {
          // ---------------- CoM role: p, v, h, theta
          const double* refp = i_com_ref() + 9 * (i >= 1 ? i - 1 : 0);         // tracking reference column i-1
          const double* ref = i_com_ref() + 9 * (has_u ? i : 0);               // dynamics / Lyapunov reference column i
          const double wz = (i >= 1) ? wz_of(C(), i - 1) : 0.0;
#pragma unroll
          for (int cidx = 0; cidx < 12; ++cidx) {
            const int j = 32 + cidx, ax = cidx % 3;
            double gcv = 0.0, dgv = 0.0, m1v = 0.0, m2v = 0.0, glv = 0.0;
            double c0 = 1.0, c1 = 0.0, c2 = 0.0, c3 = 0.0, by = 0.0;
            if (cidx < 3) {
              if (i >= 1) { const double wq = (cidx == 2) ? wz : C().w_xy; gcv = 2.0 * wq * (x[cidx] - refp[cidx]); dgv = 2.0 * wq; }
              if (cidx == 2 && (mask & (1ull << R_PZ))) {
                const double sv = s[R_PZ], lv = lam[R_PZ], inv = cmpc_rcp(sv), rg = x[IP + 2] - C().pz_max - C().relax + sv, sig = lv * inv;
                stat_row(st, rg, sv, lv); rec[Q_RG + R_PZ] = rg;
                m1v = inv; m2v = sig * rg; glv = lv; dgv += sig;
              }
              if (has_u) {
                glv += lamL * E.LG[cidx];
                c1 = d * k1 / m; c2 = d * E.F[(ax + 2) % 3]; c3 = -d * E.F[(ax + 1) % 3];
                by = c0 * yn[IP + ax] + c1 * yn[ITH + ax] + c2 * yn[IH + (ax + 1) % 3] + c3 * yn[IH + (ax + 2) % 3];
              }
            } else if (cidx < 6) {
              if (has_u) {
                glv = lamL * E.LG[cidx];
                c0 = d; c1 = 1.0; c2 = d / m;
                by = c0 * yn[IP + ax] + c1 * yn[IV + ax] + c2 * yn[ITH + ax];
              }
            } else if (cidx < 9) {
              if (has_u) { gcv = 2.0 * C().w_h * x[cidx]; dgv = 2.0 * C().w_h; by = yn[cidx]; }
            } else {
              if (has_u) { glv = lamL * E.LG[cidx - 3]; by = yn[cidx]; }
            }
            if (has_u) { double* col = rec + Q_BA + 4 * j; col[0] = c0; col[1] = c1; col[2] = c2; col[3] = c3; }
            const double r = x_free ? gcv + glv + by - yi[cidx] : 0.0;
            stat_dual(st, r);
#ifdef CMPC_TRACE
            cmpc_dbg_rd[j] = r;
#endif
            st[4] += fabs(yi[cidx]);
            rec[Q_GC + j] = gcv; rec[Q_DIAG + j] = dgv; rec[Q_M1 + j] = m1v; rec[Q_M2 + j] = m2v;
            if (has_u) {
              double xp;
              if (cidx < 3) xp = x[cidx] + d * x[IV + cidx];
              else if (cidx < 6) xp = x[cidx] + d * ((cidx == 5 ? -C().grav : 0.0) + E.F[ax] / m);
              else if (cidx < 9) xp = E.xph[ax];
              else xp = x[cidx] + (d / m) * (k1 * (x[ax] - ref[ax]) + x[IV + ax] - ref[3 + ax]);
              const double dj = xp - xn[cidx];
              rec[Q_D + cidx] = dj;
              const double ad = fabs(dj); st[0] = ad > st[0] ? ad : st[0];
            }
          }
        } else {
          // ---------------- feet role: yaw / position states of both feet, foot velocity / yaw-rate inputs
          const double* fr = i_foot_ref() + 8 * (i >= 1 ? i - 1 : 0);
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double ge = e ? gr : gl;
            const int po = e ? IPSR : IPSL, xo = e ? IPR : IPL;
            const bool box = (i >= 1) && ((mask >> (R_BOX + 6 * e)) & 1ull);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {                                   // jj = 0: yaw, 1..3: position
              const int cidx = (jj == 0) ? po : xo + jj - 1, j = 32 + cidx, ax = jj - 1;
              double gcv = 0.0, dgv = 0.0, m1v = 0.0, m2v = 0.0, glv = 0.0;
              double c1 = 0.0, c2 = 0.0, c3 = 0.0, by = 0.0;
              if (jj == 0) {
                if (i >= 1) { gcv = 2.0 * C().w_foot * ge * (x[po] - fr[6 + e]); dgv = 2.0 * C().w_foot * ge; }
                if (has_u) {
                  dgv += -d * ge * E.sum[15 + e];
                  c1 = d * ge * E.sum[9 + 3 * e]; c2 = d * ge * E.sum[10 + 3 * e]; c3 = d * ge * E.sum[11 + 3 * e];
                  by = yn[cidx] + c1 * yn[IH] + c2 * yn[IH + 1] + c3 * yn[IH + 2];
                }
              } else {
                if (i >= 1) {
                  const double err = x[cidx] - fr[3 * e + ax];
                  gcv = 2.0 * C().w_foot * ge * err; dgv = 2.0 * C().w_foot * ge;
                  if (box) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                      const int r = R_BOX + 6 * e + 2 * ax + q;
                      const double ja = q ? -1.0 : 1.0, gval = ja * err - C().box[ax];
                      const double sv = s[r], lv = lam[r], inv = cmpc_rcp(sv), rg = gval - C().relax + sv, sig = lv * inv;
                      stat_row(st, rg, sv, lv); rec[Q_RG + r] = rg;
                      m1v += ja * inv; m2v += ja * sig * rg; glv += ja * lv; dgv += sig;
                    }
                  }
                }
                if (has_u) {
                  c1 = -d * ge * E.sum[3 * e + (ax + 2) % 3]; c2 = d * ge * E.sum[3 * e + (ax + 1) % 3];
                  by = yn[cidx] + c1 * yn[IH + (ax + 1) % 3] + c2 * yn[IH + (ax + 2) % 3];
                }
              }
              if (has_u) { double* col = rec + Q_BA + 4 * j; col[0] = 1.0; col[1] = c1; col[2] = c2; col[3] = c3; }
              const double r = x_free ? gcv + glv + by - yi[cidx] : 0.0;
              stat_dual(st, r);
#ifdef CMPC_TRACE
              cmpc_dbg_rd[j] = r;
#endif
              st[4] += fabs(yi[cidx]);
              rec[Q_GC + j] = gcv; rec[Q_DIAG + j] = dgv; rec[Q_M1 + j] = m1v; rec[Q_M2 + j] = m2v;
              if (has_u) {
                // input driving this state: yaw rate u[30 + e], foot velocity u[24 + 3 e + ax]
                const int ui = (jj == 0) ? 30 + e : 24 + 3 * e + ax;
                const double bu = d * (1.0 - ge), uv = u[ui];
                const double dj = x[cidx] + bu * uv - xn[cidx];
                rec[Q_D + cidx] = dj;
                const double ad = fabs(dj); st[0] = ad > st[0] ? ad : st[0];
                double* col = rec + Q_BA + 4 * ui;
                col[0] = bu; col[1] = 0.0; col[2] = 0.0; col[3] = 0.0;
                const double gu = 2.0 * C().eps_reg * uv;
                rec[Q_GC + ui] = gu; rec[Q_DIAG + ui] = 2.0 * C().eps_reg; rec[Q_M1 + ui] = 0.0; rec[Q_M2 + ui] = 0.0;
                const double ru = gu + bu * yn[cidx];
                stat_dual(st, ru);
#ifdef CMPC_TRACE
                cmpc_dbg_rd[ui] = ru;
#endif
              }
            }
          }
        }
      }
      par.sync();
      // ---- P5: roles -> stage
      for (int il = tid; il < ns; il += nt) {
        double prim = 0, dual = 0, smax = 0, smin = 1e300, msum = 0;
        for (int r = 0; r < 11; ++r) {
          const double* q = es[il].st[r];
          prim = q[0] > prim ? q[0] : prim; dual = q[1] > dual ? q[1] : dual; smax = q[2] > smax ? q[2] : smax; smin = q[3] < smin ? q[3] : smin;
          msum += q[4];
        }
        double* a = sm.acc[i0 + il];
        a[0] = prim; a[1] = dual; a[2] = smax; a[3] = smin; a[4] = msum;
      }
      par.sync();
    }
    reduce_acc(out);
  }

  // reduction over the <= 65 stages by warp 0: lanes stride the stages, shuffle tree across the lanes, a warp vote carries
  // NaNs (a NaN never wins a comparison), lane 0 publishes
  CMPC_HD void reduce_acc(double* out) {
    Smem& sm = par.template smem<Smem>();
    const int N = C().N;
    if (par.warp() == 0) {
      double prim = 0, dual = 0, smax = 0, smin = 1e300, ssum = 0;
      bool bad = false;
      for (int i = par.lane(); i <= N; i += par.lanes()) {
        const double* a = sm.acc[i];
        bad = bad || !(a[0] == a[0]) || !(a[1] == a[1]);
        prim = a[0] > prim ? a[0] : prim; dual = a[1] > dual ? a[1] : dual;
        smax = a[2] > smax ? a[2] : smax; smin = a[3] < smin ? a[3] : smin;
        ssum += a[4];
      }
      prim = par.wmax(prim); dual = par.wmax(dual); smax = par.wmax(smax); smin = par.wmin(smin); ssum = par.wsum(ssum);
      if (par.any(bad)) prim = nan("");
      if (par.lane() == 0) { sm.red5[0] = prim; sm.red5[1] = dual; sm.red5[2] = smax; sm.red5[3] = smin; sm.red5[4] = ssum; }
    }
    par.sync();
    out[0] = sm.red5[0]; out[1] = sm.red5[1]; out[2] = sm.red5[2]; out[3] = sm.red5[3]; out[4] = sm.red5[4]; out[5] = 0.0; out[6] = 0.0; out[7] = 0.0;
    // (no barrier here: red5 / acc are next written behind at least one barrier of the following pass)
  }

  CMPC_HD int n_rows_total() const {
    Smem& sm = par.template smem<Smem>();
    int n = 0;
    for (int i = 0; i <= C().N; ++i) n += cmpc_popcount(sm.mask[i]);
    return n;
  }

  // scaled optimality error of the barrier problem (IPOPT eq. 5/6)
  CMPC_HD double kkt_error(const double* ev, double mu_t, int nrows, double* parts) const {
    const double smax_ = 100.0;
    const double nmult = (double)(nrows + (C().N + 1) * NX);
    double sd = ev[4] / nmult; sd = (sd > smax_ ? sd : smax_) / smax_;
    const double a = fabs(ev[2] - mu_t), b = fabs(ev[3] - mu_t);
    const double compl_ = (nrows > 0) ? (a > b ? a : b) / sd : 0.0;
    const double dual = ev[1] / sd;
    parts[0] = dual; parts[1] = ev[0]; parts[2] = compl_;
    double e = dual > ev[0] ? dual : ev[0];
    return e > compl_ ? e : compl_;
  }

  // M index of stage variable j of the 60-ordering z = [u ; x]  (block ordering is [u ; w ; x])
  CMPC_HD static int mz(int j) { return j < NU ? j : j + NW; }

  // ---- stage KKT block of stage i without the cost-to-go term: cost + barrier + Lagrangian curvature.
  // The Lyapunov row and the angular-momentum row touch many variables and carry multipliers of 1e3..1e6;
  // condensing them (sigma g g', sigma = lam/s up to 1e18) would destroy the Schur complements by
  // cancellation, so their new multipliers w stay explicit unknowns of the (quasi-definite) stage block:
  //   [ H   g ] [dz]   [ -grad          ]
  //   [ g' -1/sigma ] [w ] = [ -(r_g + mu/lam) ]
  // R: the stage's derivative record staged in shared memory (R[Q_xxx]).  Five barriers: the phases between them write
  // disjoint sets of entries.  The products W = P [B A] and tv = p + P d of the cost-to-go term are formed alongside
  // the first phase (they do not touch M).
  CMPC_HD void assemble_stage(int i, double reg, const double* R) {
    Smem& sm = par.template smem<Smem>();
    const int tid = par.tid(), nt = par.nt();
    const int lane = par.lane(), wid = par.warp(), nw = par.nwarps(), nl = par.lanes();
    const double* bav = R + Q_BA;
    for (int t = tid; t < MSZ; t += nt) sm.M[t] = 0.0;
    // W = P [B A]  (28 x 60): warp per row r, lanes over columns j;  tv = p + P d
    // (a lane keeps the row pattern and the values of its column over the rows it visits; empty slots read P(r, 0) times 0)
    for (int j = lane; j < NZ; j += nl) {
      int rr[4]; double bv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { bv[q] = bav[4 * j + q]; rr[q] = sm.barz[4 * j + q]; }       // (empty slots hold 0 in the record)
      // (rows in groups: every load of a group is issued before its first store -- the compiler cannot move a shared-memory
      // load above a store that might alias it, and one row at a time is one exposed load latency per row)
      constexpr int GW = CMPC_PIPE_W;
      for (int r0 = wid; r0 < NX; r0 += GW * nw) {
        double pq[GW][4];
        // (branch-free loads from clamped rows, only the store is predicated: conditionally defined array elements would send the
        // arrays to local memory)
#pragma unroll
        for (int g = 0; g < GW; ++g) {
          const int r = r0 + g * nw;
          const double* Pr = sm.P + (r < NX ? r : NX - 1) * NX;
#pragma unroll
          for (int q = 0; q < 4; ++q) pq[g][q] = Pr[rr[q]];
        }
#pragma unroll
        for (int g = 0; g < GW; ++g) {
          const int r = r0 + g * nw;
          const double v = (pq[g][0] * bv[0] + pq[g][1] * bv[1]) + (pq[g][2] * bv[2] + pq[g][3] * bv[3]);
          if (r < NX) sm.W[r * NZ + j] = v;
        }
      }
    }
    for (int r = tid; r < NX; r += nt) {
      double s = sm.pv[r];
      for (int j = 0; j < NX; ++j) s += sm.P[r * NX + j] * R[Q_D + j];
      sm.tv[r] = s;
    }
    par.sync();
    const bool has_hw = (i == 0) && (sm.mask[0] & (1ull << R_HW));
    // ---- one phase of entries that are each written once (the block is zero): gradient row and diagonal (60 items),
    // the two multiplier diagonals, friction off-diagonals (xz, yz within a vertex; the diagonal part is in Q_DIAG), symmetry-term
    // off-diagonals (-2 w_sym / 4 between same-axis components of two vertices of one foot), rate cross terms (q_v, f_z),
    // and the bilinear torque term (f_ek, p), (f_ek, p_e), (f_ek, psi_e) (rows x, cols u; cross-axis pairs only).
    // Item ranges start at multiples of 32 where it matters: a warp stays on one code path.
    {
      const double y0 = R[Q_YH], y1 = R[Q_YH + 1], y2 = R[Q_YH + 2];      // delta * y_h
      constexpr int I0 = 64, I1 = I0 + 16, I2 = I1 + 36 + 12, I3 = I2 + 8 + 24, I4 = I3 + 48, I5 = I4 + 48, I6 = I5 + 24;   // 64 | 80 | 128 | 160 | 208 | 256 | 280
      for (int t = tid; t < I6; t += nt) {
        if (t < I0) {
          if (t < NZ) {
            sm.M[mi(GR, mz(t))] = R[Q_GC + t] + mu * R[Q_M1 + t] + R[Q_M2 + t];
            sm.M[mi(mz(t), mz(t))] = R[Q_DIAG + t] + reg;
          } else if (t == NZ) {
            sm.M[mi(NU, NU)] = -1.0 / R[Q_LSIG];
            sm.M[mi(GR, NU)] = R[Q_LRG] + mu / R[Q_LLAM];
            sm.M[mi(NU + 1, NU + 1)] = has_hw ? -1.0 / R[Q_HSIG] : -1.0;
            sm.M[mi(GR, NU + 1)] = has_hw ? R[Q_HRG] + mu / R[Q_HLAM] : 0.0;
          }
        } else if (t < I1) {
          const int q = t - I0, v = q >> 1, yz = q & 1;                     // (z, x) -> record slot 2, (z, y) -> slot 4
          sm.M[mi(3 * v + 2, 3 * v + yz)] = R[Q_FRIC + 6 * v + (yz ? 4 : 2)];
        } else if (t < I2) {
          const int q = t - I1;
          if (q < 36) {
            const int e = q / 18, ax = (q % 18) / 6, pr = q % 6;
            const int ka = (pr == 0) ? 1 : (pr < 3 ? 2 : 3), kb = (pr == 0 || pr == 1 || pr == 3) ? 0 : ((pr == 2 || pr == 4) ? 1 : 2);
            sm.M[mi(12 * e + 3 * ka + ax, 12 * e + 3 * kb + ax)] = -0.5 * C().w_sym * R[Q_GAM + e];
          }
        } else if (t < I3) {
          const int v = t - I2;
          if (v < 8) sm.M[mi(XO + IQ + v, 3 * v + 2)] = -2.0 * R[Q_GAMP + v / 4];
        } else if (t < I5) {
          // (f, p) and (f, p_e) cross-axis pairs: 6 per vertex each
          const int tt = (t < I4) ? t - I3 : t - I4, v = tt / 6, pq = tt % 6, e = v / 4;
          const int a_ = pq >> 1, b_ = (a_ + 1 + (pq & 1)) % 3;              // b != a
          // Yx = [[0, -y2, y1], [y2, 0, -y0], [-y1, y0, 0]]
          const double yv = (a_ == 0) ? (b_ == 1 ? -y2 : y1) : (a_ == 1 ? (b_ == 0 ? y2 : -y0) : (b_ == 0 ? -y1 : y0));
          const double val = R[Q_GAM + e] * yv;
          if (t < I4) sm.M[mi(XO + IP + b_, 3 * v + a_)] = -val;
          else sm.M[mi(XO + (e ? IPR : IPL) + b_, 3 * v + a_)] = val;
        } else {
          const int tt = t - I5, v = tt / 3, a_ = tt % 3, e = v / 4;
          const double dx_ = R[Q_DR + 2 * v], dy_ = R[Q_DR + 2 * v + 1];
          const double cr = (a_ == 0) ? -y2 * dy_ : (a_ == 1 ? y2 * dx_ : y0 * dy_ - y1 * dx_);   // y x (R'c), R'c = (dx_, dy_, 0)
          sm.M[mi(XO + (e ? IPSR : IPSL), 3 * v + a_)] = R[Q_GAM + e] * cr;
        }
      }
    }
    par.sync();
    // Lyapunov row: its gradient as row / column NU and the curvature lam * C (x) I_3 over the same-axis pairs of the 33
    // touched variables, scattered through the table built once per solve
    {
      const double gam3[3] = {R[Q_GAM], R[Q_GAM + 1], 1.0};
      for (int e = tid; e < NLY; e += nt) {
        const int g = sm.ly_g[e];
        const int ga = g & 15, gb = g >> 4;
        const double va = ga == 0 ? gam3[0] : (ga == 1 ? gam3[1] : 1.0), vb = gb == 0 ? gam3[0] : (gb == 1 ? gam3[1] : 1.0);
        sm.M[sm.ly_m[e]] += va * vb * R[sm.ly_s[e]];
      }
    }
    par.sync();
    // angular-momentum row (stage 0): curvature 2 lam Bh'Bh in the force block, gradient as row NU+1
    if (has_hw) {
      const double lamh = R[Q_HLAM];
      for (int t = tid; t < 24 * 24 + 24; t += nt) {
        if (t >= 24 * 24) {
          const int a_ = t - 24 * 24;
          double ca[3] = {0, 0, 0};
          ca[(a_ % 3 + 1) % 3] = bav[4 * a_ + 1]; ca[(a_ % 3 + 2) % 3] = bav[4 * a_ + 2];
          sm.M[mi(NU + 1, a_)] = 2.0 * (ca[0] * R[Q_HP] + ca[1] * R[Q_HP + 1] + ca[2] * R[Q_HP + 2]);
          continue;
        }
        const int a_ = t / 24, b_ = t % 24;
        if (b_ > a_) continue;
        double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
        ca[(a_ % 3 + 1) % 3] = bav[4 * a_ + 1]; ca[(a_ % 3 + 2) % 3] = bav[4 * a_ + 2];
        cb[(b_ % 3 + 1) % 3] = bav[4 * b_ + 1]; cb[(b_ % 3 + 2) % 3] = bav[4 * b_ + 2];
        sm.M[mi(a_, b_)] += 2.0 * lamh * (ca[0] * cb[0] + ca[1] * cb[1] + ca[2] * cb[2]);
      }
      par.sync();
    }
  }

  // asynchronous copy of stage i's derivative record (without the row residuals) into a staging buffer
  CMPC_HD void record_in(int i, double* buf) {
    par.copy_async(buf, gREC() + (size_t)i * RECSZ, Q_RG);
    par.commit_async();
  }

  // ---- backward Riccati sweep.  Returns false if a pivot has the wrong sign (inputs > 0, multipliers < 0).
  // Thread mapping: one warp per matrix row (rows dealt cyclically to the warps), lanes across the columns of the
  // row: shared-memory accesses are conflict-free (row stride 63 doubles) and whole rows whose multiplier is zero
  // -- the stage block is sparse: stance-foot inputs, swing-foot forces, previous-f_z states -- are skipped
  // without divergence.
  CMPC_HD bool backward(double reg) {
    Smem& sm = par.template smem<Smem>();
    const int N = C().N, tid = par.tid(), nt = par.nt();
    const int lane = par.lane(), wid = par.warp(), nw = par.nwarps(), nl = par.lanes();
    par.wait_async();                                              // (a sweep abandoned on a bad pivot may have left a record copy in flight)
    {   // terminal stage: P_N diagonal, p_N = modified gradient (x part)
      const double* rec = gREC() + (size_t)N * RECSZ;
      for (int t = tid; t < NX * NX; t += nt) sm.P[t] = 0.0;
      par.sync();
      for (int t = tid; t < NX; t += nt) {
        sm.P[t * NX + t] = rec[Q_DIAG + 32 + t] + reg;
        sm.pv[t] = rec[Q_GC + 32 + t] + mu * rec[Q_M1 + 32 + t] + rec[Q_M2 + 32 + t];
      }
      par.sync();
    }
    record_in(N - 1, sm.recb[(N - 1) & 1]);
    for (int i = N - 1; i >= 0; --i) {
      double* fac = gFAC() + (size_t)i * FACSZ;
      CMPC_TIC(sm);
      par.wait_async();
      par.sync();                                                  // record i is in; every thread is done with stage i + 1
      if (i > 0) record_in(i - 1, sm.recb[(i - 1) & 1]);           // lands while stage i is assembled and factorised
      const double* R = sm.recb[i & 1];
      const double* bav = R + Q_BA;
      assemble_stage(i, reg, R);
      CMPC_TOC(sm, PF_ASM);
      // M += [B A]' W (lower triangle), gradient row += [B A]' tv : warp per row a, lanes over b <= a
      // (rows in groups of CMPC_PIPE_M per warp: the loads of a group -- coefficients, W entries, the M entries to be updated --
      // are all issued before its first store; one row at a time is a chain of four dependent shared-memory latencies per row,
      // sixteen rows per warp)
      {
        constexpr int GM = CMPC_PIPE_M;
        for (int a0 = wid; a0 < NZ; a0 += GM * nw) {
          // (branch-free loads from clamped rows, only the store is predicated: conditionally defined array elements would send the
          // arrays to local memory)
          int rr[GM][4], mrow[GM]; double bv[GM][4]; bool on[GM];
#pragma unroll
          for (int g = 0; g < GM; ++g) {
            const int a_ = a0 + g * nw, ac = a_ < NZ ? a_ : NZ - 1;
            on[g] = a_ < NZ && sm.barow[4 * ac] >= 0;                      // (off: structurally empty column -- previous-f_z states)
            mrow[g] = mi(mz(ac), 0);
#pragma unroll
            for (int q = 0; q < 4; ++q) { bv[g][q] = bav[4 * ac + q]; rr[g][q] = sm.baofs[4 * ac + q]; }
          }
          const int amax = (a0 + (GM - 1) * nw < NZ) ? a0 + (GM - 1) * nw : NZ - 1;
          for (int b_ = lane; b_ <= amax; b_ += nl) {
            double wv[GM][4], mv[GM];
            const int mb = mz(b_);
#pragma unroll
            for (int g = 0; g < GM; ++g) {
#pragma unroll
              for (int q = 0; q < 4; ++q) wv[g][q] = sm.W[rr[g][q] + b_];
              mv[g] = sm.M[mrow[g] + mb];
            }
#pragma unroll
            for (int g = 0; g < GM; ++g) {
              const double s = (bv[g][0] * wv[g][0] + bv[g][1] * wv[g][1]) + (bv[g][2] * wv[g][2] + bv[g][3] * wv[g][3]);
              if (on[g] && b_ <= a0 + g * nw) sm.M[mrow[g] + mb] = mv[g] + s;
            }
          }
        }
        if (wid == NZ % nw) {                                      // gradient row
          for (int b_ = lane; b_ < NZ; b_ += nl) {
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int r2 = sm.barow[4 * b_ + q]; if (r2 >= 0) s += bav[4 * b_ + q] * sm.tv[r2]; }
            sm.M[mi(GR, mz(b_))] += s;
          }
        }
      }
      par.sync();
      CMPC_TOC(sm, PF_PBA);
      // ---- partial LDL' of the [u ; w] block, right-looking, matrix held in REGISTER tiles: the 64 x 64 (padded)
      // lower triangle is cut into 4 x 4 tiles.  The 120 tiles right of tile column 0 live one per thread for the
      // whole factorisation; the 16 tiles of tile column 0 are final after the first block and are held only until
      // then (by the last 16 threads).  The 32 input columns are eliminated in BLOCKS OF FOUR (one tile column):
      //   1. the owner of the diagonal tile factors it (4 x 4, four rsqrt) and publishes L_d and the reciprocal
      //      diagonal                                                                         -- barrier A
      //   2. the owners of the tiles below it form their rows of L (4 x 4 forward substitution, registers) and
      //      publish the 64 x 4 panel                                                         -- barrier B
      //   3. every tile to the right applies the rank-4 update (64 FMAs fed by 32 shared loads).
      // The panel is stored column by column (four values per column and tile row): the update is four rank-1 steps, each
      // streaming one column of L for the tile's rows and columns (128-bit loads), eight operand registers next to the tile.
      // (the panel and the pivot-column buffer live in the W storage, idle between the P [B A] products and the gains.)
      // The two explicit multipliers (columns 32, 33; pivot sign -1) follow column by column.  The gradient row
      // (row 62) is carried along.  Inputs without coupling at a stage (stance-foot velocities, swing-foot
      // tangential forces) need no special case: their columns are zero below the diagonal.
      // Blocks of four inputs that are DECOUPLED at this stage -- the forces of a swing foot (gamma = 0: no dynamics, no
      // rows, cost 10 |f|^2 only, no force-rate term), the velocity / yaw-rate inputs in double support ((1 - gamma) = 0) -- have a diagonal
      // block and nothing below it but the gradient row: their elimination step is the reciprocal root of the diagonal.
      // Such a block takes no barrier, no panel and no update; its gradient-row entries are scaled when the tiles go back
      // to shared memory.  Forces of the left / right foot are blocks 0-2 / 3-5, blocks 6-7 hold the foot inputs.
      unsigned skipb = 0;
#if CMPC_SKIP_BLOCKS
      {
        // (the force-rate term couples f_z of stage i to the state q with the PREVIOUS stage's gamma: the first swing stage
        // after a lift-off is not decoupled)
        const bool sl_ = R[Q_GAM] < 0.5 && R[Q_GAMP] == 0.0, sr_ = R[Q_GAM + 1] < 0.5 && R[Q_GAMP + 1] == 0.0;
        const bool ds_ = R[Q_GAM] > 0.5 && R[Q_GAM + 1] > 0.5;
        skipb = (sl_ ? 0x07u : 0u) | (sr_ ? 0x38u : 0u) | (ds_ ? 0xC0u : 0u);
      }
#endif
      {
        double T[Par::TPT][16];
        int ti_[Par::TPT], tj_[Par::TPT];
#pragma unroll
        for (int sl = 0; sl < Par::TPT; ++sl) {
          ti_[sl] = tile_i[sl]; tj_[sl] = tile_j[sl];
          if (ti_[sl] >= 0) {
#pragma unroll
            for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
              for (int b_ = 0; b_ < 4; ++b_) {
                const int r = 4 * ti_[sl] + a_, cc = 4 * tj_[sl] + b_;
                T[sl][4 * a_ + b_] = (r < MROWS && cc <= r && cc < NZA) ? sm.M[mi(r, cc)] : 0.0;
              }
          }
        }
        bool okp = true;
        // factor a diagonal tile held in registers; publish {1/d0..1/d3, l10, l20, l21, l30, l31, l32} and the pivot test
        auto diag_tile = [&](double (&t)[16], int tk) -> bool {
          // two 2 x 2 steps: in each, rsqrt(a) and rsqrt(a c - b^2) are independent and branch-free, so they interleave and
          // the dependent chain holds two rsqrt latencies instead of four (second pivot of a step = (a c - b^2) / a, its
          // reciprocal root = a rsqrt(a) rsqrt(a c - b^2))
          const double p0 = t[0];
          const double det1 = p0 * t[5] - t[4] * t[4];
          const double i0 = cmpc_rsqrt_nb(p0), r1 = cmpc_rsqrt_nb(det1);
          const double p1 = det1 * (i0 * i0), i1 = r1 * (p0 * i0);
          const double l10 = t[4] * i0, l20 = t[8] * i0, l30 = t[12] * i0;
          const double l21 = (t[9] - l20 * l10) * i1, l31 = (t[13] - l30 * l10) * i1;
          const double p2 = t[10] - l20 * l20 - l21 * l21;
          const double b2 = t[14] - l30 * l20 - l31 * l21;
          const double c2 = t[15] - l30 * l30 - l31 * l31;
          const double det2 = p2 * c2 - b2 * b2;
          const double i2 = cmpc_rsqrt_nb(p2), r3 = cmpc_rsqrt_nb(det2);
          const double p3 = det2 * (i2 * i2), i3 = r3 * (p2 * i2);
          const double l32 = b2 * i2;
          // (upper bounds: the single-precision seed of cmpc_rsqrt_nb overflows beyond ~1e38 and would return 0 silently)
          const bool good = p0 > 1e-14 && p1 > 1e-14 && p2 > 1e-14 && p3 > 1e-14 && p0 < 1e30 && p2 < 1e30 && det1 < 1e36 && det2 < 1e36;
#ifdef CMPC_TRACE
          if (cmpc_trace_on && !good) printf("   pivot fail stage %d block %d piv %.3e %.3e %.3e %.3e reg %.1e\n", i, tk, p0, p1, p2, p3, reg);
#endif
          double* d = sm.dpub;
          d[0] = i0; d[1] = i1; d[2] = i2; d[3] = i3; d[4] = l10; d[5] = l20; d[6] = l21; d[7] = l30; d[8] = l31; d[9] = l32;
          sm.rdiag[4 * tk] = i0; sm.rdiag[4 * tk + 1] = i1; sm.rdiag[4 * tk + 2] = i2; sm.rdiag[4 * tk + 3] = i3;
          t[0] = p0 * i0; t[4] = l10; t[5] = p1 * i1; t[8] = l20; t[9] = l21; t[10] = p2 * i2;
          t[12] = l30; t[13] = l31; t[14] = l32; t[15] = p3 * i3;
          return good;
        };
        // rows of L of a tile below the diagonal tile (in place), published as tile row `ti` of the panel
        auto panel_tile = [&](double (&t)[16], int ti) {
          const double* d = sm.dpub;
          const double i0 = d[0], i1 = d[1], i2 = d[2], i3 = d[3], l10 = d[4], l20 = d[5], l21 = d[6], l30 = d[7], l31 = d[8], l32 = d[9];
          double* pl = (sm.W + 128) + PSTR * ti;
          // a tile that is exactly zero (rows without coupling to this block: the stage block is sparse) stays zero, is not
          // published, and every update that would read it is skipped (`pzf`)
          bool zero = true;
#pragma unroll
          for (int q = 0; q < 16; ++q) zero = zero && (t[q] == 0.0);
          sm.pzf[ti] = zero ? 1 : 0;
          if (zero) return;
#pragma unroll
          for (int a_ = 0; a_ < 4; ++a_) {
            const double v0 = t[4 * a_] * i0;
            const double v1 = (t[4 * a_ + 1] - v0 * l10) * i1;
            const double v2 = (t[4 * a_ + 2] - v0 * l20 - v1 * l21) * i2;
            const double v3 = (t[4 * a_ + 3] - v0 * l30 - v1 * l31 - v2 * l32) * i3;
            t[4 * a_] = v0; t[4 * a_ + 1] = v1; t[4 * a_ + 2] = v2; t[4 * a_ + 3] = v3;
          }
          // published column by column (pl[4 kk + a]): the update then streams one column of L per rank-1 step
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            reinterpret_cast<Pair*>(pl)[2 * kk] = Pair{t[kk], t[4 + kk]};
            reinterpret_cast<Pair*>(pl)[2 * kk + 1] = Pair{t[8 + kk], t[12 + kk]};
          }
        };
        // diagonal tile of a decoupled block: L = sqrt(diag), reciprocal published for the gains
        auto skip_diag = [&](double (&t)[16], int tk) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double dq = t[5 * q], iq = cmpc_rsqrt_nb(dq);
            sm.rdiag[4 * tk + q] = iq; t[5 * q] = dq * iq;
          }
        };
        {   // ---- block 0: tile column 0 (transient tiles)
          double C[Par::CPT][16];
          int ci_[Par::CPT];
          bool mygood = true;
          const int cbase = nt >= 16 ? nt - 16 : 0;
#pragma unroll
          for (int sl = 0; sl < Par::CPT; ++sl) {
            const int cidx = tid - cbase + sl * nt;
            ci_[sl] = (cidx >= 0 && cidx < 16) ? cidx : -1;
            if (ci_[sl] >= 0) {
#pragma unroll
              for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                for (int b_ = 0; b_ < 4; ++b_) {
                  const int r = 4 * ci_[sl] + a_;
                  C[sl][4 * a_ + b_] = (r < MROWS && b_ <= r) ? sm.M[mi(r, b_)] : 0.0;
                }
              if (ci_[sl] == 0) { if (skipb & 1u) skip_diag(C[sl], 0); else mygood = diag_tile(C[sl], 0); }
            }
          }
          if (!(skipb & 1u)) { if (!par.sync_and(mygood)) okp = false; }   // A: barrier + vote on the pivot test
          if (okp) {
#pragma unroll
            for (int sl = 0; sl < Par::CPT; ++sl) {
              if (ci_[sl] < 0) continue;
              if (ci_[sl] > 0 && !(skipb & 1u)) panel_tile(C[sl], ci_[sl]);
#pragma unroll
              for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                for (int b_ = 0; b_ < 4; ++b_) {
                  const int r = 4 * ci_[sl] + a_;
                  if (r < MROWS && b_ <= r) sm.M[mi(r, b_)] = C[sl][4 * a_ + b_];       // final: column 0..3 of L
                }
            }
          }
        }
        for (int tk = 0; tk < NU / 4 && okp; ++tk) {
          if ((skipb >> tk) & 1u) {                                 // decoupled block: diagonal only (uniform branch)
            if (tk > 0) {
#pragma unroll
              for (int sl = 0; sl < Par::TPT; ++sl)
                if (ti_[sl] == tk && tj_[sl] == tk) skip_diag(T[sl], tk);
            }
            continue;
          }
          if (tk > 0) {
            bool mygood = true;
#pragma unroll
            for (int sl = 0; sl < Par::TPT; ++sl)
              if (ti_[sl] == tk && tj_[sl] == tk) mygood = diag_tile(T[sl], tk);
            if (!par.sync_and(mygood)) { okp = false; break; }      // A: barrier + vote (uniform result)
#pragma unroll
            for (int sl = 0; sl < Par::TPT; ++sl)
              if (tj_[sl] == tk && ti_[sl] > tk) panel_tile(T[sl], ti_[sl]);
          }
          par.sync();                                               // B
#pragma unroll
          for (int sl = 0; sl < Par::TPT; ++sl) {
            const int ti = ti_[sl], tj = tj_[sl];
            if (tj <= tk) continue;                                  // (no tile: tj = 0) tiles left of / in the block: final
            if (sm.pzf[ti] | sm.pzf[tj]) continue;                   // a zero panel tile on either side: nothing to subtract
            const Pair* pr = reinterpret_cast<const Pair*>((sm.W + 128) + PSTR * ti);
            const Pair* pc = reinterpret_cast<const Pair*>((sm.W + 128) + PSTR * tj);
            // four rank-1 steps, one column of the panel each: eight operand registers live next to the tile
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const Pair r01 = pr[2 * kk], r23 = pr[2 * kk + 1], c01 = pc[2 * kk], c23 = pc[2 * kk + 1];
              const double lr[4] = {r01.x, r01.y, r23.x, r23.y}, lc[4] = {c01.x, c01.y, c23.x, c23.y};
              if (ti == tj) {
#pragma unroll
                for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                  for (int b_ = 0; b_ <= a_; ++b_) T[sl][4 * a_ + b_] -= lr[a_] * lc[b_];
              } else {
#pragma unroll
                for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                  for (int b_ = 0; b_ < 4; ++b_) T[sl][4 * a_ + b_] -= lr[a_] * lc[b_];
              }
            }
          }
        }
        if (!okp) return false;
        int nproc = 0;
        {
          const int tk = NU / 4;
#pragma unroll
          for (int kk = 0; kk < NW; ++kk) {                         // kk static: tiles stay in registers
            const int k = 4 * tk + kk;
            if (!okp) break;
            if (kk == 1 && !((i == 0) && (sm.mask[0] & (1ull << R_HW)))) {
              // the angular-momentum row exists at stage 0 only: elsewhere its multiplier is a decoupled dummy (diagonal -1)
              if (tid == 0) sm.rdiag[k] = -1.0;
              break;
            }
            double* cb = sm.W + (nproc & 1) * 64;
            ++nproc;
#pragma unroll
            for (int sl = 0; sl < Par::TPT; ++sl)
              if (ti_[sl] >= 0 && tj_[sl] == tk) {
#pragma unroll
                for (int a_ = 0; a_ < 4; ++a_) cb[4 * ti_[sl] + a_] = T[sl][4 * a_ + kk];
              }
            par.sync();
            const double piv = cb[k];
            const double sgn = -1.0;
#ifdef CMPC_TRACE
            if (cmpc_trace_on && !(sgn * piv > 0.0)) printf("   pivot fail stage %d k %d piv %.3e reg %.1e\n", i, k, piv, reg);
#endif
            // an explicit multiplier has pivot -1/sigma - g'M^-1 g < 0, as small as 1/sigma
            if (!(sgn * piv > 0.0)) { okp = false; break; }    // uniform: same value for every thread
            const double inv = cmpc_rsqrt_nb(sgn * piv);
            if (tid == 0) sm.rdiag[k] = sgn * inv;                 // reciprocal of the stored (signed) diagonal
#pragma unroll
            for (int sl = 0; sl < Par::TPT; ++sl) {
              const int ti = ti_[sl], tj = tj_[sl];
              if (ti < tk || tj < tk) continue;                       // (ti < 0 included) tile above / left of the pivot: final
              double lr[4], lc[4];
#pragma unroll
              for (int a_ = 0; a_ < 4; ++a_) { lr[a_] = cb[4 * ti + a_] * inv; lc[a_] = cb[4 * tj + a_] * inv; }
              if (tj == tk) {
                // tile holds column k itself: entries left of / on column k need the index tests
#pragma unroll
                for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                  for (int b_ = 0; b_ < 4; ++b_) {
                    const int r = 4 * ti + a_;
                    if (b_ > kk) { if (r > k && 4 * tj + b_ <= r) T[sl][4 * a_ + b_] -= sgn * lr[a_] * lc[b_]; }
                    else if (b_ == kk && r >= k) T[sl][4 * a_ + b_] = lr[a_];     // column k of L (diagonal: sqrt|piv|)
                  }
              } else if (ti == tj) {
#pragma unroll
                for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                  for (int b_ = 0; b_ <= a_; ++b_) T[sl][4 * a_ + b_] -= sgn * lr[a_] * lc[b_];
              } else {
#pragma unroll
                for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
                  for (int b_ = 0; b_ < 4; ++b_) T[sl][4 * a_ + b_] -= sgn * lr[a_] * lc[b_];
              }
            }
          }
        }
        if (!okp) return false;
        par.sync();
#pragma unroll
        for (int sl = 0; sl < Par::TPT; ++sl) {
          if (ti_[sl] < 0) continue;
          // gradient-row entries (row 62 = tile row 15, a_ = 2) of a decoupled block were left unscaled: L(62, c) = g_c / l_cc
          if (ti_[sl] == 15 && tj_[sl] < NU / 4 && ((skipb >> tj_[sl]) & 1u)) {
#pragma unroll
            for (int b_ = 0; b_ < 4; ++b_) T[sl][8 + b_] *= sm.rdiag[4 * tj_[sl] + b_];
          }
#pragma unroll
          for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
            for (int b_ = 0; b_ < 4; ++b_) {
              const int r = 4 * ti_[sl] + a_, cc = 4 * tj_[sl] + b_;
              if (r < MROWS && cc <= r && cc < NZA) sm.M[mi(r, cc)] = T[sl][4 * a_ + b_];
            }
        }
        if (skipb & 1u) for (int q = tid; q < 4; q += nt) sm.M[mi(GR, q)] *= sm.rdiag[q];   // same for block 0 (its tiles went back to shared memory earlier)
        par.sync();
      }
      CMPC_TOC(sm, PF_CHOL);
      // ---- gains: K = -L^-T L_S', k = -L^-T l_m (one right-hand side per thread, registers, L broadcast from
      // shared memory), staged in W.  The 29 right-hand sides occupy warp 0 only; the other warps copy the cost-to-go
      // P, p (final since the factorisation) out of the stage block meanwhile -- to shared memory for the next stage and
      // to the scratch for the costates of the forward sweep.  Then K, k are streamed out by everybody.
      if (Par::GAINS4) {
        // four lanes per right-hand side (8 right-hand sides per warp, all four warps): lane `sub` of a group holds the entries
        // j = 4 jj + sub of its right-hand side; step k: the lane that owns entry k scales it and hands it to the group by a
        // width-4 shuffle, every lane then updates its entries left of k.  A quarter of the FMAs per lane, on four warps.
        const int sub = lane & 3, t = wid * 8 + (lane >> 2);
        const bool live = t < NX + 1;
        const double* rowp = sm.M + mi(live ? (t < NX ? XO + t : GR) : GR, 0);
        double vl[(NA + 3) / 4];
#pragma unroll
        for (int jj = 0; jj < (NA + 3) / 4; ++jj) vl[jj] = (4 * jj + sub < NA) ? -rowp[4 * jj + sub] : 0.0;
#pragma unroll
        for (int k = NA - 1; k >= 0; --k) {
          const int jk = k >> 2, own = k & 3;
          const double zk = par.shfl4(vl[jk] * sm.rdiag[k], own);
          if (sub == own) vl[jk] = zk;
          const double* Lk = sm.M + mi(k, 0) + sub;
#pragma unroll
          for (int jj = 0; jj < jk; ++jj) vl[jj] -= Lk[4 * jj] * zk;
          if (sub < own) vl[jk] -= Lk[4 * jk] * zk;
        }
        if (live) {
#pragma unroll
          for (int jj = 0; jj < (NA + 3) / 4; ++jj) if (4 * jj + sub < NA) sm.W[t * NA + 4 * jj + sub] = vl[jj];      // K staged as 29 x 34 in W
        }
        for (int r = wid; r < NX; r += nw)
          for (int cc = lane; cc < NX; cc += nl) {
            const double v = (cc <= r) ? sm.M[mi(XO + r, XO + cc)] : sm.M[mi(XO + cc, XO + r)];
            sm.P[r * NX + cc] = v;
            fac[F_P + r * NX + cc] = v;
          }
        for (int q = tid; q < NX; q += nt) { const double v = sm.M[mi(GR, XO + q)]; sm.pv[q] = v; fac[F_PV + q] = v; }
      } else {
      if (nw == 1 || wid == 0) {
        for (int t = tid; t < NX + 1; t += nt) {
          double* rowp = sm.M + mi(t < NX ? XO + t : GR, 0);
          double v[NA];
#pragma unroll
          for (int q = 0; q < NA; ++q) v[q] = -rowp[q];
#pragma unroll
          for (int k = NA - 1; k >= 0; --k) {
            const double zk = v[k] * sm.rdiag[k];
            v[k] = zk;
            // (chunks of eight with a scheduling fence: the row of L must not be loaded whole ahead of time, v[] needs the registers)
#pragma unroll
            for (int j0 = 0; j0 < k; j0 += 8) {
#pragma unroll
              for (int j = j0; j < j0 + 8 && j < k; ++j) v[j] -= sm.M[mi(k, j)] * zk;
              CMPC_SCHED_FENCE();
            }
          }
#pragma unroll
          for (int q = 0; q < NA; ++q) sm.W[t * NA + q] = v[q];          // W (28 x 60) is free here: K staged as 29 x 34
        }
      }
      if (nw == 1 || wid > 0) {
        const int w0 = nw == 1 ? 0 : wid - 1, ws = nw == 1 ? 1 : nw - 1;
        for (int r = w0; r < NX; r += ws)
          for (int cc = lane; cc < NX; cc += nl) {
            const double v = (cc <= r) ? sm.M[mi(XO + r, XO + cc)] : sm.M[mi(XO + cc, XO + r)];
            sm.P[r * NX + cc] = v;
            fac[F_P + r * NX + cc] = v;
          }
        if (w0 == 0) for (int t = lane; t < NX; t += nl) { const double v = sm.M[mi(GR, XO + t)]; sm.pv[t] = v; fac[F_PV + t] = v; }
      }
      }
      par.sync();
      for (int t = tid; t < (NX + 1) * NA; t += nt) fac[F_K + t] = sm.W[t];
      // (no barrier here: the next stage begins with one before anything of W / P is written again)
      CMPC_TOC(sm, PF_STORE);
    }
    return true;
  }

  // ---- forward sweep: Newton step z_i = [du ; w] = k_i + K_i dx_i, dx_{i+1} = d_i + A dx_i + B du_i (sequential: two
  // barriers per stage, the next stage's K | k | d | [B A] values arrive by asynchronous copy while the current stage
  // is computed), then the full-step costates y_i = p_i + P_i dx_i of all stages at once.
  CMPC_HD void stage_in(int i, double* kb, double* bb) {
    const double* fac = gFAC() + (size_t)i * FACSZ;
    const double* rec = gREC() + (size_t)i * RECSZ;
    par.copy_async(kb, fac + F_K, KSZ);
    par.copy_async(kb + KSZ, rec + Q_D, NX);
    par.copy_async(bb, rec + Q_BA, NZ * 4);
    par.commit_async();
  }

  CMPC_HD void forward(double reg) {
    Smem& sm = par.template smem<Smem>();
    const int N = C().N, tid = par.tid(), nt = par.nt();
    double* dxall = sm.M;                                        // dx of all stages, (N + 1) x NX: the stage block is idle here
    double* kbuf[2] = {sm.W, sm.W + FWDBUF};                     // W | P storage (contiguous, idle here)
    double* bbuf[2] = {sm.recb[0], sm.W + 2 * FWDBUF};
    static_assert(2 * FWDBUF + NZ * 4 <= NX * NZ + NX * NX, "forward staging must fit in W | P");
    static_assert((NMAX + 1) * NX <= MSZ, "dx of all stages must fit in the stage block storage");
    for (int t = tid; t < NX; t += nt) { dxall[t] = 0.0; gDX()[t] = 0.0; }
    // the factors of the late stages were written first and have left L2 by now (the scratch of all resident CTAs exceeds it):
    // ask for all of them at once, so the DRAM latency is paid once and not once per stage of the sequential sweep
    par.prefetch_l2(gFAC(), N * FACSZ);
    stage_in(0, kbuf[0], bbuf[0]);
    par.wait_async();
    par.sync();
    for (int i = 0; i < N; ++i) {
      const int cur = i & 1;
      if (i + 1 < N) stage_in(i + 1, kbuf[cur ^ 1], bbuf[cur ^ 1]);
      const double* K = kbuf[cur];
      const double* bav = bbuf[cur];
      const double* dx = dxall + i * NX;
      for (int cidx = tid; cidx < NA; cidx += nt) {
        double s0 = K[NX * NA + cidx], s1 = 0.0, s2 = 0.0, s3 = 0.0;      // four accumulators: a dependent chain of 7 FMAs instead of 14
#pragma unroll
        for (int r = 0; r < NX; r += 4) {
          s0 += K[r * NA + cidx] * dx[r]; s1 += K[(r + 1) * NA + cidx] * dx[r + 1];
          s2 += K[(r + 2) * NA + cidx] * dx[r + 2]; s3 += K[(r + 3) * NA + cidx] * dx[r + 3];
        }
        const double s = (s0 + s1) + (s2 + s3);
        sm.zs[cidx] = s;
        if (cidx < NU) gDU()[i * NU + cidx] = s; else gDW()[i * NW + cidx - NU] = s;
      }
      par.sync();
      // dx_{i+1} = d + A dx + B du   (row gather over the structural pattern)
      // (the angular-momentum rows have ~25 entries, a chain of dependent index -> value loads each: four lanes per row,
      // two entries in flight per lane, summed over the group by shuffles)
      {
        constexpr int G = Par::GAINS4 ? 4 : 1;
        // (every lane of a warp takes part in the shuffles: the loop bound is rounded up to whole warps, idle lanes carry an empty row)
        const int nl_ = par.lanes(), qend = (G == 4) ? ((NX * G + nl_ - 1) / nl_) * nl_ : NX;
        for (int q = tid; q < qend; q += nt) {
          const int r = q / G, sub = q - r * G;
          const bool live = r < NX;
          const int e1 = live ? sm.csr_ptr[r + 1] : 0;
          double s = (live && sub == 0) ? K[KSZ + r] : 0.0, s2 = 0.0;
          int e = live ? sm.csr_ptr[r] + sub : 0;
          for (; e + G < e1; e += 2 * G) {
            const int jq = sm.csr_idx[e], j = jq >> 2, jq2 = sm.csr_idx[e + G], j2 = jq2 >> 2;
            s += bav[jq] * (j < NU ? sm.zs[j] : dx[j - NU]);
            s2 += bav[jq2] * (j2 < NU ? sm.zs[j2] : dx[j2 - NU]);
          }
          if (e < e1) { const int jq = sm.csr_idx[e], j = jq >> 2; s += bav[jq] * (j < NU ? sm.zs[j] : dx[j - NU]); }
          s += s2;
          if (G == 4) { s += par.shfl4x(s, 1); s += par.shfl4x(s, 2); }
          if (live && sub == 0) { dxall[(i + 1) * NX + r] = s; gDX()[(i + 1) * NX + r] = s; }
        }
      }
      par.wait_async();
      par.sync();
    }
    // costates (full step) of all stages: y_i = p_i + P_i dx_i, terminal y_N = p_N + P_N dx_N (P_N diagonal)
    for (int t = tid; t < (N + 1) * NX; t += nt) {
      const int i = t / NX, r = t - i * NX;
      const double* dx = dxall + i * NX;
      double s;
      if (i < N) {
        const double* fac = gFAC() + (size_t)i * FACSZ;
        double s0 = fac[F_PV + r], s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int j = 0; j < NX; j += 4) {
          s0 += fac[F_P + j * NX + r] * dx[j]; s1 += fac[F_P + (j + 1) * NX + r] * dx[j + 1];
          s2 += fac[F_P + (j + 2) * NX + r] * dx[j + 2]; s3 += fac[F_P + (j + 3) * NX + r] * dx[j + 3];
        }
        s = (s0 + s1) + (s2 + s3);
      } else {
        const double* rec = gREC() + (size_t)N * RECSZ;
        s = rec[Q_GC + 32 + r] + mu * rec[Q_M1 + 32 + r] + rec[Q_M2 + 32 + r] + (rec[Q_DIAG + 32 + r] + reg) * dx[r];
      }
      gYN()[t] = s;
    }
    par.sync();
  }

  // ---- slack steps ds = -(g - relax + s) - J dz row by row (the row residuals come from the eval pass), multiplier
  // steps, the fraction-to-boundary step lengths and the directional derivative of the barrier function.
  // CTA-wide: items are (stage, role) with the roles of the eval pass (8 vertices, CoM rows, foot rows).
  CMPC_HD void slack_steps(double tau, double* a_p, double* a_d, double* dphi) {
    Smem& sm = par.template smem<Smem>();
    const int N = C().N, tid = par.tid(), nt = par.nt();
    double (*st)[4] = reinterpret_cast<double (*)[4]>(sm.M);       // [(N + 1) * 10][4]: a_p, a_d, grad'dz, sum ds/s per item
    static_assert((NMAX + 1) * 10 * 4 <= MSZ + NX * NZ, "slack step scratch must fit in M | W");
    for (int t = tid; t < (N + 1) * 10; t += nt) {
      const int role = t / (N + 1), i = t - role * (N + 1);
      const bool has_u = i < N;
      const uint64_t mask = sm.mask[i];
      const double* dx = gDX() + i * NX; const double* du = gDU() + (has_u ? i : 0) * NU;
      const double* s = gS() + i * NR; const double* lam = gLAM() + i * NR; double* ds = gDS() + i * NR;
      const double* rec = gREC() + (size_t)i * RECSZ;
      double ap = 1.0, ad = 1.0, gd = 0.0, dsos = 0.0;
      auto row = [&](int r, double dsr) {
        const double sv = s[r], lv = lam[r];
        ds[r] = dsr;
        const double dl = -lv + mu / sv - lv / sv * dsr;
        if (dsr < 0.0) { const double a = -tau * sv / dsr; ap = a < ap ? a : ap;
#ifdef CMPC_TRACE
          if (cmpc_trace_on > 1 && a < 0.9) printf("      a_p %.2e stage %d row %d s %.2e ds %.2e lam %.2e rg %.2e\n", a, i, r, sv, dsr, lv, r >= 2 ? rec[Q_RG + r] : 0.0);
#endif
        }
        if (dl < 0.0) { const double a = -tau * lv / dl; ad = a < ad ? a : ad; }
        dsos += dsr / sv;
      };
      if (role < 8) {
        const int v = role;
        if (has_u) {
          const double f0 = du[3 * v], f1 = du[3 * v + 1], f2 = du[3 * v + 2], mf = C().mu_fric;
          if (mask & (1ull << (R_UNI + v))) {
            row(R_FRIC + 4 * v + 0, -rec[Q_RG + R_FRIC + 4 * v + 0] - (f0 - mf * f2));
            row(R_FRIC + 4 * v + 1, -rec[Q_RG + R_FRIC + 4 * v + 1] - (-f0 - mf * f2));
            row(R_FRIC + 4 * v + 2, -rec[Q_RG + R_FRIC + 4 * v + 2] - (f1 - mf * f2));
            row(R_FRIC + 4 * v + 3, -rec[Q_RG + R_FRIC + 4 * v + 3] - (-f1 - mf * f2));
            row(R_UNI + v, -rec[Q_RG + R_UNI + v] - (-f2));
          } else {
            for (int q = 0; q < 4; ++q) ds[R_FRIC + 4 * v + q] = 0.0;
            ds[R_UNI + v] = 0.0;
          }
          gd += rec[Q_GC + 3 * v] * f0 + rec[Q_GC + 3 * v + 1] * f1 + rec[Q_GC + 3 * v + 2] * f2;
        } else {
          for (int q = 0; q < 4; ++q) ds[R_FRIC + 4 * v + q] = 0.0;
          ds[R_UNI + v] = 0.0;
        }
        gd += rec[Q_GC + 32 + IQ + v] * dx[IQ + v];
      } else if (role == 8) {
        // explicit rows: the solve returned the new multiplier w; ds follows from s dlam + lam ds = mu - s lam
        if (mask & (1ull << R_LYAP)) { const double sv = s[R_LYAP], lv = lam[R_LYAP]; row(R_LYAP, mu / lv - sv - sv / lv * (gDW()[i * NW] - lv)); }
        else ds[R_LYAP] = 0.0;
        if (mask & (1ull << R_HW)) { const double sv = s[R_HW], lv = lam[R_HW]; row(R_HW, mu / lv - sv - sv / lv * (gDW()[i * NW + 1] - lv)); }
        else ds[R_HW] = 0.0;
        if (mask & (1ull << R_PZ)) row(R_PZ, -rec[Q_RG + R_PZ] - dx[IP + 2]);
        else ds[R_PZ] = 0.0;
        for (int j = 0; j < 12; ++j) gd += rec[Q_GC + 32 + j] * dx[j];
      } else {
        for (int e = 0; e < 2; ++e) {
          const bool box = (mask >> (R_BOX + 6 * e)) & 1ull;
          for (int j = 0; j < 3; ++j) {
            const int r = R_BOX + 6 * e + 2 * j;
            const double v = dx[(e ? IPR : IPL) + j];
            if (box) { row(r, -rec[Q_RG + r] - v); row(r + 1, -rec[Q_RG + r + 1] + v); }
            else { ds[r] = 0.0; ds[r + 1] = 0.0; }
          }
        }
        for (int j = 12; j < 20; ++j) gd += rec[Q_GC + 32 + j] * dx[j];
        if (has_u) for (int j = 24; j < 32; ++j) gd += rec[Q_GC + j] * du[j];
      }
      st[t][0] = ap; st[t][1] = ad; st[t][2] = gd; st[t][3] = dsos;
    }
    par.sync();
    // two-level reduction: per stage (threads), then over the stages (every thread, broadcast reads)
    for (int i = tid; i <= N; i += nt) {
      double ap = 1.0, ad = 1.0, gd = 0.0, dsos = 0.0;
      for (int role = 0; role < 10; ++role) {
        const double* q = st[role * (N + 1) + i];
        ap = q[0] < ap ? q[0] : ap; ad = q[1] < ad ? q[1] : ad; gd += q[2]; dsos += q[3];
      }
      sm.acc[i][0] = ap; sm.acc[i][1] = ad; sm.acc[i][2] = gd; sm.acc[i][3] = dsos;
    }
    par.sync();
    if (par.warp() == 0) {
      double ap = 1.0, ad = 1.0, gd = 0.0, dsos = 0.0;
      for (int i = par.lane(); i <= N; i += par.lanes()) {
        ap = sm.acc[i][0] < ap ? sm.acc[i][0] : ap; ad = sm.acc[i][1] < ad ? sm.acc[i][1] : ad;
        gd += sm.acc[i][2]; dsos += sm.acc[i][3];
      }
      ap = par.wmin(ap); ad = par.wmin(ad); gd = par.wsum(gd); dsos = par.wsum(dsos);
      if (par.lane() == 0) { sm.red5[0] = ap; sm.red5[1] = ad; sm.red5[2] = gd; sm.red5[3] = dsos; }
    }
    par.sync();
    *a_p = sm.red5[0]; *a_d = sm.red5[1]; *dphi = sm.red5[2] - mu * sm.red5[3];
  }

  // ---- trial point (x + a dx, u + a du, s + a ds) for the line search: constraint violation theta, cost (with and
  // without the regularisation term), sum ln s, max unrelaxed violation.  CTA-wide like the eval pass.
  CMPC_HD void trial(double alpha, double* out) {
    Smem& sm = par.template smem<Smem>();
    const int N = C().N, tid = par.tid(), nt = par.nt();
    TrialScratch* ts = reinterpret_cast<TrialScratch*>(sm.M);
    const double d = C().delta, m = IN().mass, k1 = IN().k1;
    for (int i0 = 0; i0 <= N; i0 += TCH) {
      const int ns = (N + 1 - i0) < TCH ? (N + 1 - i0) : TCH;
      for (int t = tid; t < ns * 2; t += nt) {
        const int il = t >> 1, e = t & 1, i = i0 + il, o = i * NX + (e ? IPSR : IPSL);
        const double psi = gX()[o] + alpha * gDX()[o];
        ts[il].cs[e] = cos(psi); ts[il].sn[e] = sin(psi);
      }
      par.sync();
      for (int t = tid; t < ns * 8; t += nt) {
        const int il = t >> 3, v = t & 7, i = i0 + il;
        if (i >= N) continue;
        const int e = v >> 2, k = v & 3;
        const double* X = gX() + i * NX; const double* DX = gDX() + i * NX;
        const double* U = gU() + i * NU; const double* DU = gDU() + i * NU;
        const int po = e ? IPR : IPL;
        const double cs = ts[il].cs[e], sn = ts[il].sn[e];
        double cx, cy; corner(C(), k, cx, cy);
        const double rx = cs * cx - sn * cy, ry = sn * cx + cs * cy;
        const double p0 = X[0] + alpha * DX[0], p1 = X[1] + alpha * DX[1], p2 = X[2] + alpha * DX[2];
        const double r0 = rx + X[po] + alpha * DX[po] - p0, r1 = ry + X[po + 1] + alpha * DX[po + 1] - p1, r2 = X[po + 2] + alpha * DX[po + 2] - p2;
        const double f0 = U[3 * v] + alpha * DU[3 * v], f1 = U[3 * v + 1] + alpha * DU[3 * v + 1], f2 = U[3 * v + 2] + alpha * DU[3 * v + 2];
        double* p = ts[il].part[v];
        p[0] = f0; p[1] = f1; p[2] = f2;
        p[3] = r1 * f2 - r2 * f1; p[4] = r2 * f0 - r0 * f2; p[5] = r0 * f1 - r1 * f0;
        p[6] = f0 * f0 + f1 * f1 + f2 * f2;
      }
      par.sync();
      for (int t = tid; t < ns * 11; t += nt) {
        const int il = t / 11, q = t - il * 11, i = i0 + il;
        if (i >= N) continue;
        const double gl = i_gamma()[2 * i], gr = i_gamma()[2 * i + 1];
        const double (*pt)[7] = ts[il].part;
        double s = 0.0;
        if (q < 6) { const int e = q / 3, j = q - 3 * e; for (int k = 0; k < 4; ++k) s += pt[4 * e + k][j]; }
        else if (q < 9) { const int j = q - 6; for (int v = 0; v < 8; ++v) s += (v < 4 ? gl : gr) * pt[v][3 + j]; }
        else { const int e = q - 9; for (int k = 0; k < 4; ++k) s += pt[4 * e + k][6]; }
        ts[il].sum[q] = s;
      }
      par.sync();
      for (int t = tid; t < ns * 10; t += nt) {
        const int role = t / ns, il = t - role * ns, i = i0 + il;
        const bool has_u = i < N;
        TrialScratch& E = ts[il];
        const double* X = gX() + i * NX; const double* DX = gDX() + i * NX;
        const double* U = gU() + (has_u ? i : 0) * NU; const double* DU = gDU() + (has_u ? i : 0) * NU;
        const double* Xn = gX() + (has_u ? i + 1 : i) * NX; const double* DXn = gDX() + (has_u ? i + 1 : i) * NX;
        const double* S = gS() + i * NR; const double* DS = gDS() + i * NR;
        const uint64_t mask = sm.mask[i];
        const double gl = i_gamma()[2 * i], gr = i_gamma()[2 * i + 1];
        double theta = 0.0, cost = 0.0, cref = 0.0, viol = 0.0;
        double sprod = 1.0, smin_t = 1.0;                          // sum ln s of the item's rows (<= 12) as one logarithm of their product
        auto row = [&](int r, double g) {
          const double st = S[r] + alpha * DS[r];
          theta += fabs(g - C().relax + st); sprod *= st; smin_t = st < smin_t ? st : smin_t; viol = g > viol ? g : viol;
        };
        auto defect = [&](int r, double xp) {
          const double ad = fabs(xp - (Xn[r] + alpha * DXn[r]));
          theta += ad; viol = ad > viol ? ad : viol;
        };
        if (role < 8) {
          if (has_u) {
            const int v = role, e = v >> 2;
            const double ge = e ? gr : gl, mf = C().mu_fric;
            const double* p = E.part[v];
            const double f0 = p[0], f1 = p[1], f2 = p[2];
            if (mask & (1ull << (R_UNI + v))) {
              row(R_FRIC + 4 * v + 0, f0 - mf * f2); row(R_FRIC + 4 * v + 1, -f0 - mf * f2);
              row(R_FRIC + 4 * v + 2, f1 - mf * f2); row(R_FRIC + 4 * v + 3, -f1 - mf * f2);
              row(R_UNI + v, -f2);
            }
            cost += (ge * C().w_sym + (1.0 - ge) * C().w_swing) * p[6];
            if (i >= 1) {
              const double dz = f2 - (X[IQ + v] + alpha * DX[IQ + v]);
              cost += C().w_rate * i_gamma()[2 * (i - 1) + e] * dz * dz;
            }
            cref = cost;
            defect(IQ + v, f2);
          }
        } else if (role == 8) {
          double x[12];
#pragma unroll
          for (int j = 0; j < 12; ++j) x[j] = X[j] + alpha * DX[j];
          if (i >= 1) {
            const double* ref = i_com_ref() + 9 * (i - 1);
            cost += C().w_xy * ((x[0] - ref[0]) * (x[0] - ref[0]) + (x[1] - ref[1]) * (x[1] - ref[1])) + wz_of(C(), i - 1) * (x[2] - ref[2]) * (x[2] - ref[2]);
          }
          if (mask & (1ull << R_PZ)) row(R_PZ, x[IP + 2] - C().pz_max);
          if (has_u) {
            const double* ref = i_com_ref() + 9 * i;
            double F[3], xph[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) { F[j] = gl * E.sum[j] + gr * E.sum[3 + j]; xph[j] = x[IH + j] + d * E.sum[6 + j]; }
            cost += C().w_h * (x[IH] * x[IH] + x[IH + 1] * x[IH + 1] + x[IH + 2] * x[IH + 2]);
            // symmetry term: w_sym (sum |f_k|^2 - 4 |mean|^2) per stance foot; the first part is with the vertices
            cost -= gl * C().w_sym * 0.25 * (E.sum[0] * E.sum[0] + E.sum[1] * E.sum[1] + E.sum[2] * E.sum[2]);
            cost -= gr * C().w_sym * 0.25 * (E.sum[3] * E.sum[3] + E.sum[4] * E.sum[4] + E.sum[5] * E.sum[5]);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              defect(IP + j, x[IP + j] + d * x[IV + j]);
              defect(IV + j, x[IV + j] + d * ((j == 2 ? -C().grav : 0.0) + F[j] / m));
              defect(IH + j, xph[j]);
              defect(ITH + j, x[ITH + j] + (d / m) * (k1 * (x[IP + j] - ref[j]) + x[IV + j] - ref[3 + j]));
            }
            if (mask & (1ull << R_LYAP)) {
              double q = 0.0;
#pragma unroll
              for (int j = 0; j < 3; ++j) {
                const double grav = (j == 2) ? -C().grav : 0.0;
                const double vp = x[IV + j] + d * (grav + F[j] / m);
                const double z1 = x[IP + j] + d * x[IV + j] - ref[j];
                const double z2 = k1 * z1 + vp - ref[3 + j];
                const double ae = F[j] / m + grav - ref[6 + j] + x[ITH + j] / m;
                q += -k1 * z1 * z1 + k1 * z2 * z2 + (1.0 - k1 * k1) * z1 * z2 + z2 * ae;
              }
              row(R_LYAP, q);
            }
            if (mask & (1ull << R_HW))
              row(R_HW, xph[0] * xph[0] + xph[1] * xph[1] + xph[2] * xph[2] - (x[IH] * x[IH] + x[IH + 1] * x[IH + 1] + x[IH + 2] * x[IH + 2]));
          }
          cref = cost;
        } else {
          const double* fr = i_foot_ref() + 8 * (i >= 1 ? i - 1 : 0);
          double creg = 0.0;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const double ge = e ? gr : gl;
            const int po = e ? IPSR : IPSL, xo = e ? IPR : IPL;
            const bool box = (i >= 1) && ((mask >> (R_BOX + 6 * e)) & 1ull);
            const double psi = X[po] + alpha * DX[po];
            if (i >= 1) cost += C().w_foot * ge * (psi - fr[6 + e]) * (psi - fr[6 + e]);
            if (has_u) { const double uv = U[30 + e] + alpha * DU[30 + e]; defect(po, psi + d * (1.0 - ge) * uv); creg += C().eps_reg * uv * uv; }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const double xv = X[xo + j] + alpha * DX[xo + j];
              if (i >= 1) {
                const double err = xv - fr[3 * e + j];
                cost += C().w_foot * ge * err * err;
                if (box) { row(R_BOX + 6 * e + 2 * j, err - C().box[j]); row(R_BOX + 6 * e + 2 * j + 1, -err - C().box[j]); }
              }
              if (has_u) { const double uv = U[24 + 3 * e + j] + alpha * DU[24 + 3 * e + j]; defect(xo + j, xv + d * (1.0 - ge) * uv); creg += C().eps_reg * uv * uv; }
            }
          }
          cref = cost; cost += creg;
        }
        double* st = E.st[role];
        st[0] = theta; st[1] = cost; st[2] = log(smin_t > 0.0 ? sprod : -1.0); st[3] = viol; st[4] = cref;      // NaN if a slack left the domain
      }
      par.sync();
      for (int il = tid; il < ns; il += nt) {
        double theta = 0, cost = 0, lns = 0, viol = 0, cref = 0;
        for (int r = 0; r < 10; ++r) {
          const double* q = ts[il].st[r];
          theta += q[0]; cost += q[1]; lns += q[2]; viol = q[3] > viol ? q[3] : viol; cref += q[4];
        }
        double* a = sm.acc[i0 + il];
        a[0] = theta; a[1] = cost; a[2] = lns; a[3] = viol; a[4] = cref;
      }
      par.sync();
    }
    reduce_trial(out);
  }

  CMPC_HD void reduce_trial(double* out) {
    Smem& sm = par.template smem<Smem>();
    const int N = C().N;
    if (par.warp() == 0) {
      double theta = 0, cost = 0, lns = 0, viol = 0, cref = 0;
      for (int i = par.lane(); i <= N; i += par.lanes()) {
        theta += sm.acc[i][0]; cost += sm.acc[i][1]; lns += sm.acc[i][2];
        viol = sm.acc[i][3] > viol ? sm.acc[i][3] : viol; cref += sm.acc[i][4];
      }
      theta = par.wsum(theta); cost = par.wsum(cost); lns = par.wsum(lns); viol = par.wmax(viol); cref = par.wsum(cref);
      if (par.lane() == 0) { sm.red5[0] = theta; sm.red5[1] = cost; sm.red5[2] = lns; sm.red5[3] = viol; sm.red5[4] = cref; }
    }
    par.sync();
    out[0] = sm.red5[0]; out[1] = sm.red5[1]; out[2] = sm.red5[2]; out[3] = sm.red5[3]; out[4] = sm.red5[4];
  }

  CMPC_HD void apply_step(double alpha, double a_d) {
    Smem& sm = par.template smem<Smem>();
    const int N = C().N, tid = par.tid(), nt = par.nt();
    for (int t = tid; t < (N + 1) * NX; t += nt) { gX()[t] += alpha * gDX()[t]; gY()[t] += alpha * (gYN()[t] - gY()[t]); }
    for (int t = tid; t < N * NU; t += nt) gU()[t] += alpha * gDU()[t];
    for (int t = tid; t < (N + 1) * NR; t += nt) {
      const int i = t / NR, r = t % NR;
      if (!(sm.mask[i] & (1ull << r))) continue;
      const double s0 = gS()[t], l0 = gLAM()[t], dsr = gDS()[t];
      const double dl = -l0 + mu / s0 - l0 / s0 * dsr;
      double sn = s0 + alpha * dsr, ln = l0 + a_d * dl;
      const double lo = mu / (1e10 * sn), hi = 1e10 * mu / sn;       // IPOPT eq. (16)
      ln = ln < lo ? lo : (ln > hi ? hi : ln);
      gS()[t] = sn; gLAM()[t] = ln;
    }
    par.sync();
  }

  // ---- solve with retries.  An interior-point method started next to the boundary of a changed active set can jam,
  // and the non-convex end game occasionally stalls a few 1e-8 short of the tolerance: a failed warm solve is repeated
  // from the solver's own cold start, a failed cold solve again with a ten times larger, then a ten times smaller initial barrier (other central paths).
  // ---- per-solve tables: structural pattern of [B A], Lyapunov constants and scatter table, tile map
  CMPC_HD void setup() {
    Smem& sm = par.template smem<Smem>();
    if (par.tid() == 0) { sm.cfg = c0; sm.inst = in0; sm.wk = w0; }       // (published by the barrier below)
    // structural pattern of [B A]: by column (ba_row) and by row (gather), built by the whole CTA
    for (int t = par.tid(); t < NZ * 4; t += par.nt()) {
      const int r0 = ba_row(t >> 2, t & 3);
      sm.barow[t] = (signed char)r0;
      sm.barz[t] = (unsigned char)(r0 >= 0 ? r0 : 0); sm.baofs[t] = (unsigned short)((r0 >= 0 ? r0 : 0) * NZ);
    }
    par.sync();
    for (int r = par.tid(); r < NX; r += par.nt()) {            // entries of row r, and the number of entries in the rows before it
      int before = 0, n = 0;
      for (int t = 0; t < NZ * 4; ++t) { const int rr = sm.barow[t]; before += (rr >= 0 && rr < r); n += (rr == r); }
      sm.csr_ptr[r] = (short)before;
      if (r == NX - 1) sm.csr_ptr[NX] = (short)(before + n);
      int k = before;
      for (int t = 0; t < NZ * 4; ++t) if (sm.barow[t] == r) sm.csr_idx[k++] = (unsigned char)t;
    }
    if (par.tid() == 0) lyapunov_consts(C(), IN(), sm.lyapC);
    // Lyapunov row: which entries of the stage block it touches, with which coefficient.  Touched variable a (0..32):
    // forces 0..23 (type F, scaled by gamma_e), then p, v, theta; curvature lam * C (x) I_3 couples same-axis pairs only.
    for (int e = par.tid(); e < NLY; e += par.nt()) {
      int ai, bi = -1;
      if (e < 33) ai = e;
      else {                                     // pair index within the axis: 66 pairs (bq <= aq) of 11 same-axis variables
        const int q = e - 33, xa = q / 66, pq = q % 66;
        int aq = 0; while ((aq + 1) * (aq + 2) / 2 <= pq) ++aq;
        const int bq = pq - aq * (aq + 1) / 2;
        ai = 3 * aq + xa; bi = 3 * bq + xa;
      }
      auto var = [](int a, int& mm, int& tt, int& gg) {
        if (a < 24) { mm = a; tt = 3; gg = a / 12; }
        else { const int q = a - 24; tt = q / 3; gg = 2; mm = XO + (tt == 0 ? IP : (tt == 1 ? IV : ITH)) + q % 3; }
      };
      int ma, ta, ga; var(ai, ma, ta, ga);
      if (bi < 0) {
        sm.ly_m[e] = (unsigned short)(ma < NU ? mi(NU, ma) : mi(ma, NU));
        sm.ly_s[e] = (unsigned short)(Q_LG + 3 * ta + ai % 3);
        sm.ly_g[e] = (unsigned char)(ga | (2 << 4));
      } else {
        int mb, tb, gb; var(bi, mb, tb, gb);
        sm.ly_m[e] = (unsigned short)(ma >= mb ? mi(ma, mb) : mi(mb, ma));
        sm.ly_s[e] = (unsigned short)(Q_LC + 4 * ta + tb);
        sm.ly_g[e] = (unsigned char)(ga | (gb << 4));
      }
    }
    for (int sl = 0; sl < Par::TPT; ++sl) {
      const int tile = par.tid() + sl * par.nt();                   // tile (ti, tj), ti >= tj >= 1, row-major in the triangle
      int ti = 0;
      if (tile < NTILE) { while ((ti + 1) * (ti + 2) / 2 <= tile) ++ti; }
      tile_i[sl] = tile < NTILE ? ti + 1 : -1;
      tile_j[sl] = tile < NTILE ? tile - ti * (ti + 1) / 2 + 1 : 0;
    }
    par.sync();
  }

  // one attempt (a launch pass of the CUDA path: the retries of failed instances are separate, compact launches there)
  CMPC_HD void run_pass(int warm, double scale, Stats* st, bool blend = false) {
    setup();
    mu_scale = scale; mu = 0; reg_last = 0; start_blend = blend;
    run_once(warm, st);
  }

  // all attempts in sequence (host build of the core, tests/hostsim)
  CMPC_HD void run(int warm, Stats* st) {
    setup();
    // attempts: as asked (warm or cold, mu_init) -> cold, mu_init -> cold, 10 mu_init -> cold from the reference-blended start, mu_init
    mu_scale = 1.0;
    run_once(warm, st);
    int it0 = st->iters;
    for (int attempt = (warm == 0 ? 1 : 0); attempt < 3; ++attempt) {
      if (st->status == ST_CONVERGED || st->status == ST_INFEASIBLE_X0) break;
      mu_scale = (attempt == 0) ? 1.0 : (attempt == 1 ? 10.0 : 1.0);
      start_blend = (attempt == 2);                                  // last attempt: another starting point instead of another barrier value
      par.sync();
      mu = 0; reg_last = 0;
      run_once(0, st);
      st->iters += it0; it0 = st->iters;
    }
  }

  // ---- the interior-point loop (kept a function of its own on the device: inlined into the kernel it costs kilobytes of spills)
  CMPC_HD_NOINLINE void run_once(int warm, Stats* st) {
    Smem& sm = par.template smem<Smem>();
    const int N = C().N;
    double pviol;
    if (par.tid() == 0) { build_masks(C(), IN(), sm.mask, &pviol); sm.red[0] = pviol; }
    par.sync();
    pviol = sm.red[0];
    par.sync();
    init_point(warm);
    const int nrows = n_rows_total();
    double ev[8], parts[3];
    double filt[16][2]; int nfilt = 0;
    double theta_max = 0, theta_min = 0; bool have_theta0 = false;
    int status = ST_MAXITER, it = 0, ls_fail = 0, stall_it = 0, jam = 0, crawl = 0;
    double stall_ref = -1.0;
    double kkt = 0.0;
    // merit quantities of the current point (constraint violation theta, cost, sum ln s): evaluated once here, afterwards
    // they are the values of the accepted trial point -- the eval pass needs no logarithms
    double cur[5];
    for (int t = par.tid(); t < (N + 1) * NX; t += par.nt()) gDX()[t] = 0.0;
    for (int t = par.tid(); t < N * NU; t += par.nt()) gDU()[t] = 0.0;
    for (int t = par.tid(); t < (N + 1) * NR; t += par.nt()) gDS()[t] = 0.0;
    par.sync();
    trial(0.0, cur);
    eval(ev);
    for (it = 0; it <= C().max_iter; ++it) {
      kkt = kkt_error(ev, C().mu_final, nrows, parts);
      if (!(ev[0] == ev[0]) || !(ev[1] == ev[1]) || !(cur[1] == cur[1])) { status = ST_NAN; break; }
      if (mu <= C().mu_final && kkt <= C().tol) { status = ST_CONVERGED; break; }
      if (it == C().max_iter) break;
      // monotone barrier update (IPOPT eq. 7)
      while (mu > C().mu_final) {
        double p2[3];
        const double emu = kkt_error(ev, mu, nrows, p2);
        if (emu > C().kappa_eps * mu) break;
        double m1 = C().kappa_mu * mu, m2 = pow(mu, C().theta_mu);
        mu = m1 < m2 ? m1 : m2; if (mu < C().mu_final) mu = C().mu_final;
        nfilt = 0;
        stall_it = it; stall_ref = -1.0;
      }
      // stall test: an attempt whose barrier-problem error has not halved within `stall_window` iterations at one barrier
      // value (`stall_final` at the last one: a healthy end game takes two to four iterations, an attempt parked at a
      // saddle of the non-convex NLP -- primal feasible, complementary, dual residual stuck, regularisation at every
      // iteration -- never leaves it) is abandoned (it would run to max_iter: jammed at a saddle of the non-convex NLP, or diverging); the caller
      // retries from another start
      const int win_mid = (warm != 0 && C().warm_stall_window > 0) ? C().warm_stall_window : C().stall_window;
      if (win_mid > 0 || C().stall_final > 0) {
        double p2[3];
        const double emu = kkt_error(ev, mu, nrows, p2);
        if (stall_ref < 0.0) { stall_ref = emu; stall_it = it; }
        else if ((mu <= C().mu_final ? C().stall_final : win_mid) > 0 && it - stall_it >= (mu <= C().mu_final ? C().stall_final : win_mid)) {
          if (emu > 0.5 * stall_ref) { status = ST_STALL; break; }
          stall_ref = emu; stall_it = it;
        }
      }
      const double tau = (1.0 - mu) > C().tau_min ? (1.0 - mu) : C().tau_min;
      // factorise, regularising as IPOPT does when the input block is not positive definite
      double reg = 0.0; bool ok = false;
      for (int attempt = 0; attempt < 40; ++attempt) {
        ++nfact;
        ok = backward(reg);
        par.sync();
        if (ok) break;
        ++nreg;
        if (reg == 0.0) reg = (reg_last == 0.0) ? 1e-4 : (reg_last / 3.0 > 1e-20 ? reg_last / 3.0 : 1e-20);
        else reg *= (reg_last == 0.0) ? 100.0 : 8.0;
        if (reg > 1e12) break;                              // (keeps the pivots inside the range of the single-precision rsqrt seed)
      }
      if (!ok) { status = ST_REGULARIZATION; break; }
      if (reg > 0.0) reg_last = reg;
      { CMPC_TIC(sm); forward(reg); CMPC_TOC(sm, PF_FWD); }
      double a_p, a_d, dphi;
      { CMPC_TIC(sm); slack_steps(tau, &a_p, &a_d, &dphi); CMPC_TOC(sm, PF_SLACK); }
      // short filter line search (Waechter-Biegler eq. 18-20); falls back to the full
      // fraction-to-boundary step if `ls_max` halvings are all rejected
      const double theta = cur[0], phi = cur[1] - mu * cur[2];      // (theta, cost, sum ln s) of the current point
      if (!have_theta0) { have_theta0 = true; theta_max = 1e4 * (theta > 1.0 ? theta : 1.0); theta_min = 1e-4 * (theta > 1.0 ? theta : 1.0); }
      double alpha = a_p; bool accepted = false; double tr[5];
      CMPC_TIC(sm);
      for (int ls = 0; ls < C().ls_max; ++ls) {
        trial(alpha, tr);
        const double th_t = tr[0], ph_t = tr[1] - mu * tr[2];
        bool okf = (th_t == th_t) && (ph_t == ph_t) && th_t <= theta_max;
        for (int f = 0; okf && f < nfilt; ++f) okf = (th_t < (1.0 - 1e-5) * filt[f][0]) || (ph_t < filt[f][1] - 1e-5 * filt[f][0]);
        if (okf) {
          const bool sw = theta <= theta_min && dphi < 0.0 && alpha * pow(-dphi, 2.3) > pow(theta, 1.1);
          if (sw) {
            if (ph_t <= phi + 1e-4 * alpha * dphi + 2.2e-15 * fabs(phi)) { accepted = true; break; }
          } else if (th_t <= (1.0 - 1e-5) * theta || ph_t <= phi - 1e-5 * theta) {
            accepted = true;
            if (nfilt < 16) { filt[nfilt][0] = (1.0 - 1e-5) * theta; filt[nfilt][1] = phi - 1e-5 * theta; ++nfilt; }
            break;
          }
        }
        alpha *= 0.5;
      }
      if (!accepted) {
        alpha = a_p;
        trial(alpha, tr);
        if (!(tr[0] == tr[0]) || !(tr[1] == tr[1]) || !(tr[2] == tr[2])) {
          // the full step leaves the domain: shrink until finite
          int k = 0;
          for (; k < 30; ++k) { alpha *= 0.5; trial(alpha, tr); if (tr[0] == tr[0] && tr[1] == tr[1] && tr[2] == tr[2]) break; }
          if (k == 30) { if (++ls_fail > 3) { status = ST_LINESEARCH; break; } }
        }
      }
      CMPC_TOC(sm, PF_TRIAL);
      for (int q = 0; q < 5; ++q) cur[q] = tr[q];
      const double prim_before = ev[0];
      { CMPC_TIC(sm); apply_step(alpha, a_d); CMPC_TOC(sm, PF_APPLY); eval(ev); CMPC_TOC(sm, PF_EVAL); }
      // warm-start safeguard: a warm point can sit next to the boundary of a changed active set where the fraction-to-boundary
      // rule cuts every step while the infeasibility GROWS (observed: 55 to 100 iterations with step lengths below 0.1 and the
      // primal residual rising from 0.5 to 17 before the solve takes off; a cold start of the same tick needs 25).
      // `jam_window` consecutive such steps abandon the warm attempt, the instance restarts cold.
      if (warm != 0 && C().jam_window > 0) {
        jam = (alpha < 0.1 && ev[0] > prim_before) ? jam + 1 : 0;
        if (jam >= C().jam_window) { status = ST_STALL; ++it; break; }
      }
      // second form of the same jam (tick 867 of the recorded walk, inside the push window: 50 iterations with step lengths of
      // 1e-2 .. 1e-6 while the infeasibility creeps DOWN from 0.9 to 0.5, then 24 healthy ones -- 68 to 96 iterations on the GPU
      // depending on the rounding, the longest solve of every benchmark batch; cold: 36): `crawl_window` consecutive steps shorter
      // than `crawl_alpha` abandon the warm attempt whatever the residual does.  All 1916 recorded ticks, warm (CPU build of this
      // core): longest solve 74 -> 54 iterations, mean 11.516 -> 11.511 with a window of 12 (8: 52 / 11.55), same KKT points (cost within 5e-12).  Rejected on the same
      // replay: a shorter stall window for warm attempts (mean 11.9 .. 13.8, other KKT points), refusing warm points with a large
      // primal residual (mean 15.4)
      if (warm != 0 && C().crawl_window > 0) {
        crawl = (alpha < C().crawl_alpha) ? crawl + 1 : 0;
        if (crawl >= C().crawl_window) { status = ST_STALL; ++it; break; }
      }
#ifdef CMPC_TRACE
      if (cmpc_trace_on) printf("it %3d cost %.8e prim %.2e dual %.2e smax %.2e smin %.2e mu %.1e reg %.1e a_p %.2e a_d %.2e alpha %.2e acc %d\n",
                               it, cur[1], ev[0], ev[1], ev[2], ev[3], mu, reg, a_p, a_d, alpha, (int)accepted);
#endif
    }
    // final report: reference cost (no eps_reg term) and max unrelaxed violation incl. dynamics defects
    double tr[5];
    for (int t = par.tid(); t < (N + 1) * NX; t += par.nt()) gDX()[t] = 0.0;
    for (int t = par.tid(); t < N * NU; t += par.nt()) gDU()[t] = 0.0;
    for (int t = par.tid(); t < (N + 1) * NR; t += par.nt()) gDS()[t] = 0.0;
    par.sync();
    trial(0.0, tr);
    if (status == ST_CONVERGED && pviol > 1e-6) status = ST_INFEASIBLE_X0;
    st->cost = tr[4]; st->viol = tr[3] > pviol ? tr[3] : (pviol > 0 ? pviol : tr[3]);
    st->kkt = kkt; st->mu = mu; st->iters = it; st->status = status; st->nfact = nfact; st->nreg = nreg;
  }
};

}  // namespace cmpc
