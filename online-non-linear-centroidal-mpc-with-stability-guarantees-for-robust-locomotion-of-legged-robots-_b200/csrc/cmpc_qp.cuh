// cmpc_qp.cuh -- batched dense convex QP for the whole-body inverse-dynamics step that consumes the MPC's output
// (SURVEY.md 8f row N4).  Replaces `QPSolver.set_values / solve` of the reference (`code/utils.py:40-92`: CasADi
// Opti('conic') + OSQP) as it is used by `InverseDynamics.get_joint_torques` (`code/inverse_dynamics.py:30-135`):
//
//     min 1/2 x'H x + F'x   s.t.  A_eq x = b_eq,  A_in x <= b_in          (utils.py:50-66)
//
// with n = 2 dofs + 12 variables (joint accelerations, torques, two contact wrenches), dofs equality rows (the
// equations of motion) and 16 inequality rows (centre-of-pressure and friction limits of two feet).  One CTA per QP, all
// matrices in shared memory, FP64.  Method: Mehrotra predictor-corrector interior point; each iteration factorises the
// quasi-definite matrix [[H + eps I + A_in' (z/s) A_in, A_eq'], [A_eq, -eps I]] once (right-looking LDL', no pivoting:
// the first n pivots are positive, the last m_eq negative) and solves with it twice.  H of the reference is singular
// (no cost on the torques; the six floating-base "torques" appear nowhere): eps = 1e-9 selects the minimum-norm member,
// as OSQP's own sigma-regularisation does; the returned actuated torques do not depend on it.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace cmpc_qp {

constexpr int QP_THREADS = 128;
constexpr int QP_MAXN = 96, QP_MAXE = 48, QP_MAXI = 32;

struct QpDims { int n, me, mi, max_iter; double tol, eps, piv; };

__host__ __device__ inline size_t qp_smem_doubles(int n, int me, int mi) {
  const int N = n + me;
  return (size_t)N * N + 4 * (size_t)N + 6 * (size_t)mi + 8;
}

// warp all-reduce helpers (blockDim = QP_THREADS)
__device__ __forceinline__ double qp_wmax(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  return v;
}
__device__ __forceinline__ double qp_wmin(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
  return v;
}
__device__ __forceinline__ double qp_wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block reduction of (max, min, sum) triples through a small shared buffer `red` (>= 12 doubles)
__device__ inline void qp_block_reduce3(double& vmax, double& vmin, double& vsum, double* red) {
  vmax = qp_wmax(vmax); vmin = qp_wmin(vmin); vsum = qp_wsum(vsum);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) { red[3 * w] = vmax; red[3 * w + 1] = vmin; red[3 * w + 2] = vsum; }
  __syncthreads();
  double a = red[0], b = red[1], c = red[2];
  for (int k = 1; k < nw; ++k) { a = red[3 * k] > a ? red[3 * k] : a; b = red[3 * k + 1] < b ? red[3 * k + 1] : b; c += red[3 * k + 2]; }
  vmax = a; vmin = b; vsum = c;
}

// One QP per CTA.  Global inputs of QP q: H [n][n], F [n], Aeq [me][n], beq [me], Ain [mi][n], bin [mi] (row-major).
// Outputs: x [n], status (0 converged, 1 max_iter, 3 nan pivot, 5 nan residual), iters.
__global__ void __launch_bounds__(QP_THREADS)
cmpc_qp_kernel(QpDims d, int batch, const double* __restrict__ H, const double* __restrict__ F, const double* __restrict__ Aeq,
               const double* __restrict__ beq, const double* __restrict__ Ain, const double* __restrict__ bin_, double* __restrict__ xout,
               int32_t* __restrict__ status, int32_t* __restrict__ iters) {
  extern __shared__ __align__(16) double qsm[];
  const int n = d.n, me = d.me, mi = d.mi, N = n + me, tid = threadIdx.x, nt = blockDim.x;
  double* K = qsm;                    // [N][N] lower triangle used: the KKT matrix, then its LDL' factors (D on the diagonal)
  double* rhs = K + (size_t)N * N;    // [N] right-hand side / solution
  double* xv = rhs + N;               // [N] x (n) and y (me)
  double* dxa = xv + N;               // [N] predictor step
  double* rd = dxa + N;               // [N] residuals r_d (n), r_p (me)
  double* sv = rd + N;                // [mi] slacks
  double* zv = sv + mi;               // [mi] multipliers
  double* ri = zv + mi;               // [mi] A_in x + s - b_in
  double* rc = ri + mi;               // [mi] complementarity target
  double* dsa = rc + mi;              // [mi] predictor slack step
  double* dza = dsa + mi;             // [mi] predictor multiplier step
  double* red = dza + mi;             // [8+] reductions / flags
  for (int q = blockIdx.x; q < batch; q += gridDim.x) {
    const double* Hq = H + (size_t)q * n * n; const double* Fq = F + (size_t)q * n;
    const double* Aq = Aeq + (size_t)q * me * n; const double* bq = beq + (size_t)q * me;
    const double* Gq = Ain + (size_t)q * mi * n; const double* hq = bin_ + (size_t)q * mi;
    for (int t = tid; t < N; t += nt) xv[t] = 0.0;
    for (int t = tid; t < mi; t += nt) { const double s0 = hq[t]; sv[t] = s0 > 1.0 ? s0 : 1.0; zv[t] = 1.0; }
    __syncthreads();
    int st = 1, it = 0;
    // scale of the data for the relative stopping test
    double scl;
    {
      double a = 0.0, b = 0.0, c = 0.0;
      for (int t = tid; t < n; t += nt) { const double v = fabs(Fq[t]); a = v > a ? v : a; }
      for (int t = tid; t < me; t += nt) { const double v = fabs(bq[t]); a = v > a ? v : a; }
      for (int t = tid; t < mi; t += nt) { const double v = fabs(hq[t]); a = v > a ? v : a; }
      qp_block_reduce3(a, b, c, red);
      scl = a > 1.0 ? a : 1.0;
    }
    for (it = 0; it <= d.max_iter; ++it) {
      // ---- residuals: r_d = H x + F + Aeq' y + Ain' z ; r_p = Aeq x - beq ; r_i = Ain x + s - bin
      for (int t = tid; t < n; t += nt) {
        double s = Fq[t] + d.eps * xv[t];
        for (int j = 0; j < n; ++j) s += Hq[(size_t)t * n + j] * xv[j];          // (H symmetric: row access, coalesced over j per thread is not needed at these sizes)
        for (int r = 0; r < me; ++r) s += Aq[(size_t)r * n + t] * xv[n + r];
        for (int r = 0; r < mi; ++r) s += Gq[(size_t)r * n + t] * zv[r];
        rd[t] = s;
      }
      for (int t = tid; t < me; t += nt) {
        double s = -bq[t];
        for (int j = 0; j < n; ++j) s += Aq[(size_t)t * n + j] * xv[j];
        rd[n + t] = s;
      }
      for (int t = tid; t < mi; t += nt) {
        double s = sv[t] - hq[t];
        for (int j = 0; j < n; ++j) s += Gq[(size_t)t * n + j] * xv[j];
        ri[t] = s;
      }
      __syncthreads();
      double rmax = 0.0, dummy = 0.0, comp = 0.0;
      for (int t = tid; t < N; t += nt) { const double v = fabs(rd[t]); rmax = v > rmax ? v : rmax; }
      for (int t = tid; t < mi; t += nt) { const double v = fabs(ri[t]); rmax = v > rmax ? v : rmax; comp += sv[t] * zv[t]; }
      qp_block_reduce3(rmax, dummy, comp, red);
      const double mu = mi > 0 ? comp / mi : 0.0;
      if (!(rmax == rmax)) { st = 5; break; }
      if (rmax <= d.tol * scl && mu <= d.tol) { st = 0; break; }
      if (it == d.max_iter) break;
      // ---- KKT matrix (lower triangle): [[H + eps I + Ain' (z/s) Ain, .], [Aeq, -eps I]]
      for (int e = tid; e < N * N; e += nt) {
        const int r = e / N, c = e - r * N;
        if (c > r) continue;
        double v;
        if (r < n) {
          v = Hq[(size_t)r * n + c] + (r == c ? d.eps : 0.0);
          for (int k = 0; k < mi; ++k) v += Gq[(size_t)k * n + r] * (zv[k] / sv[k]) * Gq[(size_t)k * n + c];
        } else if (c < n) v = Aq[(size_t)(r - n) * n + c];
        else v = (r == c) ? -d.eps : 0.0;
        K[(size_t)r * N + c] = v;
      }
      __syncthreads();
      // ---- LDL' (right-looking, no pivoting): column k scaled by 1/d_k, trailing update A(i,j) -= l_i d_k l_j
      bool okp = true;
      for (int k = 0; k < N; ++k) {
        double dk = K[(size_t)k * N + k];
        if (!(dk == dk)) { okp = false; break; }                             // uniform (same shared value for every thread)
        // dynamic regularisation: a pivot that lost its sign to cancellation (the wrench of a swing foot sits at the apex of
        // its cone, all eight rows active, barrier weights z/s ~ 1e12) is replaced by a tiny one of the right sign; the
        // step becomes inexact, the residuals of the next iteration are not
        if (k < n ? !(dk > d.piv) : !(dk < -d.piv)) dk = k < n ? d.piv : -d.piv;
        __syncthreads();                                                     // every thread has read the old pivot
        if (tid == 0) K[(size_t)k * N + k] = dk;
        const double inv = 1.0 / dk;
        const int m = N - k - 1;
        // trailing update first reads the unscaled column, so scale into a side buffer (rhs is free here)
        for (int t = tid; t < m; t += nt) rhs[t] = K[(size_t)(k + 1 + t) * N + k] * inv;
        __syncthreads();
        for (int e = tid; e < m * m; e += nt) {
          const int a = e / m, b = e - a * m;
          if (b > a) continue;
          K[(size_t)(k + 1 + a) * N + (k + 1 + b)] -= rhs[a] * K[(size_t)(k + 1 + b) * N + k];
        }
        __syncthreads();
        for (int t = tid; t < m; t += nt) K[(size_t)(k + 1 + t) * N + k] = rhs[t];
        __syncthreads();
      }
      if (!okp) { st = 3; break; }
      // solve K w = rhs in place (L unit lower, D diagonal), by warp 0 column by column would serialise; all threads per column instead
      auto kkt_solve = [&]() {
        for (int k = 0; k < N; ++k) {                     // forward: rhs_i -= L(i,k) rhs_k
          const double vk = rhs[k];
          for (int t = k + 1 + tid; t < N; t += nt) rhs[t] -= K[(size_t)t * N + k] * vk;
          __syncthreads();
        }
        for (int t = tid; t < N; t += nt) rhs[t] /= K[(size_t)t * N + t];
        __syncthreads();
        for (int k = N - 1; k >= 0; --k) {                // backward: rhs_j -= L(k,j) rhs_k
          const double vk = rhs[k];
          for (int t = tid; t < k; t += nt) rhs[t] -= K[(size_t)k * N + t] * vk;
          __syncthreads();
        }
      };
      // right-hand side for a complementarity target rc:  [ -r_d - Ain' ((rc + z r_i) / s) ; -r_p ]
      auto build_rhs = [&]() {
        for (int t = tid; t < n; t += nt) {
          double s = -rd[t];
          for (int k = 0; k < mi; ++k) s -= Gq[(size_t)k * n + t] * ((rc[k] + zv[k] * ri[k]) / sv[k]);
          rhs[t] = s;
        }
        for (int t = tid; t < me; t += nt) rhs[n + t] = -rd[n + t];
        __syncthreads();
      };
      // slack / multiplier steps of the solution in rhs, and the largest step that keeps (s, z) > 0
      auto slack_step = [&](double* ds, double* dz) -> double {
        double amin = 1e300, b = 0.0, c = 0.0;
        for (int t = tid; t < mi; t += nt) {
          double g = 0.0;
          for (int j = 0; j < n; ++j) g += Gq[(size_t)t * n + j] * rhs[j];
          const double dsv = -ri[t] - g, dzv = (rc[t] - zv[t] * dsv) / sv[t];
          ds[t] = dsv; dz[t] = dzv;
          if (dsv < 0.0) { const double a = -sv[t] / dsv; amin = a < amin ? a : amin; }
          if (dzv < 0.0) { const double a = -zv[t] / dzv; amin = a < amin ? a : amin; }
        }
        double mx = 0.0;
        qp_block_reduce3(mx, amin, c, red); (void)b;
        return amin;
      };
      // ---- predictor
      for (int t = tid; t < mi; t += nt) rc[t] = -sv[t] * zv[t];
      __syncthreads();
      build_rhs();
      kkt_solve();
      double a_aff = slack_step(dsa, dza);
      a_aff = a_aff < 1.0 ? a_aff : 1.0;
      double mu_aff = 0.0, d1 = 0.0, d2 = 0.0;
      for (int t = tid; t < mi; t += nt) mu_aff += (sv[t] + a_aff * dsa[t]) * (zv[t] + a_aff * dza[t]);
      qp_block_reduce3(d1, d2, mu_aff, red);
      mu_aff = mi > 0 ? mu_aff / mi : 0.0;
      const double sig = mu > 0.0 ? (mu_aff / mu) * (mu_aff / mu) * (mu_aff / mu) : 0.0;
      // ---- corrector
      for (int t = tid; t < mi; t += nt) rc[t] = sig * mu - sv[t] * zv[t] - dsa[t] * dza[t];
      __syncthreads();
      build_rhs();
      kkt_solve();
      double alpha = slack_step(dsa, dza);
      alpha = 0.995 * alpha; alpha = alpha < 1.0 ? alpha : 1.0;
      for (int t = tid; t < N; t += nt) xv[t] += alpha * rhs[t];
      for (int t = tid; t < mi; t += nt) { sv[t] += alpha * dsa[t]; zv[t] += alpha * dza[t]; }
      __syncthreads();
    }
    for (int t = tid; t < n; t += nt) xout[(size_t)q * n + t] = xv[t];
    if (tid == 0) { if (status) status[q] = st; if (iters) iters[q] = it; }
    __syncthreads();
  }
}

}  // namespace cmpc_qp
