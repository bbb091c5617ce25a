// cmpc_kernels.cu -- sm_100a kernels and the C ABI (include/cmpc.h) of the batched centroidal-MPC solver.
//
// One CTA per NLP instance; the instance's stage matrices live in shared memory, its iterate / factors in a
// device-resident workspace (kept across ticks for warm starts).  See cmpc_solver.h for the algorithm and
// cmpc_model.h for the NLP (both cite code/centroidal_mpc_vertices.py line by line).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>

#include "../../include/cmpc.h"
#include "cmpc_solver.h"

using namespace cmpc;

#ifndef CMPC_THREADS
#define CMPC_THREADS 128     // threads per instance (CTA size)
#endif
#ifndef CMPC_MIN_CTAS
#define CMPC_MIN_CTAS 3      // resident CTAs per SM the register allocation is sized for (168 registers; 4 CTAs at 128 registers measure 2 % slower: spills)
#endif

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  if (e != cudaSuccess) snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  else snprintf(g_err, sizeof(g_err), "%s", what);
  return code;
}

#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(-2, what, e_); } while (0)

extern __shared__ __align__(16) unsigned char cmpc_smem_raw[];     // the CTA's dynamic shared memory (one Smem)

struct ParCta {
  template <class T> __device__ T& smem() const { return *reinterpret_cast<T*>(cmpc_smem_raw); }
  __device__ void bind(void*) const {}
  __device__ int tid() const { return (int)threadIdx.x; }
  __device__ int nt() const { return (int)blockDim.x; }
  __device__ void sync() const { __syncthreads(); }
  __device__ int lane() const { return (int)(threadIdx.x & 31); }
  __device__ int warp() const { return (int)(threadIdx.x >> 5); }
  __device__ int nwarps() const { return (int)(blockDim.x >> 5); }
  __device__ int lanes() const { return 32; }
  __device__ void sync_warp() const { __syncwarp(); }
  // asynchronous global -> shared copy of n doubles, spread over the CTA (cp.async, 8 bytes per element)
  __device__ void copy_async(double* dst, const double* src, int n) const {
    for (int t = (int)threadIdx.x; t < n; t += (int)blockDim.x) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + t);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src + t) : "memory");
    }
  }
  __device__ void commit_async() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
  __device__ void wait_async() const { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
  static constexpr int TPT = (NTILE + CMPC_THREADS - 1) / CMPC_THREADS;      // 4x4 register tiles per thread: 120 tiles over the CTA
  static constexpr int CPT = 1;                                              // transient tiles of tile column 0 (16) per thread
};

struct Outputs {
  double *x1, *u0, *xN, *cost, *viol;
  int32_t *status, *iters;
  int32_t* counters;   // [B][2] nfact, nreg
  unsigned long long* prof;   // [PF_COUNT] phase cycles summed over CTAs (only with -DCMPC_PROFILE)
};

__global__ void __launch_bounds__(CMPC_THREADS, CMPC_MIN_CTAS)
cmpc_solve_kernel(Config c, int batch, const double* __restrict__ x0, const double* __restrict__ com_ref,
                  const double* __restrict__ foot_ref, const double* __restrict__ gamma,
                  const double* __restrict__ mass, const double* __restrict__ k1, double* work, size_t wstride,
                  int warm, Outputs out, const int32_t* __restrict__ perm, int32_t* __restrict__ last_iters) {
  Smem& sm = *reinterpret_cast<Smem*>(cmpc_smem_raw);
  const int N = c.N;
#ifdef CMPC_PROFILE
  if (threadIdx.x == 0) for (int k = 0; k < PF_COUNT; ++k) sm.prof[k] = 0;
  const long long cta_t0 = clock64();
#endif
  for (int slot = blockIdx.x; slot < batch; slot += gridDim.x) {
    const int b = perm ? perm[slot] : slot;          // longest-expected-first order (see cmpc_order_kernel)
    Instance in;
    in.x0 = x0 + (size_t)NXP * b;
    in.com_ref = com_ref + (size_t)9 * N * b;
    in.foot_ref = foot_ref + (size_t)8 * N * b;
    in.gamma = gamma + (size_t)2 * (N + 1) * b;
    in.mass = mass[b];
    in.k1 = k1[b];
    Work w = carve_work(work + wstride * b, N);
    ParCta par;
    Solver<ParCta> sol(c, in, w, sm, par);
    Stats st;
#ifdef CMPC_PROFILE
    const long long solve_t0 = clock64();
#endif
    sol.run(warm, &st);
    __syncthreads();
#ifdef CMPC_PROFILE
    if (threadIdx.x == 0) sm.prof[PF_SOLVE] += clock64() - solve_t0;
#endif
    const int tid = threadIdx.x;
    if (out.x1) for (int j = tid; j < NXP; j += blockDim.x) out.x1[(size_t)NXP * b + j] = w.X[NX + j];
    if (out.xN) for (int j = tid; j < NXP; j += blockDim.x) out.xN[(size_t)NXP * b + j] = w.X[N * NX + j];
    if (out.u0) for (int j = tid; j < NU; j += blockDim.x) out.u0[(size_t)NU * b + j] = w.U[j];
    if (tid == 0) {
      if (out.cost) out.cost[b] = st.cost;
      if (out.viol) out.viol[b] = st.viol;
      if (out.status) out.status[b] = st.status;
      if (out.iters) out.iters[b] = st.iters;
      if (out.counters) { out.counters[2 * b] = st.nfact; out.counters[2 * b + 1] = st.nreg; }
      if (last_iters) last_iters[b] = st.nfact;
    }
    __syncthreads();
  }
#ifdef CMPC_PROFILE
  if (threadIdx.x == 0) sm.prof[PF_CTA] = clock64() - cta_t0;
  if (threadIdx.x == 0 && out.prof) for (int k = 0; k < PF_COUNT; ++k) atomicAdd(&out.prof[k], (unsigned long long)sm.prof[k]);
#endif
}

// Launch order for the next tick: instances sorted by the work (factorisations) their previous solve needed, longest
// first.  CTAs are dispatched in block-index order, so the expensive instances start early and the cheap ones fill the
// tail (LPT scheduling); iteration counts of consecutive MPC ticks are strongly correlated.  A contact switch that has
// just entered the end of the horizon (schedule of the last three stages not constant) makes the previous solution a
// poor start for the new last stages: such instances are moved forward by a fixed bonus.  Counting sort, one CTA.
__device__ __forceinline__ int order_key(int b, int N, const int32_t* __restrict__ last_iters, const double* __restrict__ gamma) {
  int k = last_iters[b];
  if (gamma) {
    const double* g = gamma + (size_t)2 * (N + 1) * b + 2 * (N - 2);          // stages N-2, N-1, N
    if (N >= 2 && (g[0] != g[2] || g[1] != g[3] || g[2] != g[4] || g[3] != g[5])) k += 8;
  }
  return k < 0 ? 0 : (k > 511 ? 511 : k);
}

__global__ void cmpc_order_kernel(int batch, int N, const int32_t* __restrict__ last_iters, const double* __restrict__ gamma,
                                  int32_t* __restrict__ perm) {
  __shared__ int hist[512];
  for (int t = threadIdx.x; t < 512; t += blockDim.x) hist[t] = 0;
  __syncthreads();
  for (int b = threadIdx.x; b < batch; b += blockDim.x) atomicAdd(&hist[511 - order_key(b, N, last_iters, gamma)], 1);
  __syncthreads();
  if (threadIdx.x == 0) { int acc = 0; for (int t = 0; t < 512; ++t) { const int c = hist[t]; hist[t] = acc; acc += c; } }
  __syncthreads();
  for (int b = threadIdx.x; b < batch; b += blockDim.x) perm[atomicAdd(&hist[511 - order_key(b, N, last_iters, gamma)], 1)] = b;
}

// gather / scatter between the 28-wide internal state layout and the 20-wide reference layout
__global__ void cmpc_export_traj(int batch, int N, const double* work, size_t wstride, double* X, double* U) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  Work w = carve_work(const_cast<double*>(work) + wstride * b, N);
  for (int t = threadIdx.x; t < (N + 1) * NXP; t += blockDim.x) X[(size_t)b * (N + 1) * NXP + t] = w.X[(t / NXP) * NX + t % NXP];
  for (int t = threadIdx.x; t < N * NU; t += blockDim.x) U[(size_t)b * N * NU + t] = w.U[t];
}

__global__ void cmpc_import_traj(int batch, int N, double* work, size_t wstride, const double* X, const double* U) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  Work w = carve_work(work + wstride * b, N);
  for (int t = threadIdx.x; t < (N + 1) * NXP; t += blockDim.x) w.X[(t / NXP) * NX + t % NXP] = X[(size_t)b * (N + 1) * NXP + t];
  for (int t = threadIdx.x; t < N * NU; t += blockDim.x) w.U[t] = U[(size_t)b * N * NU + t];
  __syncthreads();
  // augmented states q_i = f_z of u_{i-1}
  for (int t = threadIdx.x; t < (N + 1) * NQ; t += blockDim.x) {
    const int i = t / NQ, v = t % NQ;
    w.X[i * NX + IQ + v] = (i >= 1) ? w.U[(i - 1) * NU + 3 * v + 2] : 0.0;
  }
}

__global__ void cmpc_fp64_probe(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, k = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, k); a1 = fma(a1, m, k); a2 = fma(a2, m, k); a3 = fma(a3, m, k);
    a4 = fma(a4, m, k); a5 = fma(a5, m, k); a6 = fma(a6, m, k); a7 = fma(a7, m, k);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace

struct cmpc_handle {
  Config cfg;
  int threads, cap, device;
  size_t wstride;            // doubles per instance
  double* work;
  // device staging for the host-buffer entry point
  double* d_in; double* d_out; int32_t* d_iout;
  double* h_in; double* h_out; int32_t* h_iout;
  int32_t* d_counters; int32_t* h_counters;
  int32_t* d_last_iters; int32_t* d_perm;     // work of the previous solve per instance, launch order of the next one
  unsigned long long* d_prof;
  size_t in_doubles, out_doubles;
  cudaStream_t stream;
  cudaEvent_t ev0, ev1;
  int last_batch, last_launches;
  size_t smem_bytes;         // dynamic shared memory per CTA (sizeof(Smem) + the occupancy-probe padding CMPC_SMEM_PAD)
  bool have_timing;
  int warm_valid;
  double* snap; int32_t* snap_iters; int snap_batch; size_t iter_doubles;   // snapshot of the warm-start part of the workspace
};

extern "C" {

const char* cmpc_last_error(void) { return g_err; }
const char* cmpc_version(void) { return "cmpc_b200 0.1 (sm_100a, fp64)"; }

int cmpc_default_config(int32_t N, cmpc_config* cfg) {
  if (!cfg || N < 1 || N > NMAX) return fail(-1, "cmpc_default_config: bad arguments (1 <= N <= 64)");
  Config c = default_config(N);
  memset(cfg, 0, sizeof(*cfg));
  cfg->N = N; cfg->max_iter = c.max_iter; cfg->ls_max = c.ls_max; cfg->threads = CMPC_THREADS;
  cfg->delta = c.delta; cfg->grav = c.grav; cfg->mu_fric = c.mu_fric;
  cfg->foot_half_len = c.hl; cfg->foot_half_wid = c.hw;
  cfg->w_h = c.w_h; cfg->w_xy = c.w_xy; cfg->w_zc = c.w_zc; cfg->w_foot = c.w_foot; cfg->w_sym = c.w_sym;
  cfg->w_swing = c.w_swing; cfg->w_rate = c.w_rate; cfg->eps_reg = c.eps_reg; cfg->pz_max = c.pz_max;
  for (int j = 0; j < 3; ++j) cfg->box[j] = c.box[j];
  cfg->relax = c.relax; cfg->mu_init = c.mu_init; cfg->mu_final = c.mu_final; cfg->mu_warm = c.mu_warm; cfg->tol = c.tol;
  cfg->kappa_eps = c.kappa_eps; cfg->kappa_mu = c.kappa_mu; cfg->theta_mu = c.theta_mu; cfg->tau_min = c.tau_min;
  cfg->bound_push = c.bound_push;
  return 0;
}

static Config to_internal(const cmpc_config* u) {
  Config c = default_config(u->N);
  c.max_iter = u->max_iter; c.ls_max = u->ls_max; c.delta = u->delta; c.grav = u->grav; c.mu_fric = u->mu_fric;
  c.hl = u->foot_half_len; c.hw = u->foot_half_wid; c.w_h = u->w_h; c.w_xy = u->w_xy; c.w_zc = u->w_zc;
  c.w_foot = u->w_foot; c.w_sym = u->w_sym; c.w_swing = u->w_swing; c.w_rate = u->w_rate; c.eps_reg = u->eps_reg;
  c.pz_max = u->pz_max; for (int j = 0; j < 3; ++j) c.box[j] = u->box[j];
  c.relax = u->relax; c.mu_init = u->mu_init; c.mu_final = u->mu_final; c.mu_warm = u->mu_warm; c.tol = u->tol;
  c.kappa_eps = u->kappa_eps; c.kappa_mu = u->kappa_mu; c.theta_mu = u->theta_mu; c.tau_min = u->tau_min;
  c.bound_push = u->bound_push;
  return c;
}

int cmpc_create(const cmpc_config* cfg, int32_t batch_capacity, int32_t device, cmpc_handle** out) {
  if (!cfg || !out || batch_capacity < 1) return fail(-1, "cmpc_create: bad arguments");
  if (cfg->N < 1 || cfg->N > NMAX) return fail(-1, "cmpc_create: horizon must satisfy 1 <= N <= 64");
  if (cfg->threads != CMPC_THREADS) return fail(-1, "cmpc_create: threads must be 128 (register-tile mapping of the stage factorisation)");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return fail(-3, "cmpc_create: no CUDA device (this library has no CPU path)", e);
  if (device < 0 || device >= ndev) return fail(-1, "cmpc_create: bad device index");
  CK(cudaSetDevice(device), "cudaSetDevice");
  cmpc_handle* h = new (std::nothrow) cmpc_handle();
  if (!h) return fail(-4, "cmpc_create: out of host memory");
  memset(h, 0, sizeof(*h));
  h->cfg = to_internal(cfg); h->threads = cfg->threads; h->cap = batch_capacity; h->device = device;
  const int N = cfg->N;
  h->wstride = (work_doubles(N) + 1) & ~(size_t)1;       // keep instances 16-byte aligned
  h->in_doubles = (size_t)NXP + 9 * N + 8 * N + 2 * (N + 1) + 2;
  h->out_doubles = (size_t)NXP + NU + NXP + 2;
  const size_t B = (size_t)batch_capacity;
  CK(cudaMalloc(&h->work, B * h->wstride * sizeof(double)), "cudaMalloc(work)");
  CK(cudaMemset(h->work, 0, B * h->wstride * sizeof(double)), "cudaMemset(work)");
  CK(cudaMalloc(&h->d_in, B * h->in_doubles * sizeof(double)), "cudaMalloc(d_in)");
  CK(cudaMalloc(&h->d_out, B * h->out_doubles * sizeof(double)), "cudaMalloc(d_out)");
  CK(cudaMalloc(&h->d_iout, B * 2 * sizeof(int32_t)), "cudaMalloc(d_iout)");
  CK(cudaMalloc(&h->d_counters, B * 2 * sizeof(int32_t)), "cudaMalloc(d_counters)");
  CK(cudaMalloc(&h->d_last_iters, B * sizeof(int32_t)), "cudaMalloc(d_last_iters)");
  CK(cudaMemset(h->d_last_iters, 0, B * sizeof(int32_t)), "cudaMemset(d_last_iters)");
  CK(cudaMalloc(&h->d_perm, B * sizeof(int32_t)), "cudaMalloc(d_perm)");
  CK(cudaMallocHost(&h->h_in, B * h->in_doubles * sizeof(double)), "cudaMallocHost(h_in)");
  CK(cudaMallocHost(&h->h_out, B * h->out_doubles * sizeof(double)), "cudaMallocHost(h_out)");
  CK(cudaMallocHost(&h->h_iout, B * 2 * sizeof(int32_t)), "cudaMallocHost(h_iout)");
  CK(cudaMallocHost(&h->h_counters, B * 2 * sizeof(int32_t)), "cudaMallocHost(h_counters)");
#ifdef CMPC_PROFILE
  CK(cudaMalloc(&h->d_prof, PF_COUNT * sizeof(unsigned long long)), "cudaMalloc(d_prof)");
#endif
  CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  CK(cudaEventCreate(&h->ev0), "cudaEventCreate");
  CK(cudaEventCreate(&h->ev1), "cudaEventCreate");
  h->smem_bytes = sizeof(Smem);
  if (const char* pad = getenv("CMPC_SMEM_PAD")) h->smem_bytes += (size_t)atol(pad);     // profiling aid: fewer resident CTAs per SM
  CK(cudaFuncSetAttribute(cmpc_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes),
     "cudaFuncSetAttribute(smem)");
#if CMPC_MIN_CTAS >= 4
  // more than three resident CTAs per SM need a shared-memory carve-out beyond the driver's default choice
  CK(cudaFuncSetAttribute(cmpc_solve_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared),
     "cudaFuncSetAttribute(carveout)");
#endif
  *out = h;
  return 0;
}

int cmpc_destroy(cmpc_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaFree(h->snap); cudaFree(h->snap_iters); cudaFree(h->work); cudaFree(h->d_in); cudaFree(h->d_out); cudaFree(h->d_iout); cudaFree(h->d_counters); cudaFree(h->d_last_iters); cudaFree(h->d_perm);
  cudaFreeHost(h->h_in); cudaFreeHost(h->h_out); cudaFreeHost(h->h_iout); cudaFreeHost(h->h_counters);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  delete h;
  return 0;
}

int cmpc_solve_device(cmpc_handle* h, int32_t batch, const double* x0, const double* com_ref,
                      const double* foot_ref, const double* gamma, const double* mass, const double* k1,
                      int32_t warm_mode, double* x1, double* u0, double* xN, double* cost, double* viol,
                      int32_t* status, int32_t* iters, void* stream) {
  if (!h) return fail(-1, "cmpc_solve_device: null handle");
  if (batch < 1 || batch > h->cap) return fail(-1, "cmpc_solve_device: batch exceeds the handle's capacity");
  if (!x0 || !com_ref || !foot_ref || !gamma || !mass || !k1) return fail(-1, "cmpc_solve_device: null input pointer");
  if (warm_mode < 0 || warm_mode > 3) return fail(-1, "cmpc_solve_device: bad warm_mode");
  if (warm_mode != CMPC_COLD && h->warm_valid < batch) warm_mode = CMPC_COLD;     // nothing to warm-start from
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  Outputs o{x1, u0, xN, cost, viol, status, iters, h->d_counters, h->d_prof};
  if (h->d_prof) CK(cudaMemsetAsync(h->d_prof, 0, PF_COUNT * sizeof(unsigned long long), s), "memset prof");
  CK(cudaEventRecord(h->ev0, s), "cudaEventRecord");
  const int32_t* perm = nullptr;
  int launches = 1;
  if (warm_mode != CMPC_COLD) {                       // the previous solve of these instances tells how expensive they are
    cmpc_order_kernel<<<1, 1024, 0, s>>>(batch, h->cfg.N, h->d_last_iters, gamma, h->d_perm);
    CK(cudaGetLastError(), "cmpc_order_kernel launch");
    perm = h->d_perm; launches = 2;
  }
  cmpc_solve_kernel<<<batch, h->threads, h->smem_bytes, s>>>(h->cfg, batch, x0, com_ref, foot_ref, gamma, mass, k1,
                                                            h->work, h->wstride, warm_mode, o, perm, h->d_last_iters);
  CK(cudaGetLastError(), "cmpc_solve_kernel launch");
  CK(cudaEventRecord(h->ev1, s), "cudaEventRecord");
  h->last_batch = batch; h->last_launches = launches; h->have_timing = true;
  h->warm_valid = batch;
  return 0;
}

int cmpc_solve_host(cmpc_handle* h, int32_t batch, const double* x0, const double* com_ref,
                    const double* foot_ref, const double* gamma, const double* mass, const double* k1,
                    int32_t warm_mode, double* x1, double* u0, double* xN, double* cost, double* viol,
                    int32_t* status, int32_t* iters) {
  if (!h) return fail(-1, "cmpc_solve_host: null handle");
  if (batch < 1 || batch > h->cap) return fail(-1, "cmpc_solve_host: batch exceeds the handle's capacity");
  if (!x0 || !com_ref || !foot_ref || !gamma || !mass || !k1) return fail(-1, "cmpc_solve_host: null input pointer");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const int N = h->cfg.N; const size_t B = (size_t)batch;
  // pack inputs into one pinned buffer -> one H2D copy
  double* p = h->h_in;
  double* hx0 = p; p += B * NXP; double* hcom = p; p += B * 9 * N; double* hfoot = p; p += B * 8 * N;
  double* hgam = p; p += B * 2 * (N + 1); double* hm = p; p += B; double* hk = p; p += B;
  memcpy(hx0, x0, B * NXP * sizeof(double)); memcpy(hcom, com_ref, B * 9 * N * sizeof(double));
  memcpy(hfoot, foot_ref, B * 8 * N * sizeof(double)); memcpy(hgam, gamma, B * 2 * (N + 1) * sizeof(double));
  memcpy(hm, mass, B * sizeof(double)); memcpy(hk, k1, B * sizeof(double));
  const size_t nin = (size_t)(p - h->h_in);
  CK(cudaMemcpyAsync(h->d_in, h->h_in, nin * sizeof(double), cudaMemcpyHostToDevice, h->stream), "H2D");
  double* d = h->d_in;
  double* dx0 = d; d += B * NXP; double* dcom = d; d += B * 9 * N; double* dfoot = d; d += B * 8 * N;
  double* dgam = d; d += B * 2 * (N + 1); double* dm = d; d += B; double* dk = d;
  double* q = h->d_out;
  double* dx1 = q; q += B * NXP; double* du0 = q; q += B * NU; double* dxN = q; q += B * NXP;
  double* dcost = q; q += B; double* dviol = q; q += B;
  const size_t nout = (size_t)(q - h->d_out);
  int rc = cmpc_solve_device(h, batch, dx0, dcom, dfoot, dgam, dm, dk, warm_mode, dx1, du0, dxN, dcost, dviol,
                             h->d_iout, h->d_iout + B, h->stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->h_out, h->d_out, nout * sizeof(double), cudaMemcpyDeviceToHost, h->stream), "D2H");
  CK(cudaMemcpyAsync(h->h_iout, h->d_iout, B * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream), "D2H");
  CK(cudaStreamSynchronize(h->stream), "cmpc_solve_host: kernel execution");
  const double* r = h->h_out;
  if (x1) memcpy(x1, r, B * NXP * sizeof(double)); r += B * NXP;
  if (u0) memcpy(u0, r, B * NU * sizeof(double)); r += B * NU;
  if (xN) memcpy(xN, r, B * NXP * sizeof(double)); r += B * NXP;
  if (cost) memcpy(cost, r, B * sizeof(double)); r += B;
  if (viol) memcpy(viol, r, B * sizeof(double));
  if (status) memcpy(status, h->h_iout, B * sizeof(int32_t));
  if (iters) memcpy(iters, h->h_iout + B, B * sizeof(int32_t));
  return 0;
}

int cmpc_get_trajectory(cmpc_handle* h, int32_t batch, double* X, double* U) {
  if (!h || !X || !U || batch < 1 || batch > h->cap) return fail(-1, "cmpc_get_trajectory: bad arguments");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const int N = h->cfg.N; const size_t B = (size_t)batch;
  double *dX, *dU;
  CK(cudaMalloc(&dX, B * (N + 1) * NXP * sizeof(double)), "cudaMalloc");
  CK(cudaMalloc(&dU, B * N * NU * sizeof(double)), "cudaMalloc");
  cmpc_export_traj<<<batch, 128, 0, h->stream>>>(batch, N, h->work, h->wstride, dX, dU);
  CK(cudaGetLastError(), "cmpc_export_traj launch");
  CK(cudaMemcpyAsync(X, dX, B * (N + 1) * NXP * sizeof(double), cudaMemcpyDeviceToHost, h->stream), "D2H");
  CK(cudaMemcpyAsync(U, dU, B * N * NU * sizeof(double), cudaMemcpyDeviceToHost, h->stream), "D2H");
  CK(cudaStreamSynchronize(h->stream), "cmpc_get_trajectory");
  cudaFree(dX); cudaFree(dU);
  return 0;
}

int cmpc_set_warm(cmpc_handle* h, int32_t batch, const double* X, const double* U) {
  if (!h || !X || !U || batch < 1 || batch > h->cap) return fail(-1, "cmpc_set_warm: bad arguments");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const int N = h->cfg.N; const size_t B = (size_t)batch;
  double *dX, *dU;
  CK(cudaMalloc(&dX, B * (N + 1) * NXP * sizeof(double)), "cudaMalloc");
  CK(cudaMalloc(&dU, B * N * NU * sizeof(double)), "cudaMalloc");
  CK(cudaMemcpyAsync(dX, X, B * (N + 1) * NXP * sizeof(double), cudaMemcpyHostToDevice, h->stream), "H2D");
  CK(cudaMemcpyAsync(dU, U, B * N * NU * sizeof(double), cudaMemcpyHostToDevice, h->stream), "H2D");
  cmpc_import_traj<<<batch, 128, 0, h->stream>>>(batch, N, h->work, h->wstride, dX, dU);
  CK(cudaGetLastError(), "cmpc_import_traj launch");
  CK(cudaStreamSynchronize(h->stream), "cmpc_set_warm");
  cudaFree(dX); cudaFree(dU);
  h->warm_valid = batch;
  return 0;
}

int cmpc_warm_save(cmpc_handle* h, int32_t batch) {
  if (!h || batch < 1 || batch > h->cap) return fail(-1, "cmpc_warm_save: bad arguments");
  if (h->warm_valid < batch) return fail(-1, "cmpc_warm_save: no warm-start state for that many instances");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const int N = h->cfg.N;
  h->iter_doubles = (size_t)(N + 1) * NX * 2 + (size_t)N * NU + (size_t)(N + 1) * NR * 2;   // X, U, Y, S, LAM
  if (!h->snap) CK(cudaMalloc(&h->snap, (size_t)h->cap * h->iter_doubles * sizeof(double)), "cudaMalloc(snapshot)");
  if (!h->snap_iters) CK(cudaMalloc(&h->snap_iters, (size_t)h->cap * sizeof(int32_t)), "cudaMalloc(snapshot iters)");
  CK(cudaMemcpyAsync(h->snap_iters, h->d_last_iters, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream), "snapshot iters");
  CK(cudaMemcpy2DAsync(h->snap, h->iter_doubles * sizeof(double), h->work, h->wstride * sizeof(double),
                       h->iter_doubles * sizeof(double), batch, cudaMemcpyDeviceToDevice, h->stream), "snapshot copy");
  CK(cudaStreamSynchronize(h->stream), "cmpc_warm_save");
  h->snap_batch = batch;
  return 0;
}

int cmpc_warm_restore(cmpc_handle* h, int32_t batch, void* stream) {
  if (!h || batch < 1 || batch > h->snap_batch) return fail(-1, "cmpc_warm_restore: no snapshot for that many instances");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  CK(cudaMemcpy2DAsync(h->work, h->wstride * sizeof(double), h->snap, h->iter_doubles * sizeof(double),
                       h->iter_doubles * sizeof(double), batch, cudaMemcpyDeviceToDevice, s), "snapshot restore");
  CK(cudaMemcpyAsync(h->d_last_iters, h->snap_iters, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToDevice, s), "snapshot restore iters");
  h->warm_valid = batch;
  return 0;
}

int cmpc_reset_warm(cmpc_handle* h) {
  if (!h) return fail(-1, "cmpc_reset_warm: null handle");
  h->warm_valid = 0;
  return 0;
}

int cmpc_last_stats(cmpc_handle* h, int64_t* iters, int64_t* nfact, int64_t* nreg, double* kernel_ms, int32_t* launches) {
  if (!h || h->last_batch < 1) return fail(-1, "cmpc_last_stats: no solve yet");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const size_t B = (size_t)h->last_batch;
  CK(cudaEventSynchronize(h->ev1), "cudaEventSynchronize");
  CK(cudaMemcpy(h->h_counters, h->d_counters, B * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost), "D2H counters");
  int64_t nf = 0, nr = 0;
  for (size_t b = 0; b < B; ++b) { nf += h->h_counters[2 * b]; nr += h->h_counters[2 * b + 1]; }
  if (nfact) *nfact = nf;
  if (nreg) *nreg = nr;
  if (iters) *iters = nf - nr;
  if (kernel_ms) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, h->ev0, h->ev1), "cudaEventElapsedTime"); *kernel_ms = ms; }
  if (launches) *launches = h->last_launches;
  return 0;
}

/* Phase cycle counters of the last solve (all zeros unless built with -DCMPC_PROFILE), 11 values: eval, assemble,
 * P[B A] products, factorisation, factor store, forward sweep, slack steps, line-search trials, step, whole solves,
 * CTA lifetimes. */
int cmpc_phase_cycles(cmpc_handle* h, uint64_t* out11) {
  if (!h || !out11) return fail(-1, "cmpc_phase_cycles: bad arguments");
  for (int k = 0; k < PF_COUNT; ++k) out11[k] = 0;
  if (h->d_prof) { CK(cudaSetDevice(h->device), "cudaSetDevice"); CK(cudaMemcpy(out11, h->d_prof, PF_COUNT * sizeof(uint64_t), cudaMemcpyDeviceToHost), "D2H prof"); }
  return 0;
}

int cmpc_footprint(const cmpc_handle* h, size_t* work_bytes_per_instance, size_t* smem_bytes_per_cta) {
  if (!h) return fail(-1, "cmpc_footprint: null handle");
  if (work_bytes_per_instance) *work_bytes_per_instance = h->wstride * sizeof(double);
  if (smem_bytes_per_cta) *smem_bytes_per_cta = sizeof(Smem);
  return 0;
}

int cmpc_measure_fp64_peak(int32_t device, double* tflops) {
  if (!tflops) return fail(-1, "cmpc_measure_fp64_peak: null pointer");
  CK(cudaSetDevice(device), "cudaSetDevice");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
  double* d;
  CK(cudaMalloc(&d, (size_t)blocks * threads * sizeof(double)), "cudaMalloc");
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a), "event"); CK(cudaEventCreate(&b), "event");
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(a), "record");
    cmpc_fp64_probe<<<blocks, threads>>>(d, iters);
    CK(cudaEventRecord(b), "record");
    CK(cudaEventSynchronize(b), "cmpc_fp64_probe");
    float ms; CK(cudaEventElapsedTime(&ms, a, b), "elapsed");
    if (ms < best) best = ms;
  }
  *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
  cudaFree(d); cudaEventDestroy(a); cudaEventDestroy(b);
  return 0;
}

}  // extern "C"
