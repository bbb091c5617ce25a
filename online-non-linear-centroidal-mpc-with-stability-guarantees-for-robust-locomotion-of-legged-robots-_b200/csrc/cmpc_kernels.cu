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
#include "cmpc_qp.cuh"

using namespace cmpc;

#ifndef CMPC_THREADS
#define CMPC_THREADS 128     // threads per instance (CTA size)
#endif
#ifndef CMPC_CP16
#define CMPC_CP16 1
#endif
#ifndef CMPC_BULK
#define CMPC_BULK 0     // 1: stage records / gains by cp.async.bulk + mbarrier (one issuing thread, 1-D TMA) instead of per-thread cp.async.
                        // Built, parity-green (all GPU tests), measured on one box: 56.5 / 56.6 ms per bench step against 55.6 / 55.6 for the
                        // 16-byte cp.async.cg path -- the blocks are 2-8 KB, prefetched one stage ahead, and the issuing thread adds a proxy
                        // fence and three serial issues to a stage that is latency bound; the default stays cp.async
#endif
#ifndef CMPC_MIN_CTAS
#define CMPC_MIN_CTAS 4      // resident CTAs per SM the register allocation is sized for (128 registers, 20 bytes of spills; measured +3 % over 3 CTAs at 168 registers; shared memory allows no fifth)
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

#ifndef CMPC_PREFETCH
#define CMPC_PREFETCH 1
#endif

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
  __device__ bool sync_and(bool p) const { return __syncthreads_and(p ? 1 : 0) != 0; }     // barrier + block vote
  __device__ bool any(bool p) const { return __any_sync(0xffffffffu, p) != 0; }            // warp vote
  // warp all-reduce by xor shuffles (every lane ends with the result)
  __device__ double wmax(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
  }
  __device__ double wmin(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
    return v;
  }
  __device__ double wsum(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  // asynchronous global -> shared copy of n doubles, spread over the CTA: 16-byte cp.async.cg (L2 only: the scratch streams
  // stay out of the L1 the local-memory traffic lives in).  Every block copied here has an even length and starts at an even
  // offset of a 16-byte aligned base (records, gains; static_asserts next to the offsets)
  __device__ void copy_async(double* dst, const double* src, int n) const {
#if CMPC_BULK
    // one thread hands the whole block to the copy engine (cp.async.bulk, 1-D TMA): no per-thread copy instructions; completion
    // is counted in bytes on the CTA's transaction barrier
    if (threadIdx.x == 0) {
      Smem& sm = smem<Smem>();
      const unsigned mb = (unsigned)__cvta_generic_to_shared(&sm.mbar), sa = (unsigned)__cvta_generic_to_shared(dst);
      const unsigned bytes = (unsigned)n * 8u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // earlier generic reads of the destination come first
      asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sa), "l"(src), "r"(bytes), "r"(mb) : "memory");
    }
#elif CMPC_CP16
    for (int t = 2 * (int)threadIdx.x; t < n; t += 2 * (int)blockDim.x) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + t);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + t) : "memory");
    }
#else
    for (int t = (int)threadIdx.x; t < n; t += (int)blockDim.x) {
      const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + t);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src + t) : "memory");
    }
#endif
  }
  // L2 prefetch of n doubles (one request per 128-byte line, spread over the CTA)
  __device__ void prefetch_l2(const double* p, int n) const {
#if CMPC_PREFETCH
    for (int t = (int)threadIdx.x * 16; t < n; t += (int)blockDim.x * 16) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + t));
#endif
  }
#if CMPC_BULK
  // a group = the copies issued since the last commit: thread 0 arrives on the barrier (phase g completes when the bytes of
  // group g have landed).  Every thread counts the groups committed and the groups it has waited for (the control flow around
  // the copies is uniform over the CTA, so the counts agree without being shared; a shared counter read before the issuing
  // thread has bumped it would make a fast thread wait for the parity of a phase that is still to come -- a deadlock), waits for
  // the parity of the latest group, and returns at once when nothing new has been committed
  mutable unsigned issued = 0, waited = 0;
  __device__ void commit_async() const {
    if (threadIdx.x == 0) {
      const unsigned mb = (unsigned)__cvta_generic_to_shared(&smem<Smem>().mbar);
      asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(mb) : "memory");
    }
    ++issued;
  }
  __device__ void wait_async() const {
    if (waited == issued) return;
    const unsigned mb = (unsigned)__cvta_generic_to_shared(&smem<Smem>().mbar), parity = (issued - 1u) & 1u;
    unsigned done;
    do {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(mb), "r"(parity) : "memory");
    } while (!done);
    waited = issued;
  }
  // once per CTA, before the first copy
  __device__ void async_setup() const {
    if (threadIdx.x == 0) {
      const unsigned mb = (unsigned)__cvta_generic_to_shared(&smem<Smem>().mbar);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    issued = 0; waited = 0;
    __syncthreads();
  }
#else
  __device__ void commit_async() const { asm volatile("cp.async.commit_group;" ::: "memory"); }
  __device__ void wait_async() const { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
  __device__ void async_setup() const {}
#endif
  static constexpr int TPT = (NTILE + CMPC_THREADS - 1) / CMPC_THREADS;      // 4x4 register tiles per thread: 120 tiles over the CTA
  static constexpr int CPT = 1;                                              // transient tiles of tile column 0 (16) per thread
#ifndef CMPC_GAINS4
#define CMPC_GAINS4 1
#endif
  static constexpr bool GAINS4 = CMPC_GAINS4 != 0 && CMPC_THREADS == 128;      // gain back substitution on four lanes per right-hand side
  __device__ double shfl4(double v, int src) const { return __shfl_sync(0xffffffffu, v, src, 4); }
  __device__ double shfl4x(double v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m, 4); }
};

struct Outputs {
  double *x1, *u0, *xN, *cost, *viol;
  int32_t *status, *iters;
  int32_t* counters;   // [B][2] nfact, nreg
  unsigned long long* prof;   // [PF_COUNT] phase cycles summed over CTAs (only with -DCMPC_PROFILE)
};

// One launch solves a tick of the whole batch, retries included.  Persistent CTAs (one per resident slot: SMs x CTAs per
// SM) pull work from a queue in global memory: the retry ring first, then the instances in launch order (atomic cursor
// `next`).  The scratch (Newton step, derivative records, stage factors) belongs to the CTA slot, only the iterate belongs
// to the instance.  An instance whose attempt does not converge is appended to the ring with its next attempt -- the
// solver's own cold start (only after a warm attempt), then a ten times larger initial barrier value, then another
// starting point (CoM states blended from x0 towards the reference) -- and is picked up by the next free slot: the
// retries of the few hard instances overlap with the tail of the batch instead of running behind it.  Nobody waits: a slot
// that finds no work leaves; a slot that queues a retry loops once more and takes it itself unless another one was faster.
struct Queue { int32_t next, head, tail, done; };      // cursor of the initial list; retry ring head / tail; resolved instances (statistics)

__device__ __forceinline__ int32_t ld_volatile(const int32_t* p) { return *reinterpret_cast<const volatile int32_t*>(p); }

__global__ void __launch_bounds__(CMPC_THREADS, CMPC_MIN_CTAS)
cmpc_solve_kernel(Config c, int batch, const double* __restrict__ x0, const double* __restrict__ com_ref,
                  const double* __restrict__ foot_ref, const double* __restrict__ gamma,
                  const double* __restrict__ mass, const double* __restrict__ k1, double* iter, size_t istride,
                  double* scratch, size_t sstride, int warm, Outputs out, const int32_t* __restrict__ order, Queue* q, int32_t* ring,
                  int32_t* __restrict__ last_iters, int32_t* __restrict__ valid) {
  Smem& sm = *reinterpret_cast<Smem*>(cmpc_smem_raw);
  const int N = c.N;
#ifdef CMPC_PROFILE
  if (threadIdx.x == 0) for (int k = 0; k < PF_COUNT; ++k) sm.prof[k] = 0;
  const long long cta_t0 = clock64();
#endif
  double* my_scratch = scratch + sstride * blockIdx.x;
  ParCta par;                                                    // (one per CTA: it counts the copy groups of the transaction barrier)
  par.async_setup();
  for (;;) {
    if (threadIdx.x == 0) {
      int item = -1;
      // the retry ring first: a retry is a long job (a cold solve after a failed warm attempt) and must not wait for the
      // initial list to drain -- it would then run alone behind the batch.  No waiting: a slot that finds the ring empty and
      // the list drained leaves -- whoever queues a retry later comes back through this loop itself and will find its own
      // entry, so every entry is taken by somebody
      for (;;) {
        const int h = ld_volatile(&q->head), t = ld_volatile(&q->tail);
        if (h >= t) break;
        if (atomicCAS(&q->head, h, h + 1) != h) continue;         // another slot took it
        int e;
        while ((e = ld_volatile(ring + h)) < 0) { }               // (the producer publishes the entry right after reserving it)
        item = e;
        break;
      }
      if (item < 0) {
        const int i = ld_volatile(&q->next) < batch ? atomicAdd(&q->next, 1) : batch;
        if (i < batch) item = order ? order[i] : i;               // attempt 0, longest-expected-first order (cmpc_order_kernel)
      }
      sm.flag = item;
    }
    __syncthreads();
    const int item = sm.flag;
    __syncthreads();
    if (item < 0) break;
    const int b = item & 0x0fffffff, attempt = (item >> 28) & 7;
    Instance in;
    in.x0 = x0 + (size_t)NXP * b;
    in.com_ref = com_ref + (size_t)9 * N * b;
    in.foot_ref = foot_ref + (size_t)8 * N * b;
    in.gamma = gamma + (size_t)2 * (N + 1) * b;
    in.mass = mass[b];
    in.k1 = k1[b];
    Work w = carve_work(iter + istride * b, my_scratch, N);
    int wm = 0;
    if (attempt == 0) wm = (warm != 0 && valid && !valid[b]) ? 0 : warm;        // no previous solution of this instance: cold
    Solver<ParCta> sol(c, in, w, sm, par);
    Stats st;
#ifdef CMPC_PROFILE
    const long long solve_t0 = clock64();
#endif
    sol.run_pass(wm, attempt == 2 ? 10.0 : 1.0, &st, attempt == 3);
    __syncthreads();
#ifdef CMPC_PROFILE
    if (threadIdx.x == 0) sm.prof[PF_SOLVE] += clock64() - solve_t0;
#endif
    const int tid = threadIdx.x;
    const bool failed = st.status != ST_CONVERGED && st.status != ST_INFEASIBLE_X0;
    const int next_attempt = (attempt == 0) ? (wm != 0 ? 1 : 2) : attempt + 1;
    const bool retry = failed && next_attempt <= 3;
    const bool accumulate = attempt > 0;
    if (!retry) {
      if (out.x1) for (int j = tid; j < NXP; j += blockDim.x) out.x1[(size_t)NXP * b + j] = w.X[NX + j];
      if (out.xN) for (int j = tid; j < NXP; j += blockDim.x) out.xN[(size_t)NXP * b + j] = w.X[N * NX + j];
      if (out.u0) for (int j = tid; j < NU; j += blockDim.x) out.u0[(size_t)NU * b + j] = w.U[j];
    }
    if (tid == 0) {
      if (out.cost) out.cost[b] = st.cost;
      if (out.viol) out.viol[b] = st.viol;
      if (out.status) out.status[b] = st.status;
      if (out.iters) out.iters[b] = (accumulate ? out.iters[b] : 0) + st.iters;
      if (out.counters) {
        out.counters[2 * b] = (accumulate ? out.counters[2 * b] : 0) + st.nfact;
        out.counters[2 * b + 1] = (accumulate ? out.counters[2 * b + 1] : 0) + st.nreg;
      }
      if (last_iters) last_iters[b] = (accumulate ? last_iters[b] : 0) + st.nfact;
      if (valid) valid[b] = 1;
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();                                            // this attempt's writes are visible before the instance is handed on / counted
      if (retry) { const int slot = atomicAdd(&q->tail, 1); ring[slot] = b | (next_attempt << 28); __threadfence(); }
      else atomicAdd(&q->done, 1);
    }
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
__global__ void cmpc_export_traj(int batch, int N, const double* iter, size_t istride, double* X, double* U) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  Work w = carve_work(const_cast<double*>(iter) + istride * b, nullptr, N);
  for (int t = threadIdx.x; t < (N + 1) * NXP; t += blockDim.x) X[(size_t)b * (N + 1) * NXP + t] = w.X[(t / NXP) * NX + t % NXP];
  for (int t = threadIdx.x; t < N * NU; t += blockDim.x) U[(size_t)b * N * NU + t] = w.U[t];
}

__global__ void cmpc_import_traj(int batch, int N, double* iter, size_t istride, const double* X, const double* U, int32_t* valid) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  if (threadIdx.x == 0 && valid) valid[b] = 1;
  Work w = carve_work(iter + istride * b, nullptr, N);
  for (int t = threadIdx.x; t < (N + 1) * NXP; t += blockDim.x) w.X[(t / NXP) * NX + t % NXP] = X[(size_t)b * (N + 1) * NXP + t];
  for (int t = threadIdx.x; t < N * NU; t += blockDim.x) w.U[t] = U[(size_t)b * N * NU + t];
  __syncthreads();
  // augmented states q_i = f_z of u_{i-1}
  for (int t = threadIdx.x; t < (N + 1) * NQ; t += blockDim.x) {
    const int i = t / NQ, v = t % NQ;
    w.X[i * NX + IQ + v] = (i >= 1) ? w.U[(i - 1) * NU + 3 * v + 2] : 0.0;
  }
}

// Per-tick parameter assembly of `centroidal_mpc.solve` (MPC file :482-600) for a batch of robots, each at its own tick:
// one CTA per robot gathers its x0, CoM / foot references, yaw references (with the column-major quirk of :599-600) and
// contact schedule from the walk's tables.  Pure gather: 8 (19 N + 22) bytes written per robot, coalesced per section.
__global__ void cmpc_assemble_kernel(cmpc_walk_tables tb, int batch, const int32_t* __restrict__ tick, const double* __restrict__ com_pos,
                                     const double* __restrict__ com_vel, const double* __restrict__ hw, const double* __restrict__ theta,
                                     const double* __restrict__ yaw, const double* __restrict__ plan, double* __restrict__ x0,
                                     double* __restrict__ com_ref, double* __restrict__ foot_ref, double* __restrict__ gamma, int32_t* __restrict__ err) {
  const int b = blockIdx.x;
  if (b >= batch) return;
  const int N = tb.N, rate = tb.rate, t = tick[b];
  int e = 0;
  if (t < 0 || t + N * rate >= tb.T_ref) e = 1;                 // the reference's IndexError (:567)
  else if (t + N * rate >= tb.T_plan) e = 2;                   // beyond the footstep plan (get_step_index_at_time returns None)
  if (threadIdx.x == 0 && err) err[b] = e;
  if (e) return;
  for (int j = threadIdx.x; j < NXP; j += blockDim.x) {
    double v;
    if (j < 3) v = com_pos[3 * b + j];
    else if (j < 6) v = com_vel[3 * b + j - 3];
    else if (j < 9) v = hw[3 * b + j - 6];
    else if (j < 12) v = theta[3 * b + j - 9];
    else if (j == IPSL) v = yaw[2 * b];
    else if (j == IPSR) v = yaw[2 * b + 1];
    else {
      const int right = j >= IPR, ax = j - (right ? IPR : IPL);
      if (t < 200) v = tb.foot_tab[(size_t)8 * t + 3 * right + ax];                                  // :493-495
      else {                                                                                         // :496-503
        const int index = tb.step_index[t - 70];
        const int a_ = index + (index % 2), b_ = index + (((index - 1) % 2 + 2) % 2);
        const int il = tb.first_swing_left ? a_ : b_, ir = tb.first_swing_left ? b_ : a_;
        v = plan[((size_t)b * tb.n_steps + (right ? ir : il)) * 3 + ax];
      }
    }
    x0[(size_t)NXP * b + j] = v;
  }
  for (int q = threadIdx.x; q < N * 9; q += blockDim.x) {
    const int i = q / 9, j = q - 9 * i;
    com_ref[(size_t)9 * N * b + q] = tb.com_tab[(size_t)9 * (t + (1 + i) * rate) + j];              // :565-575
  }
  for (int q = threadIdx.x; q < N * 8; q += blockDim.x) {
    const int i = q >> 3, j = q & 7;
    const int tt = (j < 6) ? t + (1 + i) * rate : t + (1 + i / 3) * rate;                             // yaw: column-major quirk (:599-600)
    foot_ref[(size_t)8 * N * b + q] = tb.foot_tab[(size_t)8 * tt + j];
  }
  for (int q = threadIdx.x; q < 2 * (N + 1); q += blockDim.x)
    gamma[(size_t)2 * (N + 1) * b + q] = tb.gamma_tab[(size_t)2 * (t + (q >> 1) * rate) + (q & 1)];  // :517-531
}

// forget the warm-start state of the flagged instances
__global__ void cmpc_clear_valid(int n, const uint8_t* __restrict__ mask, int32_t* __restrict__ valid) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < n && (!mask || mask[b])) valid[b] = 0;
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
  int slots;                 // resident CTA slots of the solve kernel on this GPU (SMs x CTAs per SM)
  size_t istride, sstride;   // doubles: iterate per instance, scratch per slot
  double* iter;              // [cap][istride]   device resident across ticks (warm starts)
  double* scratch;           // [slots][sstride] Newton step, derivative records, stage factors of the running solves
  // device staging for the host-buffer entry point
  double* d_in; double* d_out; int32_t* d_iout;
  double* h_in; double* h_out; int32_t* h_iout;
  int32_t* d_counters; int32_t* h_counters;
  int32_t* d_last_iters; int32_t* d_perm;     // work of the previous solve per instance, launch order of the next one
  int32_t* d_valid;                           // per instance: a previous solution exists (warm start possible)
  int32_t* d_queue;                           // work queue of a launch (struct Queue)
  int32_t* d_fail;                            // [3 cap] retry ring: instance | attempt << 28
  uint8_t* d_mask;                            // staging of cmpc_reset_warm's mask
  double* d_traj; double* h_traj; size_t traj_cap;   // trajectory staging (allocated at first use, kept)
  unsigned long long* d_prof;
  size_t in_doubles, out_doubles, traj_doubles;
  cudaStream_t stream;       // the handle's own stream (used when the caller passes stream == NULL)
  cudaStream_t last_stream;  // stream of the last operation on this handle ...
  cudaEvent_t ev_done;       // ... and the event recorded behind it: operations on another stream wait for it
  bool have_done;
  cudaEvent_t ev0, ev1;
  int last_batch, last_launches;
  size_t smem_bytes;         // dynamic shared memory per CTA (sizeof(Smem) + the occupancy-probe padding CMPC_SMEM_PAD)
  bool have_timing;
  double* snap; int32_t* snap_iters; int32_t* snap_valid; int snap_batch;   // snapshot of the warm-start state
};

namespace {

// Every operation on a handle is ordered after the previous one, whatever streams the two were issued on.
int order_after_last(cmpc_handle* h, cudaStream_t s) {
  if (h->have_done && h->last_stream != s) CK(cudaStreamWaitEvent(s, h->ev_done, 0), "cudaStreamWaitEvent");
  return 0;
}
int mark_done(cmpc_handle* h, cudaStream_t s) {
  CK(cudaEventRecord(h->ev_done, s), "cudaEventRecord");
  h->last_stream = s; h->have_done = true;
  return 0;
}

int ensure_traj(cmpc_handle* h) {
  if (h->d_traj) return 0;
  const size_t n = (size_t)h->cap * h->traj_doubles;
  CK(cudaMalloc(&h->d_traj, n * sizeof(double)), "cudaMalloc(trajectory staging)");
  CK(cudaMallocHost(&h->h_traj, n * sizeof(double)), "cudaMallocHost(trajectory staging)");
  return 0;
}

}  // namespace

extern "C" {

const char* cmpc_last_error(void) { return g_err; }
const char* cmpc_version(void) { return "cmpc_b200 0.2 (sm_100a, fp64)"; }

int cmpc_default_config(int32_t N, cmpc_config* cfg) {
  if (!cfg || N < 1 || N > NMAX) return fail(-1, "cmpc_default_config: bad arguments (1 <= N <= 64)");
  Config c = default_config(N);
  memset(cfg, 0, sizeof(*cfg));
  cfg->N = N; cfg->max_iter = c.max_iter; cfg->ls_max = c.ls_max; cfg->threads = CMPC_THREADS; cfg->stall_window = c.stall_window; cfg->stall_final = c.stall_final; cfg->jam_window = c.jam_window;
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
  c.max_iter = u->max_iter; c.ls_max = u->ls_max; c.stall_window = u->stall_window; c.stall_final = u->stall_final; c.jam_window = u->jam_window;
  c.delta = u->delta; c.grav = u->grav; c.mu_fric = u->mu_fric;
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
  h->istride = (iter_doubles(N) + 1) & ~(size_t)1;       // keep blocks 16-byte aligned
  h->sstride = (scratch_doubles(N) + 1) & ~(size_t)1;
  h->in_doubles = (size_t)NXP + 9 * N + 8 * N + 2 * (N + 1) + 2;
  h->out_doubles = (size_t)NXP + NU + NXP + 2;
  h->traj_doubles = (size_t)(N + 1) * NXP + (size_t)N * NU;
  const size_t B = (size_t)batch_capacity;
  int rc = 0;
  // (every failure below releases what has been allocated so far)
#define CKC(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { rc = fail(-2, what, e_); cmpc_destroy(h); return rc; } } while (0)
  h->smem_bytes = sizeof(Smem);
  if (const char* pad = getenv("CMPC_SMEM_PAD")) h->smem_bytes += (size_t)atol(pad);     // profiling aid: fewer resident CTAs per SM
  CKC(cudaFuncSetAttribute(cmpc_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes),
      "cudaFuncSetAttribute(smem)");
#if CMPC_MIN_CTAS >= 4
  // more than three resident CTAs per SM need a shared-memory carve-out beyond the driver's default choice
  CKC(cudaFuncSetAttribute(cmpc_solve_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared),
      "cudaFuncSetAttribute(carveout)");
#endif
  {
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    int per_sm = 0;
    CKC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cmpc_solve_kernel, h->threads, h->smem_bytes), "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    if (per_sm < 1) per_sm = 1;
    h->slots = per_sm * prop.multiProcessorCount;
    if (const char* sl = getenv("CMPC_SLOTS")) { const int v = atoi(sl); if (v > 0) h->slots = v; }          // experiments
  }
  const size_t nslots = (size_t)(h->slots < batch_capacity ? h->slots : batch_capacity);
  CKC(cudaMalloc(&h->iter, B * h->istride * sizeof(double)), "cudaMalloc(iterates)");
  CKC(cudaMemset(h->iter, 0, B * h->istride * sizeof(double)), "cudaMemset(iterates)");
  CKC(cudaMalloc(&h->scratch, nslots * h->sstride * sizeof(double)), "cudaMalloc(scratch)");
  CKC(cudaMemset(h->scratch, 0, nslots * h->sstride * sizeof(double)), "cudaMemset(scratch)");
  CKC(cudaMalloc(&h->d_in, B * h->in_doubles * sizeof(double)), "cudaMalloc(d_in)");
  CKC(cudaMalloc(&h->d_out, B * h->out_doubles * sizeof(double)), "cudaMalloc(d_out)");
  CKC(cudaMalloc(&h->d_iout, B * 2 * sizeof(int32_t)), "cudaMalloc(d_iout)");
  CKC(cudaMalloc(&h->d_counters, B * 2 * sizeof(int32_t)), "cudaMalloc(d_counters)");
  CKC(cudaMalloc(&h->d_last_iters, B * sizeof(int32_t)), "cudaMalloc(d_last_iters)");
  CKC(cudaMemset(h->d_last_iters, 0, B * sizeof(int32_t)), "cudaMemset(d_last_iters)");
  CKC(cudaMalloc(&h->d_perm, B * sizeof(int32_t)), "cudaMalloc(d_perm)");
  CKC(cudaMalloc(&h->d_valid, B * sizeof(int32_t)), "cudaMalloc(d_valid)");
  CKC(cudaMemset(h->d_valid, 0, B * sizeof(int32_t)), "cudaMemset(d_valid)");
  CKC(cudaMalloc(&h->d_queue, 8 * sizeof(int32_t)), "cudaMalloc(d_queue)");
  CKC(cudaMalloc(&h->d_fail, 3 * B * sizeof(int32_t)), "cudaMalloc(d_fail)");
  CKC(cudaMalloc(&h->d_mask, B), "cudaMalloc(d_mask)");
  CKC(cudaMallocHost(&h->h_in, B * h->in_doubles * sizeof(double)), "cudaMallocHost(h_in)");
  CKC(cudaMallocHost(&h->h_out, B * h->out_doubles * sizeof(double)), "cudaMallocHost(h_out)");
  CKC(cudaMallocHost(&h->h_iout, B * 2 * sizeof(int32_t)), "cudaMallocHost(h_iout)");
  CKC(cudaMallocHost(&h->h_counters, B * 2 * sizeof(int32_t)), "cudaMallocHost(h_counters)");
#ifdef CMPC_PROFILE
  CKC(cudaMalloc(&h->d_prof, PF_COUNT * sizeof(unsigned long long)), "cudaMalloc(d_prof)");
#endif
  CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking), "cudaStreamCreate");
  CKC(cudaEventCreate(&h->ev0), "cudaEventCreate");
  CKC(cudaEventCreate(&h->ev1), "cudaEventCreate");
  CKC(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming), "cudaEventCreate");
#undef CKC
  *out = h;
  return 0;
}

int cmpc_destroy(cmpc_handle* h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  cudaFree(h->snap); cudaFree(h->snap_iters); cudaFree(h->snap_valid); cudaFree(h->iter); cudaFree(h->scratch); cudaFree(h->d_in); cudaFree(h->d_out);
  cudaFree(h->d_iout); cudaFree(h->d_counters); cudaFree(h->d_last_iters); cudaFree(h->d_perm); cudaFree(h->d_valid); cudaFree(h->d_queue);
  cudaFree(h->d_fail); cudaFree(h->d_mask); cudaFree(h->d_traj); cudaFree(h->d_prof);
  cudaFreeHost(h->h_in); cudaFreeHost(h->h_out); cudaFreeHost(h->h_iout); cudaFreeHost(h->h_counters); cudaFreeHost(h->h_traj);
  if (h->stream) cudaStreamDestroy(h->stream);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->ev_done) cudaEventDestroy(h->ev_done);
  delete h;
  return 0;
}

// the launches of one solve on stream s (no ordering / bookkeeping)
static int launch_solve(cmpc_handle* h, int32_t batch, const double* x0, const double* com_ref, const double* foot_ref, const double* gamma,
                        const double* mass, const double* k1, int32_t warm_mode, const Outputs& o, cudaStream_t s, int* launches) {
  CK(cudaMemsetAsync(h->d_queue, 0, 8 * sizeof(int32_t), s), "memset queue");
  CK(cudaMemsetAsync(h->d_fail, 0xFF, 3 * (size_t)batch * sizeof(int32_t), s), "memset retry ring");     // -1: entry not published yet
  const int32_t* order = nullptr;
  int nl = 0;
  if (warm_mode != CMPC_COLD) {                       // the previous solve of these instances tells how expensive they are
    cmpc_order_kernel<<<1, 1024, 0, s>>>(batch, h->cfg.N, h->d_last_iters, gamma, h->d_perm);
    CK(cudaGetLastError(), "cmpc_order_kernel launch");
    order = h->d_perm; ++nl;
  }
  const int grid = batch < h->slots ? batch : h->slots;
  cmpc_solve_kernel<<<grid, h->threads, h->smem_bytes, s>>>(h->cfg, batch, x0, com_ref, foot_ref, gamma, mass, k1, h->iter, h->istride,
                                                            h->scratch, h->sstride, warm_mode, o, order, reinterpret_cast<Queue*>(h->d_queue), h->d_fail,
                                                            h->d_last_iters, h->d_valid);
  CK(cudaGetLastError(), "cmpc_solve_kernel launch");
  *launches = nl + 1;
  return 0;
}

int cmpc_solve_device(cmpc_handle* h, int32_t batch, const double* x0, const double* com_ref,
                      const double* foot_ref, const double* gamma, const double* mass, const double* k1,
                      int32_t warm_mode, double* x1, double* u0, double* xN, double* cost, double* viol,
                      int32_t* status, int32_t* iters, void* stream) {
  if (!h) return fail(-1, "cmpc_solve_device: null handle");
  if (batch < 1 || batch > h->cap) return fail(-1, "cmpc_solve_device: batch exceeds the handle's capacity");
  if (!x0 || !com_ref || !foot_ref || !gamma || !mass || !k1) return fail(-1, "cmpc_solve_device: null input pointer");
  if (warm_mode < 0 || warm_mode > 4) return fail(-1, "cmpc_solve_device: bad warm_mode");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  if (int rc = order_after_last(h, s)) return rc;
  Outputs o{x1, u0, xN, cost, viol, status, iters, h->d_counters, h->d_prof};
  if (h->d_prof) CK(cudaMemsetAsync(h->d_prof, 0, PF_COUNT * sizeof(unsigned long long), s), "memset prof");
  CK(cudaEventRecord(h->ev0, s), "cudaEventRecord");
  int launches = 0;
  if (int rc = launch_solve(h, batch, x0, com_ref, foot_ref, gamma, mass, k1, warm_mode, o, s, &launches)) return rc;
  CK(cudaEventRecord(h->ev1, s), "cudaEventRecord");
  h->last_batch = batch; h->last_launches = launches; h->have_timing = true;
  return mark_done(h, s);
}

static int solve_host_impl(cmpc_handle* h, int32_t batch, const double* x0, const double* com_ref,
                           const double* foot_ref, const double* gamma, const double* mass, const double* k1,
                           int32_t warm_mode, double* x1, double* u0, double* xN, double* cost, double* viol,
                           int32_t* status, int32_t* iters, int32_t traj_batch, double* X, double* U, const char* who) {
  if (!h) return fail(-1, "cmpc_solve_host: null handle");
  if (batch < 1 || batch > h->cap) return fail(-1, "cmpc_solve_host: batch exceeds the handle's capacity");
  if (!x0 || !com_ref || !foot_ref || !gamma || !mass || !k1) return fail(-1, "cmpc_solve_host: null input pointer");
  if (traj_batch < 0 || traj_batch > batch || (traj_batch > 0 && (!X || !U))) return fail(-1, "cmpc_solve_host_traj: bad trajectory arguments");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  if (traj_batch > 0) { if (int rc = ensure_traj(h)) return rc; }
  const int N = h->cfg.N; const size_t B = (size_t)batch;
  // pack inputs into one pinned buffer -> one H2D copy
  double* p = h->h_in;
  double* hx0 = p; p += B * NXP; double* hcom = p; p += B * 9 * N; double* hfoot = p; p += B * 8 * N;
  double* hgam = p; p += B * 2 * (N + 1); double* hm = p; p += B; double* hk = p; p += B;
  if (int rc = order_after_last(h, h->stream)) return rc;
  CK(cudaStreamSynchronize(h->stream), who);           // the pinned staging of a previous call must have been consumed
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
  const size_t TB = (size_t)traj_batch, nX = TB * (N + 1) * NXP, nU = TB * N * NU;
  if (traj_batch > 0) {                                // the full primal trajectories ride on the same synchronisation
    cmpc_export_traj<<<traj_batch, 128, 0, h->stream>>>(traj_batch, N, h->iter, h->istride, h->d_traj, h->d_traj + nX);
    CK(cudaGetLastError(), "cmpc_export_traj launch");
    CK(cudaMemcpyAsync(h->h_traj, h->d_traj, (nX + nU) * sizeof(double), cudaMemcpyDeviceToHost, h->stream), "D2H");
  }
  CK(cudaStreamSynchronize(h->stream), who);
  const double* r = h->h_out;
  if (x1) memcpy(x1, r, B * NXP * sizeof(double)); r += B * NXP;
  if (u0) memcpy(u0, r, B * NU * sizeof(double)); r += B * NU;
  if (xN) memcpy(xN, r, B * NXP * sizeof(double)); r += B * NXP;
  if (cost) memcpy(cost, r, B * sizeof(double)); r += B;
  if (viol) memcpy(viol, r, B * sizeof(double));
  if (status) memcpy(status, h->h_iout, B * sizeof(int32_t));
  if (iters) memcpy(iters, h->h_iout + B, B * sizeof(int32_t));
  if (traj_batch > 0) { memcpy(X, h->h_traj, nX * sizeof(double)); memcpy(U, h->h_traj + nX, nU * sizeof(double)); }
  return 0;
}

int cmpc_solve_host(cmpc_handle* h, int32_t batch, const double* x0, const double* com_ref,
                    const double* foot_ref, const double* gamma, const double* mass, const double* k1,
                    int32_t warm_mode, double* x1, double* u0, double* xN, double* cost, double* viol,
                    int32_t* status, int32_t* iters) {
  return solve_host_impl(h, batch, x0, com_ref, foot_ref, gamma, mass, k1, warm_mode, x1, u0, xN, cost, viol, status, iters, 0, nullptr, nullptr,
                         "cmpc_solve_host: kernel execution");
}

int cmpc_solve_host_traj(cmpc_handle* h, int32_t batch, const double* x0, const double* com_ref,
                         const double* foot_ref, const double* gamma, const double* mass, const double* k1,
                         int32_t warm_mode, double* x1, double* u0, double* xN, double* cost, double* viol,
                         int32_t* status, int32_t* iters, int32_t traj_batch, double* X, double* U) {
  return solve_host_impl(h, batch, x0, com_ref, foot_ref, gamma, mass, k1, warm_mode, x1, u0, xN, cost, viol, status, iters, traj_batch, X, U,
                         "cmpc_solve_host_traj: kernel execution");
}

int cmpc_get_trajectory(cmpc_handle* h, int32_t batch, double* X, double* U) {
  if (!h || !X || !U || batch < 1 || batch > h->cap) return fail(-1, "cmpc_get_trajectory: bad arguments");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  if (int rc = ensure_traj(h)) return rc;
  if (int rc = order_after_last(h, h->stream)) return rc;
  const int N = h->cfg.N; const size_t B = (size_t)batch, nX = B * (N + 1) * NXP, nU = B * N * NU;
  cmpc_export_traj<<<batch, 128, 0, h->stream>>>(batch, N, h->iter, h->istride, h->d_traj, h->d_traj + nX);
  CK(cudaGetLastError(), "cmpc_export_traj launch");
  CK(cudaMemcpyAsync(h->h_traj, h->d_traj, (nX + nU) * sizeof(double), cudaMemcpyDeviceToHost, h->stream), "D2H");
  if (int rc = mark_done(h, h->stream)) return rc;
  CK(cudaStreamSynchronize(h->stream), "cmpc_get_trajectory");
  memcpy(X, h->h_traj, nX * sizeof(double)); memcpy(U, h->h_traj + nX, nU * sizeof(double));
  return 0;
}

int cmpc_set_warm(cmpc_handle* h, int32_t batch, const double* X, const double* U) {
  if (!h || !X || !U || batch < 1 || batch > h->cap) return fail(-1, "cmpc_set_warm: bad arguments");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  if (int rc = ensure_traj(h)) return rc;
  if (int rc = order_after_last(h, h->stream)) return rc;
  CK(cudaStreamSynchronize(h->stream), "cmpc_set_warm");          // pinned staging free
  const int N = h->cfg.N; const size_t B = (size_t)batch, nX = B * (N + 1) * NXP, nU = B * N * NU;
  memcpy(h->h_traj, X, nX * sizeof(double)); memcpy(h->h_traj + nX, U, nU * sizeof(double));
  CK(cudaMemcpyAsync(h->d_traj, h->h_traj, (nX + nU) * sizeof(double), cudaMemcpyHostToDevice, h->stream), "H2D");
  cmpc_import_traj<<<batch, 128, 0, h->stream>>>(batch, N, h->iter, h->istride, h->d_traj, h->d_traj + nX, h->d_valid);
  CK(cudaGetLastError(), "cmpc_import_traj launch");
  if (int rc = mark_done(h, h->stream)) return rc;
  CK(cudaStreamSynchronize(h->stream), "cmpc_set_warm");
  return 0;
}

int cmpc_warm_save(cmpc_handle* h, int32_t batch) {
  if (!h || batch < 1 || batch > h->cap) return fail(-1, "cmpc_warm_save: bad arguments");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  const size_t B = (size_t)h->cap;
  if (!h->snap) CK(cudaMalloc(&h->snap, B * h->istride * sizeof(double)), "cudaMalloc(snapshot)");
  if (!h->snap_iters) CK(cudaMalloc(&h->snap_iters, B * sizeof(int32_t)), "cudaMalloc(snapshot iters)");
  if (!h->snap_valid) CK(cudaMalloc(&h->snap_valid, B * sizeof(int32_t)), "cudaMalloc(snapshot valid)");
  if (int rc = order_after_last(h, h->stream)) return rc;
  CK(cudaMemcpyAsync(h->snap_iters, h->d_last_iters, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream), "snapshot iters");
  CK(cudaMemcpyAsync(h->snap_valid, h->d_valid, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream), "snapshot valid");
  CK(cudaMemcpyAsync(h->snap, h->iter, (size_t)batch * h->istride * sizeof(double), cudaMemcpyDeviceToDevice, h->stream), "snapshot copy");
  if (int rc = mark_done(h, h->stream)) return rc;
  CK(cudaStreamSynchronize(h->stream), "cmpc_warm_save");
  h->snap_batch = batch;
  return 0;
}

int cmpc_warm_restore(cmpc_handle* h, int32_t batch, void* stream) {
  if (!h || batch < 1 || batch > h->snap_batch) return fail(-1, "cmpc_warm_restore: no snapshot for that many instances");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  cudaStream_t s = stream ? (cudaStream_t)stream : h->stream;
  if (int rc = order_after_last(h, s)) return rc;
  CK(cudaMemcpyAsync(h->iter, h->snap, (size_t)batch * h->istride * sizeof(double), cudaMemcpyDeviceToDevice, s), "snapshot restore");
  CK(cudaMemcpyAsync(h->d_last_iters, h->snap_iters, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToDevice, s), "snapshot restore iters");
  CK(cudaMemcpyAsync(h->d_valid, h->snap_valid, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToDevice, s), "snapshot restore valid");
  return mark_done(h, s);
}

int cmpc_reset_warm(cmpc_handle* h, const uint8_t* mask, int32_t n) {
  if (!h) return fail(-1, "cmpc_reset_warm: null handle");
  if (mask && (n < 1 || n > h->cap)) return fail(-1, "cmpc_reset_warm: bad mask length");
  CK(cudaSetDevice(h->device), "cudaSetDevice");
  if (int rc = order_after_last(h, h->stream)) return rc;
  const int cnt = mask ? n : h->cap;
  if (mask) CK(cudaMemcpyAsync(h->d_mask, mask, (size_t)n, cudaMemcpyHostToDevice, h->stream), "H2D mask");
  cmpc_clear_valid<<<(cnt + 255) / 256, 256, 0, h->stream>>>(cnt, mask ? h->d_mask : nullptr, h->d_valid);
  CK(cudaGetLastError(), "cmpc_clear_valid launch");
  if (int rc = mark_done(h, h->stream)) return rc;
  CK(cudaStreamSynchronize(h->stream), "cmpc_reset_warm");         // (the caller's mask buffer may be pageable)
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

int cmpc_footprint(const cmpc_handle* h, size_t* iterate_bytes_per_instance, size_t* scratch_bytes_per_slot, int32_t* slots,
                   size_t* smem_bytes_per_cta) {
  if (!h) return fail(-1, "cmpc_footprint: null handle");
  if (iterate_bytes_per_instance) *iterate_bytes_per_instance = h->istride * sizeof(double);
  if (scratch_bytes_per_slot) *scratch_bytes_per_slot = h->sstride * sizeof(double);
  if (slots) *slots = h->slots;
  if (smem_bytes_per_cta) *smem_bytes_per_cta = sizeof(Smem);
  return 0;
}

int cmpc_assemble_device(const cmpc_walk_tables* tb, int32_t device, int32_t batch, const int32_t* tick, const double* com_pos,
                         const double* com_vel, const double* hw, const double* theta, const double* yaw, const double* plan,
                         double* x0, double* com_ref, double* foot_ref, double* gamma, int32_t* err, void* stream) {
  if (!tb || batch < 1 || !tick || !com_pos || !com_vel || !hw || !theta || !yaw || !plan || !x0 || !com_ref || !foot_ref || !gamma)
    return fail(-1, "cmpc_assemble_device: null argument");
  if (tb->N < 1 || tb->N > NMAX || tb->rate < 1 || tb->n_steps < 1 || !tb->com_tab || !tb->foot_tab || !tb->gamma_tab || !tb->step_index)
    return fail(-1, "cmpc_assemble_device: bad tables");
  CK(cudaSetDevice(device), "cudaSetDevice");
  cmpc_assemble_kernel<<<batch, 128, 0, (cudaStream_t)stream>>>(*tb, batch, tick, com_pos, com_vel, hw, theta, yaw, plan, x0, com_ref, foot_ref, gamma, err);
  CK(cudaGetLastError(), "cmpc_assemble_kernel launch");
  return 0;
}

// ---- batched dense QP of the whole-body inverse-dynamics step (csrc/cmpc_qp.cuh)
int cmpc_qp_solve_device(int32_t device, int32_t batch, int32_t n, int32_t m_eq, int32_t m_in, const double* H, const double* F,
                         const double* A_eq, const double* b_eq, const double* A_in, const double* b_in, double tol, int32_t max_iter,
                         double* x, int32_t* status, int32_t* iters, void* stream) {
  using namespace cmpc_qp;
  if (batch < 1 || n < 1 || n > QP_MAXN || m_eq < 0 || m_eq > QP_MAXE || m_in < 0 || m_in > QP_MAXI)
    return fail(-1, "cmpc_qp_solve_device: bad sizes (n <= 96, m_eq <= 48, m_in <= 32)");
  if (!H || !F || !x || (m_eq && (!A_eq || !b_eq)) || (m_in && (!A_in || !b_in))) return fail(-1, "cmpc_qp_solve_device: null pointer");
  CK(cudaSetDevice(device), "cudaSetDevice");
  QpDims d{n, m_eq, m_in, max_iter > 0 ? max_iter : 60, tol > 0 ? tol : 1e-9, 1e-9, 1e-11};
  const size_t smem = qp_smem_doubles(n, m_eq, m_in) * sizeof(double);
  CK(cudaFuncSetAttribute(cmpc_qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(qp smem)");
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cmpc_qp_kernel, QP_THREADS, smem), "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
  if (per_sm < 1) return fail(-1, "cmpc_qp_solve_device: the QP does not fit in shared memory");
  const int slots = per_sm * prop.multiProcessorCount;
  cmpc_qp_kernel<<<batch < slots ? batch : slots, QP_THREADS, smem, (cudaStream_t)stream>>>(d, batch, H, F, A_eq, b_eq, A_in, b_in, x, status, iters);
  CK(cudaGetLastError(), "cmpc_qp_kernel launch");
  return 0;
}

int cmpc_qp_solve_host(int32_t device, int32_t batch, int32_t n, int32_t m_eq, int32_t m_in, const double* H, const double* F,
                       const double* A_eq, const double* b_eq, const double* A_in, const double* b_in, double tol, int32_t max_iter,
                       double* x, int32_t* status, int32_t* iters) {
  if (batch < 1 || n < 1 || !H || !F || !x) return fail(-1, "cmpc_qp_solve_host: bad arguments");
  CK(cudaSetDevice(device), "cudaSetDevice");
  const size_t B = (size_t)batch;
  const size_t sz[6] = {B * n * n, B * n, B * m_eq * n, B * m_eq, B * m_in * n, B * m_in};
  const double* src[6] = {H, F, A_eq, b_eq, A_in, b_in};
  size_t tot = 0; for (int k = 0; k < 6; ++k) tot += sz[k];
  double* dbuf = nullptr; int32_t* ibuf = nullptr;
  CK(cudaMalloc(&dbuf, (tot + B * n) * sizeof(double)), "cudaMalloc(qp)");
  if (cudaMalloc(&ibuf, 2 * B * sizeof(int32_t)) != cudaSuccess) { cudaFree(dbuf); return fail(-2, "cudaMalloc(qp status)"); }
  double* dp[6]; size_t off = 0;
  int rc = 0;
  for (int k = 0; k < 6 && !rc; ++k) {
    dp[k] = dbuf + off; off += sz[k];
    if (sz[k] && cudaMemcpy(dp[k], src[k], sz[k] * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(-2, "cmpc_qp_solve_host: H2D");
  }
  double* dx = dbuf + tot;
  if (!rc) rc = cmpc_qp_solve_device(device, batch, n, m_eq, m_in, dp[0], dp[1], dp[2], dp[3], dp[4], dp[5], tol, max_iter, dx, ibuf, ibuf + B, nullptr);
  if (!rc && cudaMemcpy(x, dx, B * n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(-2, "cmpc_qp_solve_host: kernel execution / D2H");
  if (!rc && status && cudaMemcpy(status, ibuf, B * sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(-2, "cmpc_qp_solve_host: D2H");
  if (!rc && iters && cudaMemcpy(iters, ibuf + B, B * sizeof(int32_t), cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(-2, "cmpc_qp_solve_host: D2H");
  cudaFree(dbuf); cudaFree(ibuf);
  return rc;
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
