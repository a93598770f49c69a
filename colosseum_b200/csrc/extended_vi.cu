// extended_vi.cu -- UCRL2's extended (optimistic) value iteration on the GPU.
//
//   colosseum/dynamic_programming/infinite_horizon.py:67-118   extended_value_iteration
//   colosseum/dynamic_programming/infinite_horizon.py:222-251  _max_proba
//   called every artificial episode by colosseum/agent/agents/infinite_horizon/ucrl2.py:330
//
// One iteration (synchronous, u1 -> u2, as in the reference):
//   sorted = argsort(u1)                                             (:116)
//   for every (s, a):  p2 = _max_proba(T[s,a], sorted, beta_p[s,a])  optimistic transition inside the L1 ball: the mass
//                      beta/2 is moved onto the best state (largest u1) and taken away from the worst states first
//                      v  = min(r_max, R_hat[s,a] + beta_r[s,a]) + p2 . u1 - u1[s]          (:95-102)
//                      Q[s,a] = v;  u2[s] = v + u1[s] if first action, larger, or within eps of the incumbent (:104-109)
//   V[s] = max_a Q[s,a];  stop when ptp(u2 - u1) < eps and return (ptp(u1), Q, V)           (:111-113)
//
// The same T traffic as a VI sweep (one dense pass over T per iteration, gathered in sorted order), plus an argsort of
// S keys.  Mapping: kernel 1 (one CTA) tests the stopping rule of the previous iteration and bitonic-sorts u1 in
// shared memory; kernel 2 gives every state one warp, which walks the A rows of the state in ascending-u1 order with
// a warp prefix scan (the "take mass from the worst states first" loop of _max_proba is a clamp against the running
// prefix sum: removed_j = clamp(excess - prefix_j, 0, p_j)).  The iteration state (current buffer, done flag, span,
// iteration count) lives on the device, kernels of later iterations exit immediately once `done` is set, and the
// host looks at the flag every few iterations only.
#include "common.cuh"

namespace colo {

struct EviState {
  int cur;         // which of u[2] is u1
  int done;        // 1 = the stopping rule fired; Q, V and span are final
  long long iters; // iterations completed
  double span;     // ptp(u1) at the stopping iteration
};

constexpr int kEviSortThreads = 1024;

template <typename TV>
__global__ void __launch_bounds__(kEviSortThreads) evi_check_and_sort_kernel(TV* __restrict__ u /*[2][S]*/, int S,
                                                                            int n_pow2, double eps, int first,
                                                                            int* __restrict__ sorted_idx,
                                                                            EviState* __restrict__ st) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_lo[kEviSortThreads / 32], s_hi[kEviSortThreads / 32], s_ulo[kEviSortThreads / 32],
      s_uhi[kEviSortThreads / 32];
  __shared__ int s_stop;
  if (st->done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int cur = st->cur;
  if (!first) {
    // stopping rule of the iteration just computed: ptp(u2 - u1) < eps  (infinite_horizon.py:111)
    const TV* u1 = u + (size_t)cur * S;
    const TV* u2 = u + (size_t)(cur ^ 1) * S;
    double lo = INFINITY, hi = -INFINITY, ulo = INFINITY, uhi = -INFINITY;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
      const double d = (double)u2[i] - (double)u1[i];
      lo = d < lo ? d : lo;
      hi = d > hi ? d : hi;
      const double x = (double)u1[i];
      ulo = x < ulo ? x : ulo;
      uhi = x > uhi ? x : uhi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = fmin(lo, __shfl_xor_sync(FULL, lo, o));
      hi = fmax(hi, __shfl_xor_sync(FULL, hi, o));
      ulo = fmin(ulo, __shfl_xor_sync(FULL, ulo, o));
      uhi = fmax(uhi, __shfl_xor_sync(FULL, uhi, o));
    }
    if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; s_ulo[warp] = ulo; s_uhi[warp] = uhi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < nw; ++w) {
        s_lo[0] = fmin(s_lo[0], s_lo[w]); s_hi[0] = fmax(s_hi[0], s_hi[w]);
        s_ulo[0] = fmin(s_ulo[0], s_ulo[w]); s_uhi[0] = fmax(s_uhi[0], s_uhi[w]);
      }
      st->iters += 1;
      s_stop = (s_hi[0] - s_lo[0]) < eps;
      if (s_stop) {
        st->done = 1;
        st->span = s_uhi[0] - s_ulo[0];  // np.ptp(u1)  (:112)
      } else {
        st->cur = cur ^ 1;  // u1 = u2  (:114)
      }
    }
    __syncthreads();
    if (s_stop) return;
    cur ^= 1;
  }
  // sorted_indices = argsort(u1)  (:116; initially arange, which is a valid argsort of the all-zero u1)
  TV* key = reinterpret_cast<TV*>(smem_raw);
  int* idx = reinterpret_cast<int*>(smem_raw + (size_t)n_pow2 * sizeof(TV));
  const TV* u1 = u + (size_t)cur * S;
  for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
    key[i] = i < S ? u1[i] : (TV)INFINITY;
    idx[i] = i;
  }
  __syncthreads();
  for (int k = 2; k <= n_pow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const bool up = (i & k) == 0;
          const TV a = key[i], b = key[l];
          const int ia = idx[i], ib = idx[l];
          // total order (value, index): deterministic for ties
          const bool gt = a > b || (a == b && ia > ib);
          if (gt == up) { key[i] = b; key[l] = a; idx[i] = ib; idx[l] = ia; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < S; i += blockDim.x) sorted_idx[i] = idx[i];
}

// One state of one iteration: the A optimistic rows of state s, walked by ONE WARP in ascending-u1 order.  Shared by the
// per-iteration kernel (u1 / sorted_idx in global memory) and the one-CTA-per-instance kernel (both in shared memory):
// same operations in the same order, so the two give the same bits.  Returns (u2[s], V[s]) in every lane; lane 0
// stores Q[s, :].
template <typename TV>
__device__ __forceinline__ void evi_state(const float* __restrict__ Ts /* T[s] = A rows */,
                                          const float* __restrict__ est_s, const double* __restrict__ br_s,
                                          const double* __restrict__ bp_s, int S, int A, int s, double r_max, double eps,
                                          const TV* __restrict__ u1, const int* __restrict__ sorted_idx,
                                          TV* __restrict__ Qs, int lane, TV* u2_out, TV* v_out) {
  const int best = sorted_idx[S - 1];
  const TV u_best = u1[best];
  const double u_s = (double)u1[s];
  double u2s = 0.0, vmax = -INFINITY;
  for (int a = 0; a < A; ++a) {
    const float* __restrict__ p = Ts + (size_t)a * S;
    const double pbest = (double)p[best];
    const double min1 = fmin(1.0, pbest + 0.5 * bp_s[a]);  // (:230)
    TV dot = 0;
    if (min1 == 1.0) {
      dot = u_best;  // p2 = e_best  (:231-233)
    } else {
      // p2 = p with p2[best] = min1; the surplus min1 - p[best] is taken from the worst states first (:235-250)
      const double excess = min1 - pbest;
      double carry = 0.0;
      for (int j0 = 0; j0 < S; j0 += 32) {
        const int j = j0 + lane;
        int id = 0;
        double pj = 0.0;
        if (j < S) {
          id = sorted_idx[j];
          pj = (double)p[id];
        }
        double incl = pj;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const double t = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += t;
        }
        const double excl = carry + incl - pj;
        double removed = excess - excl;
        removed = removed < 0.0 ? 0.0 : (removed > pj ? pj : removed);
        double p2 = pj - removed;
        if (j < S && id == best) p2 = min1;
        // p2 is a float32 array in the reference
        if (j < S) dot += (TV)(float)p2 * u1[id];
        carry += __shfl_sync(FULL, incl, 31);
      }
      dot = warp_sum(dot);
    }
    const double r_opt = fmin((double)(float)r_max, (double)est_s[a] + br_s[a]);  // (:96-99)
    const double v = r_opt + ((double)dot - u_s);  // vec[s] -= 1  (:95,:100)
    const TV q = (TV)v;
    if (lane == 0) Qs[a] = q;
    vmax = (double)q > vmax ? (double)q : vmax;
    const double cand = v + u_s;
    if (a == 0 || cand > u2s || fabs(cand - u2s) < eps) u2s = (double)(TV)cand;  // (:102-109), stored as float32
  }
  *u2_out = (TV)u2s;
  *v_out = (TV)vmax;  // (:110)
}

template <typename TV>
__global__ void __launch_bounds__(256) evi_rows_kernel(const float* __restrict__ T, const float* __restrict__ est_r,
                                                       const double* __restrict__ beta_r,
                                                       const double* __restrict__ beta_p, int S, int A, double r_max,
                                                       double eps, TV* __restrict__ u, const int* __restrict__ sorted_idx,
                                                       TV* __restrict__ Q, TV* __restrict__ V,
                                                       const EviState* __restrict__ st) {
  if (st->done) return;
  const int lane = threadIdx.x & 31;
  const int s = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (s >= S) return;
  const int cur = st->cur;
  const TV* __restrict__ u1 = u + (size_t)cur * S;
  TV* __restrict__ u2 = u + (size_t)(cur ^ 1) * S;
  const size_t sa = (size_t)s * A;
  TV u2s, vs;
  evi_state<TV>(T + sa * S, est_r + sa, beta_r + sa, beta_p + sa, S, A, s, r_max, eps, u1, sorted_idx, Q + sa, lane, &u2s,
                &vs);
  if (lane == 0) {
    u2[s] = u2s;
    V[s] = vs;
  }
}

// ---- batched: one CTA per listed instance, the whole solve in one launch ------------------------------------------------
// UCRL2's artificial episodes end at loop-dependent times (ucrl2.py:173-181), so a batch of loops asks for the extended
// VI of an arbitrary SUBSET of its models at once: `index[k]` names the instance whose T / est_rewards / Q / V are used
// (stride = one instance), the bounds and the outputs span / iters / status are compact ([m]).  u1, u2 and the argsort
// live in shared memory; iterations are separated by __syncthreads only.
constexpr int kEviBatchThreads = 1024;

struct EviBatchArgs {
  const float* T;
  const float* est_r;
  const double* beta_r;
  const double* beta_p;
  const int* index;
  int m, S, A, n_pow2;
  double r_max, eps;
  long long max_iter;
  void* Q;
  void* V;
  double* span;
  long long* iters;
  int* status;
};

template <typename TV>
__global__ void __launch_bounds__(kEviBatchThreads) evi_batched_kernel(const EviBatchArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[4][kEviBatchThreads / 32];
  __shared__ int s_stop;
  const int S = a.S, A = a.A, n_pow2 = a.n_pow2;
  TV* ua = reinterpret_cast<TV*>(smem_raw);
  TV* ub = ua + S;
  TV* key = ub + S;
  int* idx = reinterpret_cast<int*>(key + n_pow2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = blockIdx.x; k < a.m; k += gridDim.x) {
    const size_t inst = a.index ? (size_t)a.index[k] : (size_t)k;
    const float* __restrict__ T = a.T + inst * S * A * S;
    const float* __restrict__ est = a.est_r + inst * S * A;
    const double* __restrict__ br = a.beta_r + (size_t)k * S * A;
    const double* __restrict__ bp = a.beta_p + (size_t)k * S * A;
    TV* __restrict__ Q = reinterpret_cast<TV*>(a.Q) + inst * S * A;
    TV* __restrict__ V = reinterpret_cast<TV*>(a.V) + inst * S;
    TV* u1 = ua;
    TV* u2 = ub;
    __syncthreads();  // the previous instance is done with the shared arrays
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
      ua[i] = 0;
      ub[i] = 0;
      idx[i] = i;  // arange is a valid argsort of the all-zero u1 (:88)
    }
    __syncthreads();
    long long it = 0;
    int rc = COLO_MAX_ITER;
    double span = 0.0;
    for (; it < a.max_iter;) {
      for (int s = warp; s < S; s += nw) {
        const size_t sa = (size_t)s * A;
        TV u2s, vs;
        evi_state<TV>(T + sa * S, est + sa, br + sa, bp + sa, S, A, s, a.r_max, a.eps, u1, idx, Q + sa, lane, &u2s, &vs);
        if (lane == 0) {
          u2[s] = u2s;
          V[s] = vs;
        }
      }
      __syncthreads();
      // stopping rule: ptp(u2 - u1) < eps, then span = ptp(u1)  (:111-112)
      double lo = INFINITY, hi = -INFINITY, ulo = INFINITY, uhi = -INFINITY;
      for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const double d = (double)u2[i] - (double)u1[i];
        lo = d < lo ? d : lo;
        hi = d > hi ? d : hi;
        const double x = (double)u1[i];
        ulo = x < ulo ? x : ulo;
        uhi = x > uhi ? x : uhi;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(FULL, lo, o));
        hi = fmax(hi, __shfl_xor_sync(FULL, hi, o));
        ulo = fmin(ulo, __shfl_xor_sync(FULL, ulo, o));
        uhi = fmax(uhi, __shfl_xor_sync(FULL, uhi, o));
      }
      if (lane == 0) { s_red[0][warp] = lo; s_red[1][warp] = hi; s_red[2][warp] = ulo; s_red[3][warp] = uhi; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < nw; ++w) {
          s_red[0][0] = fmin(s_red[0][0], s_red[0][w]); s_red[1][0] = fmax(s_red[1][0], s_red[1][w]);
          s_red[2][0] = fmin(s_red[2][0], s_red[2][w]); s_red[3][0] = fmax(s_red[3][0], s_red[3][w]);
        }
        s_stop = (s_red[1][0] - s_red[0][0]) < a.eps;
      }
      __syncthreads();
      ++it;
      if (s_stop) {
        span = s_red[3][0] - s_red[2][0];
        rc = COLO_OK;
        break;
      }
      TV* t = u1; u1 = u2; u2 = t;  // u1 = u2  (:114)
      // sorted_indices = argsort(u1)  (:116), total order (value, index)
      for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
        key[i] = i < S ? u1[i] : (TV)INFINITY;
        idx[i] = i;
      }
      __syncthreads();
      for (int kk = 2; kk <= n_pow2; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
          for (int i = threadIdx.x; i < n_pow2; i += blockDim.x) {
            const int l = i ^ j;
            if (l > i) {
              const bool up = (i & kk) == 0;
              const TV x = key[i], y = key[l];
              const int ia = idx[i], ib = idx[l];
              const bool gt = x > y || (x == y && ia > ib);
              if (gt == up) { key[i] = y; key[l] = x; idx[i] = ib; idx[l] = ia; }
            }
          }
          __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
      a.span[k] = span;
      a.iters[k] = it;
      a.status[k] = rc;
    }
  }
}

template <typename TV>
int extended_vi_batched(const float* T, const float* est_r, const double* beta_r, const double* beta_p,
                        const int* index, int m, int S, int A, double r_max, double eps, long long max_iter, TV* Q, TV* V,
                        double* span, long long* iters, int* status, void* stream) {
  COLO_ARG_CHECK(T && est_r && beta_r && beta_p && Q && V && span && iters && status, "null argument");
  COLO_ARG_CHECK(m >= 0 && S > 0 && A > 0 && S <= 8192 && max_iter >= 1, "m >= 0, S in [1, 8192], A > 0, max_iter >= 1");
  if (m == 0) return COLO_OK;
  int n_pow2 = 1;
  while (n_pow2 < S) n_pow2 <<= 1;
  const size_t smem = (size_t)2 * S * sizeof(TV) + (size_t)n_pow2 * (sizeof(TV) + sizeof(int));
  COLO_ARG_CHECK(smem <= 200 * 1024, "S too large for the one-CTA-per-instance extended VI");
  auto kern = evi_batched_kernel<TV>;
  { const int _es = ensure_dynamic_smem((const void*)kern, smem); if (_es != COLO_OK) return _es; }
  EviBatchArgs a;
  a.T = T; a.est_r = est_r; a.beta_r = beta_r; a.beta_p = beta_p; a.index = index;
  a.m = m; a.S = S; a.A = A; a.n_pow2 = n_pow2;
  a.r_max = r_max; a.eps = eps; a.max_iter = max_iter;
  a.Q = Q; a.V = V; a.span = span; a.iters = iters; a.status = status;
  // one warp per state and iteration: small models take small CTAs (cheaper barriers, more models resident per SM)
  const int threads = S > 256 ? kEviBatchThreads : (S > 64 ? 512 : 256);
  const int per_sm = 2048 / threads;
  const int grid = m < per_sm * 2 * sm_count() ? m : per_sm * 2 * sm_count();
  kern<<<grid, threads, smem, (cudaStream_t)stream>>>(a);
  return check_launch("evi_batched_kernel");
}

template <typename TV>
int extended_vi(const float* T, const float* est_r, const double* beta_r, const double* beta_p, int S, int A,
                double r_max, double eps, long long max_iter, TV* Q, TV* V, double* out_host, void* work, void* stream) {
  COLO_ARG_CHECK(T && est_r && beta_r && beta_p && Q && V && out_host && work, "null argument");
  COLO_ARG_CHECK(S > 0 && A > 0 && S <= 8192, "S in [1, 8192], A > 0");
  cudaStream_t st = (cudaStream_t)stream;
  int n_pow2 = 1;
  while (n_pow2 < S) n_pow2 <<= 1;
  // work = u[2][S] | sorted_idx[S] | state
  char* w = (char*)work;
  TV* u = (TV*)w;
  w += ((size_t)2 * S * sizeof(TV) + 255) / 256 * 256;
  int* sorted_idx = (int*)w;
  w += ((size_t)S * sizeof(int) + 255) / 256 * 256;
  EviState* state = (EviState*)w;
  COLO_CUDA_TRY(cudaMemsetAsync(work, 0, (size_t)(w - (char*)work) + sizeof(EviState), st));
  const size_t smem = (size_t)n_pow2 * (sizeof(TV) + sizeof(int));
  auto sort_kern = evi_check_and_sort_kernel<TV>;
  { const int _es = ensure_dynamic_smem((const void*)sort_kern, smem); if (_es != COLO_OK) return _es; }
  const int row_blocks = (int)(((long long)S * 32 + 255) / 256);
  EviState h = {};
  long long launched = 0;
  int check = 4;
  while (launched <= max_iter) {
    long long burst = check;
    if (launched + burst > max_iter + 1) burst = max_iter + 1 - launched;
    for (long long i = 0; i < burst; ++i) {
      sort_kern<<<1, kEviSortThreads, smem, st>>>(u, S, n_pow2, eps, launched == 0, sorted_idx, state);
      int r = check_launch("evi_check_and_sort_kernel");
      if (r != COLO_OK) return r;
      evi_rows_kernel<TV><<<row_blocks, 256, 0, st>>>(T, est_r, beta_r, beta_p, S, A, r_max, eps, u, sorted_idx, Q, V, state);
      r = check_launch("evi_rows_kernel");
      if (r != COLO_OK) return r;
      ++launched;
    }
    COLO_CUDA_TRY(cudaMemcpyAsync(&h, state, sizeof(EviState), cudaMemcpyDeviceToHost, st));
    COLO_CUDA_TRY(cudaStreamSynchronize(st));
    if (h.done) break;
    if (check < 64) check *= 2;
  }
  out_host[0] = h.span;
  out_host[1] = (double)h.iters;
  return h.done ? COLO_OK : COLO_MAX_ITER;
}

}  // namespace colo

extern "C" {

size_t colo_extended_vi_work_bytes(int S, int f64) {
  const size_t e = f64 ? 8 : 4;
  return ((size_t)2 * S * e + 255) / 256 * 256 + ((size_t)S * 4 + 255) / 256 * 256 + 256;
}
int colo_extended_vi_f32(const float* T, const float* est_rewards, const double* beta_r, const double* beta_p, int S,
                         int A, double r_max, double eps, long long max_iter, float* Q, float* V, double* out_host,
                         void* work, void* stream) {
  return colo::extended_vi<float>(T, est_rewards, beta_r, beta_p, S, A, r_max, eps, max_iter, Q, V, out_host, work, stream);
}
int colo_extended_vi_f64acc(const float* T, const float* est_rewards, const double* beta_r, const double* beta_p, int S,
                            int A, double r_max, double eps, long long max_iter, double* Q, double* V, double* out_host,
                            void* work, void* stream) {
  return colo::extended_vi<double>(T, est_rewards, beta_r, beta_p, S, A, r_max, eps, max_iter, Q, V, out_host, work,
                                   stream);
}
int colo_extended_vi_batched_f32(const float* T, const float* est_rewards, const double* beta_r, const double* beta_p,
                                 const int* index, int m, int S, int A, double r_max, double eps, long long max_iter,
                                 float* Q, float* V, double* span, long long* iters, int* status, void* stream) {
  return colo::extended_vi_batched<float>(T, est_rewards, beta_r, beta_p, index, m, S, A, r_max, eps, max_iter, Q, V,
                                          span, iters, status, stream);
}
int colo_extended_vi_batched_f64acc(const float* T, const float* est_rewards, const double* beta_r, const double* beta_p,
                                    const int* index, int m, int S, int A, double r_max, double eps, long long max_iter,
                                    double* Q, double* V, double* span, long long* iters, int* status, void* stream) {
  return colo::extended_vi_batched<double>(T, est_rewards, beta_r, beta_p, index, m, S, A, r_max, eps, max_iter, Q, V,
                                           span, iters, status, stream);
}

}  // extern "C"
