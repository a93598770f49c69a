// markov.cu -- the Markov chain a policy induces on an MDP, for the average-reward indicators.
//
//   colosseum/mdp/utils/markov_chain.py:34-51   get_average_rewards / get_transition_probabilities
//       r[s] = sum_a R[s,a] pi[s,a];   P[s,j] = min(1, sum_a T[s,a,j] pi[s,a])
//   colosseum/mdp/utils/markov_chain.py:64-137  get_stationary_distribution (consumed by get_average_reward :12-31,
//       BaseMDP.{optimal,worst,random}_average_reward mdp/base.py:895-941 and the regret computation of
//       experiment/agent_mdp_interaction.py:518-578)
//
// policy_chain_kernel streams T once (one warp per state, lanes over the next-state index: coalesced reads of the A
// rows, coalesced write of the P row) -- the same HBM-bound pass as a value-iteration sweep.  The stationary
// distribution reached from the start distribution x0 is x0 * lim L^(2^k), L = (I + P)/2 the lazy chain (aperiodic,
// same stationary distributions): repeated squaring in fp64 with the rows rescaled to unit sum after every product.
// ~50 squarings cover 2^50 steps, which is what nearly reducible chains need (the optimal policy of SimpleGrid with
// p_rand = 0.01 has two metastable corners: plain power iteration changes by < 1e-8 per step while still 0.2 away
// from the limit -- measured, see tests/test_gpu_markov.py).  lazy_transpose_kernel + colo_power_iteration_f64
// (x <- M x / |M x| on the backup kernels) remain available for well-mixing chains.
#include <string.h>

#include "common.cuh"

namespace colo {

__global__ void __launch_bounds__(256) policy_chain_kernel(const float* __restrict__ T, const float* __restrict__ R,
                                                           const float* __restrict__ pi, int S, int A,
                                                           float* __restrict__ P, float* __restrict__ r) {
  // blockIdx.y = policy of a batch evaluated on ONE MDP (T, R shared; pi, P, r per policy)
  pi += (size_t)blockIdx.y * S * A;
  P += (size_t)blockIdx.y * S * S;
  if (r != nullptr) r += (size_t)blockIdx.y * S;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long s = warp; s < S; s += n_warps) {
    const float* Ts = T + (size_t)s * A * S;
    const float* ps = pi + (size_t)s * A;
    for (int j = lane; j < S; j += 32) {
      float acc = 0.f;
      for (int a = 0; a < A; ++a) acc += ldg_stream1(Ts + (size_t)a * S + j) * __ldg(ps + a);
      P[(size_t)s * S + j] = fminf(1.0f, acc);  // markov_chain.py:51
    }
    if (lane == 0 && r != nullptr) {
      float acc = 0.f;
      for (int a = 0; a < A; ++a) acc += R[(size_t)s * A + a] * ps[a];
      r[s] = acc;  // :41
    }
  }
}

__global__ void lazy_transpose_kernel(const float* __restrict__ P, int S, float* __restrict__ M) {
  // M[j, s] = 0.5 * P[s, j] + 0.5 * (s == j); 32 x 32 tiles through shared memory, both sides coalesced
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int s = by + i, j = bx + threadIdx.x;
    tile[i][threadIdx.x] = (s < S && j < S) ? P[(size_t)s * S + j] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int j = bx + i, s = by + threadIdx.x;
    if (j < S && s < S) M[(size_t)j * S + s] = 0.5f * tile[threadIdx.x][i] + (s == j ? 0.5f : 0.f);
  }
}

// ---- limiting matrix by repeated squaring (robust for nearly reducible chains, where power iteration stalls) ----
__global__ void lazy_matrix_f64_kernel(const float* __restrict__ P, int S, double* __restrict__ L) {
  // L = (I + P)/2 with rows rescaled to sum exactly (to rounding) 1; one warp per row; blockIdx.y = chain of a batch
  P += (size_t)blockIdx.y * S * S;
  L += (size_t)blockIdx.y * S * S;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long s = warp; s < S; s += n_warps) {
    double sum = 0.0;
    for (int j = lane; j < S; j += 32) sum += 0.5 * (double)P[(size_t)s * S + j] + (j == s ? 0.5 : 0.0);
    sum = warp_sum(sum);
    const double inv = 1.0 / sum;
    for (int j = lane; j < S; j += 32) L[(size_t)s * S + j] = (0.5 * (double)P[(size_t)s * S + j] + (j == s ? 0.5 : 0.0)) * inv;
  }
}

constexpr int kGemmTile = 64, kGemmBK = 16;

__global__ void __launch_bounds__(256) dsquare_kernel(const double* __restrict__ A, int S, double* __restrict__ C,
                                                      const int* __restrict__ frozen = nullptr) {
  // C = A * A (row-major fp64, S x S): 64 x 64 tile per CTA, 4 x 4 per thread, k-major shared tiles;
  // blockIdx.z = chain of a batch (frozen[z] != 0: that chain has converged, its buffers are left alone)
  if (frozen != nullptr && frozen[blockIdx.z]) return;
  A += (size_t)blockIdx.z * S * S;
  C += (size_t)blockIdx.z * S * S;
  __shared__ __align__(16) double As[kGemmBK][kGemmTile + 2];
  __shared__ __align__(16) double Bs[kGemmBK][kGemmTile + 2];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int m0 = blockIdx.y * kGemmTile, n0 = blockIdx.x * kGemmTile;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int k0 = 0; k0 < S; k0 += kGemmBK) {
    // A tile [64 rows x 16 k] -> As[k][m];  B tile [16 k x 64 cols] -> Bs[k][n]
    for (int i = threadIdx.x; i < kGemmTile * kGemmBK; i += 256) {
      const int m = i / kGemmBK, k = i % kGemmBK;
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < S && gk < S) ? A[(size_t)gm * S + gk] : 0.0;
      const int kk = i / kGemmTile, n = i % kGemmTile;
      const int gk2 = k0 + kk, gn = n0 + n;
      Bs[kk][n] = (gk2 < S && gn < S) ? A[(size_t)gk2 * S + gn] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kGemmBK; ++k) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gm = m0 + ty * 4 + i, gn = n0 + tx * 4 + j;
      if (gm < S && gn < S) C[(size_t)gm * S + gn] = acc[i][j];
    }
}

__global__ void row_normalize_diff_kernel(double* __restrict__ C, const double* __restrict__ Prev, int S,
                                          unsigned long long* __restrict__ diff, const int* __restrict__ frozen = nullptr) {
  // rows of C rescaled to unit sum (the powers stay stochastic); diff = max |C - Prev| after the rescale;
  // blockIdx.y = chain of a batch
  if (frozen != nullptr && frozen[blockIdx.y]) return;
  C += (size_t)blockIdx.y * S * S;
  Prev += (size_t)blockIdx.y * S * S;
  diff += blockIdx.y;
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  double d = 0.0;
  for (long long s = warp; s < S; s += n_warps) {
    double sum = 0.0;
    for (int j = lane; j < S; j += 32) sum += C[(size_t)s * S + j];
    sum = warp_sum(sum);
    const double inv = 1.0 / sum;
    for (int j = lane; j < S; j += 32) {
      const double v = C[(size_t)s * S + j] * inv;
      C[(size_t)s * S + j] = v;
      const double dd = fabs(v - Prev[(size_t)s * S + j]);
      d = dd > d ? dd : d;
    }
  }
  d = warp_max(d);
  if (lane == 0 && d > 0.0) atomic_max_nonneg(diff, d);
}

__global__ void vecmat_kernel(const double* __restrict__ x0, const double* __restrict__ M, int S, double* __restrict__ x) {
  // x[j] = sum_s x0[s] M[s,j], then rescaled to unit sum by the caller-visible single CTA pass below
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= S) return;
  double acc = 0.0;
  for (int s = 0; s < S; ++s) acc += x0[s] * M[(size_t)s * S + j];
  x[j] = acc;
}

__global__ void unit_sum_kernel(double* __restrict__ x, int S) {
  __shared__ double sm[32];
  __shared__ double tot;
  double part = 0.0;
  for (int i = threadIdx.x; i < S; i += blockDim.x) part += x[i];
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    tot = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S; i += blockDim.x) x[i] /= tot;
}

// ---- a batch of policies on one MDP (the continuous regret tick of N agent loops) -------------------------------------
// after every squaring: chains whose product moved by less than tol freeze (their limit stays in the buffer the
// squaring wrote); `left` counts the chains still moving
__global__ void chain_freeze_kernel(const unsigned long long* __restrict__ diff, int B, double tol, int k,
                                    int* __restrict__ frozen, int* __restrict__ where, int* __restrict__ squarings,
                                    int* __restrict__ left) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B || frozen[b]) return;
  const double d = __longlong_as_double((long long)diff[b]);
  if (d < tol) {
    frozen[b] = 1;
    where[b] = (k + 1) & 1;  // the buffer squaring k wrote
    squarings[b] = k + 1;
  } else {
    atomicAdd(left, 1);
  }
}

// the average reward of a chain from its limiting matrix M = lim L^n: with ONE recurrent class every row of M is the
// stationary distribution (markov_chain.py:12-31: (average_rewards * sd).sum()); rows that disagree mean several
// classes, whose weights follow the reference's first-reachable-class rule -- flagged for the caller.  One CTA per chain.
__global__ void __launch_bounds__(256) chain_average_reward_kernel(const double* __restrict__ M0, const double* __restrict__ M1,
                                                                   const int* __restrict__ where, const float* __restrict__ r,
                                                                   int S, double row_tol, double* __restrict__ ar,
                                                                   int* __restrict__ multichain) {
  __shared__ double s_val[8], s_dev[8];
  const int b = blockIdx.x;
  const double* M = (where[b] ? M1 : M0) + (size_t)b * S * S;
  const float* rb = r + (size_t)b * S;
  double val = 0.0, sum = 0.0, dev = 0.0;
  for (int j = threadIdx.x; j < S; j += blockDim.x) {
    const double m0 = M[j];
    val += m0 * (double)rb[j];
    sum += m0;
  }
  for (size_t e = threadIdx.x; e < (size_t)S * S; e += blockDim.x) {
    const double d = fabs(M[e] - M[e % S]);
    dev = d > dev ? d : dev;
  }
  __shared__ double s_sum[8];
  val = warp_sum(val);
  sum = warp_sum(sum);
  dev = warp_max(dev);
  if ((threadIdx.x & 31) == 0) { s_val[threadIdx.x >> 5] = val; s_sum[threadIdx.x >> 5] = sum; s_dev[threadIdx.x >> 5] = dev; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0, t = 0.0, d = 0.0;
    for (int w = 0; w < 8; ++w) { v += s_val[w]; t += s_sum[w]; d = s_dev[w] > d ? s_dev[w] : d; }
    ar[b] = v / t;
    multichain[b] = d > row_tol;
  }
}

}  // namespace colo

extern "C" {

size_t colo_average_rewards_work_bytes(int B, int S) {
  return (size_t)B * S * S * (4 + 8 + 8) + (size_t)B * S * 4 + (size_t)B * (8 + 4 * 3) + 1024;
}

int colo_average_rewards_f64(const float* T, const float* R, const float* pi, int B, int S, int A, double tol,
                             int max_squarings, double* ar_out, int* multichain_out, int* squarings_out, void* work,
                             void* stream) {
  COLO_ARG_CHECK(T && R && pi && ar_out && multichain_out && work && B >= 0 && S > 0 && A > 0 && max_squarings > 0,
                 "T, R, pi, ar_out, multichain_out, work, B, S, A");
  if (B == 0) return COLO_OK;
  COLO_ARG_CHECK(B <= 65535, "B <= 65535 per call");
  cudaStream_t st = (cudaStream_t)stream;
  char* w = (char*)work;
  double* M0 = (double*)w; w += (size_t)B * S * S * 8;
  double* M1 = (double*)w; w += (size_t)B * S * S * 8;
  unsigned long long* diff = (unsigned long long*)w; w += (size_t)B * 8;
  float* P = (float*)w; w += (size_t)B * S * S * 4;
  float* r = (float*)w; w += (size_t)B * S * 4;
  int* frozen = (int*)w; w += (size_t)B * 4;
  int* where = (int*)w; w += (size_t)B * 4;
  int* squarings = (int*)w; w += (size_t)B * 4;
  w = (char*)(((uintptr_t)w + 255) / 256 * 256);
  int* left = (int*)w;
  COLO_CUDA_TRY(cudaMemsetAsync(frozen, 0, (size_t)B * 12, st));
  const long long wb = ((long long)S + 7) / 8;
  const int wgrid = (int)(wb < 1024 ? wb : 1024);
  colo::policy_chain_kernel<<<dim3(wgrid, B), 256, 0, st>>>(T, R, pi, S, A, P, r);
  int rc = colo::check_launch("policy_chain_kernel");
  if (rc != COLO_OK) return rc;
  colo::lazy_matrix_f64_kernel<<<dim3(wgrid, B), 256, 0, st>>>(P, S, M0);
  rc = colo::check_launch("lazy_matrix_f64_kernel");
  if (rc != COLO_OK) return rc;
  const int tiles = (S + colo::kGemmTile - 1) / colo::kGemmTile;
  int status = COLO_MAX_ITER;
  for (int k = 0; k < max_squarings; ++k) {
    double* cur = (k & 1) ? M1 : M0;
    double* nxt = (k & 1) ? M0 : M1;
    COLO_CUDA_TRY(cudaMemsetAsync(diff, 0, (size_t)B * 8, st));
    COLO_CUDA_TRY(cudaMemsetAsync(left, 0, sizeof(int), st));
    colo::dsquare_kernel<<<dim3(tiles, tiles, B), 256, 0, st>>>(cur, S, nxt, frozen);
    rc = colo::check_launch("dsquare_kernel");
    if (rc != COLO_OK) return rc;
    colo::row_normalize_diff_kernel<<<dim3(wgrid, B), 256, 0, st>>>(nxt, cur, S, diff, frozen);
    rc = colo::check_launch("row_normalize_diff_kernel");
    if (rc != COLO_OK) return rc;
    colo::chain_freeze_kernel<<<(B + 127) / 128, 128, 0, st>>>(diff, B, tol, k, frozen, where, squarings, left);
    rc = colo::check_launch("chain_freeze_kernel");
    if (rc != COLO_OK) return rc;
    if (k >= 8) {  // no chain of interest converges in fewer squarings; afterwards look every time
      int h = 0;
      COLO_CUDA_TRY(cudaMemcpyAsync(&h, left, sizeof(int), cudaMemcpyDeviceToHost, st));
      COLO_CUDA_TRY(cudaStreamSynchronize(st));
      if (h == 0) { status = COLO_OK; break; }
    }
  }
  if (status != COLO_OK) return status;
  colo::chain_average_reward_kernel<<<B, 256, 0, st>>>(M0, M1, where, r, S, 1e-9, ar_out, multichain_out);
  rc = colo::check_launch("chain_average_reward_kernel");
  if (rc != COLO_OK) return rc;
  if (squarings_out) COLO_CUDA_TRY(cudaMemcpyAsync(squarings_out, squarings, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  return COLO_OK;
}


size_t colo_stationary_distribution_work_bytes(int S) { return (size_t)2 * S * S * sizeof(double) + 512; }

int colo_stationary_distribution_f64(const float* P, int S, const double* x0, double tol, int max_squarings,
                                     double* x_out, int* squarings_out_host, void* work, void* stream) {
  COLO_ARG_CHECK(P && x0 && x_out && work && S > 0 && max_squarings > 0, "P, x0, x_out, work, S");
  cudaStream_t st = (cudaStream_t)stream;
  double* M0 = (double*)work;
  double* M1 = M0 + (size_t)S * S;
  unsigned long long* diff = (unsigned long long*)(M1 + (size_t)S * S);
  const long long wb = ((long long)S + 7) / 8;
  const int wgrid = (int)(wb < (long long)colo::sm_count() * 16 ? wb : (long long)colo::sm_count() * 16);
  colo::lazy_matrix_f64_kernel<<<wgrid, 256, 0, st>>>(P, S, M0);
  int r = colo::check_launch("lazy_matrix_f64_kernel");
  if (r != COLO_OK) return r;
  dim3 grid((S + colo::kGemmTile - 1) / colo::kGemmTile, (S + colo::kGemmTile - 1) / colo::kGemmTile);
  double* cur = M0;
  double* nxt = M1;
  int k = 0, rc = COLO_MAX_ITER;
  for (; k < max_squarings; ++k) {
    COLO_CUDA_TRY(cudaMemsetAsync(diff, 0, sizeof(unsigned long long), st));
    colo::dsquare_kernel<<<grid, 256, 0, st>>>(cur, S, nxt);
    r = colo::check_launch("dsquare_kernel");
    if (r != COLO_OK) return r;
    colo::row_normalize_diff_kernel<<<wgrid, 256, 0, st>>>(nxt, cur, S, diff);
    r = colo::check_launch("row_normalize_diff_kernel");
    if (r != COLO_OK) return r;
    unsigned long long h = 0;
    COLO_CUDA_TRY(cudaMemcpyAsync(&h, diff, sizeof(h), cudaMemcpyDeviceToHost, st));
    COLO_CUDA_TRY(cudaStreamSynchronize(st));
    double d;
    memcpy(&d, &h, sizeof(d));
    double* t = cur; cur = nxt; nxt = t;
    if (d < tol) { rc = COLO_OK; ++k; break; }
  }
  colo::vecmat_kernel<<<(S + 127) / 128, 128, 0, st>>>(x0, cur, S, x_out);
  r = colo::check_launch("vecmat_kernel");
  if (r != COLO_OK) return r;
  colo::unit_sum_kernel<<<1, 1024, 0, st>>>(x_out, S);
  r = colo::check_launch("unit_sum_kernel");
  if (r != COLO_OK) return r;
  if (squarings_out_host) *squarings_out_host = k;
  return rc;
}

int colo_policy_chain(const float* T, const float* R, const float* pi, int S, int A, float* P_out, float* r_out,
                      void* stream) {
  COLO_ARG_CHECK(T && pi && P_out && S > 0 && A > 0 && (r_out == nullptr || R != nullptr), "T, pi, P_out, S, A");
  const long long blocks = ((long long)S + 7) / 8;
  const long long cap = (long long)colo::sm_count() * 16;
  colo::policy_chain_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(T, R, pi, S, A, P_out, r_out);
  return colo::check_launch("policy_chain_kernel");
}

int colo_lazy_transpose(const float* P, int S, float* M_out, void* stream) {
  COLO_ARG_CHECK(P && M_out && S > 0, "P, M_out, S");
  dim3 grid((S + 31) / 32, (S + 31) / 32), block(32, 8);
  colo::lazy_transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(P, S, M_out);
  return colo::check_launch("lazy_transpose_kernel");
}

}  // extern "C"
