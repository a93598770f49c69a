// posterior.cu -- PSRL's transition-model sample on the GPU.
//
//   colosseum/agent/mdp_models/bayesian_models/conjugate_transitions.py:48-60  M_DIR._sample / sample:
//       r = rng.standard_gamma(hyper_params).astype(float32);  T = r / (1e-5 + r.sum(-1))      (sic: 1e-5 in the denominator)
//   feeds value iteration every episode: colosseum/agent/agents/episodic/posterior_sampling.py:142-144,
//   colosseum/agent/agents/infinite_horizon/posterior_sampling.py:177,374.
//
// S*A*S gamma draws per sample (Marsaglia-Tsang squeeze, alpha < 1 boosted with u^(1/alpha); fp64 like numpy's
// standard_gamma, rounded to float32 before the normalisation like the reference's .astype).  One warp per row: draws
// are written unnormalised while the row sum is reduced, then the row is rescaled in place.  Randomness: Philox4x32-10
// keyed by `seed`, counter (row*S + j, draw index t, attempt) -- a pure function of its arguments, so a model sharded by
// rows over GPUs samples exactly what one GPU would.  Parity with the reference is distributional (its numpy stream
// cannot be reproduced): moments and KS tests against scipy in tests/test_gpu_posterior.py.
#include "gamma_device.cuh"

namespace colo {

template <bool FAST>
__global__ void __launch_bounds__(256) dirichlet_rows_kernel(const float* __restrict__ hyper, long long rows, int S,
                                                             long long row0, unsigned long long seed,
                                                             unsigned long long t, float* __restrict__ T) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < rows; r += n_warps) {
    const float* h = hyper + (size_t)r * S;
    float* out = T + (size_t)r * S;
    float sum = 0.f;
    for (int j = lane; j < S; j += 32) {
      const uint64_t elem = (uint64_t)(row0 + r) * (uint64_t)S + (uint64_t)j;
      const float g = FAST ? gamma_draw_fast(h[j], seed, elem, t) : (float)gamma_draw((double)h[j], seed, elem, t);
      out[j] = g;
      sum += g;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / (1e-5f + sum);  // conjugate_transitions.py:53
    for (int j = lane; j < S; j += 32) out[j] *= inv;
  }
}

// N_NIG.sample (bayesian_models/conjugate_rewards.py:76-92): tau ~ Gamma(alpha, scale 1/beta) -> float32,
// var = 1/(lambda*tau), mean ~ Normal(mu, sqrt(var)) -> float32.  One thread per (s,a) row of (mu, lambda, alpha, beta).
__global__ void __launch_bounds__(256) nig_rows_kernel(const float* __restrict__ hyper, long long rows, long long row0,
                                                       unsigned long long seed, unsigned long long t,
                                                       float* __restrict__ R) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float4 h = __ldg(reinterpret_cast<const float4*>(hyper) + r);
    const uint64_t elem = (uint64_t)(row0 + r);
    const float tau = (float)(gamma_draw((double)h.z, seed ^ 0xD1B54A32D192ED03ULL, elem, t) / (double)h.w);
    const float var = 1.0f / (h.y * tau);
    const Philox4 w = philox4x32_10(seed ^ 0x8CB92BA72F3D8DD7ULL, elem, t);
    const double u1 = u53(w.w[0], w.w[1]) + 1.1102230246251565e-16, u2 = (double)w.w[2] * (1.0 / 4294967296.0);
    const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    R[r] = (float)((double)h.x + sqrt((double)var) * z);
  }
}

// Emission-map noise (colosseum/noises/*.py added in EmissionMap.get_observation, emission_maps/base.py:136-138): every
// observation row that is not the all-zero terminal observation gets `period` fresh draws, element j of the row using
// draw j % period.  kind 1: GaussianUncorrelated (noises/gaussian_uncorrelated.py:11-12; period = D);
// kind 2: StudentTUncorrelated (noises/student_t_uncorrelated.py:11-12 -- sic: the reference samples ONE array of the
// observation's shape and hands out its slices along the first axis, so a vector observation receives the same scalar
// on every feature: period = prod(shape[1:])).
__global__ void __launch_bounds__(256) emit_noise_kernel(float* __restrict__ out, const unsigned char* __restrict__ step_type,
                                                         const int* __restrict__ h, long long N, int H, int D, int period,
                                                         int kind, double param, unsigned long long seed,
                                                         unsigned long long t, unsigned long long env0) {
  const long long total = N * D;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long e = idx / D;
    const int j = (int)(idx - e * D);
    if (H > 0 && (h[e] >= H || step_type[e] == COLO_STEP_LAST)) continue;  // zeros past the horizon carry no noise
    const uint64_t elem = (env0 + (uint64_t)e) * (uint64_t)period + (uint64_t)(j % period);
    const Philox4 w = philox4x32_10(seed ^ 0x5851F42D4C957F2DULL, elem, t);
    const double u1 = u53(w.w[0], w.w[1]) + 1.1102230246251565e-16, u2 = (double)w.w[2] * (1.0 / 4294967296.0);
    const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    double x;
    if (kind == 1) {
      x = z * param;
    } else {  // standard_t(df) = z / sqrt(chi2_df / df), chi2_df = 2 * Gamma(df / 2)
      const double g = gamma_draw(0.5 * param, seed ^ 0x14057B7EF767814FULL, elem, t);
      x = z / sqrt(2.0 * g / param);
    }
    out[idx] += (float)x;
  }
}

// Correlated emission noises (colosseum/noises/gaussian_correlated.py:9-17, student_t_correlated.py:9-17): one covariance
// W ~ Wishart(df = D, scale * I) is drawn ONCE per emission map -- on the host, with the reference's own scipy call --
// and every observation then receives x ~ N(0, W) (kind 1) or the multivariate Student-t with shape W and df degrees of
// freedom, x = L z / sqrt(chi2_df / df) (kind 2; scipy's default df = 1).  L = chol(W) lower triangular f32 [D,D]; one
// CTA per env: the D standard normals of the env go to shared memory, thread i forms row i of L z.
__global__ void __launch_bounds__(256) emit_noise_correlated_kernel(float* __restrict__ out,
                                                                    const unsigned char* __restrict__ step_type,
                                                                    const int* __restrict__ h, long long N, int H, int D,
                                                                    const float* __restrict__ L, int kind, double df,
                                                                    unsigned long long seed, unsigned long long t,
                                                                    unsigned long long env0) {
  extern __shared__ float zs[];
  for (long long e = blockIdx.x; e < N; e += gridDim.x) {
    const bool skip = H > 0 && (h[e] >= H || step_type[e] == COLO_STEP_LAST);  // zeros past the horizon carry no noise
    if (!skip) {
      for (int j = threadIdx.x; j < D; j += blockDim.x) {
        const uint64_t elem = (env0 + (uint64_t)e) * (uint64_t)D + (uint64_t)j;
        const Philox4 w = philox4x32_10(seed ^ 0x2545F4914F6CDD1DULL, elem, t);
        const double u1 = u53(w.w[0], w.w[1]) + 1.1102230246251565e-16, u2 = (double)w.w[2] * (1.0 / 4294967296.0);
        zs[j] = (float)(sqrt(-2.0 * log(u1)) * cospi(2.0 * u2));
      }
    }
    __syncthreads();
    if (!skip) {
      float scale = 1.f;
      if (kind == 2) {  // one chi-square per observation (scipy multivariate_t: x = z_W / sqrt(chi2_df / df))
        const double g = gamma_draw(0.5 * df, seed ^ 0x9FB21C651E98DF25ULL, env0 + (uint64_t)e, t);
        scale = (float)(1.0 / sqrt(2.0 * g / df));
      }
      for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const float* row = L + (size_t)i * D;
        float acc = 0.f;
        for (int j = 0; j <= i; ++j) acc = fmaf(__ldg(row + j), zs[j], acc);
        out[(size_t)e * D + i] += acc * scale;
      }
    }
    __syncthreads();
  }
}

// N_N.sample (conjugate_rewards.py:119-127): Normal(mu, scale = tau) -> float32, rows laid out as (mu, tau, -, -)
__global__ void __launch_bounds__(256) nn_rows_kernel(const float* __restrict__ hyper, long long rows, long long row0,
                                                      unsigned long long seed, unsigned long long t,
                                                      float* __restrict__ R) {
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x) {
    const float4 h = __ldg(reinterpret_cast<const float4*>(hyper) + r);
    const Philox4 w = philox4x32_10(seed ^ 0x8CB92BA72F3D8DD7ULL, (uint64_t)(row0 + r), t);
    const double u1 = u53(w.w[0], w.w[1]) + 1.1102230246251565e-16, u2 = (double)w.w[2] * (1.0 / 4294967296.0);
    const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    R[r] = (float)((double)h.x + (double)h.y * z);
  }
}

}  // namespace colo

extern "C" int colo_sample_nn_rewards(const float* hyper, long long rows, long long row0, unsigned long long seed,
                                      unsigned long long t, float* R_out, void* stream) {
  COLO_ARG_CHECK(hyper && R_out && rows >= 0 && row0 >= 0 && (uintptr_t)hyper % 16 == 0, "hyper (16-byte aligned), R_out, rows");
  if (rows == 0) return COLO_OK;
  const long long blocks = (rows + 255) / 256;
  const long long cap = (long long)colo::sm_count() * 8;
  colo::nn_rows_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(hyper, rows, row0, seed, t, R_out);
  return colo::check_launch("nn_rows_kernel");
}

extern "C" int colo_emit_noise(float* out, const unsigned char* step_type, const int* h, long long N, int H, int D,
                               int period, int kind, double param, unsigned long long seed, unsigned long long t,
                               unsigned long long env0, void* stream) {
  COLO_ARG_CHECK(out && step_type && h && N >= 0 && D > 0 && period > 0 && period <= D && (kind == 1 || kind == 2) && param > 0,
                 "out, step_type, h, N, D, period, kind in {1,2}, param > 0");
  if (N == 0) return COLO_OK;
  const long long blocks = (N * D + 255) / 256;
  const long long cap = (long long)colo::sm_count() * 16;
  colo::emit_noise_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(out, step_type, h, N, H, D, period,
                                                                                                kind, param, seed, t, env0);
  return colo::check_launch("emit_noise_kernel");
}

extern "C" int colo_emit_noise_correlated(float* out, const unsigned char* step_type, const int* h, long long N, int H, int D,
                                          const float* chol, int kind, double df, unsigned long long seed,
                                          unsigned long long t, unsigned long long env0, void* stream) {
  COLO_ARG_CHECK(out && step_type && h && chol && N >= 0 && D > 0 && D <= 12288 && (kind == 1 || kind == 2) && df > 0,
                 "out, step_type, h, chol, N, 0 < D <= 12288, kind in {1,2}, df > 0");
  if (N == 0) return COLO_OK;
  const long long cap = (long long)colo::sm_count() * 8;
  const size_t smem = (size_t)D * sizeof(float);
  { const int es = colo::ensure_dynamic_smem((const void*)colo::emit_noise_correlated_kernel, smem); if (es != COLO_OK) return es; }
  colo::emit_noise_correlated_kernel<<<(int)(N < cap ? N : cap), 256, smem, (cudaStream_t)stream>>>(out, step_type, h, N, H, D, chol,
                                                                                                  kind, df, seed, t, env0);
  return colo::check_launch("emit_noise_correlated_kernel");
}

extern "C" int colo_sample_nig_rewards(const float* hyper, long long rows, long long row0, unsigned long long seed,
                                       unsigned long long t, float* R_out, void* stream) {
  COLO_ARG_CHECK(hyper && R_out && rows >= 0 && row0 >= 0 && (uintptr_t)hyper % 16 == 0, "hyper (16-byte aligned), R_out, rows");
  if (rows == 0) return COLO_OK;
  const long long blocks = (rows + 255) / 256;
  const long long cap = (long long)colo::sm_count() * 8;
  colo::nig_rows_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(hyper, rows, row0, seed, t, R_out);
  return colo::check_launch("nig_rows_kernel");
}

static int dirichlet_launch(const float* hyper, long long rows, int S, long long row0, unsigned long long seed,
                            unsigned long long t, float* T_out, void* stream, bool fast) {
  COLO_ARG_CHECK(hyper && T_out && rows >= 0 && S > 0 && row0 >= 0, "hyper, T_out, rows, S, row0");
  if (rows == 0) return COLO_OK;
  const long long blocks = (rows + 7) / 8;
  const long long cap = (long long)colo::sm_count() * 16;
  const int grid = (int)(blocks < cap ? blocks : cap);
  if (fast)
    colo::dirichlet_rows_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(hyper, rows, S, row0, seed, t, T_out);
  else
    colo::dirichlet_rows_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(hyper, rows, S, row0, seed, t, T_out);
  return colo::check_launch("dirichlet_rows_kernel");
}

extern "C" int colo_sample_dirichlet_rows(const float* hyper, long long rows, int S, long long row0,
                                          unsigned long long seed, unsigned long long t, float* T_out, void* stream) {
  return dirichlet_launch(hyper, rows, S, row0, seed, t, T_out, stream, false);
}

extern "C" int colo_sample_dirichlet_rows_fast(const float* hyper, long long rows, int S, long long row0,
                                               unsigned long long seed, unsigned long long t, float* T_out,
                                               void* stream) {
  return dirichlet_launch(hyper, rows, S, row0, seed, t, T_out, stream, true);
}
