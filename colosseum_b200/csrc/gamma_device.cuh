// gamma_device.cuh -- the gamma draws behind every posterior sample (posterior.cu, continuous_agents.cu): Marsaglia-Tsang
// squeeze on Philox4x32-10 counters, fp64 like numpy's standard_gamma and a single-precision variant for the hot loops.
#pragma once
#include "common.cuh"

namespace colo {

// Philox stream separation: every consumer of the caller's `seed` keys its blocks with its own constant, so that no
// (key, counter) pair is shared with the interaction kernels -- the env step draws with key = seed and counter (env
// index, step), the agents with key = seed ^ 0x9E3779B97F4A7C15 (agents.cu) -- nor between the gamma body, the
// alpha < 1 boost, the NIG draws and the emission noise (callers fold their own constant into `seed` on top).
constexpr uint64_t kGammaBodyKey = 0xA0761D6478BD642FULL, kGammaBoostKey = 0xE7037ED1A0B428DBULL;

__device__ __forceinline__ double gamma_draw(double alpha, uint64_t seed, uint64_t elem, uint64_t t) {
  if (!(alpha > 0.0)) return 0.0;
  const double a = alpha < 1.0 ? alpha + 1.0 : alpha;
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  double g = 0.0;
  for (uint64_t attempt = 0; attempt < 64; ++attempt) {
    const Philox4 w = philox4x32_10(seed ^ kGammaBodyKey, elem, (t << 8) | attempt);
    // Box-Muller normal from two 53/32-bit uniforms, one more uniform for the squeeze test
    const double u1 = (u53(w.w[0], w.w[1]) + 1.1102230246251565e-16), u2 = (double)w.w[2] * (1.0 / 4294967296.0);
    const double x = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    const double v0 = 1.0 + c * x;
    if (v0 <= 0.0) continue;
    const double v = v0 * v0 * v0;
    const double u = ((double)w.w[3] + 0.5) * (1.0 / 4294967296.0);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) {
      g = d * v;
      break;
    }
  }
  if (alpha < 1.0) {  // gamma(alpha) = gamma(alpha + 1) * U^(1/alpha)
    const Philox4 w = philox4x32_10(seed ^ kGammaBoostKey, elem, t);
    const double u = u53(w.w[0], w.w[1]) + 1.1102230246251565e-16;
    g *= pow(u, 1.0 / alpha);
  }
  return g;
}

// Single-precision variant of gamma_draw for the hot loop of the batched PSRL agents (N*S*A*S draws per episode): one
// Philox block per attempt feeds the normal (Box-Muller on two 24/32-bit uniforms), the squeeze test and the
// alpha < 1 boost; log / pow / cos go through the SFU.  The normal is truncated at 5.8 sigma (24-bit uniform) and the
// transcendental error is ~1e-6 relative: invisible to the float32 result the reference keeps, but not the fp64
// arithmetic numpy uses -- hence a separate entry point (colo_sample_dirichlet_rows_fast).
__device__ __forceinline__ float gamma_draw_fast(float alpha, uint64_t seed, uint64_t elem, uint64_t t) {
  if (!(alpha > 0.f)) return 0.f;
  const float a = alpha < 1.f ? alpha + 1.f : alpha;
  const float d = a - (1.f / 3.f), c = rsqrtf(9.f * d);
  for (uint64_t attempt = 0; attempt < 64; ++attempt) {
    const Philox4 w = philox4x32_10(seed ^ kGammaBodyKey, elem, (t << 8) | attempt);
    const float u1 = ((float)(w.w[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = (float)w.w[1] * (1.0f / 4294967296.0f);
    const float x = sqrtf(-2.f * __logf(u1)) * cospif(2.f * u2);
    const float v0 = 1.f + c * x;
    if (v0 <= 0.f) continue;
    const float v = v0 * v0 * v0;
    const float u = ((float)(w.w[2] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    if (__logf(u) < 0.5f * x * x + d - d * v + d * __logf(v)) {
      float g = d * v;
      if (alpha < 1.f) {  // gamma(alpha) = gamma(alpha + 1) * U^(1/alpha)
        const float ub = ((float)(w.w[3] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        g *= exp2f(__log2f(ub) / alpha);
      }
      return g;
    }
  }
  return 0.f;
}

}  // namespace colo
