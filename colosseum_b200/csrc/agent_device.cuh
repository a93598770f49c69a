// agent_device.cuh -- device helpers shared by the agent-loop kernels (agents.cu, continuous_agents.cu): the env step on the
// reference's own successor sampler (colosseum/mdp/utils/custom_samplers.py:49-72), the reward draw
// (colosseum/mdp/base.py:1187-1207 as a quantile table) and QValuesActor.select_action
// (colosseum/agent/actors/Q_values_actor.py:58-82).  Round-to-nearest intrinsics only: the CPU oracle computes the same
// IEEE operations in the same order.
#pragma once
#include "common.cuh"

namespace colo {

constexpr unsigned long long kAgentKey = 0x9E3779B97F4A7C15ULL;  // the agent's Philox key = seed ^ kAgentKey

struct Step {
  int nxt, cls;
};

__device__ __forceinline__ int bisect_count_d(const double* __restrict__ cum, int n, double x) {
  int pos = 0;
  for (int k = 0; k < n - 1; ++k) pos += (__ldg(cum + k) <= x) ? 1 : 0;
  return pos;
}

__device__ __forceinline__ Step env_succ(const colo_mdp_tables& tb, int s, int a, double u) {
  const size_t sa = (size_t)s * tb.A + a;
  const size_t base = sa * tb.Ksucc;
  const int n = __ldg(tb.succ_len + sa);
  int pos = 0;
  if (n > 1) {
    const double total = __ldg(tb.succ_cum + base + n - 1) + 0.0;
    pos = bisect_count_d(tb.succ_cum + base, n, __dmul_rn(u, total));
  }
  Step r;
  r.nxt = __ldg(tb.succ_idx + base + pos);
  r.cls = tb.rew_cls_succ ? __ldg(tb.rew_cls_succ + base + pos) : 0;
  return r;
}

__device__ __forceinline__ float reward_from_class(const colo_mdp_tables& tb, int cls, float u) {
  const float* q = tb.rew_q + (size_t)cls * tb.nq;
  const float t = __fmul_rn(u, (float)(tb.nq - 1));
  int i = (int)t;
  i = i > tb.nq - 2 ? tb.nq - 2 : i;
  const float f = __fsub_rn(t, (float)i);
  const float q0 = __ldg(q + i), q1 = __ldg(q + i + 1);
  const float r0 = fmaf(f, __fsub_rn(q1, q0), q0);
  return fmaf(r0, __fsub_rn(tb.rmax, tb.rmin), -tb.rmin);
}

__device__ __forceinline__ int start_state(const colo_mdp_tables& tb, double u) {
  if (tb.n_start == 1) return __ldg(tb.start_idx);
  const double total = __ldg(tb.start_cum + tb.n_start - 1) + 0.0;
  return __ldg(tb.start_idx + bisect_count_d(tb.start_cum, tb.n_start, __dmul_rn(u, total)));
}

// QValuesActor.select_action on one row of q-values
__device__ __forceinline__ int select_action(const float* __restrict__ q, int A, double eps, const Philox4& w) {
  if (eps >= 0.0 && (double)u24(w.w[0]) < eps) return act_from_word(w.w[1], A);
  float best = q[0];
  int ties = 1;
  for (int a = 1; a < A; ++a) {
    const float v = q[a];
    if (v > best) {
      best = v;
      ties = 1;
    } else if (v == best) {
      ++ties;
    }
  }
  int k = act_from_word(w.w[1], ties);
  for (int a = 0; a < A; ++a)
    if (q[a] == best && k-- == 0) return a;
  return A - 1;
}

constexpr unsigned long long kBoltzmannKey = 0x94D049BB133111EBULL;  // the Boltzmann draw's Philox key = seed ^ this

// the exploration parameters at the actor's interaction count `total` (colo_actor_args: schedule tables or constants)
__device__ __forceinline__ double actor_epsilon(const colo_actor_args& ac, double eps_const, long long total) {
  if (ac.epsilon_schedule == nullptr) return eps_const;
  long long k = total - ac.t0;
  k = k < 0 ? 0 : (k >= ac.len ? ac.len - 1 : k);
  return __ldg(ac.epsilon_schedule + k);
}
__device__ __forceinline__ double actor_temperature(const colo_actor_args& ac, long long total) {
  if (ac.temperature_schedule == nullptr) return ac.boltzmann_temperature;
  long long k = total - ac.t0;
  k = k < 0 ? 0 : (k >= ac.len ? ac.len - 1 : k);
  return __ldg(ac.temperature_schedule + k);
}

// Boltzmann exploration (Q_values_actor.py:73-78): q = exp(temperature * q) in float32, p = q / q.sum() in float32,
// numpy's choice(p): cdf = cumsum(p) in float64, normalised by its last entry, index = #(cdf <= u)
__device__ __forceinline__ float boltz_weight(float temp, float q) {
  return (float)exp((double)__fmul_rn(temp, q));
}
__device__ __forceinline__ int boltzmann_action(const float* __restrict__ q, int A, double temperature, double u) {
  const float temp = (float)temperature;
  float sum = 0.f;
  for (int a = 0; a < A; ++a) sum = __fadd_rn(sum, boltz_weight(temp, q[a]));
  double tot = 0.0;
  for (int a = 0; a < A; ++a) tot = __dadd_rn(tot, (double)__fdiv_rn(boltz_weight(temp, q[a]), sum));
  double run = 0.0;
  int idx = 0;
  for (int a = 0; a < A; ++a) {
    run = __dadd_rn(run, (double)__fdiv_rn(boltz_weight(temp, q[a]), sum));
    idx += (__ddiv_rn(run, tot) <= u) ? 1 : 0;
  }
  return idx < A ? idx : A - 1;
}

// QValuesActor.select_action with the full exploration set; A_random = the range of the epsilon-greedy draw
__device__ __forceinline__ int actor_select(const float* __restrict__ q, int A, int A_random, double eps_const,
                                            const colo_actor_args& ac, long long total, const Philox4& w,
                                            unsigned long long seed, unsigned long long loop) {
  const double eps = actor_epsilon(ac, eps_const, total);
  if (eps >= 0.0 && (double)u24(w.w[0]) < eps) return act_from_word(w.w[1], A_random);
  if (ac.boltzmann) {
    const Philox4 wb = philox4x32_10(seed ^ kBoltzmannKey, loop, (uint64_t)total);
    return boltzmann_action(q, A, actor_temperature(ac, total), u53(wb.w[0], wb.w[1]));
  }
  return select_action(q, A, -1.0, w);
}

}  // namespace colo
