// continuous_agents.cu -- the model-based agents of the continuous (infinite-horizon) setting as N device loops
// (SURVEY.md section 8(f)-1/-4):
//
//   UCRL2Continuous   colosseum/agent/agents/infinite_horizon/ucrl2.py:34-357
//   PSRLContinuous    colosseum/agent/agents/infinite_horizon/posterior_sampling.py:117-452
//
// Both agents cut the interaction into ARTIFICIAL EPISODES that end at loop-dependent times (a state-action pair doubled
// its visit count), and re-plan at every episode end (extended value iteration / posterior sample + discounted value
// iteration).  The batch therefore advances in rounds:
//
//   1. *_steps kernel      one thread per loop runs MDPLoop.run's body (experiment/agent_mdp_interaction.py:238-298):
//                          select_action, BaseMDP.step, step_update, is_episode_end -- until the loop reaches the target
//                          time or its episode ends (`ended[i] = 1`: the loop then waits);
//   2. the host reads `ended`, lists the loops that wait, and launches the episode-end kernels FOR THAT LIST ONLY
//                          (bounds / posterior sample -> batched planner -> model update), which clear the flag;
//   3. repeat until every loop is at the target time.
//
// Every loop keeps its own interaction time t[i] (the Philox counter of its draws and the reference's `time` argument),
// so a loop's trajectory does not depend on which other loops share its batch or on how the rounds fall.
// Arithmetic follows numpy's types sub-expression by sub-expression (NEP 50), with round-to-nearest intrinsics, so the
// CPU restatement the tests compare with agrees bit for bit on trajectories and model tables.
#include "agent_device.cuh"

namespace colo {

// ------------------------------------------------------------------------------------------------ UCRL2Continuous
// step_update (ucrl2.py:183-199): N[s,a,s'] += 1, the reward and (never LAST in a continuous MDP) the next state are
// appended to the episode's per-(s,a) lists -- here one time-ordered log per loop, which model_update reads back in the
// same per-(s,a) order.  is_episode_end (:173-181): nu_k >= max(1, N[s,a].sum() - nu_k).
__global__ void __launch_bounds__(128) ucrl2_steps_kernel(const colo_mdp_tables tb, const colo_ucrl2_args p,
                                                          long long t_target) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N) return;
  if (p.ended[i] != 0) return;
  const int S = tb.S, A = tb.A;
  const size_t SA = (size_t)S * A;
  const float* Q = p.Q + i * SA;
  int* Nsas = p.Nsas + i * SA * S;
  int* Nsa = p.Nsa + i * SA;
  int* nu = p.nu + i * SA;
  int* log = p.ep_log + (size_t)i * p.log_cap * 2;
  int s = p.state[i];
  long long t = p.t[i];
  int len = p.ep_len[i];
  double cum = p.cum_reward[i];
  int flag = 0;
  while (t < t_target) {
    if (len >= p.log_cap) {  // an episode longer than the caller's optimisation horizon allows
      flag = 2;
      break;
    }
    const Philox4 we = philox4x32_10(p.seed, p.env0 + (uint64_t)i, (uint64_t)t);
    const Philox4 wa = philox4x32_10(p.seed ^ kAgentKey, p.env0 + (uint64_t)i, (uint64_t)t);
    const int a = select_action(Q + (size_t)s * A, A, p.epsilon_greedy, wa);
    const size_t sa = (size_t)s * A + a;
    const int n_sa = Nsa[sa] + 1, nu_k = nu[sa] + 1;
    const Step st = env_succ(tb, s, a, u53(we.w[0], we.w[1]));
    const float r = reward_from_class(tb, st.cls, u24(we.w[2]));
    Nsas[sa * S + st.nxt] += 1;
    Nsa[sa] = n_sa;
    nu[sa] = nu_k;
    log[2 * len] = (int)sa;
    log[2 * len + 1] = __float_as_int(r);
    ++len;
    cum = __dadd_rn(cum, (double)r);
    if (p.trace) {
      const long long k = t - p.trace_t0;
      if (k >= 0 && k < p.trace_steps) {
        int* tr = p.trace + ((size_t)k * p.N + i) * 4;
        tr[0] = s; tr[1] = a; tr[2] = st.nxt; tr[3] = __float_as_int(r);
      }
    }
    s = st.nxt;
    ++t;
    const int before = n_sa - nu_k;
    if (nu_k >= (before > 1 ? before : 1)) {
      flag = 1;
      break;
    }
  }
  p.state[i] = s;
  p.t[i] = t;
  p.ep_len[i] = len;
  p.cum_reward[i] = cum;
  if (flag) p.ended[i] = flag;
}

// episode_end_update, first half (ucrl2.py:183-192, :252-311): episode += 1, delta = 1 / sqrt(iteration + 1), then the
// confidence bounds solve_optimistic_model passes to extended_value_iteration -- computed from the counts of THIS
// episode's end but the iteration counter, P and variance proxies of the previous model update (the reference updates
// the model after planning).  beta_p is the [s, a, 0] entry of the reference's array: _max_proba reads
// `(p[...] + beta / 2)[0]` (dynamic_programming/infinite_horizon.py:230), the first entry of the slice it is handed.
__global__ void __launch_bounds__(128) ucrl2_bounds_kernel(const colo_ucrl2_args p, int S, int A, const int* __restrict__ index,
                                                           int m, double alpha_r, double alpha_p, double r_max,
                                                           int bernstein_p, double* __restrict__ beta_r,
                                                           double* __restrict__ beta_p) {
  const size_t SA = (size_t)S * A;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)m * (long long)SA) return;
  const int k = (int)(gid / (long long)SA);
  const size_t sa = (size_t)(gid % (long long)SA);
  const size_t i = (size_t)index[k];
  const long long it = p.iteration[i];
  const double delta = __ddiv_rn(1.0, __dsqrt_rn((double)(it + 1)));
  const int nb = p.Nsa[i * SA + sa];
  const double n1 = (double)(nb > 1 ? nb : 1);
  // _chernoff (ucrl2.py:22-24): range * sqrt(sqrt_C * log(log_C * (it + 1) / delta) / max(1, N))
  const double Lr = log(__ddiv_rn((double)(2LL * S * A * (it + 1)), delta));
  beta_r[(size_t)k * SA + sa] = __dmul_rn(alpha_r, __dmul_rn(r_max, __dsqrt_rn(__ddiv_rn(__dmul_rn(3.5, Lr), n1))));
  double bp;
  if (!bernstein_p) {
    const double Lp = log(__ddiv_rn((double)(2LL * A * (it + 1)), delta));
    bp = __dmul_rn(alpha_p, __dsqrt_rn(__ddiv_rn(__dmul_rn((double)(14LL * S), Lp), n1)));
  } else {
    // bernstein (ucrl2.py:27-30, :294-309): var_p = P (1 - P) in float32, 14 var_p (float32) / N (float64)
    const double nm1 = (double)(nb - 1 > 1 ? nb - 1 : 1);
    const float P0 = p.P[(i * SA + sa) * S];
    const float var_p = __fmul_rn(P0, __fsub_rn(1.0f, P0));
    const double L = log(__ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(2.0, (double)S), (double)A), (double)(it + 1)), delta));
    const double Aterm = __dmul_rn(__ddiv_rn((double)__fmul_rn(14.0f, var_p), n1), L);
    const double Bterm = __dmul_rn(__ddiv_rn(49.0, __dmul_rn(3.0, nm1)), L);
    bp = __dadd_rn(__dmul_rn(__dsqrt_rn(alpha_p), __dsqrt_rn(Aterm)), __dmul_rn(alpha_p, Bterm));
  }
  beta_p[(size_t)k * SA + sa] = bp;
  if (sa == 0) {
    p.delta[i] = delta;
    p.episode[i] += 1;
  }
}

// model_update, the transition half (ucrl2.py:220-221): P[s,a] = N[s,a] / N[s,a].sum() (int32 / int64 = float64, stored
// float32) for the pairs visited in the episode.  One warp per (listed loop, s, a); clears the episode's visit count.
__global__ void __launch_bounds__(256) ucrl2_rows_kernel(const colo_ucrl2_args p, int S, int A,
                                                         const int* __restrict__ index, int m) {
  const size_t SA = (size_t)S * A;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)m * (long long)SA) return;
  const size_t i = (size_t)index[(int)(w / (long long)SA)];
  const size_t sa = (size_t)(w % (long long)SA);
  int* nu = p.nu + i * SA + sa;
  if (*nu == 0) return;
  const double tot = (double)p.Nsa[i * SA + sa];
  const int* __restrict__ n = p.Nsas + (i * SA + sa) * S;
  float* __restrict__ P = p.P + (i * SA + sa) * S;
  for (int j = lane; j < S; j += 32) P[j] = (float)__ddiv_rn((double)n[j], tot);
  __syncwarp();
  if (lane == 0) *nu = 0;
}

// model_update, the reward half (ucrl2.py:201-218), one thread per listed loop over its time-ordered log: for the j-th
// reward of (s, a) in the episode, scale_f = N[s,a].sum() + j (the count ALREADY holds the episode's visits, sic), then
//   est *= scale_f / (scale_f + 1.0);  est += r / (scale_f + 1.0)          float32 element (*|+) float64, stored float32
//   var += (r - old_est) * (r - est)                                        python float and float32: float32 throughout
//   hold *= scale_f / (scale_f + 1.0); hold += 1 / (scale_f + 1)
// `seen` (i32 [N,S,A], all zero between calls) counts the rewards of (s, a) already consumed.  Clears the episode.
__global__ void __launch_bounds__(64) ucrl2_rewards_kernel(const colo_ucrl2_args p, int S, int A,
                                                           const int* __restrict__ index, int m) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= m) return;
  const size_t i = (size_t)index[k];
  const size_t SA = (size_t)S * A;
  const int* __restrict__ log = p.ep_log + i * (size_t)p.log_cap * 2;
  const int len = p.ep_len[i];
  const int* __restrict__ Nsa = p.Nsa + i * SA;
  int* seen = p.seen + i * SA;
  float* est = p.est_r + i * SA;
  float* var = p.var_r + i * SA;
  float* hold = p.hold + i * SA;
  for (int e = 0; e < len; ++e) {
    const int sa = log[2 * e];
    const float r = __int_as_float(log[2 * e + 1]);
    const int j = seen[sa] + 1;
    seen[sa] = j;
    const double sf = (double)((long long)Nsa[sa] + j);
    const double sf1 = __dadd_rn(sf, 1.0);
    const double ratio = __ddiv_rn(sf, sf1);
    const float old = est[sa];
    float x = (float)__dmul_rn((double)old, ratio);
    x = (float)__dadd_rn((double)x, __ddiv_rn((double)r, sf1));
    est[sa] = x;
    var[sa] = __fadd_rn(var[sa], __fmul_rn(__fsub_rn(r, old), __fsub_rn(r, x)));
    float hd = (float)__dmul_rn((double)hold[sa], ratio);
    hd = (float)__dadd_rn((double)hd, __ddiv_rn(1.0, sf1));
    hold[sa] = hd;
  }
  for (int e = 0; e < len; ++e) seen[log[2 * e]] = 0;
  p.iteration[i] += len;
  p.ep_len[i] = 0;
  p.ended[i] = 0;
}

static int ucrl2_check(const colo_ucrl2_args* a) {
  COLO_ARG_CHECK(a, "args are NULL");
  COLO_ARG_CHECK(a->N >= 0, "N >= 0");
  if (a->N == 0) return COLO_OK;
  COLO_ARG_CHECK(a->state && a->t && a->cum_reward && a->Q && a->Nsas && a->Nsa && a->P && a->est_r && a->var_r &&
                     a->hold && a->nu && a->seen && a->ep_len && a->ep_log && a->ended && a->iteration && a->episode &&
                     a->delta,
                 "every table pointer of colo_ucrl2_args must be set");
  COLO_ARG_CHECK(a->log_cap >= 1, "log_cap >= 1");
  return COLO_OK;
}

}  // namespace colo

extern "C" {

int colo_ucrl2_steps(const colo_mdp_tables* tb, const colo_ucrl2_args* a, long long t_target, void* stream) {
  COLO_ARG_CHECK(tb, "tables are NULL");
  int r = colo::ucrl2_check(a);
  if (r != COLO_OK || a->N == 0) return r;
  COLO_ARG_CHECK(tb->S > 0 && tb->A > 0 && tb->H == 0 && tb->rew_q && tb->n_cls > 0 && tb->nq >= 2,
                 "UCRL2Continuous needs the tables of a continuous MDP (H == 0)");
  COLO_ARG_CHECK(tb->succ_cum && tb->succ_idx && tb->succ_len && tb->Ksucc > 0, "successor tables");
  const int grid = (int)((a->N + 127) / 128);
  colo::ucrl2_steps_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*tb, *a, t_target);
  return colo::check_launch("ucrl2_steps_kernel");
}

int colo_ucrl2_bounds(const colo_ucrl2_args* a, int S, int A, const int* index, int m, double alpha_r, double alpha_p,
                      double r_max, int bernstein_p, double* beta_r, double* beta_p, void* stream) {
  int r = colo::ucrl2_check(a);
  if (r != COLO_OK) return r;
  COLO_ARG_CHECK(S > 0 && A > 0 && m >= 0 && (m == 0 || (index && beta_r && beta_p)), "S, A, m, index, beta_r, beta_p");
  if (m == 0) return COLO_OK;
  const long long n = (long long)m * S * A;
  colo::ucrl2_bounds_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      *a, S, A, index, m, alpha_r, alpha_p, r_max, bernstein_p, beta_r, beta_p);
  return colo::check_launch("ucrl2_bounds_kernel");
}

int colo_ucrl2_model_update(const colo_ucrl2_args* a, int S, int A, const int* index, int m, void* stream) {
  int r = colo::ucrl2_check(a);
  if (r != COLO_OK) return r;
  COLO_ARG_CHECK(S > 0 && A > 0 && m >= 0 && (m == 0 || index), "S, A, m, index");
  if (m == 0) return COLO_OK;
  const long long warps = (long long)m * S * A;
  colo::ucrl2_rows_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a, S, A, index, m);
  r = colo::check_launch("ucrl2_rows_kernel");
  if (r != COLO_OK) return r;
  colo::ucrl2_rewards_kernel<<<(m + 63) / 64, 64, 0, (cudaStream_t)stream>>>(*a, S, A, index, m);
  return colo::check_launch("ucrl2_rewards_kernel");
}

}  // extern "C"
