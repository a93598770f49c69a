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
#include "gamma_device.cuh"

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
    const int a = actor_select(Q + (size_t)s * A, A, A, p.epsilon_greedy, p.actor, t, wa, p.seed, p.env0 + (uint64_t)i);
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

// ------------------------------------------------------------------------------------------------- PSRLContinuous
// select_action (posterior_sampling.py:389-392) = QValuesActor.select_action on the EXTENDED action set (A * psi columns:
// psi sampled transition rows per real action), then extended_action_to_real: real = int(action / psi) -- applied to the
// epsilon-greedy draw too, which the actor takes from range(A) (sic).  step_update (:394-412): BayesianMDPModel.step_update
// (N_NIG / N_N reward posterior, Dirichlet count += 1; bayesian_model.py:78-92), N[s,a,s'] += 1, the episode's
// transition list (here its length nu).  is_episode_end (:333-345, min_steps_before_new_episode = 0):
// N_tau >= 2 (N_tau - nu_k).  (self.M is written by the reference and never read: not kept.)
__global__ void __launch_bounds__(128) psrlc_steps_kernel(const colo_mdp_tables tb, const colo_psrlc_args p,
                                                          long long t_target) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.N) return;
  if (p.ended[i] != 0) return;
  const int S = tb.S, A = tb.A, psi = p.psi, AE = A * psi;
  const size_t SA = (size_t)S * A;
  const float* Q = p.Q + i * (size_t)S * AE;
  float* dir = p.dir_hyper + i * SA * S;
  float* nig = p.nig_hyper + i * SA * 4;
  int* Nsas = p.Nsas + i * SA * S;
  int* Nsa = p.Nsa + i * SA;
  int* nu = p.nu + i * SA;
  int s = p.state[i];
  long long t = p.t[i];
  double cum = p.cum_reward[i];
  int flag = 0;
  while (t < t_target) {
    const Philox4 we = philox4x32_10(p.seed, p.env0 + (uint64_t)i, (uint64_t)t);
    const Philox4 wa = philox4x32_10(p.seed ^ kAgentKey, p.env0 + (uint64_t)i, (uint64_t)t);
    const int a_ext = actor_select(Q + (size_t)s * AE, AE, A, p.epsilon_greedy, p.actor, t, wa, p.seed, p.env0 + (uint64_t)i);
    const int a = a_ext / psi;
    const size_t sa = (size_t)s * A + a;
    float* hp = nig + sa * 4;
    const float4 hp0 = *reinterpret_cast<const float4*>(hp);
    const int n_sa = Nsa[sa] + 1, nu_k = nu[sa] + 1;
    const Step st = env_succ(tb, s, a, u53(we.w[0], we.w[1]));
    const float r = reward_from_class(tb, st.cls, u24(we.w[2]));
    const float mu0 = hp0.x, l0 = hp0.y, a0 = hp0.z, b0 = hp0.w;
    if (p.reward_model == 1) {  // N_N.update_sa (conjugate_rewards.py:112-117): float32 throughout
      const float t1 = __fadd_rn(l0, 1.0f);
      hp[0] = __fdiv_rn(__fadd_rn(__fmul_rn(mu0, l0), r), t1);
      hp[1] = t1;
    } else {  // N_NIG.update_sa with one reward (conjugate_rewards.py:56-74)
      const double y = (double)r;
      const float l1 = __fadd_rn(l0, 1.0f);
      const double mu1 = __ddiv_rn(__dadd_rn((double)__fmul_rn(l0, mu0), y), (double)l1);
      const double dy = __dsub_rn(y, (double)mu0);
      const double disc = __ddiv_rn(__dmul_rn((double)l0, __dmul_rn(dy, dy)), (double)l1);
      hp[0] = (float)mu1;
      hp[1] = l1;
      hp[2] = __fadd_rn(a0, 0.5f);
      hp[3] = (float)__dadd_rn((double)b0, __dmul_rn(0.5, __dadd_rn(0.0, disc)));
    }
    float* dc = dir + sa * S + st.nxt;
    *dc = __fadd_rn(*dc, 1.0f);
    Nsas[sa * S + st.nxt] += 1;
    Nsa[sa] = n_sa;
    nu[sa] = nu_k;
    cum = __dadd_rn(cum, (double)r);
    if (p.trace) {
      const long long k = t - p.trace_t0;
      if (k >= 0 && k < p.trace_steps) {
        int* tr = p.trace + ((size_t)k * p.N + i) * 4;
        tr[0] = s; tr[1] = a_ext; tr[2] = st.nxt; tr[3] = __float_as_int(r);
      }
    }
    s = st.nxt;
    ++t;
    if (n_sa >= 2 * (n_sa - nu_k)) {
      flag = 1;
      break;
    }
  }
  p.state[i] = s;
  p.t[i] = t;
  p.cum_reward[i] = cum;
  if (flag) p.ended[i] = flag;
}

constexpr uint64_t kSimpleSamplingKey = 0xC2B2AE3D27D4EB4FULL;  // the z draws of optimistic_sampling

// optimistic_sampling (posterior_sampling.py:414-447) for the listed loops: T_ext[k, s, a * psi + q, :] for q < psi.
//   N[s,a].sum() >= eta : a Dirichlet posterior sample of the row (M_DIR._sample, conjugate_transitions.py:48-54:
//                         float32 gammas, r / (1e-5 + sum r));
//   otherwise           : "simple sampling", P_minus = P_hat - min(sqrt(3 P_hat log(4S) / N) + 3 log(4S) / N, P_hat) with
//                         the missing mass 1 - sum(P_minus) put on ONE state z_q drawn per sample (the same for every row
//                         of that sample; the reference's self._rng.randint(S)), in float64, stored float32.
// One warp per (listed loop, s, a); the psi samples of a row reuse the row statistics.  Draw counters: posterior rows
// (seed; ((env0+i) S A + sa) S + j, episode * 64 + q), z (seed ^ kSimpleSamplingKey; env0+i, episode * 64 + q).
template <bool FAST>
__global__ void __launch_bounds__(256) psrlc_sample_kernel(const colo_psrlc_args p, int S, int A,
                                                           const int* __restrict__ index, int m, double eta,
                                                           float* __restrict__ T_ext) {
  const size_t SA = (size_t)S * A;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= (long long)m * (long long)SA) return;
  const int k = (int)(w / (long long)SA);
  const size_t sa = (size_t)(w % (long long)SA);
  const size_t s = sa / A, a = sa % A;
  const size_t i = (size_t)index[k];
  const int psi = p.psi;
  const uint64_t ep = (uint64_t)p.episode[i];
  const int nsum = p.Nsa[i * SA + sa];
  float* out0 = T_ext + (((size_t)k * S + s) * ((size_t)A * psi) + a * psi) * S;
  if ((double)nsum >= eta) {
    const float* __restrict__ h = p.dir_hyper + (i * SA + sa) * S;
    for (int q = 0; q < psi; ++q) {
      float* out = out0 + (size_t)q * S;
      float sum = 0.f;
      for (int j = lane; j < S; j += 32) {
        const uint64_t elem = ((p.env0 + (uint64_t)i) * SA + sa) * (uint64_t)S + (uint64_t)j;
        const float g = FAST ? gamma_draw_fast(h[j], p.seed, elem, ep * 64 + q)
                             : (float)gamma_draw((double)h[j], p.seed, elem, ep * 64 + q);
        out[j] = g;
        sum += g;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / (1e-5f + sum);
      for (int j = lane; j < S; j += 32) out[j] *= inv;
    }
  } else {
    const int* __restrict__ n = p.Nsas + (i * SA + sa) * S;
    const double tot = (double)(nsum > 1 ? nsum : 1);
    const double L = log((double)(4 * S));
    double part = 0.0;
    for (int j = lane; j < S; j += 32) {
      const int nj = n[j];
      const double ph = __ddiv_rn((double)nj, tot);
      const double nn = (double)(nj > 1 ? nj : 1);
      const double rad = __dadd_rn(__dsqrt_rn(__ddiv_rn(__dmul_rn(__dmul_rn(3.0, ph), L), nn)),
                                   __ddiv_rn(__dmul_rn(3.0, L), nn));
      const double pm = __dsub_rn(ph, rad < ph ? rad : ph);
      part += pm;
    }
    const double summing = 1.0 - warp_sum(part);
    for (int q = 0; q < psi; ++q) {
      const Philox4 wz = philox4x32_10(p.seed ^ kSimpleSamplingKey, p.env0 + (uint64_t)i, ep * 64 + q);
      const int z = act_from_word(wz.w[0], S);
      float* out = out0 + (size_t)q * S;
      for (int j = lane; j < S; j += 32) {
        const int nj = n[j];
        const double ph = __ddiv_rn((double)nj, tot);
        const double nn = (double)(nj > 1 ? nj : 1);
        const double rad = __dadd_rn(__dsqrt_rn(__ddiv_rn(__dmul_rn(__dmul_rn(3.0, ph), L), nn)),
                                     __ddiv_rn(__dmul_rn(3.0, L), nn));
        double pm = __dsub_rn(ph, rad < ph ? rad : ph);
        if (j == z) pm += summing;
        out[j] = (float)pm;
      }
    }
  }
}

// sample_R (N_NIG.sample / N_N.sample, conjugate_rewards.py:76-92, :119-127) for the listed loops, tiled psi times along
// the extended action axis: R_ext[k, s, c] = R[s, c % A] (np.tile(R, (1, psi)), posterior_sampling.py:371-372 -- sic:
// transitions are laid out action-major (c // psi is the real action), rewards sample-major); optionally
// R = maximum(r_max, R) first (truncate_reward_with_max, :369-370, sic).  One thread per (listed loop, s, a).
__global__ void __launch_bounds__(128) psrlc_rewards_kernel(const colo_psrlc_args p, int S, int A,
                                                            const int* __restrict__ index, int m, int truncate,
                                                            float r_max, float* __restrict__ R_ext) {
  const size_t SA = (size_t)S * A;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)m * (long long)SA) return;
  const int k = (int)(gid / (long long)SA);
  const size_t sa = (size_t)(gid % (long long)SA);
  const size_t s = sa / A, a = sa % A;
  const size_t i = (size_t)index[k];
  const uint64_t ep = (uint64_t)p.episode[i];
  const float4 h = *reinterpret_cast<const float4*>(p.nig_hyper + (i * SA + sa) * 4);
  const uint64_t elem = (p.env0 + (uint64_t)i) * SA + sa;
  const Philox4 w = philox4x32_10(p.seed ^ 0x8CB92BA72F3D8DD7ULL, elem, ep);
  const double u1 = u53(w.w[0], w.w[1]) + 1.1102230246251565e-16, u2 = (double)w.w[2] * (1.0 / 4294967296.0);
  const double zn = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
  float r;
  if (p.reward_model == 1) {
    r = (float)((double)h.x + (double)h.y * zn);
  } else {
    const float tau = (float)(gamma_draw((double)h.z, p.seed ^ 0xD1B54A32D192ED03ULL, elem, ep) / (double)h.w);
    const float var = 1.0f / (h.y * tau);
    r = (float)((double)h.x + sqrt((double)var) * zn);
  }
  if (truncate) r = fmaxf(r_max, r);
  float* out = R_ext + ((size_t)k * S + s) * ((size_t)A * p.psi);
  for (int c = (int)a; c < A * p.psi; c += A) out[c] = r;
}

// end of episode_end_update for the listed loops (:376): the episode's transition lists are dropped; the flag clears.
__global__ void __launch_bounds__(256) psrlc_finish_kernel(const colo_psrlc_args p, int S, int A,
                                                           const int* __restrict__ index, int m) {
  const size_t SA = (size_t)S * A;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)m * (long long)SA) return;
  const size_t i = (size_t)index[(int)(gid / (long long)SA)];
  const size_t sa = (size_t)(gid % (long long)SA);
  p.nu[i * SA + sa] = 0;
  if (sa == 0) {
    p.ended[i] = 0;
    p.episode[i] += 1;
  }
}

static int psrlc_check(const colo_psrlc_args* a) {
  COLO_ARG_CHECK(a, "args are NULL");
  COLO_ARG_CHECK(a->N >= 0 && a->psi >= 1 && a->psi <= 64, "N >= 0, 1 <= psi <= 64");
  if (a->N == 0) return COLO_OK;
  COLO_ARG_CHECK(a->state && a->t && a->cum_reward && a->Q && a->dir_hyper && a->nig_hyper && a->Nsas && a->Nsa &&
                     a->nu && a->ended && a->episode,
                 "every table pointer of colo_psrlc_args must be set");
  COLO_ARG_CHECK((uintptr_t)a->nig_hyper % 16 == 0, "nig_hyper must be 16-byte aligned (rows are read as one float4)");
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

int colo_psrlc_steps(const colo_mdp_tables* tb, const colo_psrlc_args* a, long long t_target, void* stream) {
  COLO_ARG_CHECK(tb, "tables are NULL");
  int r = colo::psrlc_check(a);
  if (r != COLO_OK || a->N == 0) return r;
  COLO_ARG_CHECK(tb->S > 0 && tb->A > 0 && tb->H == 0 && tb->rew_q && tb->n_cls > 0 && tb->nq >= 2,
                 "PSRLContinuous needs the tables of a continuous MDP (H == 0)");
  COLO_ARG_CHECK(tb->succ_cum && tb->succ_idx && tb->succ_len && tb->Ksucc > 0, "successor tables");
  const int grid = (int)((a->N + 127) / 128);
  colo::psrlc_steps_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(*tb, *a, t_target);
  return colo::check_launch("psrlc_steps_kernel");
}

int colo_psrlc_sample_models(const colo_psrlc_args* a, int S, int A, const int* index, int m, double eta, int truncate,
                             float r_max, int fast, float* T_ext, float* R_ext, void* stream) {
  int r = colo::psrlc_check(a);
  if (r != COLO_OK) return r;
  COLO_ARG_CHECK(S > 0 && A > 0 && m >= 0 && (m == 0 || (index && T_ext && R_ext)), "S, A, m, index, T_ext, R_ext");
  if (m == 0) return COLO_OK;
  const long long warps = (long long)m * S * A;
  const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
  if (fast)
    colo::psrlc_sample_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(*a, S, A, index, m, eta, T_ext);
  else
    colo::psrlc_sample_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(*a, S, A, index, m, eta, T_ext);
  r = colo::check_launch("psrlc_sample_kernel");
  if (r != COLO_OK) return r;
  colo::psrlc_rewards_kernel<<<(unsigned)((warps + 127) / 128), 128, 0, (cudaStream_t)stream>>>(*a, S, A, index, m, truncate,
                                                                                             r_max, R_ext);
  return colo::check_launch("psrlc_rewards_kernel");
}

int colo_psrlc_finish_episode(const colo_psrlc_args* a, int S, int A, const int* index, int m, void* stream) {
  int r = colo::psrlc_check(a);
  if (r != COLO_OK) return r;
  COLO_ARG_CHECK(S > 0 && A > 0 && m >= 0 && (m == 0 || index), "S, A, m, index");
  if (m == 0) return COLO_OK;
  const long long n = (long long)m * S * A;
  colo::psrlc_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*a, S, A, index, m);
  return colo::check_launch("psrlc_finish_kernel");
}

}  // extern "C"
