/*
 * colo_oracle.c -- CPU restatement of the Colosseum hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle and the `cpu_baseline` of bench.py.  Nothing under colosseum_b200/ may link,
 * import or call it: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against golden vectors that
 * tests/golden/make_golden.py produced by running the unmodified Python reference (numba DP, NextStateSampler,
 * BaseMDP.step, hardness measures) and against the reference's own cached_hardness_measures .txt files and executed
 * notebook outputs.
 *
 * Each function cites the reference lines it restates (paths relative to /root/reference/colosseum/).
 * Build:  make -C oracle   ->  oracle/_build/libcolo_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_OVERFLOW 1
#define ORC_MAX_ITER 2
#define ORC_NEEDS_RESET 3

#define FOLD_MAX 0
#define FOLD_PI 1
#define FOLD_MIN 2

/* ------------------------------------------------------------------------------------------------------------
 * (B1) the reference iterate, fp32, in place (Gauss-Seidel):
 * dynamic_programming/infinite_horizon.py:121-142 (_discounted_value_iteration) and :167-184
 * (_discounted_policy_evaluation).  V starts at 0; per sweep, for s in order: Q[s]=R[s]+gamma*T[s]@V (V already
 * holds this sweep's updates for s'<s); V[s]=max_a Q[s,a] (or sum_a Q*pi); overflow -> "None"; stop when
 * max|V_old-V| < eps.  This is what bench.py times as the CPU baseline ("port" of the numba kernel).
 * ---------------------------------------------------------------------------------------------------------- */
int orc_discounted_gs_f32(const float* T, const float* R, const float* pi, int S, int A, float gamma, float eps,
                          float max_abs, long long max_iter, float* Q, float* V, long long* iters) {
  float* Vold = (float*)malloc(sizeof(float) * (size_t)S);
  memset(V, 0, sizeof(float) * (size_t)S);
  memset(Q, 0, sizeof(float) * (size_t)S * A);
  for (long long it = 0; it < max_iter; ++it) {
    memcpy(Vold, V, sizeof(float) * (size_t)S);
    for (int s = 0; s < S; ++s) {
      float best = -INFINITY, mix = 0.f;
      for (int a = 0; a < A; ++a) {
        const float* row = T + ((size_t)s * A + a) * S;
        float acc = 0.f;
        for (int j = 0; j < S; ++j) acc += row[j] * V[j];
        float q = R[(size_t)s * A + a] + gamma * acc;
        Q[(size_t)s * A + a] = q;
        if (q > best) best = q;
        if (pi) mix += q * pi[(size_t)s * A + a];
      }
      V[s] = pi ? mix : best;
      if (max_abs > 0.f && fabsf(V[s]) > max_abs) {
        free(Vold);
        if (iters) *iters = it + 1;
        return ORC_OVERFLOW;
      }
    }
    float diff = 0.f;
    for (int s = 0; s < S; ++s) {
      float d = fabsf(Vold[s] - V[s]);
      if (d > diff) diff = d;
    }
    if (diff < eps) {
      free(Vold);
      if (iters) *iters = it + 1;
      return ORC_OK;
    }
  }
  free(Vold);
  if (iters) *iters = max_iter;
  return ORC_MAX_ITER;
}

/* Batch of independent MDPs through (B1), OpenMP over instances -- the reference's own scaling mechanism is a
 * process pool over instances (hardness/measures/diameter.py:109-124, experiment/experiment_instances.py:160). */
int orc_discounted_gs_f32_batch(const float* T, const float* R, int B, int S, int A, float gamma, float eps,
                                long long max_iter, float* Q, float* V, long long* iters) {
  int rc = ORC_OK;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    int r = orc_discounted_gs_f32(T + (size_t)b * S * A * S, R + (size_t)b * S * A, NULL, S, A, gamma, eps, 0.f,
                                  max_iter, Q + (size_t)b * S * A, V + (size_t)b * S, iters ? iters + b : NULL);
    if (r != ORC_OK) {
#pragma omp critical
      rc = r;
    }
  }
  return rc;
}

/* Fixed number of fp32 in-place sweeps over a batch (throughput leg of the CPU baseline: sweeps/s).  Timing only, never
 * a parity oracle: the row product is a SIMD reduction (reassociated partial sums), as the BLAS sgemv behind the
 * reference's numba `T[s] @ V` (infinite_horizon.py:133) is -- a scalar in-order loop would understate the CPU. */
void orc_sweeps_gs_f32_batch(const float* T, const float* R, int B, int S, int A, float gamma, int n_sweeps,
                             float* V) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b) {
    const float* Tb = T + (size_t)b * S * A * S;
    const float* Rb = R + (size_t)b * S * A;
    float* Vb = V + (size_t)b * S;
    for (int it = 0; it < n_sweeps; ++it)
      for (int s = 0; s < S; ++s) {
        float best = -INFINITY;
        for (int a = 0; a < A; ++a) {
          const float* row = Tb + ((size_t)s * A + a) * S;
          float acc = 0.f;
#pragma omp simd reduction(+ : acc)
          for (int j = 0; j < S; ++j) acc += row[j] * Vb[j];
          float q = Rb[(size_t)s * A + a] + gamma * acc;
          if (q > best) best = q;
        }
        Vb[s] = best;
      }
  }
}

/* ------------------------------------------------------------------------------------------------------------
 * (B2) fixed-point oracle, fp64: the same recurrence (infinite_horizon.py:131-135 / :176-179) iterated
 * synchronously in double to |dV|_inf <= tol.  Both sweep orders converge to the same V* (gamma<1), which is
 * where parity is defined ("Gauss-Seidel vs Jacobi").  gauss_seidel!=0 uses the reference's
 * in-place order instead.  Q is the Q of the last sweep, as in the reference.
 * ---------------------------------------------------------------------------------------------------------- */
int orc_discounted_f64(const float* T, const float* R, const float* pi, int S, int A, double gamma, double tol,
                       double max_abs, long long max_iter, int fold, int gauss_seidel, double* Q, double* V,
                       long long* iters) {
  double* Vn = (double*)malloc(sizeof(double) * (size_t)S);
  memset(V, 0, sizeof(double) * (size_t)S);
  int rc = ORC_MAX_ITER;
  long long it = 0;
  for (; it < max_iter; ++it) {
    double diff = 0.0;
    for (int s = 0; s < S; ++s) {
      double best = fold == FOLD_MIN ? INFINITY : -INFINITY, mix = 0.0;
      for (int a = 0; a < A; ++a) {
        const float* row = T + ((size_t)s * A + a) * S;
        double acc = 0.0;
        for (int j = 0; j < S; ++j) acc += (double)row[j] * V[j];
        double q = (R ? (double)R[(size_t)s * A + a] : 0.0) + gamma * acc;
        if (Q) Q[(size_t)s * A + a] = q;
        if (fold == FOLD_MAX && q > best) best = q;
        if (fold == FOLD_MIN && q < best) best = q;
        if (fold == FOLD_PI) mix += q * (double)pi[(size_t)s * A + a];
      }
      double v = fold == FOLD_PI ? mix : best;
      double d = fabs(v - V[s]);
      if (d > diff) diff = d;
      if (gauss_seidel)
        V[s] = v;
      else
        Vn[s] = v;
      if (max_abs > 0.0 && fabs(v) > max_abs) {
        rc = ORC_OVERFLOW;
        goto done;
      }
    }
    if (!gauss_seidel) memcpy(V, Vn, sizeof(double) * (size_t)S);
    if (diff < tol) {
      rc = ORC_OK;
      ++it;
      goto done;
    }
  }
done:
  if (iters) *iters = it;
  free(Vn);
  return rc;
}

/* ------------------------------------------------------------------------------------------------------------
 * (B3) episodic backward induction: dynamic_programming/finite_horizon.py:11-26 (VI), :29-42 (PE).
 * Q[H+1,S,A], V[H+1,S]; row H is zero; no discount; pi is [H,S,A].  fp64 accumulation of fp32 inputs.
 * ---------------------------------------------------------------------------------------------------------- */
int orc_episodic_f64(const float* T, const float* R, const float* pi, int S, int A, int H, double max_value,
                     double* Q, double* V) {
  memset(Q, 0, sizeof(double) * (size_t)(H + 1) * S * A);
  memset(V, 0, sizeof(double) * (size_t)(H + 1) * S);
  for (int h = H - 1; h >= 0; --h) {
    const double* Vn = V + (size_t)(h + 1) * S;
    for (int s = 0; s < S; ++s) {
      double best = -INFINITY, mix = 0.0;
      for (int a = 0; a < A; ++a) {
        const float* row = T + ((size_t)s * A + a) * S;
        double acc = 0.0;
        for (int j = 0; j < S; ++j) acc += (double)row[j] * Vn[j];
        double q = (double)R[(size_t)s * A + a] + acc;
        Q[((size_t)h * S + s) * A + a] = q;
        if (q > best) best = q;
        if (pi) mix += q * (double)pi[((size_t)h * S + s) * A + a];
      }
      V[(size_t)h * S + s] = pi ? mix : best;
      if (!pi && max_value > 0.0 && best > max_value) return ORC_OVERFLOW;
    }
  }
  return ORC_OK;
}

/* same in the reference's arithmetic type (fp32), for the CPU baseline timing of config C1 */
int orc_episodic_f32(const float* T, const float* R, const float* pi, int S, int A, int H, float* Q, float* V) {
  memset(Q, 0, sizeof(float) * (size_t)(H + 1) * S * A);
  memset(V, 0, sizeof(float) * (size_t)(H + 1) * S);
  for (int h = H - 1; h >= 0; --h) {
    const float* Vn = V + (size_t)(h + 1) * S;
    for (int s = 0; s < S; ++s) {
      float best = -INFINITY, mix = 0.f;
      for (int a = 0; a < A; ++a) {
        const float* row = T + ((size_t)s * A + a) * S;
        float acc = 0.f;
        for (int j = 0; j < S; ++j) acc += row[j] * Vn[j];
        float q = R[(size_t)s * A + a] + acc;
        Q[((size_t)h * S + s) * A + a] = q;
        if (q > best) best = q;
        if (pi) mix += q * pi[((size_t)h * S + s) * A + a];
      }
      V[(size_t)h * S + s] = pi ? mix : best;
    }
  }
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------------------
 * (B4) continuous diameter: hardness/measures/diameter.py:76-106.  For target es: make es absorbing, reward -1
 * elsewhere, VI with gamma=1; d_es = -min V.  Restated as the equivalent hitting-time recurrence
 * (diameter.py:321-346): E[es]=0, E[j] = min_a (1 + sum_ns T[j,a,ns] E[ns]); iterated synchronously in fp64 to
 * tolerance `tol` (no early exit -- the reference's early exits make its own value path-dependent at ~1e-5).
 * E_out (may be NULL) receives the K x S table.  Returns max_k max_s E.
 * ---------------------------------------------------------------------------------------------------------- */
int orc_diameter_continuous_f64(const float* T, const int* targets, int K, int S, int A, double tol,
                                double max_value, long long max_iter, double* E_out, double* diameter,
                                long long* sweeps) {
  int rc = ORC_OK;
  double diam = 0.0;
  long long max_sweeps = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < K; ++k) {
    int es = targets[k];
    double* E = (double*)calloc((size_t)S, sizeof(double));
    double* En = (double*)calloc((size_t)S, sizeof(double));
    long long it = 0;
    int local_rc = ORC_MAX_ITER;
    for (; it < max_iter; ++it) {
      double diff = 0.0, mx = 0.0;
      for (int j = 0; j < S; ++j) {
        if (j == es) {
          En[j] = 0.0;
          continue;
        }
        double best = INFINITY;
        for (int a = 0; a < A; ++a) {
          const float* row = T + ((size_t)j * A + a) * S;
          double acc = 1.0;
          for (int ns = 0; ns < S; ++ns) acc += (double)row[ns] * E[ns];
          if (acc < best) best = acc;
        }
        En[j] = best;
        double d = fabs(best - E[j]);
        if (d > diff) diff = d;
        if (best > mx) mx = best;
      }
      double* t = E;
      E = En;
      En = t;
      if (max_value > 0.0 && mx > max_value) {
        local_rc = ORC_OVERFLOW;
        break;
      }
      if (diff < tol) {
        local_rc = ORC_OK;
        ++it;
        break;
      }
    }
    double mx = 0.0;
    for (int j = 0; j < S; ++j)
      if (E[j] > mx) mx = E[j];
    if (E_out) memcpy(E_out + (size_t)k * S, E, sizeof(double) * (size_t)S);
#pragma omp critical
    {
      if (mx > diam) diam = mx;
      if (it > max_sweeps) max_sweeps = it;
      if (local_rc != ORC_OK) rc = local_rc;
    }
    free(E);
    free(En);
  }
  *diameter = diam;
  if (sweeps) *sweeps = max_sweeps;
  return rc;
}

/* the reference's own single-target kernel in ITS arithmetic (fp32, in place, its early exit):
 * hardness/measures/diameter.py:321-346 -- used for the CPU baseline timing of the diameter. */
float orc_diameter_target_ref_f32(const float* T, int es, int S, int A, float max_diam, float eps) {
  float* E = (float*)calloc((size_t)S, sizeof(float));
  float* Eo = (float*)calloc((size_t)S, sizeof(float));
  float mx = 0.f;
  for (long long t = 0; t < 1000000; ++t) {
    memcpy(Eo, E, sizeof(float) * (size_t)S);
    for (int j = 0; j < S; ++j) {
      if (j == es) continue;
      float best = INFINITY;
      for (int a = 0; a < A; ++a) {
        const float* row = T + ((size_t)j * A + a) * S;
        float acc = 0.f;
        for (int ns = 0; ns < S; ++ns)
          if (ns != es) acc += row[ns] * (1.f + E[ns]);
        acc += row[es];
        if (acc < best) best = acc;
      }
      E[j] = best;
    }
    float diff = 0.f;
    mx = 0.f;
    for (int j = 0; j < S; ++j) {
      float d = fabsf(Eo[j] - E[j]);
      if (d > diff) diff = d;
      if (E[j] > mx) mx = E[j];
    }
    if (diff < eps || (diff < 0.05f && mx - 1.f < max_diam)) break;
  }
  free(E);
  free(Eo);
  return mx > max_diam ? mx : max_diam;
}

/* ------------------------------------------------------------------------------------------------------------
 * (B5) episodic diameter in the augmented (h,s) space: hardness/measures/diameter.py:285-318 on
 * T_epi[H,S,A,S] (mdp/utils/mdp_creation.py:98-128).  Per target es: ETs[H-1,:] = T[H-1,0,0,:]@(1+ETs[0,:]);
 * for h=H-1..1, j!=es: ETs[h-1,j] = min_a( T[h-1,j,a,es] + sum_{ns!=es} T[h-1,j,a,ns]*(1+ETs[h,ns]) ).
 * Iterated in the reference's in-place order, fp64, to tolerance `tol` (no early exit).  Per state the minimum
 * over h of the POSITIVE entries, then the max over states; diameter = max over targets.
 * ---------------------------------------------------------------------------------------------------------- */
int orc_diameter_episodic_f64(const float* T, const int* targets, int K, int H, int S, int A, double tol,
                              long long max_iter, double* diameter, long long* sweeps) {
  double diam = -INFINITY;
  long long max_sweeps = 0;
  int rc = ORC_OK;
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < K; ++k) {
    int es = targets[k];
    double* E = (double*)calloc((size_t)H * S, sizeof(double));
    long long it = 0;
    int local_rc = ORC_MAX_ITER;
    for (; it < max_iter; ++it) {
      double diff = 0.0;
      {
        const float* row = T + ((size_t)(H - 1) * S * A) * S; /* T[H-1,0,0,:] */
        double acc = 0.0;
        for (int ns = 0; ns < S; ++ns) acc += (double)row[ns] * (1.0 + E[ns]);
        for (int j = 0; j < S; ++j) {
          double d = fabs(acc - E[(size_t)(H - 1) * S + j]);
          if (d > diff) diff = d;
          E[(size_t)(H - 1) * S + j] = acc;
        }
      }
      for (int h = H - 1; h >= 1; --h) {
        const double* En = E + (size_t)h * S;
        for (int j = 0; j < S; ++j) {
          if (j == es) continue;
          double best = INFINITY;
          for (int a = 0; a < A; ++a) {
            const float* row = T + (((size_t)(h - 1) * S + j) * A + a) * S;
            double acc = (double)row[es];
            for (int ns = 0; ns < S; ++ns)
              if (ns != es) acc += (double)row[ns] * (1.0 + En[ns]);
            if (acc < best) best = acc;
          }
          double d = fabs(best - E[(size_t)(h - 1) * S + j]);
          if (d > diff) diff = d;
          E[(size_t)(h - 1) * S + j] = best;
        }
      }
      if (diff < tol) {
        local_rc = ORC_OK;
        ++it;
        break;
      }
    }
    double cur = -INFINITY;
    for (int s = 0; s < S; ++s) {
      double mn = INFINITY;
      for (int h = 0; h < H; ++h) {
        double v = E[(size_t)h * S + s];
        if (v > 0.0 && v < mn) mn = v;
      }
      if (mn > cur) cur = mn;
    }
#pragma omp critical
    {
      if (cur > diam) diam = cur;
      if (it > max_sweeps) max_sweeps = it;
      if (local_rc != ORC_OK) rc = local_rc;
    }
    free(E);
  }
  *diameter = diam;
  if (sweeps) *sweeps = max_sweeps;
  return rc;
}

/* ------------------------------------------------------------------------------------------------------------
 * (B6) environmental value norm: hardness/measures/value_norm.py:55-61,85-87.
 * Ev[i,a]=sum_j T[i,a,j]V[j]; norm = max_{i,a} sqrt(sum_j T[i,a,j]*(V[j]-Ev[j,a])^2)  -- Ev indexed by the NEXT
 * state j (reference behaviour, kept).
 * ---------------------------------------------------------------------------------------------------------- */
double orc_value_norm_f64(const float* T, const double* V, int S, int A) {
  double* Ev = (double*)malloc(sizeof(double) * (size_t)S * A);
  for (int i = 0; i < S; ++i)
    for (int a = 0; a < A; ++a) {
      const float* row = T + ((size_t)i * A + a) * S;
      double acc = 0.0;
      for (int j = 0; j < S; ++j) acc += (double)row[j] * V[j];
      Ev[(size_t)i * A + a] = acc;
    }
  double best = 0.0;
  for (int i = 0; i < S; ++i)
    for (int a = 0; a < A; ++a) {
      const float* row = T + ((size_t)i * A + a) * S;
      double acc = 0.0;
      for (int j = 0; j < S; ++j) {
        double d = V[j] - Ev[(size_t)j * A + a];
        acc += (double)row[j] * d * d;
      }
      double n = sqrt(acc);
      if (n > best) best = n;
    }
  free(Ev);
  return best;
}

/* (B7) gaps: hardness/measures/sum_reciprocals_suboptimality_gaps.py:6-28 */
double orc_gaps_f64(const double* Q, const double* V, const unsigned char* mask, long long NS, int A, double reg) {
  double acc = 0.0;
  for (long long n = 0; n < NS; ++n) {
    if (mask && !mask[n]) continue;
    for (int a = 0; a < A; ++a) acc += 1.0 / (V[n] - Q[n * A + a] + reg);
  }
  return acc;
}

/* ------------------------------------------------------------------------------------------------------------
 * (A) interaction step.
 * Philox4x32-10 (Salmon et al., SC'11) counter RNG: counter = (env lo, env hi, t lo, t hi), key = seed.
 * ---------------------------------------------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0;
    c[1] = n1;
    c[2] = n2;
    c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

void orc_philox(uint64_t seed, uint64_t env, uint64_t t, uint32_t out[4]) {
  uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), (uint32_t)t, (uint32_t)(t >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  memcpy(out, c, sizeof(uint32_t) * 4);
}

/* word -> uniform conventions shared with the CUDA kernels */
static inline double u53(uint32_t a, uint32_t b) { /* CPython random(): (a>>5, b>>6) -> 53 bits */
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
static inline float u24(uint32_t a) { return (float)(a >> 8) * (1.0f / 16777216.0f); }
static inline int act_from_word(uint32_t w, int A) { return (int)(((uint64_t)w * (uint64_t)A) >> 32); }

typedef struct {
  int S, A, H, ld;
  const void* cdf;
  const double* succ_cum;
  const int* succ_idx;
  const int* succ_len;
  int Ksucc;
  const unsigned char* rew_cls_sas;
  const int* rew_cls_sa;
  const int* rew_cls_succ;
  const float* rew_q;
  int n_cls, nq;
  float rmin, rmax;
  const double* start_cum;
  const int* start_idx;
  int n_start;
} orc_tables;

/* reward draw: quantile-table interpolation at u, then mdp/base.py:1205-1207's rescale r*(max-min) - min (sic) */
static inline float reward_draw(const orc_tables* tb, int cls, float u) {
  const float* q = tb->rew_q + (size_t)cls * tb->nq;
  float t = u * (float)(tb->nq - 1);
  int i = (int)t;
  if (i > tb->nq - 2) i = tb->nq - 2;
  float f = t - (float)i;
  float r0 = fmaf(f, q[i + 1] - q[i], q[i]);
  return fmaf(r0, tb->rmax - tb->rmin, -tb->rmin);
}

/* CPython random.choices / NextStateSampler.sample (mdp/utils/custom_samplers.py:49-72):
 * bisect_right(cum, u*total, 0, n-1) over the sampler's own successor order. */
static inline int bisect_pos(const double* cum, int n, double x) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (x < cum[mid])
      hi = mid;
    else
      lo = mid + 1;
  }
  return lo;
}

static int sample_start(const orc_tables* tb, double u) {
  if (tb->n_start == 1) return tb->start_idx[0];
  double total = tb->start_cum[tb->n_start - 1] + 0.0;
  return tb->start_idx[bisect_pos(tb->start_cum, tb->n_start, u * total)];
}

/* dense row search: first j with cdf[j] > x, clamped to the first j where the row reaches its total
 * (= the last positive-probability index = bisect's hi=n-1 clamp on the sampler's successor list) */
#define DENSE_SEARCH(TYPE, row, S, u, out)             \
  do {                                                 \
    TYPE total = (row)[(S)-1];                         \
    TYPE x = (TYPE)(u) * total;                        \
    int j1 = (S), j2 = (S)-1;                          \
    for (int j = 0; j < (S); ++j)                      \
      if ((row)[j] > x) {                              \
        j1 = j;                                        \
        break;                                         \
      }                                                \
    for (int j = 0; j < (S); ++j)                      \
      if ((row)[j] >= total) {                         \
        j2 = j;                                        \
        break;                                         \
      }                                                \
    (out) = j1 < j2 ? j1 : j2;                         \
  } while (0)

/* mode: 0 dense f32 cdf (u_next float), 1 dense f64 cdf (u_next double), 2 successor lists (u_next double).
 * Restates BaseMDP.reset / BaseMDP.step (mdp/base.py:1268-1317) for N independent envs; see SURVEY.md B.1/B.2. */
int orc_env_step(const orc_tables* tb, int mode, long long N, int* action, int random_actions, const void* u_next,
                 const float* u_rew, uint64_t seed, uint64_t t, uint64_t env0, int auto_reset, int* state, int* h,
                 unsigned char* step_type, float* reward, int* obs, unsigned long long* visits_s,
                 unsigned long long* visits_sa) {
  int status = ORC_OK;
  const int S = tb->S, A = tb->A;
#pragma omp parallel for schedule(static)
  for (long long e = 0; e < N; ++e) {
    uint32_t w[4];
    orc_philox(seed, env0 + (uint64_t)e, t, w);
    double un64 = u_next ? (mode == 0 ? (double)((const float*)u_next)[e] : ((const double*)u_next)[e]) : u53(w[0], w[1]);
    float un32 = u_next ? (mode == 0 ? ((const float*)u_next)[e] : 0.f) : u24(w[0]);
    float ur = u_rew ? u_rew[e] : u24(w[2]);
    if (step_type[e] == 2) {
      if (!auto_reset) {
#pragma omp critical
        status = ORC_NEEDS_RESET;
        continue;
      }
      int s0 = sample_start(tb, un64);
      state[e] = s0;
      h[e] = 0;
      step_type[e] = 0;
      reward[e] = NAN;
      obs[e] = s0;
      if (visits_s) {
#pragma omp atomic
        visits_s[s0] += 1ULL;
      }
      continue;
    }
    int a = random_actions ? act_from_word(w[3], A) : action[e];
    if (random_actions) action[e] = a;
    int s = state[e];
    int nxt, cls;
    if (mode == 2) {
      size_t base = ((size_t)s * A + a) * tb->Ksucc;
      int n = tb->succ_len[(size_t)s * A + a];
      int pos = 0;
      if (n > 1) {
        double total = tb->succ_cum[base + n - 1] + 0.0;
        pos = bisect_pos(tb->succ_cum + base, n, un64 * total);
      }
      nxt = tb->succ_idx[base + pos];
      cls = tb->rew_cls_succ ? tb->rew_cls_succ[base + pos] : 0;
    } else {
      size_t base = ((size_t)s * A + a) * tb->ld;
      if (mode == 0) {
        const float* row = (const float*)tb->cdf + base;
        DENSE_SEARCH(float, row, S, un32, nxt);
      } else {
        const double* row = (const double*)tb->cdf + base;
        DENSE_SEARCH(double, row, S, un64, nxt);
      }
      cls = tb->rew_cls_sas ? tb->rew_cls_sas[((size_t)s * A + a) * S + nxt]
                            : (tb->rew_cls_sa ? tb->rew_cls_sa[(size_t)s * A + a] : 0);
    }
    int hh = h[e] + 1;
    h[e] = hh;
    state[e] = nxt;
    reward[e] = reward_draw(tb, cls, ur);
    if (visits_s) {
#pragma omp atomic
      visits_s[nxt] += 1ULL;
    }
    if (visits_sa) {
#pragma omp atomic
      visits_sa[(size_t)nxt * A + a] += 1ULL;
    }
    if (tb->H > 0 && hh >= tb->H) {
      step_type[e] = 2;
      obs[e] = -1;
    } else {
      step_type[e] = 1;
      obs[e] = nxt;
    }
  }
  return status;
}

int orc_env_reset(const orc_tables* tb, long long N, const double* u_next, uint64_t seed, uint64_t t, uint64_t env0,
                  int* state,
                  int* h, unsigned char* step_type, int* obs, unsigned long long* visits_s) {
  for (long long e = 0; e < N; ++e) {
    uint32_t w[4];
    orc_philox(seed, env0 + (uint64_t)e, t, w);
    double u = u_next ? u_next[e] : u53(w[0], w[1]);
    int s0 = sample_start(tb, u);
    state[e] = s0;
    h[e] = 0;
    step_type[e] = 0;
    obs[e] = s0;
    if (visits_s) visits_s[s0] += 1ULL;
  }
  return ORC_OK;
}

/* dense cdf builder shared with the product's definition: sequential fp64 running sum, rounded to storage type,
 * padding [S,ld) filled with the row total */
void orc_build_dense_cdf(const float* T, int S, int A, int ld, void* cdf, int out_is_f64) {
  for (size_t r = 0; r < (size_t)S * A; ++r) {
    double acc = 0.0;
    for (int j = 0; j < ld; ++j) {
      if (j < S) acc += (double)T[r * S + j];
      if (out_is_f64)
        ((double*)cdf)[r * ld + j] = acc;
      else
        ((float*)cdf)[r * ld + j] = (float)acc;
    }
  }
}

/* ------------------------------------------------------------------------------------------------------------
 * extended value iteration (UCRL2): colosseum/dynamic_programming/infinite_horizon.py:67-118 and _max_proba :222-251,
 * restated loop for loop.  Types as numba infers them: T, estimated rewards, u1, u2, p2, Q, V are float32; beta_r,
 * beta_p, min1, s, s2, max1, r_optimal, v are float64; np.dot(vec, u1) of two float32 vectors is float32 (accumulated
 * here in index order).  argsort ties are broken by index (numpy's quicksort leaves them unspecified).
 * Returns ORC_OK with *span = ptp(u1), or ORC_MAX_ITER.
 */
static const float* g_sort_key;
static int cmp_idx(const void* a, const void* b) {
  int ia = *(const int*)a, ib = *(const int*)b;
  float ka = g_sort_key[ia], kb = g_sort_key[ib];
  if (ka < kb) return -1;
  if (ka > kb) return 1;
  return ia - ib;
}

int orc_extended_vi_f32(const float* T, const float* est, const double* beta_r, const double* beta_p, int S, int A,
                        double r_max, double eps, long long max_iter, float* Q, float* V, double* span,
                        long long* iters) {
  float* u1 = (float*)calloc((size_t)S, sizeof(float));
  float* u2 = (float*)calloc((size_t)S, sizeof(float));
  float* p2 = (float*)calloc((size_t)S, sizeof(float));
  int* sorted = (int*)malloc((size_t)S * sizeof(int));
  for (int i = 0; i < S; ++i) sorted[i] = i;
  int rc = ORC_MAX_ITER;
  long long it = 0;
  for (; it < max_iter; ++it) {
    for (int s = 0; s < S; ++s) {
      float qmax = -INFINITY;
      for (int a = 0; a < A; ++a) {
        const float* p = T + ((size_t)s * A + a) * S;
        const double beta = beta_p[(size_t)s * A + a];
        const int best = sorted[S - 1];
        /* _max_proba */
        double min1 = p[best] + beta / 2;
        if (min1 > 1.0) min1 = 1.0;
        for (int j = 0; j < S; ++j) p2[j] = 0.f;
        if (min1 == 1.0) {
          p2[best] = 1.f;
        } else {
          for (int j = 0; j < S; ++j) p2[j] = p[j]; /* p2[support_p] = restricted_sorted_p */
          p2[best] = (float)min1;
          double sm = 1.0 - p[best] + min1, s2 = sm;
          for (int j = 0; j < S; ++j) {
            int id = sorted[j];
            float proba = p[id];
            if (proba == 0.f) continue; /* only the support is visited */
            double max1 = 1.0 - sm + proba;
            if (max1 < 0.0) max1 = 0.0;
            s2 += max1 - proba;
            p2[id] = (float)max1;
            sm = s2;
            if (sm <= 1.0) break;
          }
        }
        p2[s] -= 1.f; /* vec[s] -= 1 */
        double r_opt = (double)est[(size_t)s * A + a] + beta_r[(size_t)s * A + a];
        if ((double)(float)r_max < r_opt) r_opt = (double)(float)r_max;
        float dot = 0.f;
        for (int j = 0; j < S; ++j) dot += p2[j] * u1[j];
        double v = r_opt + (double)dot;
        Q[(size_t)s * A + a] = (float)v;
        if (a == 0 || v + u1[s] > u2[s] || fabs(v + u1[s] - u2[s]) < eps) u2[s] = (float)(v + u1[s]);
        if ((float)v > qmax) qmax = (float)v;
      }
      V[s] = qmax;
    }
    float lo = INFINITY, hi = -INFINITY;
    for (int i = 0; i < S; ++i) {
      float d = u2[i] - u1[i];
      if (d < lo) lo = d;
      if (d > hi) hi = d;
    }
    if ((double)(hi - lo) < eps) {
      float ulo = INFINITY, uhi = -INFINITY;
      for (int i = 0; i < S; ++i) {
        if (u1[i] < ulo) ulo = u1[i];
        if (u1[i] > uhi) uhi = u1[i];
      }
      *span = (double)(uhi - ulo);
      rc = ORC_OK;
      ++it;
      break;
    }
    float* t = u1; u1 = u2; u2 = t;
    g_sort_key = u1;
    qsort(sorted, (size_t)S, sizeof(int), cmp_idx);
  }
  if (iters) *iters = it;
  free(u1); free(u2); free(p2); free(sorted);
  return rc;
}

/* ---------------------------------------------------------------------------------------------------------------
 * N independent Q-learning agent/MDP loops.  Restates, per loop and per step, MDPLoop.run's body
 * (colosseum/experiment/agent_mdp_interaction.py:238-298): QValuesActor.select_action
 * (colosseum/agent/actors/Q_values_actor.py:58-82), BaseMDP.step on the reference's own successor sampler, then
 *   episodic   QValuesModel.step_update  (colosseum/agent/agents/episodic/q_learning.py:53-103)
 *   continuous _QValuesModel.step_update (colosseum/agent/agents/infinite_horizon/q_learning.py:86-111)
 * with numpy's (NEP 50) type of every sub-expression: float32 where the reference's operands are float32 arrays and
 * python scalars, float64 as soon as a numpy float64 scalar (alpha_t, self.i) takes part.  Randomness is supplied by
 * Philox (env key = seed, agent key = seed ^ 0x9E3779B97F4A7C15) because numpy's RandomState stream cannot be
 * reproduced on a GPU; the update rule itself is pinned against the reference's model classes replayed on this
 * function's trace (tests/golden/make_qlearning_golden.py). */
/* QValuesActor's exploration (colosseum/agent/actors/Q_values_actor.py:20-82): constant or tabulated epsilon and Boltzmann
 * temperature as functions of the actor's interaction counter; see orc_actor_select below. */
typedef struct {
  const double* epsilon_schedule;
  const double* temperature_schedule;
  long long t0;
  int len;
  int boltzmann;
  double boltzmann_temperature;
} orc_actor_args;

typedef struct {
  long long N;
  uint64_t seed, env0;
  int* state;
  int* h;
  int* cnt;
  float* Q;
  float* Q_main;
  float* V;
  float* mu;
  float* sigma;
  float* beta;
  int ucb_type;
  double c_1, c_2, min_at, log_term, sqrt_h7sa;
  double H_eff, gamma, span_approx;
  double epsilon_greedy;
  double* cum_reward;
  long long* n_episodes;
  int* trace;
  orc_actor_args actor;
} orc_qlearning_args;

static int orc_select_action(const float* q, int A, double eps, const uint32_t w[4]) {
  if (eps >= 0.0 && (double)u24(w[0]) < eps) return act_from_word(w[1], A);
  float best = q[0];
  for (int a = 1; a < A; ++a)
    if (q[a] > best) best = q[a];
  int ties = 0;
  for (int a = 0; a < A; ++a) ties += q[a] == best;
  int k = act_from_word(w[1], ties); /* np.where(q == q.max())[0][k] */
  for (int a = 0; a < A; ++a)
    if (q[a] == best && k-- == 0) return a;
  return A - 1;
}

static double orc_actor_value(const double* schedule, long long t0, int len, double constant, long long total) {
  if (!schedule) return constant;
  long long k = total - t0;
  k = k < 0 ? 0 : (k >= len ? len - 1 : k);
  return schedule[k];
}

/* Boltzmann exploration (Q_values_actor.py:73-78): q = np.exp(temperature * q) (float32: a python float times a float32
 * array), p = q / q.sum() (float32), rng.choice(range(A), p=p): numpy takes cdf = p.cumsum() in float64, cdf /= cdf[-1]
 * and returns cdf.searchsorted(u, side="right").  exp is evaluated in double and rounded (one rounding, like a correctly
 * rounded float32 exp). */
static float orc_boltz_weight(float temp, float q) {
  const float x = temp * q;
  return (float)exp((double)x);
}
int orc_boltzmann_action(const float* q, int A, double temperature, double u) {
  const float temp = (float)temperature;
  float sum = 0.f;
  for (int a = 0; a < A; ++a) sum = sum + orc_boltz_weight(temp, q[a]);
  double tot = 0.0;
  for (int a = 0; a < A; ++a) {
    const float pa = orc_boltz_weight(temp, q[a]) / sum;
    tot += (double)pa;
  }
  double run = 0.0;
  int idx = 0;
  for (int a = 0; a < A; ++a) {
    const float pa = orc_boltz_weight(temp, q[a]) / sum;
    run += (double)pa;
    idx += (run / tot <= u) ? 1 : 0;
  }
  return idx < A ? idx : A - 1;
}

/* QValuesActor.select_action (:58-82) at interaction count `total`: epsilon-greedy draw from range(A_random), else
 * Boltzmann (if enabled), else greedy with uniform tie-breaking */
static int orc_actor_select(const float* q, int A, int A_random, double eps_const, const orc_actor_args* ac,
                            long long total, const uint32_t w[4], uint64_t seed, uint64_t loop) {
  const double eps = orc_actor_value(ac->epsilon_schedule, ac->t0, ac->len, eps_const, total);
  if (eps >= 0.0 && (double)u24(w[0]) < eps) return act_from_word(w[1], A_random);
  if (ac->boltzmann) {
    uint32_t wb[4];
    orc_philox(seed ^ 0x94D049BB133111EBULL, loop, (uint64_t)total, wb);
    const double temp = orc_actor_value(ac->temperature_schedule, ac->t0, ac->len, ac->boltzmann_temperature, total);
    return orc_boltzmann_action(q, A, temp, u53(wb[0], wb[1]));
  }
  return orc_select_action(q, A, -1.0, w);
}

int orc_qlearning_steps(const orc_tables* tb, const orc_qlearning_args* p, int episodic, int n_steps, uint64_t t0) {
  const int S = tb->S, A = tb->A, H = tb->H;
  const size_t per_q = (size_t)(episodic ? H : 1) * S * A, per_v = (size_t)(episodic ? H + 1 : 1) * S;
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < p->N; ++i) {
    int* cnt = p->cnt + i * per_q;
    float* Q = p->Q + i * per_q;
    float* V = p->V + i * per_v;
    float* Qm = episodic ? NULL : p->Q_main + i * per_q;
    float* mu = p->mu ? p->mu + i * per_q : NULL;
    float* sg = p->sigma ? p->sigma + i * per_q : NULL;
    float* be = p->beta ? p->beta + i * per_q : NULL;
    int s = p->state[i], h = p->h[i];
    double cum = p->cum_reward[i];
    const double Hd = episodic ? (double)H : p->H_eff;
    const double H3 = (double)H * H * H;
    for (int step = 0; step < n_steps; ++step) {
      const uint64_t t = t0 + (uint64_t)step;
      uint32_t we[4], wa[4];
      orc_philox(p->seed, p->env0 + (uint64_t)i, t, we);
      orc_philox(p->seed ^ 0x9E3779B97F4A7C15ULL, p->env0 + (uint64_t)i, t, wa);
      const size_t row = ((size_t)(episodic ? h : 0) * S + s) * A;
      const int a = orc_actor_select(Q + row, A, A, p->epsilon_greedy, &p->actor, (long long)t, wa, p->seed, p->env0 + (uint64_t)i);
      /* BaseMDP.step: NextStateSampler.sample + sample_reward (as orc_env_step, mode 2) */
      const size_t base = ((size_t)s * A + a) * tb->Ksucc;
      const int nsucc = tb->succ_len[(size_t)s * A + a];
      int pos = 0;
      if (nsucc > 1) {
        const double total = tb->succ_cum[base + nsucc - 1] + 0.0;
        pos = bisect_pos(tb->succ_cum + base, nsucc, u53(we[0], we[1]) * total);
      }
      const int nxt = tb->succ_idx[base + pos];
      const int cls = tb->rew_cls_succ ? tb->rew_cls_succ[base + pos] : 0;
      const float r = reward_draw(tb, cls, u24(we[2]));
      const int hh = h + 1;
      const int last = episodic && hh >= H;
      const int obs = last ? -1 : nxt;
      const int sp = obs < 0 ? S - 1 : obs; /* numpy negative index */
      const size_t idx = row + a;
      const int n = cnt[idx] + 1;
      cnt[idx] = n;
      double alpha = (Hd + 1.0) / (Hd + (double)n);
      /* python's max(min_at, ratio) returns min_at -- a PYTHON float -- unless ratio > min_at.  A python float is a
       * weak scalar under NEP 50: multiplied with a float32 table entry it is rounded to float32 and the product is
       * a float32 one, where the np.float64 ratio promotes to float64 (q_learning.py:66, :92-102). */
      const int py_alpha = !(alpha > p->min_at);
      if (py_alpha) alpha = p->min_at;
      const double om = 1.0 - alpha;
      if (episodic) {
        const float vnext = V[(size_t)hh * S + sp];
        double b = 0.0;
        float b32 = 0.f;
        int b_is_f32 = 0;
        if (p->ucb_type == 0) {
          b = p->c_1 * sqrt(H3 * p->log_term / (double)n);
        } else {
          const float m = mu[idx] + vnext;
          const float g = sg[idx] + vnext * vnext;
          mu[idx] = m;
          sg[idx] = g;
          const float old_beta = be[idx];
          const float d = g - m;
          const float hd2 = (float)H * (d * d);
          const int n2 = (int)((unsigned)n * (unsigned)n);
          const double x = (double)hd2 / (double)n2 + (double)H;
          const double first = sqrt(x * p->log_term);
          const double second = p->sqrt_h7sa * p->log_term / (double)n;
          const double v1 = p->c_1 * (first + second);
          const double v2 = p->c_2 * sqrt(H3 * p->log_term / (double)n);
          const float nb = (float)(v2 < v1 ? v2 : v1);
          be[idx] = nb;
          if (py_alpha) { /* every operand is float32 or a python scalar: the whole bonus is float32 arithmetic */
            const float t1 = (float)om * old_beta;
            const float t2 = nb - t1;
            const float t3 = t2 / 2.0f;
            b32 = t3 / (float)alpha;
            b_is_f32 = 1;
          } else {
            b = ((double)nb - om * (double)old_beta) / 2.0 / alpha;
          }
        }
        /* python float + np.float32 is a float32 sum under NEP 50; the np.float64 bonus then promotes */
        const float rv = r + vnext;
        if (b_is_f32) { /* float32 bonus, python-float alpha: the update never leaves float32 */
          const float target32 = rv + b32;
          const float lhs = (float)alpha * Q[idx];
          const float rhs = (float)om * target32;
          Q[idx] = lhs + rhs;
        } else {
          const double target = (double)rv + b;
          const double lhs = py_alpha ? (double)((float)alpha * Q[idx]) : alpha * (double)Q[idx];
          Q[idx] = (float)(lhs + om * target); /* sic: alpha weighs the old estimate */
        }
        float mx = Q[row];
        for (int k = 1; k < A; ++k)
          if (Q[row + k] > mx) mx = Q[row + k];
        V[(size_t)h * S + s] = mx < (float)H ? mx : (float)H;
      } else {
        const double b = 4.0 * p->span_approx * sqrt(Hd / (double)n * p->log_term);
        const double target = (double)r + p->gamma * (double)V[sp] + b;
        const float qm = (float)(om * (double)Q[idx] + alpha * target);
        Qm[idx] = qm;
        if (qm < Q[idx]) Q[idx] = qm;
        const size_t rp = (size_t)sp * A;
        float mx = Q[rp];
        for (int k = 1; k < A; ++k)
          if (Q[rp + k] > mx) mx = Q[rp + k];
        V[sp] = mx;
      }
      cum += (double)r;
      if (p->trace) {
        int* tr = p->trace + ((size_t)step * p->N + i) * 4;
        union { float f; int i; } u;
        u.f = r;
        tr[0] = s; tr[1] = a; tr[2] = obs; tr[3] = u.i;
      }
      if (last) {
        if (p->n_episodes) p->n_episodes[i] += 1;
        h = 0;
        s = sample_start(tb, u53(wa[2], wa[3]));
      } else {
        h = hh;
        s = nxt;
      }
    }
    p->state[i] = s;
    p->h[i] = h;
    p->cum_reward[i] = cum;
  }
  return ORC_OK;
}

/* PSRLEpisodic between two posterior samples (colosseum/agent/agents/episodic/posterior_sampling.py:142-147): greedy
 * action on Q[N,H+1,S,A], BaseMDP.step, BayesianMDPModel.step_update (agent/mdp_models/bayesian_model.py:78-92) =
 * N_NIG.update_sa with one reward (bayesian_models/conjugate_rewards.py:56-74) + M_DIR.update_sa unless the episode
 * ended (conjugate_transitions.py:43-45), in numpy's types. */
typedef struct {
  long long N;
  uint64_t seed, env0;
  int* state;
  int* h;
  const float* Q;
  float* dir_hyper;
  float* nig_hyper;
  double epsilon_greedy;
  double* cum_reward;
  long long* n_episodes;
  int* trace;
  int reward_model; /* 0 N_NIG, 1 N_N (conjugate_rewards.py:112-117) */
  orc_actor_args actor;
} orc_psrl_args;

int orc_psrl_steps(const orc_tables* tb, const orc_psrl_args* p, int n_steps, uint64_t t0) {
  const int S = tb->S, A = tb->A, H = tb->H;
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < p->N; ++i) {
    const float* Q = p->Q + (size_t)i * (H + 1) * S * A;
    float* dir = p->dir_hyper + (size_t)i * S * A * S;
    float* nig = p->nig_hyper + (size_t)i * S * A * 4;
    int s = p->state[i], h = p->h[i];
    double cum = p->cum_reward[i];
    for (int step = 0; step < n_steps; ++step) {
      const uint64_t t = t0 + (uint64_t)step;
      uint32_t we[4], wa[4];
      orc_philox(p->seed, p->env0 + (uint64_t)i, t, we);
      orc_philox(p->seed ^ 0x9E3779B97F4A7C15ULL, p->env0 + (uint64_t)i, t, wa);
      const int a = orc_actor_select(Q + ((size_t)h * S + s) * A, A, A, p->epsilon_greedy, &p->actor, (long long)t, wa, p->seed,
                                     p->env0 + (uint64_t)i);
      const size_t base = ((size_t)s * A + a) * tb->Ksucc;
      const int nsucc = tb->succ_len[(size_t)s * A + a];
      int pos = 0;
      if (nsucc > 1) {
        const double total = tb->succ_cum[base + nsucc - 1] + 0.0;
        pos = bisect_pos(tb->succ_cum + base, nsucc, u53(we[0], we[1]) * total);
      }
      const int nxt = tb->succ_idx[base + pos];
      const int cls = tb->rew_cls_succ ? tb->rew_cls_succ[base + pos] : 0;
      const float r = reward_draw(tb, cls, u24(we[2]));
      const int hh = h + 1;
      const int last = hh >= H;
      float* hp = nig + ((size_t)s * A + a) * 4;
      const float mu0 = hp[0], l0 = hp[1], a0 = hp[2], b0 = hp[3];
      if (p->reward_model == 1) { /* N_N.update_sa: float32 operands and weak python scalars -> float32 arithmetic */
        const float t1 = l0 + 1.0f;
        const float num = mu0 * l0 + r;
        hp[0] = num / t1;
        hp[1] = t1;
      } else {
        const double y = (double)r; /* np.mean([r]) */
        const float l1 = l0 + 1.0f;
        const float lm = l0 * mu0;
        const double mu1 = ((double)lm + y) / (double)l1;
        const double dy = y - (double)mu0;
        const double disc = (double)l0 * (dy * dy) / (double)l1;
        hp[0] = (float)mu1;
        hp[1] = l1;
        hp[2] = a0 + 0.5f;
        hp[3] = (float)((double)b0 + 0.5 * (0.0 + disc));
      }
      if (!last) dir[((size_t)s * A + a) * S + nxt] += 1.0f;
      cum += (double)r;
      if (p->trace) {
        int* tr = p->trace + ((size_t)step * p->N + i) * 4;
        union { float f; int i; } u;
        u.f = r;
        tr[0] = s; tr[1] = a; tr[2] = last ? -1 : nxt; tr[3] = u.i;
      }
      if (last) {
        if (p->n_episodes) p->n_episodes[i] += 1;
        h = 0;
        s = sample_start(tb, u53(wa[2], wa[3]));
      } else {
        h = hh;
        s = nxt;
      }
    }
    p->state[i] = s;
    p->h[i] = h;
    p->cum_reward[i] = cum;
  }
  return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------------------------
 * UCRL2Continuous (colosseum/agent/agents/infinite_horizon/ucrl2.py:34-357) for N independent loops, restated per loop:
 *   orc_ucrl2_steps         MDPLoop.run's body until the loop's time reaches t_target or its artificial episode ends:
 *                           QValuesActor.select_action on Q[s], BaseMDP.step, step_update (:183-199: N[s,a,s'] += 1 and the
 *                           episode's per-(s,a) reward / next-state lists, kept here as one time-ordered log per loop),
 *                           is_episode_end (:173-181: nu_k >= max(1, N[s,a].sum() - nu_k));
 *   orc_ucrl2_bounds        episode_end_update's first lines (:183-186) and the bounds of solve_optimistic_model
 *                           (:223-311): delta = 1/sqrt(iteration+1), beta_r (_chernoff), beta_p (_chernoff / bernstein;
 *                           only the [s,a,0] entry, the one _max_proba reads, infinite_horizon.py:230);
 *   orc_ucrl2_model_update  model_update (:201-221) in numpy's types: a float32 table entry times / plus a float64 is
 *                           computed in float64 and stored float32; `r - old_estimate` with a python-float reward is a
 *                           float32 difference (NEP 50).
 * Pinned against the unmodified reference class replayed on this restatement's trace (tests/golden/make_ucrl2_golden.py).
 */
typedef struct {
  long long N;
  uint64_t seed, env0;
  int* state;
  long long* t;
  double* cum_reward;
  const float* Q;
  int* Nsas;
  int* Nsa;
  float* P;
  float* est_r;
  float* var_r;
  float* hold;
  int* nu;
  int* seen;
  int* ep_len;
  int* ep_log;
  int log_cap;
  int* ended;
  long long* iteration;
  long long* episode;
  double* delta;
  double epsilon_greedy;
  int* trace;
  long long trace_t0;
  int trace_steps;
  orc_actor_args actor;
} orc_ucrl2_args;

int orc_ucrl2_steps(const orc_tables* tb, const orc_ucrl2_args* p, long long t_target) {
  const int S = tb->S, A = tb->A;
  const size_t SA = (size_t)S * A;
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < p->N; ++i) {
    if (p->ended[i] != 0) continue;
    const float* Q = p->Q + (size_t)i * SA;
    int* Nsas = p->Nsas + (size_t)i * SA * S;
    int* Nsa = p->Nsa + (size_t)i * SA;
    int* nu = p->nu + (size_t)i * SA;
    int* log = p->ep_log + (size_t)i * p->log_cap * 2;
    int s = p->state[i];
    long long t = p->t[i];
    int len = p->ep_len[i];
    double cum = p->cum_reward[i];
    int flag = 0;
    while (t < t_target) {
      if (len >= p->log_cap) { flag = 2; break; }
      uint32_t we[4], wa[4];
      orc_philox(p->seed, p->env0 + (uint64_t)i, (uint64_t)t, we);
      orc_philox(p->seed ^ 0x9E3779B97F4A7C15ULL, p->env0 + (uint64_t)i, (uint64_t)t, wa);
      const int a = orc_actor_select(Q + (size_t)s * A, A, A, p->epsilon_greedy, &p->actor, t, wa, p->seed, p->env0 + (uint64_t)i);
      const size_t sa = (size_t)s * A + a;
      const size_t base = sa * tb->Ksucc;
      const int nsucc = tb->succ_len[sa];
      int pos = 0;
      if (nsucc > 1) {
        const double total = tb->succ_cum[base + nsucc - 1] + 0.0;
        pos = bisect_pos(tb->succ_cum + base, nsucc, u53(we[0], we[1]) * total);
      }
      const int nxt = tb->succ_idx[base + pos];
      const int cls = tb->rew_cls_succ ? tb->rew_cls_succ[base + pos] : 0;
      const float r = reward_draw(tb, cls, u24(we[2]));
      Nsas[sa * S + nxt] += 1;
      Nsa[sa] += 1;
      nu[sa] += 1;
      union { float f; int i; } u;
      u.f = r;
      log[2 * len] = (int)sa;
      log[2 * len + 1] = u.i;
      ++len;
      cum += (double)r;
      if (p->trace) {
        const long long k = t - p->trace_t0;
        if (k >= 0 && k < p->trace_steps) {
          int* tr = p->trace + ((size_t)k * p->N + i) * 4;
          tr[0] = s; tr[1] = a; tr[2] = nxt; tr[3] = u.i;
        }
      }
      s = nxt;
      ++t;
      const int before = Nsa[sa] - nu[sa];
      if (nu[sa] >= (before > 1 ? before : 1)) { flag = 1; break; }
    }
    p->state[i] = s;
    p->t[i] = t;
    p->ep_len[i] = len;
    p->cum_reward[i] = cum;
    if (flag) p->ended[i] = flag;
  }
  return ORC_OK;
}

int orc_ucrl2_bounds(const orc_ucrl2_args* p, int S, int A, const int* index, int m, double alpha_r, double alpha_p,
                     double r_max, int bernstein_p, double* beta_r, double* beta_p) {
  const size_t SA = (size_t)S * A;
  for (int k = 0; k < m; ++k) {
    const size_t i = (size_t)index[k];
    const long long it = p->iteration[i];
    const double delta = 1.0 / sqrt((double)(it + 1));
    for (size_t sa = 0; sa < SA; ++sa) {
      const int nb = p->Nsa[i * SA + sa];
      const double n1 = (double)(nb > 1 ? nb : 1);
      const double Lr = log((double)(2LL * S * A * (it + 1)) / delta);
      beta_r[(size_t)k * SA + sa] = alpha_r * (r_max * sqrt(3.5 * Lr / n1));
      double bp;
      if (!bernstein_p) {
        const double Lp = log((double)(2LL * A * (it + 1)) / delta);
        bp = alpha_p * sqrt((double)(14LL * S) * Lp / n1);
      } else {
        const double nm1 = (double)(nb - 1 > 1 ? nb - 1 : 1);
        const float P0 = p->P[(i * SA + sa) * S];
        const float one_m = 1.0f - P0;
        const float var_p = P0 * one_m;
        const float v14 = 14.0f * var_p;
        const double L = log(2.0 * (double)S * (double)A * (double)(it + 1) / delta);
        const double At = (double)v14 / n1 * L;
        const double Bt = 49.0 / (3.0 * nm1) * L;
        bp = sqrt(alpha_p) * sqrt(At) + alpha_p * Bt;
      }
      beta_p[(size_t)k * SA + sa] = bp;
    }
    p->delta[i] = delta;
    p->episode[i] += 1;
  }
  return ORC_OK;
}

int orc_ucrl2_model_update(const orc_ucrl2_args* p, int S, int A, const int* index, int m) {
  const size_t SA = (size_t)S * A;
  for (int k = 0; k < m; ++k) {
    const size_t i = (size_t)index[k];
    const int* log = p->ep_log + i * (size_t)p->log_cap * 2;
    const int len = p->ep_len[i];
    const int* Nsa = p->Nsa + i * SA;
    int* seen = p->seen + i * SA;
    int* nu = p->nu + i * SA;
    float* est = p->est_r + i * SA;
    float* var = p->var_r + i * SA;
    float* hold = p->hold + i * SA;
    for (int e = 0; e < len; ++e) {
      const int sa = log[2 * e];
      union { float f; int i; } u;
      u.i = log[2 * e + 1];
      const float r = u.f;
      const int j = ++seen[sa];
      const double sf = (double)((long long)Nsa[sa] + j);
      const double sf1 = sf + 1.0;
      const double ratio = sf / sf1;
      const float old = est[sa];
      float x = (float)((double)old * ratio);
      x = (float)((double)x + (double)r / sf1);
      est[sa] = x;
      const float d0 = r - old, d1 = r - x;
      const float pr = d0 * d1;
      var[sa] = var[sa] + pr;
      float hd = (float)((double)hold[sa] * ratio);
      hd = (float)((double)hd + 1.0 / sf1);
      hold[sa] = hd;
    }
    for (int e = 0; e < len; ++e) seen[log[2 * e]] = 0;
    for (size_t sa = 0; sa < SA; ++sa) {
      if (nu[sa] == 0) continue;
      const double tot = (double)Nsa[sa];
      const int* n = p->Nsas + (i * SA + sa) * S;
      float* P = p->P + (i * SA + sa) * S;
      for (int j = 0; j < S; ++j) P[j] = (float)((double)n[j] / tot);
      nu[sa] = 0;
    }
    p->iteration[i] += len;
    p->ep_len[i] = 0;
    p->ended[i] = 0;
  }
  return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------------------------
 * PSRLContinuous between two re-plannings (colosseum/agent/agents/infinite_horizon/posterior_sampling.py:333-345,
 * :389-412, :449-452) for N independent loops: QValuesActor.select_action on the EXTENDED q-values Q[N,S,A*psi], real
 * action = int(action / psi) (also for the epsilon-greedy draw, which comes from range(A), sic),
 * BayesianMDPModel.step_update (N_NIG / N_N posterior, Dirichlet count += 1), N[s,a,s'] += 1, and is_episode_end:
 * N_tau >= 2 (N_tau - nu_k).  Each loop stops at t_target or when its artificial episode ends (ended[i] = 1).
 */
typedef struct {
  long long N;
  uint64_t seed, env0;
  int* state;
  long long* t;
  double* cum_reward;
  const float* Q;
  int psi;
  float* dir_hyper;
  float* nig_hyper;
  int reward_model;
  int* Nsas;
  int* Nsa;
  int* nu;
  int* ended;
  long long* episode;
  double epsilon_greedy;
  int* trace;
  long long trace_t0;
  int trace_steps;
  orc_actor_args actor;
} orc_psrlc_args;

int orc_psrlc_steps(const orc_tables* tb, const orc_psrlc_args* p, long long t_target) {
  const int S = tb->S, A = tb->A, psi = p->psi, AE = A * psi;
  const size_t SA = (size_t)S * A;
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < p->N; ++i) {
    if (p->ended[i] != 0) continue;
    const float* Q = p->Q + (size_t)i * S * AE;
    float* dir = p->dir_hyper + (size_t)i * SA * S;
    float* nig = p->nig_hyper + (size_t)i * SA * 4;
    int* Nsas = p->Nsas + (size_t)i * SA * S;
    int* Nsa = p->Nsa + (size_t)i * SA;
    int* nu = p->nu + (size_t)i * SA;
    int s = p->state[i];
    long long t = p->t[i];
    double cum = p->cum_reward[i];
    int flag = 0;
    while (t < t_target) {
      uint32_t we[4], wa[4];
      orc_philox(p->seed, p->env0 + (uint64_t)i, (uint64_t)t, we);
      orc_philox(p->seed ^ 0x9E3779B97F4A7C15ULL, p->env0 + (uint64_t)i, (uint64_t)t, wa);
      const int a_ext = orc_actor_select(Q + (size_t)s * AE, AE, A, p->epsilon_greedy, &p->actor, t, wa, p->seed,
                                         p->env0 + (uint64_t)i);
      const int a = a_ext / psi;
      const size_t sa = (size_t)s * A + a;
      const size_t base = sa * tb->Ksucc;
      const int nsucc = tb->succ_len[sa];
      int pos = 0;
      if (nsucc > 1) {
        const double total = tb->succ_cum[base + nsucc - 1] + 0.0;
        pos = bisect_pos(tb->succ_cum + base, nsucc, u53(we[0], we[1]) * total);
      }
      const int nxt = tb->succ_idx[base + pos];
      const int cls = tb->rew_cls_succ ? tb->rew_cls_succ[base + pos] : 0;
      const float r = reward_draw(tb, cls, u24(we[2]));
      float* hp = nig + sa * 4;
      const float mu0 = hp[0], l0 = hp[1], a0 = hp[2], b0 = hp[3];
      if (p->reward_model == 1) {
        const float t1 = l0 + 1.0f;
        const float num = mu0 * l0 + r;
        hp[0] = num / t1;
        hp[1] = t1;
      } else {
        const double y = (double)r;
        const float l1 = l0 + 1.0f;
        const float lm = l0 * mu0;
        const double mu1 = ((double)lm + y) / (double)l1;
        const double dy = y - (double)mu0;
        const double disc = (double)l0 * (dy * dy) / (double)l1;
        hp[0] = (float)mu1;
        hp[1] = l1;
        hp[2] = a0 + 0.5f;
        hp[3] = (float)((double)b0 + 0.5 * (0.0 + disc));
      }
      dir[sa * S + nxt] += 1.0f;
      Nsas[sa * S + nxt] += 1;
      Nsa[sa] += 1;
      nu[sa] += 1;
      cum += (double)r;
      if (p->trace) {
        const long long k = t - p->trace_t0;
        if (k >= 0 && k < p->trace_steps) {
          int* tr = p->trace + ((size_t)k * p->N + i) * 4;
          union { float f; int i; } u;
          u.f = r;
          tr[0] = s; tr[1] = a_ext; tr[2] = nxt; tr[3] = u.i;
        }
      }
      s = nxt;
      ++t;
      if (Nsa[sa] >= 2 * (Nsa[sa] - nu[sa])) { flag = 1; break; }
    }
    p->state[i] = s;
    p->t[i] = t;
    p->cum_reward[i] = cum;
    if (flag) p->ended[i] = flag;
  }
  return ORC_OK;
}
