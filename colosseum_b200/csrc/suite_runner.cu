// suite_runner.cu -- the C3 work item (BASELINE.json configs[2]) run natively: for each benchmark MDP instance, N parallel
// episodes x n random-agent steps through the step kernel, then the three hardness measures with the property layer's
// choices of the reference (colosseum/hardness/analysis.py:327-421, colosseum/mdp/base.py:996-1114,
// mdp/base_finite.py:167-178) -- the sequence colosseum_b200/suite.py runs from Python, as ONE call per shard.
//
// Why native: an instance is a chain of small latency-bound solves (one CTA, or one CTA per few targets, for an MDP
// with a few hundred states), so the GPU is filled by running many instances at once, each on its own stream.  Python
// worker threads cannot do that: the glue between the solves holds the GIL (measured: 1 / 4 / 8 / 16 Python workers ->
// 125 / 169 / 120 / 116 instances/s).  Here the workers are C++ threads; every solver call is the public C ABI of
// include/colosseum_b200.h, exactly what the Python path calls, so the numbers are the same numbers.
#include <math.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace colo {

template <typename T>
__global__ void suite_gather_nodes_kernel(const T* __restrict__ V_hs, const int* __restrict__ node_h,
                                          const int* __restrict__ node_s, int n, int S, T* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = V_hs[(size_t)node_h[i] * S + node_s[i]];
}

__global__ void suite_iota_kernel(int* x, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = i;
}

// Large tensors of an instance (T_epi, T_cf up to gigabytes, the diameter's work buffers) do not go through the
// stream-ordered pool every time: a worker keeps its k-th large buffer from instance to instance (grow-only; the worker
// runs its instances one after the other on ONE stream, so reuse is ordered).  Measured reason: with 8 workers carving
// multi-gigabyte requests out of the shared pool, an instance that takes 10 ms alone sporadically took 0.3-1.6 s (the
// pool grows / defragments under its lock and every other worker's allocation waits behind it).
struct WorkerCache {
  std::vector<std::pair<void*, size_t>> bufs;
  int dev = -1;
  void drop(cudaStream_t st) {
    for (auto& b : bufs)
      if (b.first) cudaFreeAsync(b.first, st);
    bufs.clear();
  }
};
constexpr size_t kCachedAllocBytes = (size_t)4 << 20;

struct StreamArena {  // stream-ordered allocations of one instance, released together
  cudaStream_t st;
  WorkerCache* cache;
  size_t next_cached = 0;
  std::vector<void*> ptrs;
  explicit StreamArena(cudaStream_t s, WorkerCache* c = nullptr) : st(s), cache(c) {}
  template <typename T>
  T* alloc(size_t n) {
    void* p = nullptr;
    const size_t bytes = (n ? n : 1) * sizeof(T);
    if (cache && bytes >= kCachedAllocBytes) {
      if (next_cached == cache->bufs.size()) cache->bufs.emplace_back(nullptr, 0);
      auto& slot = cache->bufs[next_cached++];
      if (slot.second < bytes) {
        if (slot.first) cudaFreeAsync(slot.first, st);
        slot = {nullptr, 0};
        const size_t cap = bytes + bytes / 4;
        if (cudaMallocAsync(&p, cap, st) != cudaSuccess) return nullptr;
        slot = {p, cap};
      }
      return (T*)slot.first;
    }
    if (cudaMallocAsync(&p, bytes, st) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    return (T*)p;
  }
  template <typename T>
  T* upload(const T* host, size_t n) {
    T* p = alloc<T>(n);
    if (p && cudaMemcpyAsync(p, host, n * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess) return nullptr;
    return p;
  }
  void release() {
    for (void* p : ptrs) cudaFreeAsync(p, st);
    ptrs.clear();
  }
  ~StreamArena() { release(); }
};

#define SUITE_TRY(expr)            \
  do {                             \
    const int _r = (expr);         \
    if (_r != COLO_OK) return _r;  \
  } while (0)
#define SUITE_PTR(p)                                                \
  do {                                                              \
    if ((p) == nullptr) {                                           \
      set_error("suite: device allocation / upload failed: %s", cudaGetErrorString(cudaGetLastError())); \
      return COLO_ERR_CUDA;                                         \
    }                                                               \
  } while (0)

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int d2h_double(const double* d, double* h, cudaStream_t st) {
  COLO_CUDA_TRY(cudaMemcpyAsync(h, d, sizeof(double), cudaMemcpyDeviceToHost, st));
  COLO_CUDA_TRY(cudaStreamSynchronize(st));
  return COLO_OK;
}

static int run_one(const colo_suite_instance& in, const colo_suite_config& cfg, cudaStream_t st, WorkerCache* cache,
                   colo_suite_result* out) {
  const int S = in.S, A = in.A, H = in.H, K = in.K;
  const size_t SA = (size_t)S * A;
  StreamArena ar(st, cache);
  cudaGetLastError();  // a failure of this thread's previous instance must not be reported against this one
  const double t0 = now_s();
  // ---------------------------------------------------------------- tables
  float* T = ar.upload(in.T, SA * S);
  float* R = ar.upload(in.R, SA);
  double* succ_cum = ar.upload(in.succ_cum, SA * K);
  int* succ_idx = ar.upload(in.succ_idx, SA * K);
  int* succ_len = ar.upload(in.succ_len, SA);
  int* rew_cls = ar.upload(in.rew_cls_succ, SA * K);
  float* rew_q = ar.upload(in.rew_q, (size_t)in.n_cls * in.nq);
  double* start_cum = ar.upload(in.start_cum, (size_t)in.n_start);
  int* start_idx = ar.upload(in.start_idx, (size_t)in.n_start);
  SUITE_PTR(T); SUITE_PTR(R); SUITE_PTR(succ_cum); SUITE_PTR(succ_idx); SUITE_PTR(succ_len); SUITE_PTR(rew_cls);
  SUITE_PTR(rew_q); SUITE_PTR(start_cum); SUITE_PTR(start_idx);
  // ---------------------------------------------------------------- step phase (BaseMDP.random_steps, base.py:1319-1355)
  {
    const long long N = cfg.n_envs;
    const int copies = 16;
    colo_mdp_tables tb;
    memset(&tb, 0, sizeof(tb));
    tb.S = S; tb.A = A; tb.H = H;
    tb.succ_cum = succ_cum; tb.succ_idx = succ_idx; tb.succ_len = succ_len; tb.Ksucc = K; tb.rew_cls_succ = rew_cls;
    tb.rew_q = rew_q; tb.n_cls = in.n_cls; tb.nq = in.nq; tb.rmin = in.rmin; tb.rmax = in.rmax;
    tb.start_cum = start_cum; tb.start_idx = start_idx; tb.n_start = in.n_start;
    colo_env_batch b;
    memset(&b, 0, sizeof(b));
    b.N = N; b.seed = cfg.seed; b.env0 = 0;
    b.state = ar.alloc<int>(N); b.h = ar.alloc<int>(N); b.step_type = ar.alloc<unsigned char>(N);
    b.action = ar.alloc<int>(N); b.reward = ar.alloc<float>(N); b.obs = ar.alloc<int>(N);
    b.visits_s = ar.alloc<unsigned long long>((size_t)copies * S);
    b.visits_sa = ar.alloc<unsigned long long>((size_t)copies * SA);
    b.visits_copies = copies;
    b.status = ar.alloc<int>(1);
    SUITE_PTR(b.state); SUITE_PTR(b.h); SUITE_PTR(b.step_type); SUITE_PTR(b.action); SUITE_PTR(b.reward); SUITE_PTR(b.obs);
    SUITE_PTR(b.visits_s); SUITE_PTR(b.visits_sa); SUITE_PTR(b.status);
    COLO_CUDA_TRY(cudaMemsetAsync(b.step_type, COLO_STEP_LAST, (size_t)N, st));
    COLO_CUDA_TRY(cudaMemsetAsync(b.visits_s, 0, (size_t)copies * S * 8, st));
    COLO_CUDA_TRY(cudaMemsetAsync(b.visits_sa, 0, (size_t)copies * SA * 8, st));
    COLO_CUDA_TRY(cudaMemsetAsync(b.status, 0, sizeof(int), st));
    SUITE_TRY(colo_env_reset(&tb, &b, nullptr, 0, st));
    SUITE_TRY(colo_env_random_steps(&tb, &b, 2, cfg.n_steps, 1, 1, st));
    std::vector<unsigned long long> vs((size_t)copies * S);
    std::vector<float> rw((size_t)N);
    COLO_CUDA_TRY(cudaMemcpyAsync(vs.data(), b.visits_s, vs.size() * 8, cudaMemcpyDeviceToHost, st));
    COLO_CUDA_TRY(cudaMemcpyAsync(rw.data(), b.reward, rw.size() * 4, cudaMemcpyDeviceToHost, st));
    COLO_CUDA_TRY(cudaStreamSynchronize(st));
    unsigned long long tot = 0;
    for (auto v : vs) tot += v;
    double rs = 0.0;
    for (float r : rw) rs += isnan(r) ? 0.0 : (double)r;
    out->visits_total = (double)tot;
    out->mean_reward_last_step = rs / (double)N;
  }
  const double t1 = now_s();
  out->step_s = t1 - t0;
  // ---------------------------------------------------------------- hardness (f64acc)
  const double gamma = (double)0.99f;  // cast to float32 first, as the reference does (infinite_horizon.py:127)
  const double eps = cfg.eps;
  double* scal = ar.alloc<double>(2);
  SUITE_PTR(scal);
  int* targets = ar.alloc<int>(S);
  SUITE_PTR(targets);
  suite_iota_kernel<<<(S + 255) / 256, 256, 0, st>>>(targets, S);
  SUITE_TRY(check_launch("suite_iota_kernel"));
  out->gaps = out->value_norm = out->diameter = out->diameter_sweeps = NAN;
  // COLO_SUITE_VERBOSE: wall time of every phase of the instance on stderr (the stream is synchronised at each mark)
  static const bool verbose = getenv("COLO_SUITE_VERBOSE") != nullptr;
  double tm = t1;
  char phases[512];
  int plen = 0;
  phases[0] = 0;
  auto mark = [&](const char* what) {
    if (!verbose) return;
    cudaStreamSynchronize(st);
    const double t = now_s();
    plen += snprintf(phases + plen, sizeof(phases) - (size_t)plen, " %s %.1f", what, (t - tm) * 1e3);
    if (plen > (int)sizeof(phases) - 32) plen = (int)sizeof(phases) - 32;
    tm = t;
  };
  if (H == 0) {
    // continuous MDP: T, R and discounted VI (base.py:635-647, :1042-1100)
    double* Q = ar.alloc<double>(SA);
    double* V = ar.alloc<double>(S);
    void* work = ar.alloc<unsigned char>(colo_solve_work_bytes(1, S, 1));
    SUITE_PTR(Q); SUITE_PTR(V); SUITE_PTR(work);
    long long iters = 0;
    SUITE_TRY(colo_solve_discounted_f64acc(T, R, nullptr, 1, S, A, gamma, eps, 0.0, 1000000, COLO_FOLD_MAX, Q, V, &iters, work, st));
    mark("vi");
    SUITE_TRY(colo_gaps_f64(Q, V, nullptr, S, A, 0.1, scal, st));
    SUITE_TRY(d2h_double(scal, &out->gaps, st));
    if (in.deterministic) {
      out->value_norm = 0.0;  // base.py:1069-1074
    } else {
      void* w2 = ar.alloc<unsigned char>(colo_value_norm_work_bytes(S, A, 1));
      SUITE_PTR(w2);
      SUITE_TRY(colo_value_norm_f64acc(T, V, S, A, w2, scal + 1, st));
      SUITE_TRY(d2h_double(scal + 1, &out->value_norm, st));
    }
    mark("gaps+norm");
    if (cfg.diameter) {
      void* w3 = ar.alloc<unsigned char>(colo_diameter_continuous_work_bytes(S, S, 1));
      SUITE_PTR(w3);
      double dh[2] = {0, 0};
      SUITE_TRY(colo_diameter_continuous_f64acc(T, targets, S, S, A, eps, 0.0, 1000000, w3, dh, st));
      out->diameter = dh[0];
      out->diameter_sweeps = dh[1];
      mark("diameter");
    }
  } else {
    // episodic MDP: backward induction + reachable (h,s) pairs for the gaps (base.py:1018-1040), the episodic tensor for
    // the diameter (base.py:996-1016), the continuous form for the value norm (base.py:1049-1056)
    const int n = in.n_nodes;
    double* Q = ar.alloc<double>((size_t)(H + 1) * SA);
    double* V = ar.alloc<double>((size_t)(H + 1) * S);
    SUITE_PTR(Q); SUITE_PTR(V);
    SUITE_TRY(colo_episodic_f64acc(T, R, nullptr, 1, S, A, H, COLO_FOLD_MAX, 0.0, Q, V, st));
    mark("episodic_vi");
    double* start_prob = ar.upload(in.start_prob, (size_t)in.n_start);
    float* T_epi = ar.alloc<float>((size_t)H * SA * S);
    unsigned char* reach = ar.alloc<unsigned char>((size_t)H * S);
    SUITE_PTR(start_prob); SUITE_PTR(T_epi); SUITE_PTR(reach);
    SUITE_TRY(colo_build_episodic_tensor(T, R, start_idx, start_prob, in.n_start, H, S, A, T_epi, nullptr, reach, st));
    mark("T_epi");
    std::vector<unsigned char> mask((size_t)(H + 1) * S, 0);
    std::vector<int> pos((size_t)H * S, -1);
    for (int i = 0; i < n; ++i) {
      mask[(size_t)in.node_h[i] * S + in.node_s[i]] = 1;
      pos[(size_t)in.node_h[i] * S + in.node_s[i]] = i;
    }
    unsigned char* mask_d = ar.upload(mask.data(), mask.size());
    SUITE_PTR(mask_d);
    SUITE_TRY(colo_gaps_f64(Q, V, mask_d, (long long)(H + 1) * S, A, 0.1, scal, st));
    SUITE_TRY(d2h_double(scal, &out->gaps, st));
    mark("gaps");
    if (in.deterministic) {
      out->value_norm = 0.0;
    } else if ((size_t)4 * n * n * A <= cfg.max_cf_bytes) {
      int* node_h = ar.upload(in.node_h, (size_t)n);
      int* node_s = ar.upload(in.node_s, (size_t)n);
      int* pos_d = ar.upload(pos.data(), pos.size());
      float* T_cf = ar.alloc<float>((size_t)n * A * n);
      float* R_cf = ar.alloc<float>((size_t)n * A);
      int* flag = ar.alloc<int>(1);
      SUITE_PTR(node_h); SUITE_PTR(node_s); SUITE_PTR(pos_d); SUITE_PTR(T_cf); SUITE_PTR(R_cf); SUITE_PTR(flag);
      COLO_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
      SUITE_TRY(colo_build_continuous_form(T, R, node_h, node_s, n, pos_d, start_idx, start_prob, in.n_start, H, S, A, T_cf,
                                           R_cf, flag, st));
      mark("T_cf");
      // V* of the continuous form from its structure (colo_continuous_form_values_*): the node at list position start_k
      std::vector<int> ph((size_t)in.n_start), ps((size_t)in.n_start);
      std::vector<float> p32((size_t)in.n_start);
      for (int k = 0; k < in.n_start; ++k) {
        if (in.start_idx[k] >= n) {
          set_error("suite: start index %d beyond the node list (%d nodes)", in.start_idx[k], n);
          return COLO_ERR_ARG;
        }
        ph[k] = in.node_h[in.start_idx[k]];
        ps[k] = in.node_s[in.start_idx[k]];
        p32[k] = (float)in.start_prob[k];
      }
      int* ph_d = ar.upload(ph.data(), ph.size());
      int* ps_d = ar.upload(ps.data(), ps.size());
      float* p_d = ar.upload(p32.data(), p32.size());
      double* V_hs = ar.alloc<double>((size_t)H * S);
      double* V_cf = ar.alloc<double>((size_t)n);
      SUITE_PTR(ph_d); SUITE_PTR(ps_d); SUITE_PTR(p_d); SUITE_PTR(V_hs); SUITE_PTR(V_cf);
      double cf_out[2] = {0, 0};
      SUITE_TRY(colo_continuous_form_values_f64acc(T, R, S, A, H, gamma, ph_d, ps_d, p_d, in.n_start, eps, 200, V_hs, cf_out, st));
      mark("cf_values");
      suite_gather_nodes_kernel<double><<<(n + 255) / 256, 256, 0, st>>>(V_hs, node_h, node_s, n, S, V_cf);
      SUITE_TRY(check_launch("suite_gather_nodes_kernel"));
      void* w2 = ar.alloc<unsigned char>(colo_value_norm_work_bytes(n, A, 1));
      SUITE_PTR(w2);
      SUITE_TRY(colo_value_norm_f64acc(T_cf, V_cf, n, A, w2, scal + 1, st));
      SUITE_TRY(d2h_double(scal + 1, &out->value_norm, st));
      int hflag = 0;
      COLO_CUDA_TRY(cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
      COLO_CUDA_TRY(cudaStreamSynchronize(st));
      if (hflag) {
        set_error("suite: a positive-probability successor is missing from the node list");
        return COLO_ERR_ARG;
      }
      mark("cf_norm");
    }  // else: the reference raises "Its continuous form is too large" (mdp_creation.py:152-155): NaN
    if (cfg.diameter) {
      void* w3 = ar.alloc<unsigned char>(colo_diameter_episodic_work_bytes(S, H, S, A, 1));
      SUITE_PTR(w3);
      double dh[2] = {0, 0};
      SUITE_TRY(colo_diameter_episodic_f64acc(T_epi, targets, S, H, S, A, eps, 0.0, 1000000, w3, dh, st));
      out->diameter = dh[0];
      out->diameter_sweeps = dh[1];
      mark("diameter");
    }
  }
  // synchronise BEFORE the blocks go back to the pool: a block freed behind pending work can be handed to another
  // worker's stream with a dependency on that work (the pool's "internal dependencies" reuse), i.e. the other
  // instance would wait for this one's solves
  COLO_CUDA_TRY(cudaStreamSynchronize(st));
  ar.release();
  out->hardness_s = now_s() - t1;
  if (verbose)
    fprintf(stderr, "[suite] S=%d A=%d H=%d nodes=%d step %.1f ms hardness %.1f ms:%s\n", S, A, H, in.n_nodes, out->step_s * 1e3,
            out->hardness_s * 1e3, phases);
  return COLO_OK;
}

std::mutex& suite_run_mutex() {
  static std::mutex m;
  return m;
}
std::vector<WorkerCache>& suite_caches() {
  static std::vector<WorkerCache> c;
  return c;
}

}  // namespace colo

extern "C" {

int colo_suite_release_caches(void) {
  std::lock_guard<std::mutex> run_lock(colo::suite_run_mutex());
  int dev = 0;
  COLO_CUDA_TRY(cudaGetDevice(&dev));
  for (auto& c : colo::suite_caches()) {
    if (c.dev >= 0 && cudaSetDevice(c.dev) == cudaSuccess) {
      c.drop((cudaStream_t)0);
      cudaStreamSynchronize((cudaStream_t)0);
    }
    c.bufs.clear();
    c.dev = -1;
  }
  COLO_CUDA_TRY(cudaSetDevice(dev));
  cudaMemPool_t mp;
  if (cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) {
    cudaDeviceSynchronize();
    cudaMemPoolTrimTo(mp, 0);  // colo_suite_run keeps freed blocks in the pool: hand them back to the device
  }
  cudaGetLastError();
  return COLO_OK;
}

int colo_suite_run(const colo_suite_instance* inst, int n, const colo_suite_config* cfg, colo_suite_result* out,
                   int n_workers) {
  COLO_ARG_CHECK(inst && cfg && out && n >= 0 && n_workers >= 1, "inst, cfg, out, n_workers >= 1");
  COLO_ARG_CHECK(cfg->n_envs >= 1 && cfg->n_steps >= 1 && cfg->eps > 0.0, "n_envs, n_steps, eps");
  int dev = 0;
  COLO_CUDA_TRY(cudaGetDevice(&dev));
  // The workers spend their time waiting in cudaStreamSynchronize.  Spinning there (the default schedule) wakes a worker
  // at once -- measured on 128 instances, 8 workers: 0.25 s spinning against 0.31-0.36 s with blocking waits, whose
  // wake-up latency stretches the most expensive episodic instances (dozens of host round trips) from ~60 ms to ~230 ms --
  // but n_workers x ranks spinning threads need as many cores.  So: spin when this process may run on at least n_workers
  // cores (sched_getaffinity), otherwise ask the runtime to block for the duration of the call (honoured when the
  // primary context accepts a flag change).  COLO_SUITE_SPIN=1 / COLO_SUITE_BLOCK=1 force either.
  int cores = 0;
  {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
  }
  const bool want_block = getenv("COLO_SUITE_BLOCK") != nullptr || (getenv("COLO_SUITE_SPIN") == nullptr && cores < n_workers);
  unsigned old_flags = 0;
  const bool have_flags = cudaGetDeviceFlags(&old_flags) == cudaSuccess;
  const bool blocking = have_flags && want_block &&
                        cudaSetDeviceFlags((old_flags & ~(unsigned)cudaDeviceScheduleMask) | cudaDeviceScheduleBlockingSync) == cudaSuccess;
  cudaGetLastError();
  if (getenv("COLO_SUITE_VERBOSE")) fprintf(stderr, "[colo_suite_run] %d cores, %d workers: %s\n", cores, n_workers, blocking ? "blocking waits" : "spinning waits");
  // Warm the stream-ordered memory pool: every instance allocates its tensors with cudaMallocAsync (T_epi, T_cf: up
  // to gigabytes), and a pool that has to grow from the OS in the middle of the run stalls every worker (measured: the
  // first pass over a shard ran at a third of the rate of the second).  Keep freed blocks in the pool and reserve the
  // footprint of the n_workers largest instances once.
  {
    cudaMemPool_t mp;
    if (cudaDeviceGetDefaultMemPool(&mp, dev) == cudaSuccess) {
      unsigned long long thr = ~0ULL;
      cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &thr);
      // never let an allocation of one worker's stream wait for the pending work of another worker's stream
      int no = 0;
      cudaMemPoolSetAttribute(mp, cudaMemPoolReuseAllowInternalDependencies, &no);
      std::vector<size_t> need((size_t)n);
      for (int i = 0; i < n; ++i) {
        const size_t S = inst[i].S, A = inst[i].A, H = inst[i].H, nn = inst[i].n_nodes;
        size_t b = 4 * S * A * S * 2 + (64u << 20);
        if (H > 0) {
          b += 4 * H * S * A * S + 8 * 2 * S * S * H / (H > 0 ? 1 : 1);  // T_epi + the diameter's F [K,H,S] ping-pong
          if (4 * nn * nn * A <= cfg->max_cf_bytes) b += 4 * nn * nn * A;
        } else {
          b += 8 * 4 * S * S;  // the diameter's E buffers
        }
        need[(size_t)i] = b;
      }
      std::sort(need.begin(), need.end(), [](size_t a, size_t b) { return a > b; });
      size_t reserve = 0;
      for (int w = 0; w < n_workers && w < n; ++w) reserve += need[(size_t)w];
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && reserve > free_b / 2) reserve = free_b / 2;
      unsigned long long have = 0;
      cudaMemPoolGetAttribute(mp, cudaMemPoolAttrReservedMemCurrent, &have);
      if (reserve > have) {
        void* blk = nullptr;
        if (cudaMallocAsync(&blk, reserve - have, (cudaStream_t)0) == cudaSuccess) cudaFreeAsync(blk, (cudaStream_t)0);
        cudaStreamSynchronize((cudaStream_t)0);
      }
    }
    cudaGetLastError();
  }
  std::atomic<int> next{0};
  std::atomic<int> first_error{COLO_OK};
  std::vector<std::thread> pool;
  const int W = n_workers < n ? n_workers : (n > 0 ? n : 1);
  // the workers' large-buffer caches outlive the call (a warm-up pass leaves them filled); one call at a time
  std::mutex& run_mutex = colo::suite_run_mutex();
  std::vector<colo::WorkerCache>& caches = colo::suite_caches();
  std::lock_guard<std::mutex> run_lock(run_mutex);
  if ((int)caches.size() < W) caches.resize((size_t)W);
  for (auto& c : caches)
    if (c.dev != dev) {
      if (c.dev >= 0 && cudaSetDevice(c.dev) == cudaSuccess) {
        c.drop((cudaStream_t)0);
        cudaStreamSynchronize((cudaStream_t)0);
      }
      c.bufs.clear();
      c.dev = dev;
    }
  cudaSetDevice(dev);
  for (int w = 0; w < W; ++w)
    pool.emplace_back([&, dev, w]() {
      cudaSetDevice(dev);
      cudaStream_t st = nullptr;
      if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) {
        first_error.store(COLO_ERR_CUDA);
        return;
      }
      for (;;) {
        const int i = next.fetch_add(1);
        if (i >= n) break;
        memset(&out[i], 0, sizeof(out[i]));
        const int r = colo::run_one(inst[i], *cfg, st, &caches[(size_t)w], &out[i]);
        out[i].status = r;
        if (r != COLO_OK) {
          strncpy(out[i].error, colo_last_error(), sizeof(out[i].error) - 1);
          int expected = COLO_OK;
          first_error.compare_exchange_strong(expected, r);
          cudaStreamSynchronize(st);
        }
      }
      cudaStreamSynchronize(st);
      cudaStreamDestroy(st);
    });
  for (auto& t : pool) t.join();
  if (blocking) cudaSetDeviceFlags(old_flags);
  cudaGetLastError();
  const int r = first_error.load();
  if (r != COLO_OK) colo::set_error("colo_suite_run: at least one instance failed (see colo_suite_result.error)");
  return r;
}

}  // extern "C"
