// builders.cu -- device builders for the two tensors that gate the episodic hardness measures:
//   * the episodic transition tensor T_epi[H,S,A,S] (+ R_epi[H,S,A]) that the episodic diameter runs on
//       colosseum/mdp/utils/mdp_creation.py:98-128   get_episodic_transition_matrix_and_rewards
//   * the "continuous form" of an episodic MDP over its reachable (h,s) nodes, T_cf[n,A,n] / R_cf[n,A], that the
//     episodic value norm and the continuous-form value functions run on
//       colosseum/mdp/utils/mdp_creation.py:131-176  get_continuous_form_episodic_transition_matrix_and_rewards
//       colosseum/mdp/base_finite.py:138-150          reachable_states
// The reference builds both with Python loops (the second one is O(nodes^2) `list.index` calls, seconds to minutes
// per MDP); here they are layer-by-layer scatter kernels over the dense T that is already resident in HBM.
// Reachability is the numeric one of mdp_creation.py:123-125: state s' is reachable at h+1 iff some reachable
// (s at h, a) has T[s,a,s'] > 0 -- on every golden instance this equals the reference's graph reachability.
#include "common.cuh"

namespace colo {

__global__ void epi_init_kernel(const int* __restrict__ start_idx, const double* __restrict__ start_prob, int n_start,
                                int H, int S, int A, float* __restrict__ T_epi, unsigned char* __restrict__ reach) {
  // T_epi[H-1, :, :, sn] = p   (mdp_creation.py:119-120) and reach[0, sn] = 1
  const long long rows = (long long)S * A;
  float* last = T_epi + (size_t)(H - 1) * rows * S;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (long long)gridDim.x * blockDim.x)
    for (int k = 0; k < n_start; ++k) last[r * S + start_idx[k]] = (float)start_prob[k];
  if (blockIdx.x == 0)
    for (int k = threadIdx.x; k < n_start; k += blockDim.x) reach[start_idx[k]] = 1;
}

__global__ void epi_layer_kernel(const float* __restrict__ T, int h, int H, int S, int A, float* __restrict__ T_epi,
                                 unsigned char* __restrict__ reach) {
  // block = one state s reachable at h: T_epi[h, s] = T[s] (h <= H-2), and its successors become reachable at h+1
  const int s = blockIdx.x;
  if (!reach[(size_t)h * S + s]) return;
  const long long n = (long long)A * S;
  const float* src = T + (size_t)s * n;
  float* dst = (h <= H - 2) ? T_epi + ((size_t)h * S + s) * n : nullptr;
  unsigned char* nxt = (h + 1 < H) ? reach + (size_t)(h + 1) * S : nullptr;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float t = src[i];
    if (dst) dst[i] = t;
    if (nxt && t > 0.f) nxt[i % S] = 1;
  }
}

__global__ void epi_rewards_kernel(const float* __restrict__ R, int H, long long SA, float* __restrict__ R_epi) {
  // R_epi = tile(R, H); R_epi[-1] = 0   (mdp_creation.py:126-127)
  const long long n = (long long)H * SA;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    R_epi[i] = (i / SA == H - 1) ? 0.f : R[i % SA];
}

__global__ void cf_rows_kernel(const float* __restrict__ T, const float* __restrict__ R,
                               const int* __restrict__ node_h, const int* __restrict__ node_s,
                               const int* __restrict__ pos /*[H,S] node index or -1*/,
                               const int* __restrict__ start_idx, const double* __restrict__ start_prob, int n_start,
                               int H, int S, int A, int n, float* __restrict__ T_cf, float* __restrict__ R_cf,
                               int* __restrict__ lost_mass) {
  // block = one (node i, action a) row of T_cf (zeroed by the caller)
  const int i = blockIdx.x / A, a = blockIdx.x % A;
  const int h = node_h[i], s = node_s[i];
  float* row = T_cf + ((size_t)i * A + a) * n;
  if (threadIdx.x == 0) R_cf[(size_t)i * A + a] = R[(size_t)s * A + a];
  if (h == H - 1) {
    // mdp_creation.py:166-168: the last layer jumps to the start distribution.  Sic: the reference writes column
    // node_to_index[sn] -- the start state's index in the ORIGINAL MDP, not the position of node (0, sn) in the node
    // list.  Kept: the reference's episodic value norm is computed on exactly this tensor.
    for (int k = threadIdx.x; k < n_start; k += blockDim.x) {
      const int j = start_idx[k];
      if (j < n) row[j] = (float)start_prob[k];
      else atomicExch(lost_mass, 1);
    }
  } else {  // :170-172
    const float* src = T + ((size_t)s * A + a) * S;
    const int* p1 = pos + (size_t)(h + 1) * S;
    for (int j = threadIdx.x; j < S; j += blockDim.x) {
      const float t = src[j];
      const int c = p1[j];
      if (c >= 0) row[c] = t;
      else if (t > 0.f) atomicExch(lost_mass, 1);  // a positive-probability successor is missing from the node list
    }
  }
}

}  // namespace colo

extern "C" {

int colo_build_episodic_tensor(const float* T, const float* R, const int* start_idx, const double* start_prob, int n_start,
                          int H, int S, int A, float* T_epi, float* R_epi, unsigned char* reach, void* stream) {
  COLO_ARG_CHECK(T && start_idx && start_prob && T_epi && reach, "T, start_idx, start_prob, T_epi, reach");
  COLO_ARG_CHECK(H >= 2 && S > 0 && A > 0 && n_start > 0, "H >= 2, S, A, n_start");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t layer = (size_t)S * A * S;
  COLO_CUDA_TRY(cudaMemsetAsync(T_epi, 0, (size_t)H * layer * sizeof(float), st));
  COLO_CUDA_TRY(cudaMemsetAsync(reach, 0, (size_t)H * S, st));
  const long long rows = (long long)S * A;
  colo::epi_init_kernel<<<(int)((rows + 255) / 256), 256, 0, st>>>(start_idx, start_prob, n_start, H, S, A, T_epi, reach);
  int r = colo::check_launch("epi_init_kernel");
  if (r != COLO_OK) return r;
  for (int h = 0; h < H - 1; ++h) {  // layer h fills T_epi[h] (h <= H-2) and reach[h+1]
    colo::epi_layer_kernel<<<S, 256, 0, st>>>(T, h, H, S, A, T_epi, reach);
    r = colo::check_launch("epi_layer_kernel");
    if (r != COLO_OK) return r;
  }
  if (R_epi) {
    COLO_ARG_CHECK(R != nullptr, "R is required when R_epi is requested");
    const long long n = (long long)H * S * A;
    colo::epi_rewards_kernel<<<(int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, st>>>(R, H, (long long)S * A,
                                                                                                    R_epi);
    r = colo::check_launch("epi_rewards_kernel");
  }
  return r;
}

int colo_build_continuous_form(const float* T, const float* R, const int* node_h, const int* node_s, int n,
                               const int* pos, const int* start_idx, const double* start_prob, int n_start, int H,
                               int S, int A, float* T_cf, float* R_cf, int* lost_mass_flag, void* stream) {
  COLO_ARG_CHECK(T && R && node_h && node_s && pos && start_idx && start_prob && T_cf && R_cf && lost_mass_flag,
                 "null argument");
  COLO_ARG_CHECK(H >= 1 && S > 0 && A > 0 && n > 0 && n_start > 0, "H, S, A, n, n_start");
  cudaStream_t st = (cudaStream_t)stream;
  COLO_CUDA_TRY(cudaMemsetAsync(T_cf, 0, (size_t)n * A * n * sizeof(float), st));
  COLO_CUDA_TRY(cudaMemsetAsync(lost_mass_flag, 0, sizeof(int), st));
  colo::cf_rows_kernel<<<n * A, 128, 0, st>>>(T, R, node_h, node_s, pos, start_idx, start_prob, n_start, H, S, A, n,
                                              T_cf, R_cf, lost_mass_flag);
  return colo::check_launch("cf_rows_kernel");
}

}  // extern "C"
