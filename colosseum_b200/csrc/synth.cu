// synth.cu -- on-device generator for the synthetic dense MDP of BASELINE.json config C5 (S=40,000, A=8, ~51 GB of
// T per MDP, never materialised on the host).  One warp per (row, action): pass 1 draws the weights and reduces
// their sum, pass 2 regenerates them (Philox is a pure function of its counter) and stores the normalised row.
#include "common.cuh"

namespace colo {

__device__ __forceinline__ float synth_weight(uint64_t seed, uint64_t rid, int j) {
  Philox4 w = philox4x32_10(seed, rid, (uint64_t)(j >> 2));
  const float u = u24(w.w[j & 3]);
  const float u2 = u * u, u4 = u2 * u2;
  return u4 * u4 + 1e-12f;
}

__global__ void synth_rows_kernel(float* __restrict__ T, float* __restrict__ R, int row0, int nrows, int S, int A,
                                  unsigned long long seed) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < (long long)nrows * A; r += n_warps) {
    const uint64_t rid = (uint64_t)row0 * (uint64_t)A + (uint64_t)r;
    float* row = T + (size_t)r * S;
    float sum = 0.f;
    // each lane owns whole Philox blocks (4 consecutive columns) so a block is generated once per pass
    for (int j0 = lane * 4; j0 < S; j0 += 128) {
      Philox4 w = philox4x32_10(seed, rid, (uint64_t)(j0 >> 2));
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (j0 + m < S) {
          const float u = u24(w.w[m]);
          const float u2 = u * u, u4 = u2 * u2;
          sum += u4 * u4 + 1e-12f;
        }
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int j0 = lane * 4; j0 < S; j0 += 128) {
      Philox4 w = philox4x32_10(seed, rid, (uint64_t)(j0 >> 2));
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (j0 + m < S) {
          const float u = u24(w.w[m]);
          const float u2 = u * u, u4 = u2 * u2;
          row[j0 + m] = (u4 * u4 + 1e-12f) * inv;
        }
    }
    if (lane == 0 && R) {
      Philox4 w = philox4x32_10(seed ^ 0x5bd1e995ULL, rid, 0);
      R[r] = u24(w.w[0]);
    }
  }
}

}  // namespace colo

extern "C" int colo_synth_dense_rows(float* T_rows, float* R_rows, int row0, int nrows, int S, int A,
                                     unsigned long long seed, void* stream) {
  COLO_ARG_CHECK(T_rows && nrows >= 0 && S > 0 && A > 0 && row0 >= 0, "T_rows, row0, nrows, S, A");
  if (nrows == 0) return COLO_OK;
  long long warps = (long long)nrows * A;
  long long blocks = (warps + 7) / 8;
  long long cap = (long long)colo::sm_count() * 16;
  colo::synth_rows_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(T_rows, R_rows, row0,
                                                                                               nrows, S, A, seed);
  return colo::check_launch("synth_rows_kernel");
}
