// tma_device.cuh -- 1-D TMA (cp.async.bulk) + mbarrier helpers shared by the kernels that stage rows of T in shared
// memory (gauss_seidel.cu: the in-place solver's prefetch ring; backup.cu: the TMA-staged variant of the synchronous sweep).
#pragma once
#include "common.cuh"

namespace colo {

__device__ __forceinline__ uint32_t gs_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gs_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gs_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   gs_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(gs_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void gs_bar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "GSW_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra GSD_%=;\n\t"
      "bra GSW_%=;\n\t"
      "GSD_%=:\n\t"
      "}" ::"r"(gs_smem_u32(bar)), "r"(parity)
      : "memory");
}

}  // namespace colo
