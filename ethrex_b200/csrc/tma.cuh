// tma.cuh -- the bulk-copy half of the Tensor Memory Accelerator (cp.async.bulk + mbarrier, SASS: UBLKCP /
// SYNCS): whole tiles are fetched global -> shared memory by the copy engine, with no register staging and no
// LSU instructions, and their arrival is signalled on a shared-memory mbarrier.  1-D bulk copies are enough
// here: every tile this library stages is a run (or a few hundred runs) of contiguous bytes.
#pragma once
#include <cstdint>

namespace b200zk {
namespace tma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void barrier_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals));
}
// make the initialised barrier visible to the async proxy before any copy names it
__device__ __forceinline__ void barrier_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// one arrival + announce `bytes` of incoming copy traffic
__device__ __forceinline__ void barrier_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void barrier_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// bytes: multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}

}  // namespace tma
}  // namespace b200zk
