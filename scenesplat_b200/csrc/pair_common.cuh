// tcgen05 cta_group::2 (CTA pair) primitives shared by the pair kernels (gemm2cta.cu, conv_gemm3.cu).
// PTX forms follow the vendored CUTLASS headers (cute/arch/copy_sm100_tma.hpp, cutlass/arch/barrier.h,
// cute/arch/tmem_allocator_sm100.hpp): both CTAs issue the TMA copies with the peer bit of the barrier address cleared,
// so the transaction bytes are counted on the even (leader) CTA's barrier; commits are multicast to both CTAs.
#pragma once
#include "tc_common.cuh"

namespace ss {
namespace pair {
__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {  // the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// TMA tile load of one CTA of the pair; the transaction bytes are counted on the LEADER's barrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
  const uint32_t b = tc::smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          dst_smem),
      "l"(m), "r"(b), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs once the pair's previous MMAs have completed
__device__ __forceinline__ void umma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}" ::"r"(
          tc::smem_u32(bar))
      : "memory");
}
// arrive on the barrier at this offset in CTA `target` of the cluster
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t target) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(tc::smem_u32(bar)),
      "r"(target)
      : "memory");
}
// same, without the release fence (a flag hand-over: the data it announces was written by asynchronous copies whose
// completion the caller has already observed through a local mbarrier)
__device__ __forceinline__ void mbar_arrive_cta_relaxed(uint64_t* bar, uint32_t target) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(tc::smem_u32(bar)),
      "r"(target)
      : "memory");
}
}  // namespace pair

}  // namespace ss
