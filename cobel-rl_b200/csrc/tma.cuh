// tma.cuh -- bulk asynchronous copies (the TMA's 1-D path: cp.async.bulk + mbarrier; SASS UBLKCP / SYNCS) used for
// row streaming (sr.cu) and for staging per-agent tables in and out of shared memory (pma.cu, dynaq.cu, qagent.cu).
// Addresses and sizes must be multiples of 16 bytes.  sm_100a only.
#pragma once
#include "common.cuh"

COBEL_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
COBEL_DEV void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
COBEL_DEV void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
COBEL_DEV void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, completion signalled on the mbarrier (SASS: UBLKCP.S.G + SYNCS)
COBEL_DEV void bulk_load(void* dst_smem, const void* src_gmem, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global (SASS: UBLKCP.G.S)
COBEL_DEV void bulk_store(void* dst_gmem, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
COBEL_DEV void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
COBEL_DEV void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// shared -> global without waiting for completion: commit, then bulk_store_wait() before the source is reused
// or the CTA exits
COBEL_DEV void bulk_store_issue(void* dst_gmem, const void* src_smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
COBEL_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
COBEL_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
COBEL_DEV bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
