// Minimal sm_100a async-copy toolkit: mbarrier + 1-D bulk copies (TMA engine, SASS UBLKCP).
//
// All projection kernels move contiguous byte ranges, so the 1-D bulk form
// (cp.async.bulk, no tensor map) is the right TMA flavour: per-plane tensor maps would be
// illegal for the padded output anyway (row pitch (W+2p)*4 B is not a 16 B multiple).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cp360 {
namespace tma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals)
               : "memory");
}

// Predicated form: every lane executes the instruction, only lanes with `on` initialise — no branch, so the warp
// cannot be lane-divergent at a following barrier (compute-sanitizer synccheck, profiles/README.md).
__device__ __forceinline__ void mbar_init_if(bool on, uint64_t* bar, uint32_t arrivals) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p mbarrier.init.shared::cta.b64 [%0], %1;\n\t}"
               ::"r"(smem_addr(bar)), "r"(arrivals), "r"((uint32_t)on)
               : "memory");
}

// Make barrier initialisation visible to the async proxy before the first bulk copy.
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Order generic-proxy shared-memory writes before subsequent async-proxy (bulk) reads.
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// global -> shared, completion signalled on `bar` (bytes % 16 == 0, both addresses 16 B aligned).
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}

// shared -> global, tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_addr(smem_src)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// Wait until at most N of this thread's bulk groups still have un-read shared-memory sources.
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace tma
}  // namespace cp360
