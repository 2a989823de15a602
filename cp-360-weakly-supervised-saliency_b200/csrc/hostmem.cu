// Host staging buffers for the end-to-end path (frames arrive in host memory, SphericalPipeline.process_host):
// page-locked allocations the copy engines can read at PCIe speed. Three flavours, because which one feeds
// eight GPUs best depends on the host (IOMMU page size, memory placement) and is measured per box by
// tools/h2d_probe.py and bench.py's e2e.h2d_ceiling probe:
//   0  cudaHostAlloc(portable)                   what torch.Tensor.pin_memory() gives
//   1  cudaHostAlloc(portable | write-combined)  no CPU cache snooping on the way to the device; CPU reads are slow
//   2  2 MB-aligned anonymous mmap + MADV_HUGEPAGE, touched, then cudaHostRegister(portable): transparent huge
//      pages -> up to 512x fewer IOMMU / GPU-MMU translations per byte streamed
#include <errno.h>
#include <string.h>
#include <sys/mman.h>

#include <map>
#include <mutex>

#include "common.cuh"

using namespace cp360;

namespace {
struct Mapping { void* raw; uint64_t maplen; int mode; };
std::map<void*, Mapping> g_host_allocs;          // window base -> how to release it
std::mutex g_host_mutex;
constexpr uint64_t kHuge = 2ull << 20;
}  // namespace

extern "C" {

int cp360_host_alloc(uint64_t bytes, int mode, void** out_host) {
  CP360_CHECK_ARG(out_host != nullptr && bytes > 0, CP360_ERR_BAD_ARG, "null result pointer or zero size");
  CP360_CHECK_ARG(mode >= 0 && mode <= 2, CP360_ERR_BAD_ARG, "unknown host allocation mode %d", mode);
  *out_host = nullptr;
  int rc = require_device();
  if (rc != CP360_OK) return rc;
  Mapping m = {nullptr, 0, mode};
  void* base = nullptr;
  if (mode == 0 || mode == 1) {
    CP360_CUDA_OK(cudaHostAlloc(&base, bytes, cudaHostAllocPortable | (mode == 1 ? cudaHostAllocWriteCombined : 0)));
  } else {
    const uint64_t len = (bytes + kHuge - 1) / kHuge * kHuge;
    // one huge page of slack so that a 2 MB-aligned window of `len` bytes exists (the slack is never touched)
    void* raw = mmap(nullptr, len + kHuge, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    CP360_CHECK_ARG(raw != MAP_FAILED, CP360_ERR_BAD_ARG, "mmap of %llu bytes failed: %s",
                    (unsigned long long)(len + kHuge), strerror(errno));
    uint8_t* win = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + kHuge - 1) / kHuge * kHuge);
    madvise(win, len, MADV_HUGEPAGE);                   // advisory: 4 KB pages when THP is unavailable
    for (uint64_t off = 0; off < len; off += 4096) win[off] = 0;   // fault the pages in before pinning them
    cudaError_t e = cudaHostRegister(win, len, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
      munmap(raw, len + kHuge);
      cudaGetLastError();
      set_error("cudaHostRegister of %llu bytes failed: %s", (unsigned long long)len, cudaGetErrorString(e));
      return CP360_ERR_CUDA;
    }
    base = win;
    m.raw = raw;
    m.maplen = len + kHuge;
  }
  {
    std::lock_guard<std::mutex> lock(g_host_mutex);
    g_host_allocs[base] = m;
  }
  *out_host = base;
  return CP360_OK;
}

int cp360_host_free(void* host_ptr) {
  if (!host_ptr) return CP360_OK;
  Mapping m;
  {
    std::lock_guard<std::mutex> lock(g_host_mutex);
    auto it = g_host_allocs.find(host_ptr);
    CP360_CHECK_ARG(it != g_host_allocs.end(), CP360_ERR_BAD_ARG, "pointer was not returned by cp360_host_alloc");
    m = it->second;
    g_host_allocs.erase(it);
  }
  if (m.mode == 2) {
    cudaHostUnregister(host_ptr);
    cudaGetLastError();
    munmap(m.raw, m.maplen);
  } else {
    CP360_CUDA_OK(cudaFreeHost(host_ptr));
  }
  return CP360_OK;
}

}  // extern "C"
