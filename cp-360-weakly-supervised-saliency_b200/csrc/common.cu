// libcp360 plumbing: status strings, thread-local last error, launch counter.
#include "common.cuh"

#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <string.h>

namespace cp360 {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int require_device() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("no usable CUDA device: %s (libcp360 has no CPU fallback)", cudaGetErrorString(e));
    return CP360_ERR_CUDA;
  }
  return CP360_OK;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* v = getenv("CP360_PDL");
    return !(v && *v == '0');
  }();
  return on;
}

namespace {
constexpr int kMaxDevices = 64;
constexpr uint32_t kEagerPairs = 8192, kCapturePairs = 57344;
struct CounterPool {
  uint32_t* base = nullptr;
  std::atomic<uint32_t> eager_seq{0}, capture_next{0};
};
CounterPool g_pools[kMaxDevices];
std::mutex g_pool_mutex;
}  // namespace

uint32_t* acquire_work_counter(cudaStream_t st) {
  static const bool enabled = [] {
    const char* v = getenv("CP360_DYNAMIC");
    return !(v && *v == '0');
  }();
  int dev = 0;
  if (!enabled || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  const bool capturing = cs != cudaStreamCaptureStatusNone;
  CounterPool& pool = g_pools[dev];
  if (!pool.base) {
    if (capturing) return nullptr;                     // no allocation inside a capture
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!pool.base) {
      uint32_t* p = nullptr;
      const size_t bytes = (size_t)(kEagerPairs + kCapturePairs) * 2 * sizeof(uint32_t);
      if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess) {
        cudaGetLastError();
        if (p) cudaFree(p);
        return nullptr;
      }
      pool.base = p;
    }
  }
  uint32_t idx;
  if (capturing) {
    const uint32_t k = pool.capture_next.fetch_add(1, std::memory_order_relaxed);
    if (k >= kCapturePairs) return nullptr;
    idx = kEagerPairs + k;
  } else {
    idx = pool.eager_seq.fetch_add(1, std::memory_order_relaxed) % kEagerPairs;
  }
  return pool.base + 2 * (size_t)idx;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached_sms = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached_dev = dev;
    cached_sms = n;
  }
  return cached_sms;
}

}  // namespace cp360

extern "C" {

int cp360_version(void) { return CP360_VERSION; }

const char* cp360_status_string(int s) {
  switch (s) {
    case CP360_OK: return "ok";
    case CP360_ERR_BAD_ARG: return "bad argument";
    case CP360_ERR_GROUP: return "CubePad size mismatch! (batch is not a multiple of 6)";
    case CP360_ERR_SHAPE: return "unsupported shape";
    case CP360_ERR_RANGE: return "value out of range";
    case CP360_ERR_CUDA: return "CUDA error";
    case CP360_ERR_ALIGN: return "misaligned pointer";
    default: return "unknown status";
  }
}

const char* cp360_last_error(void) { return cp360::g_err; }

uint64_t cp360_launch_count(void) { return cp360::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
