// libcp360 plumbing: status strings, thread-local last error, launch counter.
#include "common.cuh"

#include <atomic>
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
