// Shared plumbing of libcp360: status codes, thread-local error text, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cp360.h"

namespace cp360 {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

// Validates that a CUDA device is usable; sets the error text otherwise.
int require_device();

#define CP360_CHECK_ARG(cond, status, ...)       \
  do {                                            \
    if (!(cond)) {                                \
      ::cp360::set_error(__VA_ARGS__);            \
      return (status);                            \
    }                                             \
  } while (0)

#define CP360_CUDA_OK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::cp360::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                       \
      return CP360_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

// After a <<<>>> launch: count it and surface launch-configuration errors.
#define CP360_LAUNCHED()                                                        \
  do {                                                                           \
    ::cp360::count_launch();                                                     \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      ::cp360::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                    \
      return CP360_ERR_CUDA;                                                     \
    }                                                                            \
  } while (0)

int sm_count();   // cached multiProcessorCount of the current device (148 on B200)
bool pdl_enabled();   // CP360_PDL != 0 (default on)

// Work counters for kernels that deal their tiles dynamically: a pair of zeroed device words
// {next unit, finished CTAs} per launch; the last CTA of a launch zeroes its pair again. Eager
// launches cycle through a ring (far longer than any launch queue), launches recorded by a stream
// capture get pairs of their own that are never handed out again (a graph may be replayed at any
// time). Returns nullptr when no pair can be provided (pool not yet created and the stream is
// capturing, or capture pairs exhausted): the caller then falls back to its static partition.
uint32_t* acquire_work_counter(cudaStream_t st);

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-serialization
// attribute may become resident while its predecessor on the stream is still draining. Every
// forward kernel calls pdl_trigger() first (lets ITS successor be scheduled early) and
// pdl_wait() before its first global-memory access (blocks until the predecessor grid has
// completed and its writes are visible), so only launch latency and the smem prologue (tables,
// barriers) overlap the predecessor's tail — stream order is otherwise intact. Every thread of
// every CTA must pass pdl_wait() so that completion stays transitive along the stream.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- optional device-side timeline (build with -DCP360_TRACE; tools/trace_chain.py) ----------
// One record per CTA: %globaltimer at entry, after pdl_wait(), when the first tile has landed and
// at exit, plus the SM id — enough to see launch gaps, prologue, pipeline fill and tail spread of
// back-to-back kernels inside a CUDA graph. Compiled out of the product build.
#ifdef CP360_TRACE
struct TraceRec { unsigned long long t[4]; unsigned kid, cta, smid, nctas; };
struct TraceBuf { TraceRec* rec; unsigned cap; unsigned* n; };
static __device__ TraceBuf g_tb;
__device__ __forceinline__ unsigned long long trace_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CP360_TRACE_BEGIN(KID)                                                                   \
  __shared__ TraceRec* s_trace;                                                                  \
  if (threadIdx.x == 0 && threadIdx.y == 0) {                                                    \
    TraceRec* r_ = nullptr;                                                                      \
    if (g_tb.rec) {                                                                              \
      const unsigned i_ = atomicAdd(g_tb.n, 1u);                                                 \
      if (i_ < g_tb.cap) {                                                                       \
        r_ = g_tb.rec + i_;                                                                      \
        unsigned sm_;                                                                            \
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));                                         \
        r_->t[0] = trace_now(); r_->t[1] = r_->t[2] = r_->t[3] = 0;                              \
        r_->kid = (KID); r_->smid = sm_;                                                         \
        r_->cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);                \
        r_->nctas = gridDim.x * gridDim.y * gridDim.z;                                           \
      }                                                                                          \
    }                                                                                            \
    s_trace = r_;                                                                                \
  }
// thread (0,0) only, no block sync needed
#define CP360_TRACE_T0(I) do { if (threadIdx.x == 0 && threadIdx.y == 0 && s_trace) s_trace->t[I] = trace_now(); } while (0)
// any thread after a block-wide sync that followed CP360_TRACE_BEGIN: earliest (I=2) / latest (I=3) wins
#define CP360_TRACE_MIN(I) do { if (s_trace) atomicMin(&s_trace->t[I], trace_now()); } while (0)
#define CP360_TRACE_MAX(I) do { if (s_trace) atomicMax(&s_trace->t[I], trace_now()); } while (0)
#define CP360_TRACE_INIT_MIN(I) do { if (threadIdx.x == 0 && threadIdx.y == 0 && s_trace) s_trace->t[I] = ~0ull; } while (0)
#else
#define CP360_TRACE_BEGIN(KID)
#define CP360_TRACE_T0(I) do {} while (0)
#define CP360_TRACE_MIN(I) do {} while (0)
#define CP360_TRACE_MAX(I) do {} while (0)
#define CP360_TRACE_INIT_MIN(I) do {} while (0)
#endif

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                 cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Same, as thread-block clusters of `cluster_x` consecutive CTAs (grid.x must be a multiple of it): the CTAs of a
// cluster are co-scheduled on one GPC and can read each other's shared memory (DSMEM).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel_cluster(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         cudaStream_t st, unsigned cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_x;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}
#endif

}  // namespace cp360
