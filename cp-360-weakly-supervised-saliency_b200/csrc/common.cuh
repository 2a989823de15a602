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

}  // namespace cp360
