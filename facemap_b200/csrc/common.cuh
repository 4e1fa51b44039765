// Shared helpers for the facemap_b200 CUDA sources (error plumbing, launch counting).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/facemap_b200.h"

namespace fm {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define FM_CUDA_TRY(expr)                                                                    \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      fm::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FM_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

#define FM_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      fm::set_error(__VA_ARGS__);      \
      return FM_ERR_INVALID;           \
    }                                  \
  } while (0)

#define FM_LAUNCH_CHECK()                                                                    \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      fm::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FM_ERR_CUDA;                                                                    \
    }                                                                                        \
    fm::count_launch();                                                                      \
  } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace fm
