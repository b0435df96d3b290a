// Shared host/device helpers: error handling, thread-local error string, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/rbm_b200.h"

namespace rbm {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define RBM_CUDA_TRY(expr)                                         \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return ::rbm::cuda_fail(_e, #expr);     \
  } while (0)

#define RBM_REQUIRE(cond, code, ...)                               \
  do {                                                             \
    if (!(cond)) { ::rbm::set_error(__VA_ARGS__); return (code); } \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return cuda_fail(cudaGetLastError(), what);
  return 0;
}

// Cached per-device attributes (cudaDeviceGetAttribute costs ~1 us per call on the launch path).
struct DeviceInfo { int device; int sm_count; int cc_major; int l2_bytes; };
int get_device_info(DeviceInfo* out);

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device, size high-water mark).
// Keyed by the kernel's ADDRESS: distinct instantiations share one function-pointer type.
int ensure_dynamic_smem_impl(const void* kern, int device, int bytes);
template <class Kern>
inline int ensure_dynamic_smem(Kern kern, int device, int bytes) {
  return ensure_dynamic_smem_impl(reinterpret_cast<const void*>(kern), device, bytes);
}

__host__ __device__ constexpr int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

}  // namespace rbm
