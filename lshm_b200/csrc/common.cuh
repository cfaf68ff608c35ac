// Shared helpers for liblshm_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include "../../include/lshm.h"

namespace lshm {

void set_error(const char* fmt, ...);

#define LSHM_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::lshm::set_error(__VA_ARGS__);           \
      return LSHM_ERR_ARG;                      \
    }                                           \
  } while (0)

// Call after a launch: report launch-configuration errors without synchronising.
#define LSHM_CHECK_LAUNCH(name)                                                   \
  do {                                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess) {                                                     \
      ::lshm::set_error("%s: CUDA error %s", name, cudaGetErrorString(e__));      \
      return LSHM_ERR_CUDA;                                                       \
    }                                                                             \
  } while (0)

#define LSHM_CUDA(call, name)                                                     \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess) {                                                     \
      ::lshm::set_error("%s: CUDA error %s", name, cudaGetErrorString(e__));      \
      return LSHM_ERR_CUDA;                                                       \
    }                                                                             \
  } while (0)

int sm_count();

static inline cudaStream_t as_stream(lshm_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float elu_f(float z) { return z > 0.f ? z : expm1f(z); }
// ELU for the conv epilogues: exp(z) - 1 with the hardware exponential (abs. error ~1e-7, i.e. far
// below the bf16x3 conv error of ~3e-6 relative).  ~5 instructions instead of ~45 for expm1f: the
// epilogue warps of the tensor-core kernels are a serial chain per tile and were the bottleneck.
__device__ __forceinline__ float elu_fast(float z) { return z > 0.f ? z : __expf(z) - 1.f; }
// derivative of ELU expressed through the *output* a = ELU(z): 1 if a>0 else a+1
__device__ __forceinline__ float delu_from_out(float a) { return a > 0.f ? 1.f : a + 1.f; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; result valid in thread 0.  `red` must hold >= 32 elements.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? red[threadIdx.x] : T(0);
  if (wid == 0) v = warp_sum(v);
  return v;
}

__device__ __forceinline__ float4 ld_nc_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace lshm
