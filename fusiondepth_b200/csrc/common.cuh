// Shared helpers for the fusiondepth_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fd {

void set_error(const char* fmt, ...);
extern long g_launches;   // kernels launched through the library since load

#define FD_CHECK_LAUNCH()                                                        \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    ++fd::g_launches;                                                            \
    if (e__ != cudaSuccess) {                                                    \
      fd::set_error("%s:%d launch failed: %s", __FILE__, __LINE__,              \
                    cudaGetErrorString(e__));                                    \
      return 1;                                                                  \
    }                                                                            \
  } while (0)

#define FD_REQUIRE(cond, ...)                                                    \
  do {                                                                           \
    if (!(cond)) {                                                               \
      fd::set_error(__VA_ARGS__);                                                \
      return 2;                                                                  \
    }                                                                            \
  } while (0)

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

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

// Block-wide sum of NV values per thread; result valid in thread 0 (returned in v[]).
// smem must hold NV * 32 floats.
template <int NV, typename T>
__device__ __forceinline__ void block_sum(T (&v)[NV], T* smem) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * 32 + wid] = v[i];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      T x = lane < nw ? smem[i * 32 + lane] : T(0);
      v[i] = warp_sum(x);
    }
  }
}

}  // namespace fd
