// Shared helpers for the fvgp_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define FVGP_ERR_CUDA (-1)
#define FVGP_ERR_ARG (-2)

#define FVGP_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      fprintf(stderr, "[fvgp_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), \
              __FILE__, __LINE__, cudaGetErrorString(_e));                              \
      return FVGP_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

#define FVGP_LAUNCH_OK() FVGP_CUDA_OK(cudaGetLastError())

#define FVGP_REQUIRE(cond)                                                           \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      fprintf(stderr, "[fvgp_b200] bad argument: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
      return FVGP_ERR_ARG;                                                           \
    }                                                                                \
  } while (0)

#include <utility>

namespace fvgp {

// Every kernel of the library is launched through this helper so that the number of launches
// (bench.py's "gpu_launches") is counted, not estimated.
extern unsigned long long g_launches;
template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  kernel<<<grid, block, smem, st>>>(std::forward<Args>(args)...);
  __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED);  // launches come from more than one host thread
}

constexpr int kMaxDim = 8;  // input-space dimensionality supported by the fused kernels

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic CTA-wide sum; result valid in thread 0. `red` must hold >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  if (wid == 0) {
    v = lane < nw ? red[lane] : 0.0;
    v = warp_sum(v);
  }
  return v;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace fvgp
