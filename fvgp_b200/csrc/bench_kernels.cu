// Micro-benchmarks that establish the FP64 roofline denominators on the box
// (MEASURED_PEAKS.json carries only HBM and BF16 figures): register-resident DMMA.8x8x4
// issue rate and DFMA issue rate.  Not part of the product path.
#include "../../include/fvgp_b200.h"
#include "common.cuh"

namespace fvgp {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(acc[i][0]), "+d"(acc[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = i;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace fvgp

using namespace fvgp;

extern "C" {

// which: 0 = DMMA, 1 = DFMA.  ctas_per_sm x 256 threads per SM.  Returns TFLOP/s in *h_tflops.
int fvgp_bench_fp64_peak(int which, int ctas_per_sm, int iters, double* d_scratch, double* h_tflops, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = sm_count() * ctas_per_sm;
  cudaEvent_t e0, e1;
  FVGP_CUDA_OK(cudaEventCreate(&e0));
  FVGP_CUDA_OK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; ++rep) {  // last repetition is the timed one
    FVGP_CUDA_OK(cudaEventRecord(e0, st));
    if (which == 0) launch(dmma_peak_kernel, grid, 256, 0, st, d_scratch, iters);
    else launch(dfma_peak_kernel, grid, 256, 0, st, d_scratch, iters);
    FVGP_CUDA_OK(cudaEventRecord(e1, st));
    FVGP_CUDA_OK(cudaEventSynchronize(e1));
  }
  FVGP_LAUNCH_OK();
  float ms = 0.f;
  FVGP_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  const double per_thread = which == 0 ? 16.0 * iters * (512.0 / 32.0) : 16.0 * iters * 2.0;
  *h_tflops = per_thread * 256.0 * grid / (ms * 1e-3) / 1e12;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return 0;
}

}  // extern "C"
