// Bench-only probes (not part of the planning path): live measurements of the two
// rooflines the solver and the sampler are reported against.
//   ltp_probe_fp64_tflops   dependent DFMA chains, 8-way ILP per thread, all SMs
//   ltp_probe_hbm_write_gbs  streaming 16-byte stores over a buffer larger than L2
// Both time with CUDA events on the stream they launch on and return the best of `reps`.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__global__ void __launch_bounds__(256) fp64_fma_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void __launch_bounds__(256) write_kernel(double* dst, size_t n2, double v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n2; i += stride)
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(dst + 2 * i), "d"(v), "d"(v) : "memory");
}

}  // namespace

extern "C" {

int ltp_probe_fp64_tflops(int device, int reps, double* tflops) {
  if (cudaSetDevice(device) != cudaSuccess) return -2;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  double* out = nullptr;
  if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return -2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < reps + 1; ++r) {
    cudaEventRecord(e0);
    fp64_fma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * iters * (double)blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (r > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

int ltp_probe_hbm_write_gbs(int device, int reps, double* gbs) {
  if (cudaSetDevice(device) != cudaSuccess) return -2;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2;
  const size_t bytes = (size_t)2 << 30;  // 2 GiB, far larger than the 126 MB L2
  double* buf = nullptr;
  if (cudaMalloc(&buf, bytes) != cudaSuccess) return -2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int r = 0; r < reps + 1; ++r) {
    cudaEventRecord(e0);
    write_kernel<<<prop.multiProcessorCount * 16, 256>>>(buf, bytes / 16, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double g = bytes / (ms * 1e-3) / 1e9;
    if (r > 0 && g > best) best = g;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(buf);
  *gbs = best;
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
