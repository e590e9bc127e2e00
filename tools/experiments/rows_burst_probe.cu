// experiment: does the rows layout (one row per (problem, joint), 16 KB apart per field) reach HBM
// bandwidth when every visit to a row writes a longer burst? Pure stores, no arithmetic: a warp
// owns 32 rows and walks the sample axis in tiles; per tile, field and row it writes BURST bytes
// contiguously (32 lanes x 8 B = 256 B per instruction, or 16 lanes x 8 B for the 128 B case).
// Rows pattern of configs[2]: 28672 rows x 2004 doubles x 4 fields = 1.84 GB.
//   nvcc -arch=sm_100a -O3 rows_burst_probe.cu -o rows_burst_probe && ./rows_burst_probe
#include <cuda_runtime.h>
#include <cstdio>
template <int BURST, int CS>
__global__ void __launch_bounds__(32) kburst(double* base, long rows, long stride, int nsamp) {
  const long r0 = (long)blockIdx.x * 32;
  const int lane = threadIdx.x;
  constexpr int K = BURST / 8;  // samples per burst
  double x = (double)r0;
  for (int s = 0; s + K <= nsamp; s += K) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      for (int row = 0; row < 32; ++row) {
        const long r = r0 + row;
        if (r >= rows) break;
        double* p = base + ((long)f * rows + r) * stride + s;
        if (K >= 32) {
#pragma unroll
          for (int u = 0; u < K; u += 32) {
            if (CS) asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p + u + lane), "d"(x) : "memory");
            else p[u + lane] = x;
          }
        } else if (lane < K) {
          if (CS) asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p + lane), "d"(x) : "memory");
          else p[lane] = x;
        }
      }
    }
    x += 1.0;
  }
}
int main() {
  long rows = 28672, stride = 2004;
  int ns = 2000;
  size_t bytes = (size_t)4 * rows * stride * 8;
  double* d;
  cudaMalloc(&d, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
#define RUN(NAME, K, ...)                                                                                   \
  {                                                                                                         \
    float best = 1e9;                                                                                       \
    double useful = (double)4 * rows * (ns / (K) * (K)) * 8;                                                \
    for (int r = 0; r < 4; r++) {                                                                           \
      cudaEventRecord(e0);                                                                                  \
      __VA_ARGS__;                                                                                          \
      cudaEventRecord(e1);                                                                                  \
      cudaEventSynchronize(e1);                                                                             \
      float ms;                                                                                             \
      cudaEventElapsedTime(&ms, e0, e1);                                                                    \
      if (r && ms < best) best = ms;                                                                        \
    }                                                                                                       \
    printf("%-28s %.3f ms  %.0f GB/s  (%s)\n", NAME, best, useful / best / 1e6, cudaGetErrorString(cudaGetLastError())); \
  }
  const unsigned g = (unsigned)((rows + 31) / 32);
  RUN("burst 128 B", 16, (kburst<128, 0><<<g, 32>>>(d, rows, stride, ns)));
  RUN("burst 256 B", 32, (kburst<256, 0><<<g, 32>>>(d, rows, stride, ns)));
  RUN("burst 512 B", 64, (kburst<512, 0><<<g, 32>>>(d, rows, stride, ns)));
  RUN("burst 1024 B", 128, (kburst<1024, 0><<<g, 32>>>(d, rows, stride, ns)));
  RUN("burst 2048 B", 256, (kburst<2048, 0><<<g, 32>>>(d, rows, stride, ns)));
  RUN("burst 4000 B", 500, (kburst<4000, 0><<<g, 32>>>(d, rows, stride, ns)));
  RUN("burst 128 B cs", 16, (kburst<128, 1><<<g, 32>>>(d, rows, stride, ns)));
  RUN("burst 256 B cs", 32, (kburst<256, 1><<<g, 32>>>(d, rows, stride, ns)));
  RUN("burst 512 B cs", 64, (kburst<512, 1><<<g, 32>>>(d, rows, stride, ns)));
  RUN("burst 1024 B cs", 128, (kburst<1024, 1><<<g, 32>>>(d, rows, stride, ns)));
  RUN("burst 2048 B cs", 256, (kburst<2048, 1><<<g, 32>>>(d, rows, stride, ns)));
  return 0;
}
