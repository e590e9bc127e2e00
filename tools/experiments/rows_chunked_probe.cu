// experiment: the rows layout written ONE ROW PER WARP at a time, the sample axis of the row cut
// into 32 chunks, lane l owning chunk l (a sequential recurrence can start anywhere once the
// state at the chunk boundary is known). Per step every lane stores one whole 32-byte sector
// (4 samples) of its chunk per field, so a warp's store instruction covers 32 sectors 512 B
// apart inside one 16 KB row, and the row is complete after 16 steps. Pure stores, no arithmetic.
// Compared with: the pattern of ltp_sample_kernel (lane = row, one sector per row and store,
// rows 16 KB apart) and the best case for a row-at-a-time warp (1 KB contiguous per store).
// Rows pattern of configs[2]: 28672 rows x 2004 doubles x 4 fields = 1.84 GB.
//   nvcc -arch=sm_100a -O3 rows_chunked_probe.cu -o rows_chunked_probe && ./rows_chunked_probe
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ void st4(double* p, double x) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p), "d"(x) : "memory");
}

// LPR lanes per row: the warp works on 32 / LPR of its 32 rows at once, lane (l % LPR) owning
// chunk (l % LPR) of row (l / LPR) of the group; LPR = 1 is today's kernel (lane = row).
// CONTIG: warp = row, lanes interleaved by sector (1 KB contiguous per store instruction), the
// best case for a row-at-a-time warp (not reachable by a sequential recurrence).
// Dynamic shared memory only limits the number of resident warps per SM.
template <int LPR, int WARPS, bool CONTIG>
__global__ void __launch_bounds__(32 * WARPS) kprobe(double* base, long rows, long stride, int nsamp) {
  const int lane = threadIdx.x & 31;
  const long tile = (long)blockIdx.x * WARPS + (threadIdx.x >> 5);
  const long r0 = tile * 32;
  if (r0 >= rows) return;
  double x = (double)r0;
  if (CONTIG) {
    for (int row = 0; row < 32; ++row) {
      const long r = r0 + row;
      if (r >= rows) break;
      for (int s = lane * 4; s + 4 <= nsamp; s += 128) {
#pragma unroll
        for (int f = 0; f < 4; ++f) st4(base + ((long)f * rows + r) * stride + s, x);
        x += 1.0;
      }
    }
    return;
  }
  const int C = ((nsamp + LPR - 1) / LPR + 3) & ~3;  // chunk length, a multiple of 4
  constexpr int RPG = 32 / LPR;                      // rows per group
  const int chunk = lane % LPR, sub = lane / LPR;
  for (int g = 0; g < LPR; ++g) {
    const long r = r0 + g * RPG + sub;
    if (r >= rows) continue;
    const int s0 = chunk * C;
    for (int i = 0; i < C; i += 4) {
      const int s = s0 + i;
      if (s + 4 <= nsamp) {
#pragma unroll
        for (int f = 0; f < 4; ++f) st4(base + ((long)f * rows + r) * stride + s, x);
      }
      x += 1.0;
    }
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
#define RUN(NAME, ...)                                                                                      \
  {                                                                                                         \
    float best = 1e9;                                                                                       \
    double useful = (double)4 * rows * ns * 8;                                                              \
    for (int r = 0; r < 5; r++) {                                                                           \
      cudaEventRecord(e0);                                                                                  \
      __VA_ARGS__;                                                                                          \
      cudaEventRecord(e1);                                                                                  \
      cudaEventSynchronize(e1);                                                                             \
      float ms;                                                                                             \
      cudaEventElapsedTime(&ms, e0, e1);                                                                    \
      if (r && ms < best) best = ms;                                                                        \
    }                                                                                                       \
    printf("%-44s %.3f ms  %.0f GB/s  (%s)\n", NAME, best, useful / best / 1e6, cudaGetErrorString(cudaGetLastError())); \
  }
  const unsigned tiles = (unsigned)((rows + 31) / 32);
  cudaFuncSetAttribute(kprobe<32, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(kprobe<16, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(kprobe<8, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  RUN("lane = row (today), 1 warp/CTA", (kprobe<1, 1, false><<<tiles, 32>>>(d, rows, stride, ns)));
  RUN("32 lanes per row", (kprobe<32, 1, false><<<tiles, 32>>>(d, rows, stride, ns)));
  RUN("32 lanes per row, <= 5 warps/SM (44 KB)", (kprobe<32, 1, false><<<tiles, 32, 44 * 1024>>>(d, rows, stride, ns)));
  RUN("32 lanes per row, <= 3 warps/SM (70 KB)", (kprobe<32, 1, false><<<tiles, 32, 70 * 1024>>>(d, rows, stride, ns)));
  RUN("16 lanes per row", (kprobe<16, 1, false><<<tiles, 32>>>(d, rows, stride, ns)));
  RUN("16 lanes per row, <= 8 warps/SM (27 KB)", (kprobe<16, 1, false><<<tiles, 32, 27 * 1024>>>(d, rows, stride, ns)));
  RUN("8 lanes per row", (kprobe<8, 1, false><<<tiles, 32>>>(d, rows, stride, ns)));
  RUN("8 lanes per row, <= 11 warps/SM (20 KB)", (kprobe<8, 1, false><<<tiles, 32, 20 * 1024>>>(d, rows, stride, ns)));
  RUN("4 lanes per row", (kprobe<4, 1, false><<<tiles, 32>>>(d, rows, stride, ns)));
  RUN("2 lanes per row", (kprobe<2, 1, false><<<tiles, 32>>>(d, rows, stride, ns)));
  RUN("32 lanes per row, 4 warps/CTA", (kprobe<32, 4, false><<<(tiles + 3) / 4, 128>>>(d, rows, stride, ns)));
  RUN("warp = row, 1 KB contiguous per store", (kprobe<32, 1, true><<<tiles, 32>>>(d, rows, stride, ns)));
  return 0;
}
