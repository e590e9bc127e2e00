// experiment: how many warps does it take to saturate HBM with streaming stores?
// (a) plain sequential fill with W warps total; (b) time-major pattern with time chunks.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void fill(double* dst, size_t n4, double v){  // 32-byte stores, grid-stride, warp-contiguous
  size_t i=(size_t)blockIdx.x*blockDim.x+threadIdx.x, st=(size_t)gridDim.x*blockDim.x;
  for(;i<n4;i+=st) asm volatile("st.global.cs.v4.f64 [%0], {%1,%1,%1,%1};"::"l"(dst+4*i),"d"(v):"memory");
}
__global__ void fill8(double* dst, size_t n, double v){  // 8-byte stores
  size_t i=(size_t)blockIdx.x*blockDim.x+threadIdx.x, st=(size_t)gridDim.x*blockDim.x;
  for(;i<n;i+=st) asm volatile("st.global.cs.f64 [%0], %1;"::"l"(dst+i),"d"(v):"memory");
}
// time-major with time chunks: block b handles rows tile (b / chunks) and samples [c*per, (c+1)*per)
__global__ void __launch_bounds__(32) ktm(double* base, long rows, long cap, int nsamp, int chunks){
  long r = (long)(blockIdx.x/chunks)*32 + threadIdx.x; if (r>=rows) return;
  int c = blockIdx.x%chunks, per=nsamp/chunks; double x=r;
  for (int s=c*per; s<(c+1)*per; s+=4){
    #pragma unroll
    for(int f=0;f<4;f++){ double* fb = base + (long)f*cap*rows;
      #pragma unroll
      for(int u=0;u<4;u++) asm volatile("st.global.cs.f64 [%0], %1;"::"l"(fb + (long)(s+u)*rows + r),"d"(x):"memory"); }
    x+=1.0; }
}
int main(){
  size_t bytes=(size_t)2<<30; double* d; cudaMalloc(&d,bytes);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  #define RUN(NAME, USEFUL, ...) { float best=1e9; for(int r=0;r<4;r++){ cudaEventRecord(e0); __VA_ARGS__; cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(r&&ms<best)best=ms;} printf("%-34s %.3f ms  %.0f GB/s (%s)\n",NAME,best,(USEFUL)/best/1e6,cudaGetErrorString(cudaGetLastError())); }
  int cfg[][2]={{148,32},{296,32},{592,32},{1024,32},{2048,32},{4096,32},{148*8,256},{148*16,256}};
  for(auto&c:cfg){ char nm[64]; snprintf(nm,64,"fill32B %dx%d (2GiB)",c[0],c[1]); RUN(nm,(double)bytes,(fill<<<c[0],c[1]>>>(d,bytes/32,1.0))); }
  for(auto&c:cfg){ char nm[64]; snprintf(nm,64,"fill8B  %dx%d (2GiB)",c[0],c[1]); RUN(nm,(double)bytes,(fill8<<<c[0],c[1]>>>(d,bytes/8,1.0))); }
  long rows=28672, cap=2000; double useful=(double)4*rows*cap*8;
  for(int ch: {1,2,4,8,16}){ char nm[64]; snprintf(nm,64,"time-major chunks=%d (%ld warps)",ch,(rows/32)*ch); RUN(nm,useful,(ktm<<<(rows/32)*ch,32>>>(d,rows,cap,2000,ch))); }
  // same total bytes as the sampler but sequential fill with 1024 warps
  RUN("fill32B 1024x32 (1.84GB)",useful,(fill<<<1024,32>>>(d,(size_t)useful/32,1.0)));
  RUN("fill32B 2368x256 (1.84GB)",useful,(fill<<<2368,256>>>(d,(size_t)useful/32,1.0)));
  return 0; }
