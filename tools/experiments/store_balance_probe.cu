// experiment: is the time-major store pattern limited by how CTAs spread over the SMs?
//  (a) histogram of CTAs per SM for one-warp CTAs at several grid sizes
//  (b) time-major pure-store kernel at row-tile counts that are / are not multiples of 148
//  (c) the same stores from a persistent grid (one CTA per SM, tiles assigned statically)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <algorithm>
__global__ void __launch_bounds__(32) ktm(double* base, long rows, long cap, int nsamp, int* smid_count, int spin){
  if (smid_count && threadIdx.x == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); atomicAdd(&smid_count[s], 1); }
  long r = (long)blockIdx.x*32 + threadIdx.x; if (r>=rows) return;
  double x=r;
  for (int s=0; s<nsamp; s+=4){
    #pragma unroll
    for(int u=0;u<4;u++){
      #pragma unroll
      for(int f=0;f<4;f++){ double* fb = base + (long)f*cap*rows;
        asm volatile("st.global.cs.f64 [%0], %1;"::"l"(fb + (long)(s+u)*rows + r),"d"(x):"memory"); }
      x+=1.0;
    }
  }
}
// persistent: CTA b (one per SM, forced by dynamic smem) has W warps; warp w takes tiles w*gridDim.x + b, + W*gridDim.x ...
__global__ void kpersist(double* base, long rows, long cap, int nsamp, long tiles){
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
  for (long t = (long)w*gridDim.x + blockIdx.x; t < tiles; t += (long)W*gridDim.x) {
    long r = t*32 + lane; if (r>=rows) continue;
    double x=r;
    for (int s=0; s<nsamp; s+=4){
      #pragma unroll
      for(int u=0;u<4;u++){
        #pragma unroll
        for(int f=0;f<4;f++){ double* fb = base + (long)f*cap*rows;
          asm volatile("st.global.cs.f64 [%0], %1;"::"l"(fb + (long)(s+u)*rows + r),"d"(x):"memory"); }
        x+=1.0;
      }
    }
  }
}
// persistent, equal share of the (tile, sample) rectangle per SM and per warp (what a chunked sampler would store)
__global__ void kshare(double* base, long rows, long cap, int nsamp, long tiles){
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
  const long total = tiles * nsamp, slots = (long)gridDim.x * W, slot = (long)blockIdx.x * W + w;
  long g0 = total * slot / slots, g1 = total * (slot + 1) / slots;
  while (g0 < g1) {
    long t = g0 / nsamp; int s0 = (int)(g0 - t*nsamp); long e_ = s0 + (g1-g0); int s1 = (int)(e_ < nsamp ? e_ : nsamp);
    long r = t*32 + lane; double x = r;
    if (r < rows) for (int s=s0; s<s1; ++s){
      #pragma unroll
      for(int f=0;f<4;f++){ double* fb = base + (long)f*cap*rows;
        asm volatile("st.global.cs.f64 [%0], %1;"::"l"(fb + (long)s*rows + r),"d"(x):"memory"); }
      x+=1.0;
    }
    g0 += s1 - s0;
  }
}
int main(){
  size_t bytes=(size_t)3<<30; double* d; cudaMalloc(&d,bytes);
  int* cnt; cudaMalloc(&cnt, 256*4); int h[256];
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  #define RUN(NAME, USEFUL, ...) { float best=1e9,sum=0; for(int r=0;r<6;r++){ cudaEventRecord(e0); __VA_ARGS__; cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(r){sum+=ms; if(ms<best)best=ms;}} printf("%-44s best %.3f ms %.0f GB/s | mean %.0f GB/s (%s)\n",NAME,best,(USEFUL)/best/1e6,(USEFUL)/(sum/5)/1e6,cudaGetErrorString(cudaGetLastError())); }
  int cap=2000;
  for (int tiles : {592, 740, 888, 896, 1036, 1184, 1480, 1776}) {
    long rows = (long)tiles*32; double useful=(double)4*rows*cap*8;
    cudaMemset(cnt,0,1024); ktm<<<tiles,32>>>(d,rows,cap,cap,cnt,0); cudaMemcpy(h,cnt,1024,cudaMemcpyDeviceToHost);
    int mn=1<<30,mx=0,used=0; for(int i=0;i<148;i++){ mn=std::min(mn,h[i]); mx=std::max(mx,h[i]); used+=h[i]>0; }
    char nm[96]; snprintf(nm,96,"ktm tiles=%d (CTAs/SM min %d max %d)",tiles,mn,mx);
    RUN(nm,useful,(ktm<<<tiles,32>>>(d,rows,cap,cap,nullptr,0)));
  }
  cudaFuncSetAttribute(kpersist, cudaFuncAttributeMaxDynamicSharedMemorySize, 120*1024);
  cudaFuncSetAttribute(kshare, cudaFuncAttributeMaxDynamicSharedMemorySize, 120*1024);
  for (int tiles : {888, 896}) for (int W : {6, 7, 8}) {
    long rows = (long)tiles*32; double useful=(double)4*rows*cap*8;
    char nm[96]; snprintf(nm,96,"persistent tiles=%d 148 CTAs x %d warps",tiles,W);
    RUN(nm,useful,(kpersist<<<148,32*W,120*1024>>>(d,rows,cap,cap,tiles)));
  }
  for (int W : {4, 6, 8, 12, 16, 24, 32}) {
    int tiles=896; long rows = (long)tiles*32; double useful=(double)4*rows*cap*8;
    char nm[96]; snprintf(nm,96,"equal-share tiles=%d 148 CTAs x %d warps",tiles,W);
    RUN(nm,useful,(kshare<<<148,32*W,120*1024>>>(d,rows,cap,cap,tiles)));
  }
  return 0; }
