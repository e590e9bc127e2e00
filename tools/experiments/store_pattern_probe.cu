// experiment: ceiling of the sampler's store pattern (one thread per row, 32B sectors, row stride 16 KB)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
template<int HINT> __device__ __forceinline__ void st256(double* p, double v){
  if (HINT==0) asm volatile("st.global.v4.f64 [%0], {%1,%1,%1,%1};"::"l"(p),"d"(v):"memory");
  if (HINT==1) asm volatile("st.global.cs.v4.f64 [%0], {%1,%1,%1,%1};"::"l"(p),"d"(v):"memory");
  if (HINT==2) asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%1,%1,%1};"::"l"(p),"d"(v):"memory");
  if (HINT==3) asm volatile("st.global.cg.v4.f64 [%0], {%1,%1,%1,%1};"::"l"(p),"d"(v):"memory");
  if (HINT==4) asm volatile("st.global.wt.v4.f64 [%0], {%1,%1,%1,%1};"::"l"(p),"d"(v):"memory");
}
// rows x 4 fields; thread t handles row t for all 4 fields; burst = sectors written back-to-back per field
template<int HINT,int BURST> __global__ void __launch_bounds__(32) k(double* base, long rows, long stride, int nsamp, int chunks=1){
  long r = (long)(blockIdx.x/chunks)*28 + threadIdx.x; if (threadIdx.x>=28 || r>=rows) return;
  int c = blockIdx.x % chunks; int per = nsamp/chunks/ (4*BURST) * (4*BURST);
  double* f[4]; for(int i=0;i<4;i++) f[i]=base + ((long)i*rows + r)*stride + c*per;
  nsamp = per;
  double x=r;
  for (int s=0; s+4*BURST<=nsamp; s+=4*BURST){
    #pragma unroll
    for(int i=0;i<4;i++){
      #pragma unroll
      for(int b=0;b<BURST;b++) st256<HINT>(f[i]+s+4*b, x);
    }
    x+=1.0;
  }
}
// smem-transposed style: warp writes 128B-contiguous pieces: lane group of 4 lanes covers one row-line (4 x 32B)
template<int HINT> __global__ void __launch_bounds__(32) kline(double* base, long rows, long stride, int nsamp){
  // 28 rows per warp; each iteration covers 16 samples (128B) per row per field: 28 rows*4 fields = 112 lines = 112*4 sectors / 32 lanes = 14 stores per lane
  long r0 = (long)blockIdx.x*28; int lane=threadIdx.x; double x=r0;
  for (int s=0; s+16<=nsamp; s+=16){
    for (int it=0; it<14; ++it){
      int item = it*32+lane;            // 0..447 : (field, row, sector)
      int sec = item & 3; int row = (item>>2) % 28; int fld = (item>>2)/28;
      long r = r0+row; if (r<rows) st256<HINT>(base + ((long)fld*rows + r)*stride + s + 4*sec, x);
    }
    x+=1.0;
  }
}
// time-major layout [sample][rows]: lane r of a warp writes element (s, r); 28 consecutive doubles per instruction
template<int HINT> __global__ void __launch_bounds__(32) ktm(double* base, long rows, long cap, int nsamp){
  long r = (long)blockIdx.x*28 + threadIdx.x; if (threadIdx.x>=28 || r>=rows) return;
  double x=r;
  for (int s=0; s<nsamp; s+=4){
    #pragma unroll
    for(int f=0;f<4;f++){
      double* fb = base + (long)f*cap*rows;
      #pragma unroll
      for(int u=0;u<4;u++){
        double* p = fb + (long)(s+u)*rows + r;
        if (HINT==1) asm volatile("st.global.cs.f64 [%0], %1;"::"l"(p),"d"(x):"memory");
        else asm volatile("st.global.f64 [%0], %1;"::"l"(p),"d"(x):"memory");
      }
    }
    x+=1.0;
  }
}
int main(){
  long rows=28672, stride=2004; int ns=2000; size_t bytes=(size_t)4*rows*stride*8; double* d; cudaMalloc(&d,bytes);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double useful=(double)4*rows*ns*8;
  #define RUN(NAME, ...) { float best=1e9; for(int r=0;r<4;r++){ cudaEventRecord(e0); __VA_ARGS__; cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); if(r&&ms<best)best=ms;} printf("%-28s %.3f ms  %.0f GB/s  (%s)\n",NAME,best,useful/best/1e6,cudaGetErrorString(cudaGetLastError())); }
  RUN("default burst1", (k<0,1><<<1024,32>>>(d,rows,stride,ns)));
  RUN("cs burst1", (k<1,1><<<1024,32>>>(d,rows,stride,ns)));
  RUN("no_alloc burst1", (k<2,1><<<1024,32>>>(d,rows,stride,ns)));
  RUN("cg burst1", (k<3,1><<<1024,32>>>(d,rows,stride,ns)));
  RUN("wt burst1", (k<4,1><<<1024,32>>>(d,rows,stride,ns)));
  RUN("default burst2", (k<0,2><<<1024,32>>>(d,rows,stride,ns)));
  RUN("cs burst2", (k<1,2><<<1024,32>>>(d,rows,stride,ns)));
  RUN("default burst4", (k<0,4><<<1024,32>>>(d,rows,stride,ns)));
  RUN("cs burst4", (k<1,4><<<1024,32>>>(d,rows,stride,ns)));
  RUN("cs burst1 chunks2", (k<1,1><<<1024*2,32>>>(d,rows,stride,ns,2)));
  RUN("cs burst1 chunks4", (k<1,1><<<1024*4,32>>>(d,rows,stride,ns,4)));
  RUN("cs burst1 chunks8", (k<1,1><<<1024*8,32>>>(d,rows,stride,ns,8)));
  RUN("cs burst1 chunks16", (k<1,1><<<1024*16,32>>>(d,rows,stride,ns,16)));
  RUN("default burst1 chunks8", (k<0,1><<<1024*8,32>>>(d,rows,stride,ns,8)));
  RUN("cs burst4 chunks5", (k<1,4><<<1024*5,32>>>(d,rows,stride,ns,5)));
  RUN("time-major default", (ktm<0><<<1024,32>>>(d,rows,stride,ns)));
  RUN("time-major cs", (ktm<1><<<1024,32>>>(d,rows,stride,ns)));
  RUN("default line-coalesced", (kline<0><<<1024,32>>>(d,rows,stride,ns)));
  RUN("cs line-coalesced", (kline<1><<<1024,32>>>(d,rows,stride,ns)));
  return 0; }
