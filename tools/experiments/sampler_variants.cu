// experiment: what separates the time-major sampler (5.2 TB/s) from the pure-store kernel of
// the same pattern (6.0 TB/s)? Synthetic rows, real per-sample arithmetic, several schedules.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I longtermplanner_b200/csrc
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ltp_math.cuh"
using namespace ltp;

namespace legacy {  // the 4-word table / fma cursor of the first sampler
__device__ __forceinline__ double seg_pack(int next, int cruise) {
  const long long b = (long long)(((unsigned long long)(unsigned)cruise << 32) | (unsigned)next);
  return __longlong_as_double(b);
}
template <int STRIDE> struct SegTableT {
  double* base;
  __device__ void build(const RowSampler& R, int limit) const {
    int cur = 0;
#pragma unroll 1
    for (int m = 0; m < kMaxSeg; ++m) {
      const int at = cur < limit ? cur : 0;
      const double j = R.jerk_at(at);
      const bool az = R.a_zero(at);
      const int vc = R.v_cruise(at) ? 1 : 0;
      if (cur < limit) { cur = R.next_break(cur); if (cur >= limit) cur = 0x7fffffff; }
      double* e = base + (4 * m) * STRIDE;
      e[0] = az ? 0.0 : R.Ts * j; e[STRIDE] = j; e[2 * STRIDE] = seg_pack(cur, vc); e[3 * STRIDE] = az ? 0.0 : 1.0;
    }
  }
};
template <int STRIDE> struct SegCursorT {
  double Ts, vcruise, a, v, q; int m, next;
  __device__ void begin(const RowSampler& R) { Ts = R.Ts; vcruise = R.vcruise; a = R.a; v = R.v; q = R.q; m = 0; next = 0x7fffffff; }
  __device__ __forceinline__ void step(const SegTableT<STRIDE>& T, int i, double& jo, double& ao, double& vo, double& qo) {
    m += (i == next) ? 1 : 0;
    const double* e = T.base + (4 * m) * STRIDE;
    const double tsj = e[0]; const double jv = e[STRIDE];
    const double pk = e[2 * STRIDE]; next = __double2loint(pk); const bool vc = __double2hiint(pk) != 0;
    const double keep = e[3 * STRIDE];
    a = fma(a, keep, tsj);
    const double ta = Ts * a;
    const double vz = fma(v, keep, ta);
    v = vc ? vcruise : vz;
    q = q + Ts * v;
    jo = jv; ao = a; vo = v; qo = q;
  }
};
}  // namespace legacy

__device__ __forceinline__ void store1(double* dst, double x) { asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(dst), "d"(x) : "memory"); }

struct Rows { const double* t; const double* dir; const double* q0; const double* v0; const double* a0; double ts, J, vdrive; long rows; };

// MODE 0: legacy cursor; 1: new cursor (2-word, select). STORE: emit stores. W warps per CTA share a tile (time chunks)
template <int MODE, bool STORE>
__global__ void __launch_bounds__(256) ksample(Rows X, int horizon, double* q, double* v, double* a, double* j, int UNROLLDUMMY) {
  __shared__ __align__(16) double s_tab[kMaxSeg * 4 * 32];
  const int lane = threadIdx.x, w = threadIdx.y, W = blockDim.y;
  const long r = (long)blockIdx.x * 32 + lane;
  if (r >= X.rows) return;
  double t[7];
  for (int k = 0; k < 7; ++k) t[k] = X.t[k * X.rows + r];
  RowSampler R; R.init(X.ts, X.J, t, X.dir[r], 0, X.q0[r], X.v0[r], X.a0[r], X.vdrive, horizon);
  const int b0 = (int)((long)horizon * w / W), b1 = (int)((long)horizon * (w + 1) / W);
  const unsigned step = (unsigned)X.rows * 8u;
  unsigned off = (unsigned)r * 8u + (unsigned)b0 * step;
  char* qb = (char*)q; char* vb = (char*)v; char* ab = (char*)a; char* jb = (char*)j;
  double jj = 0, aa = 0, vv = 0, qq = 0;
  if (MODE == 0) {
    const legacy::SegTableT<32> T{&s_tab[lane]};
    if (w == 0) T.build(R, horizon);
    __syncthreads();
    legacy::SegCursorT<32> C; C.begin(R);
    int i = 0;
#pragma unroll 4
    for (; i < b0; ++i) C.step(T, i, jj, aa, vv, qq);
#pragma unroll 4
    for (; i < b1; ++i) {
      C.step(T, i, jj, aa, vv, qq);
      if (STORE) { store1((double*)(qb + off), qq); store1((double*)(vb + off), vv); store1((double*)(ab + off), aa); store1((double*)(jb + off), jj); }
      off += step;
    }
    if (!STORE && qq == 123.456) q[r] = qq + aa + vv + jj;
  } else {
    const SegTableT<64> T{&s_tab[lane * 2]};
    if (w == 0) T.build(R, horizon);
    __syncthreads();
    SegCursorT<64> C; C.begin(R);
    int i = 0;
#pragma unroll 4
    for (; i < b0; ++i) C.step(T, i, jj, aa, vv, qq);
#pragma unroll 4
    for (; i < b1; ++i) {
      C.step(T, i, jj, aa, vv, qq);
      if (STORE) { store1((double*)(qb + off), qq); store1((double*)(vb + off), vv); store1((double*)(ab + off), aa); store1((double*)(jb + off), jj); }
      off += step;
    }
    if (!STORE && qq == 123.456) q[r] = qq + aa + vv + jj;
  }
}

// register-cached piece: the table is read only when the piece changes (divergent but rare branch)
template <bool STORE>
__global__ void __launch_bounds__(256) kcached(Rows X, int horizon, double* q, double* v, double* a, double* j) {
  __shared__ __align__(16) double s_tab[kMaxSeg * 2 * 32];
  const int lane = threadIdx.x, w = threadIdx.y, W = blockDim.y;
  const long r = (long)blockIdx.x * 32 + lane;
  if (r >= X.rows) return;
  double t[7];
  for (int k = 0; k < 7; ++k) t[k] = X.t[k * X.rows + r];
  RowSampler R; R.init(X.ts, X.J, t, X.dir[r], 0, X.q0[r], X.v0[r], X.a0[r], X.vdrive, horizon);
  const SegTableT<64> T{&s_tab[lane * 2]};
  if (w == 0) T.build(R, horizon);
  __syncthreads();
  const int b0 = (int)((long)horizon * w / W), b1 = (int)((long)horizon * (w + 1) / W);
  const unsigned step = (unsigned)X.rows * 8u;
  unsigned off = (unsigned)r * 8u + (unsigned)b0 * step;
  char* qb = (char*)q; char* vb = (char*)v; char* ab = (char*)a; char* jb = (char*)j;
  const double Ts = X.ts, vcruise = R.vcruise;
  double ca = R.a, cv = R.v, cq = R.q;
  int m = -1, next = 0; double jv = 0, tsj = 0; bool vc = false, live = true;
  for (int i = 0; i < b1; ++i) {
    if (i == next) {
      ++m; const double2 e = *reinterpret_cast<const double2*>(T.base + m * 64);
      jv = e.x; seg_unpack(e.y, next, vc, live); tsj = Ts * jv;
    }
    const double a1 = ca + tsj; ca = live ? a1 : 0.0;
    const double v1 = cv + Ts * ca; cv = vc ? vcruise : (live ? v1 : 0.0);
    cq = cq + Ts * cv;
    if (i >= b0) {
      if (STORE) { store1((double*)(qb + off), cq); store1((double*)(vb + off), cv); store1((double*)(ab + off), ca); store1((double*)(jb + off), jv); }
      off += step;
    }
  }
  if (!STORE && cq == 123.456) q[r] = cq + ca + cv + jv;
}

// pure stores, same pattern
__global__ void __launch_bounds__(256) kstore(long rows, int horizon, double* q, double* v, double* a, double* j) {
  const int lane = threadIdx.x, w = threadIdx.y, W = blockDim.y;
  const long r = (long)blockIdx.x * 32 + lane;
  if (r >= rows) return;
  const int b0 = (int)((long)horizon * w / W), b1 = (int)((long)horizon * (w + 1) / W);
  const unsigned step = (unsigned)rows * 8u;
  unsigned off = (unsigned)r * 8u + (unsigned)b0 * step;
  char* qb = (char*)q; char* vb = (char*)v; char* ab = (char*)a; char* jb = (char*)j;
  double x = r;
#pragma unroll 4
  for (int i = b0; i < b1; ++i) {
    store1((double*)(qb + off), x); store1((double*)(vb + off), x); store1((double*)(ab + off), x); store1((double*)(jb + off), x);
    off += step; x += 1.0;
  }
}

int main() {
  const long rows = 28672; const int H = 2001; const double ts = 0.001;
  std::vector<double> t(7 * rows), dir(rows), q0(rows), v0(rows), a0(rows);
  srand(1);
  for (long r = 0; r < rows; ++r) {
    double acc = 0;
    for (int k = 0; k < 7; ++k) { double d = 0.002 + (rand() / (double)RAND_MAX) * (k == 3 ? 1.2 : 0.1); acc += d; t[k * rows + r] = acc; }
    dir[r] = (rand() & 1) ? 1.0 : -1.0; q0[r] = 0.1; v0[r] = 0.01; a0[r] = 0.02;
  }
  double *dt, *dd, *dq0, *dv0, *da0, *out;
  cudaMalloc(&dt, t.size() * 8); cudaMalloc(&dd, rows * 8); cudaMalloc(&dq0, rows * 8); cudaMalloc(&dv0, rows * 8); cudaMalloc(&da0, rows * 8);
  cudaMemcpy(dt, t.data(), t.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dd, dir.data(), rows * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dq0, q0.data(), rows * 8, cudaMemcpyHostToDevice); cudaMemcpy(dv0, v0.data(), rows * 8, cudaMemcpyHostToDevice); cudaMemcpy(da0, a0.data(), rows * 8, cudaMemcpyHostToDevice);
  const size_t field = (size_t)rows * H * 8;
  cudaMalloc(&out, 4 * field);
  double *q = out, *v = out + field / 8, *a = out + 2 * field / 8, *j = out + 3 * field / 8;
  Rows X{dt, dd, dq0, dv0, da0, ts, 7500.0, 1.5, rows};
  const unsigned grid = (unsigned)(rows / 32);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const double useful = 4.0 * field;
#define RUN(NAME, ...) { float best = 1e9, sum = 0; for (int r_ = 0; r_ < 8; r_++) { cudaEventRecord(e0); __VA_ARGS__; cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (r_ > 1) { sum += ms; if (ms < best) best = ms; } } printf("%-40s best %.4f ms (%.0f GB/s) mean %.4f ms (%.0f GB/s) %s\n", NAME, best, useful / best / 1e6, sum / 6, useful / (sum / 6) / 1e6, cudaGetErrorString(cudaGetLastError())); }
  for (int W : {1, 2, 4, 8}) {
    char nm[64]; dim3 blk(32, W);
    snprintf(nm, 64, "pure store          W=%d", W); RUN(nm, (kstore<<<grid, blk>>>(rows, H, q, v, a, j)));
    snprintf(nm, 64, "legacy fma/4-word   W=%d", W); RUN(nm, (ksample<0, true><<<grid, blk>>>(X, H, q, v, a, j, 0)));
    snprintf(nm, 64, "legacy, no stores   W=%d", W); RUN(nm, (ksample<0, false><<<grid, blk>>>(X, H, q, v, a, j, 0)));
    snprintf(nm, 64, "select/2-word       W=%d", W); RUN(nm, (ksample<1, true><<<grid, blk>>>(X, H, q, v, a, j, 0)));
    snprintf(nm, 64, "select, no stores   W=%d", W); RUN(nm, (ksample<1, false><<<grid, blk>>>(X, H, q, v, a, j, 0)));
    snprintf(nm, 64, "reg-cached piece    W=%d", W); RUN(nm, (kcached<true><<<grid, blk>>>(X, H, q, v, a, j)));
    snprintf(nm, 64, "reg-cached no store W=%d", W); RUN(nm, (kcached<false><<<grid, blk>>>(X, H, q, v, a, j)));
  }
  return 0;
}
