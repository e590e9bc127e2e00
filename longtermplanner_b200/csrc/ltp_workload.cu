// Bench / test tooling, NOT part of the planning path (nothing in libltp_b200.so depends on
// it): device-side twins of longtermplanner_b200/workloads.py and a trajectory read-back
// reducer, so that workloads too large for the host (BASELINE.json configs[4]: 2^26 problems
// x 12 joints, ~60 TB of samples) can be generated, planned and verified chunk by chunk
// without the data ever leaving the GPU.
//
//   ltp_wl_random_states  counter-based splitmix64 start/goal states, bit-identical to
//                         workloads.random_states (recipe of the reference's
//                         tests/randomConfiguration.m:14-34)
//   ltp_wl_row_stats      per (problem, joint) row of a sampled trajectory: sequential sums
//                         of q, v, a, j, max |v|, max |a|, last q, last v -- a checksum of
//                         checksums that the CPU oracle can reproduce bit for bit, plus the
//                         quantities of the domain properties (limits respected, goal reached)
//
// Compiled with -fmad=false: the generator must round like numpy.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace {

struct WlLimits {
  double q_min[32], q_max[32], v_max[32], a_max[32], j_max[32];
};

__device__ __forceinline__ double uniform01(uint64_t counter, uint64_t seed) {
  uint64_t z = seed + (counter + 1ull) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * 1.1102230246251565e-16;  // 2^-53
}

__global__ void __launch_bounds__(256)
random_states_kernel(const __grid_constant__ WlLimits L, int dof, int64_t n, int64_t start, uint64_t seed,
                     double margin, double* __restrict__ q_goal, double* __restrict__ q_0,
                     double* __restrict__ v_0, double* __restrict__ a_0) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int jt = blockIdx.y;
  if (i >= n) return;
  const double e = 1e-6;
  const uint64_t base = ((uint64_t)(start + i) * (uint64_t)dof + (uint64_t)jt) * 4ull;
  const double u0 = uniform01(base + 0, seed), u1 = uniform01(base + 1, seed);
  const double u2 = uniform01(base + 2, seed), u3 = uniform01(base + 3, seed);
  const double q_min = L.q_min[jt], q_max = L.q_max[jt], v_max = L.v_max[jt], a_max = L.a_max[jt],
               j_max = L.j_max[jt];
  const double q0 = q_min + u0 * (q_max - q_min);
  const double qg = (q_min + margin) + u1 * ((q_max - margin) - (q_min + margin));
  const double v0 = -(v_max - e) + u2 * (2.0 * (v_max - e));
  const bool pos = v0 >= 0;
  const double root = sqrt(2.0 * j_max * (v_max - fabs(v0)));
  const double a_lb = pos ? -(a_max - e) : fmax(-(a_max - e), -root);
  const double a_ub = pos ? fmin(a_max - e, root) : a_max;
  double a0 = a_lb + u3 * (a_ub - a_lb);
  a0 = fmin(fmax(a0, -a_max), a_max);
  const int64_t at = (int64_t)jt * n + i;
  q_goal[at] = qg;
  q_0[at] = q0;
  v_0[at] = v0;
  a_0[at] = a0;
}

// one thread per (problem, joint) row; out[row * 8 + k]
__global__ void __launch_bounds__(128)
row_stats_kernel(int layout, int64_t n, int dof, int64_t stride, int horizon, const int32_t* __restrict__ traj_len,
                 const double* __restrict__ q, const double* __restrict__ v, const double* __restrict__ a,
                 const double* __restrict__ j, double* __restrict__ out) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n * dof) return;
  const int64_t p = r / dof;
  int len = horizon > 0 ? horizon : traj_len[p];
  if (len > stride) len = (int)stride;
  const int64_t first = layout == 1 ? r : r * stride;
  const int64_t step = layout == 1 ? n * dof : 1;
  double sq = 0, sv = 0, sa = 0, sj = 0, mv = 0, ma = 0, ql = 0, vl = 0;
  for (int i = 0; i < len; ++i) {
    const int64_t at = first + (int64_t)i * step;
    const double qq = q[at], vv = v[at], aa = a[at], jj = j[at];
    sq = sq + qq; sv = sv + vv; sa = sa + aa; sj = sj + jj;
    mv = fmax(mv, fabs(vv));
    ma = fmax(ma, fabs(aa));
    ql = qq;
    vl = vv;
  }
  double* o = out + r * 8;
  o[0] = sq; o[1] = sv; o[2] = sa; o[3] = sj; o[4] = mv; o[5] = ma; o[6] = ql; o[7] = vl;
}

}  // namespace

extern "C" {

// limit vectors: host, dof doubles each. outputs: device, joint-major [dof][n].
int ltp_wl_random_states(int dof, const double* q_min, const double* q_max, const double* v_max,
                         const double* a_max, const double* j_max, int64_t n, int64_t start, uint64_t seed,
                         double margin, double* q_goal, double* q_0, double* v_0, double* a_0, void* stream) {
  if (dof < 1 || dof > 32 || n < 0) return -1;
  if (n == 0) return 0;
  WlLimits L;
  for (int i = 0; i < dof; ++i) {
    L.q_min[i] = q_min[i]; L.q_max[i] = q_max[i]; L.v_max[i] = v_max[i]; L.a_max[i] = a_max[i];
    L.j_max[i] = j_max[i];
  }
  dim3 grid((unsigned)((n + 255) / 256), dof);
  random_states_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(L, dof, n, start, seed, margin, q_goal, q_0, v_0, a_0);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// layout 0 rows q[(p*dof+jt)*stride + i], 1 time-major q[(i*n+p)*dof+jt]. out: [n*dof][8] device.
int ltp_wl_row_stats(int layout, int64_t n, int dof, int64_t stride, int horizon, const int32_t* traj_len,
                     const double* q, const double* v, const double* a, const double* j, double* out,
                     void* stream) {
  if (n < 0 || dof < 1) return -1;
  if (n == 0) return 0;
  const int64_t rows = n * dof;
  row_stats_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(layout, n, dof, stride, horizon,
                                                                                     traj_len, q, v, a, j, out);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

}  // extern "C"
