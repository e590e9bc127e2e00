// C++ host-side test of the batched entry points of the drop-in class (the reference's host
// language is C++): planTrajectories over device buffers must give, for every problem, the
// bits that the single-plan planTrajectory (reference signature, long_term_planner.h:144-150)
// gives; planStream must visit every chunk once and its device-side totals must add up;
// advance must hand back the sample at the tick. Exit code 0 = all checks passed.
// Build: tests/build_cpp_tests.sh (g++, links liblong_term_planner + libltp_b200 + cudart).
#include <cuda_runtime_api.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "long_term_planner/long_term_planner.h"

using long_term_planner::BatchPlan;
using long_term_planner::LongTermPlanner;
using long_term_planner::Trajectory;

static int failures = 0;
#define CHECK(cond, ...)                 \
  do {                                   \
    if (!(cond)) {                       \
      ++failures;                        \
      std::printf("FAILED %s:%d: ", __FILE__, __LINE__); \
      std::printf(__VA_ARGS__);          \
      std::printf("\n");                 \
    }                                    \
  } while (0)

template <typename T>
static T* dev_alloc(size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) { std::printf("cudaMalloc failed\n"); std::exit(2); }
  return static_cast<T*>(p);
}

struct Seen { int64_t chunks = 0, problems = 0, next_first = 0; bool ordered = true; };
static int consumer(void* user, const ltp_chunk* c, void*) {
  Seen* s = static_cast<Seen*>(user);
  s->ordered &= (c->first == s->next_first);
  s->next_first = c->first + c->count;
  s->chunks++;
  s->problems += c->count;
  return 0;
}

int main() {
  const int dof = 7;
  const double ts = 0.001;
  const std::vector<double> q_min{-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973};
  const std::vector<double> q_max{2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973};
  const std::vector<double> v_max{2.175, 2.175, 2.175, 2.175, 2.61, 2.61, 2.61};
  const std::vector<double> a_max{15, 7.5, 10, 12.5, 15, 20, 20};
  const std::vector<double> j_max{7500, 3750, 5000, 6250, 7500, 10000, 10000};
  LongTermPlanner ltp(dof, ts, q_min, q_max, v_max, a_max, j_max);

  const int64_t n = 300;
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> u(0.0, 1.0);
  // joint-major host arrays x[jt * n + p]
  std::vector<double> qg(dof * n), q0(dof * n), v0(dof * n), a0(dof * n);
  for (int64_t p = 0; p < n; ++p)
    for (int jt = 0; jt < dof; ++jt) {
      const int64_t at = jt * n + p;
      q0[at] = q_min[jt] + u(rng) * (q_max[jt] - q_min[jt]);
      qg[at] = q_min[jt] + 0.05 + u(rng) * (q_max[jt] - q_min[jt] - 0.1);
      v0[at] = (2 * u(rng) - 1) * 0.5 * v_max[jt];
      a0[at] = (2 * u(rng) - 1) * 0.3 * a_max[jt];
    }
  double *d_qg = dev_alloc<double>(dof * n), *d_q0 = dev_alloc<double>(dof * n), *d_v0 = dev_alloc<double>(dof * n),
         *d_a0 = dev_alloc<double>(dof * n);
  cudaMemcpy(d_qg, qg.data(), dof * n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d_q0, q0.data(), dof * n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d_v0, v0.data(), dof * n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(d_a0, a0.data(), dof * n * 8, cudaMemcpyHostToDevice);

  const int64_t cap = 4096;
  BatchPlan plan;
  std::memset(&plan.solution, 0, sizeof plan.solution);
  plan.solution.t_scaled = dev_alloc<double>(8 * dof * n);  // 64-byte records: t[0..6], v_drive
  plan.solution.dir = dev_alloc<double>(dof * n);
  plan.solution.v_drive = dev_alloc<double>(dof * n);
  plan.solution.mod = dev_alloc<uint8_t>(dof * n);
  plan.solution.slowest = dev_alloc<int32_t>(n);
  plan.solution.traj_len = dev_alloc<int32_t>(n);
  plan.solution.reached = dev_alloc<uint8_t>(n);
  plan.q = dev_alloc<double>(cap * n * dof);
  plan.v = dev_alloc<double>(cap * n * dof);
  plan.a = dev_alloc<double>(cap * n * dof);
  plan.j = dev_alloc<double>(cap * n * dof);
  plan.success = dev_alloc<uint8_t>(n + 4);
  plan.horizon = 0;
  plan.layout = LTP_LAYOUT_TIME_MAJOR;
  plan.stride = cap;
  int rc = ltp.planTrajectories(n, d_qg, d_q0, d_v0, d_a0, plan, nullptr);
  CHECK(rc == LTP_OK, "planTrajectories rc=%d", rc);
  cudaDeviceSynchronize();
  std::vector<int32_t> len(n);
  std::vector<uint8_t> succ(n);
  cudaMemcpy(len.data(), plan.solution.traj_len, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(succ.data(), plan.success, n, cudaMemcpyDeviceToHost);
  std::vector<double> hq(cap * n * dof), hj(cap * n * dof);
  cudaMemcpy(hq.data(), plan.q, hq.size() * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(hj.data(), plan.j, hj.size() * 8, cudaMemcpyDeviceToHost);

  // every 13th problem through the reference-signature single-plan call
  int compared = 0;
  for (int64_t p = 0; p < n; p += 13) {
    std::vector<double> g(dof), s0(dof), sv(dof), sa(dof);
    for (int jt = 0; jt < dof; ++jt) { g[jt] = qg[jt * n + p]; s0[jt] = q0[jt * n + p]; sv[jt] = v0[jt * n + p]; sa[jt] = a0[jt * n + p]; }
    Trajectory traj;
    const bool ok = ltp.planTrajectory(g, s0, sv, sa, traj);
    CHECK(ok == (succ[p] != 0), "problem %ld: success %d vs batched %d", (long)p, (int)ok, (int)succ[p]);
    CHECK(traj.length == len[p], "problem %ld: length %d vs batched %d", (long)p, traj.length, len[p]);
    for (int jt = 0; jt < dof && traj.length == len[p]; ++jt)
      for (int i = 0; i < traj.length; ++i) {
        const size_t at = ((size_t)i * n + p) * dof + jt;
        if (traj.q[jt][i] != hq[at] || traj.j[jt][i] != hj[at]) {
          CHECK(false, "problem %ld joint %d sample %d differs", (long)p, jt, i);
          i = traj.length;
        }
      }
    ++compared;
  }
  CHECK(compared > 20, "compared %d", compared);

  // streamed run: 300 problems in chunks of 128 -> 3 chunks, in order
  Seen seen;
  ltp_stream_stats st;
  rc = ltp.planStream(n, d_qg, d_q0, d_v0, d_a0, 128, 0, cap, consumer, &seen, &st);
  CHECK(rc == LTP_OK, "planStream rc=%d", rc);
  CHECK(seen.chunks == 3 && seen.problems == n && seen.ordered, "chunks %ld problems %ld", (long)seen.chunks, (long)seen.problems);
  int64_t samples = 0, ok_count = 0;
  for (int64_t p = 0; p < n; ++p) { samples += (int64_t)len[p] * dof; ok_count += succ[p]; }
  CHECK(st.problems == n && st.chunks == 3, "stats problems %ld chunks %ld", (long)st.problems, (long)st.chunks);
  CHECK(st.samples == samples && st.bytes == samples * 32, "stats samples %ld vs %ld", (long)st.samples, (long)samples);
  CHECK(st.success == ok_count, "stats success %ld vs %ld", (long)st.success, (long)ok_count);

  // receding horizon: the state at sample 9 of the batched plan
  double *n_q = dev_alloc<double>(dof * n), *n_v = dev_alloc<double>(dof * n), *n_a = dev_alloc<double>(dof * n);
  rc = ltp.advance(n, 9, plan.stride, plan.solution.traj_len, plan.solution.reached, plan.q, plan.v, plan.a, n_q, n_v, n_a, nullptr);
  CHECK(rc == LTP_OK, "advance rc=%d", rc);
  cudaDeviceSynchronize();
  std::vector<double> hn(dof * n);
  cudaMemcpy(hn.data(), n_q, dof * n * 8, cudaMemcpyDeviceToHost);
  for (int64_t p = 0; p < n; p += 7)
    for (int jt = 0; jt < dof; ++jt) {
      const int i = len[p] > 9 ? 9 : len[p] - 1;
      const double want = hq[((size_t)i * n + p) * dof + jt];
      const double lo = q_min[jt], hi = q_max[jt];
      const double clamped = want < lo ? lo : (want > hi ? hi : want);
      CHECK(hn[jt * n + p] == clamped, "advance problem %ld joint %d", (long)p, jt);
    }

  std::printf("%d checks failed\n", failures);
  return failures == 0 ? 0 : 1;
}
