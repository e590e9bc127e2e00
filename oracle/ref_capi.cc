// TEST INFRASTRUCTURE ONLY (oracle/): C entry points around the UNMODIFIED reference
// sources, compiled where they lie:
//   /root/reference/src/long_term_planner.cc
//   /root/reference/include/long_term_planner/{long_term_planner.h,roots.h}
// against oracle/eigen_shim (Eigen 3.4 is absent from this container; see
// oracle/eigen_shim/Eigen/Eigenvalues). Output: oracle/_ref/libltp_ref.so (git-ignored).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library. It is the checker, never the product.
//
// The protected methods are reached exactly the way the reference's own test fixture
// does it (tests/include/long_term_planner_fixture.h:34-57): a subclass with
// using-declarations.
#include "long_term_planner/long_term_planner.h"

#include <cstdint>
#include <cstring>
#include <iostream>
#include <thread>
#include <vector>

namespace {

class Exposed : public long_term_planner::LongTermPlanner {
 public:
  using LongTermPlanner::getTrajectory;
  using LongTermPlanner::optBraking;
  using LongTermPlanner::optSwitchTimes;
  using LongTermPlanner::timeScaling;
  Exposed(int dof, double ts, std::vector<double> q_min, std::vector<double> q_max,
          std::vector<double> v_max, std::vector<double> a_max, std::vector<double> j_max)
      : LongTermPlanner(dof, ts, q_min, q_max, v_max, a_max, j_max),
        dof(dof), ts(ts), q_min(q_min), q_max(q_max), v_max(v_max), a_max(a_max), j_max(j_max) {}
  // private in the reference; mirrored here for the solve-only leg
  int dof;
  double ts;
  std::vector<double> q_min, q_max, v_max, a_max, j_max;
};

typedef std::array<double, 7> T7;

// Stages 1-3 of planTrajectory (reference long_term_planner.cc:14-55) re-expressed
// through the exposed protected methods, i.e. the same calls minus cc:58-61.
// Returns 1 when the reference would reach getTrajectory, 0 on its early `return false`.
int solve_one(Exposed* L, const double* q_goal, const double* q_0, const double* v_0,
              const double* a_0, double* t_opt, double* t_scaled, double* dir, double* v_drive,
              unsigned char* mod, int* slowest, int* ts_ok) {
  const int dof = L->dof;
  std::vector<double> q0(q_0, q_0 + dof), v0(v_0, v_0 + dof), a0(a_0, a_0 + dof);
  *slowest = -1;
  for (int i = 0; i < dof; ++i) { v_drive[i] = L->v_max[i]; mod[i] = 0; dir[i] = 0; ts_ok[i] = -1; }
  std::memset(t_opt, 0, sizeof(double) * 7 * dof);
  std::memset(t_scaled, 0, sizeof(double) * 7 * dof);
  if (!L->checkInputs(q0, v0, a0)) return 0;                                   // cc:14-15
  std::vector<T7> topt(dof), tsc(dof);
  std::vector<char> m(dof);
  int ok = 1;
  for (int i = 0; i < dof && ok; ++i) {                                        // cc:27-30
    bool s = L->optSwitchTimes(i, q_goal[i], q_0[i], v_0[i], a_0[i], L->v_max[i], topt[i], dir[i], m[i]);
    if (!s) ok = 0;
  }
  for (int i = 0; i < dof; ++i) { std::memcpy(t_opt + 7 * i, topt[i].data(), 56); mod[i] = (unsigned char)m[i]; }
  if (!ok) return 0;
  double t_required = -1;                                                      // cc:31-39
  for (int i = 0; i < dof; ++i)
    if (topt[i][6] > t_required) { t_required = topt[i][6]; *slowest = i; }
  if (*slowest == -1) return 0;
  for (int i = 0; i < dof; ++i) {                                              // cc:42-48
    if (i == *slowest) continue;
    bool s = L->timeScaling(i, q_goal[i], q_0[i], v_0[i], a_0[i], dir[i], t_required, tsc[i], v_drive[i], m[i]);
    ts_ok[i] = s ? 1 : 0;
  }
  for (int i = 0; i < dof; ++i) {                                              // cc:50-55
    if (*std::max_element(tsc[i].begin(), tsc[i].end()) <= 0.0) tsc[i] = topt[i];
    std::memcpy(t_scaled + 7 * i, tsc[i].data(), 56);
    mod[i] = (unsigned char)m[i];
  }
  return 1;
}

template <class F>
void parallel_for(int64_t n, int threads, F f) {
  if (threads <= 1 || n < 2) { f(0, n, 0); return; }
  std::vector<std::thread> th;
  for (int t = 0; t < threads; ++t) {
    int64_t lo = n * t / threads, hi = n * (t + 1) / threads;
    th.emplace_back([=] { f(lo, hi, t); });
  }
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

void* ref_create(int dof, double t_sample, const double* q_min, const double* q_max,
                 const double* v_max, const double* a_max, const double* j_max) {
  auto v = [dof](const double* p) { return std::vector<double>(p, p + dof); };
  // the reference prints one diagnostic line per failed safety check (cc:343); mute it,
  // batched runs would otherwise emit millions of lines
  std::cerr.setstate(std::ios_base::failbit);
  return new Exposed(dof, t_sample, v(q_min), v(q_max), v(v_max), v(a_max), v(j_max));
}
void ref_destroy(void* h) { delete static_cast<Exposed*>(h); }

// ---- per-joint primitives, array-of-items; every item uses limit set `joint[i]` ----
void ref_opt_braking(void* h, int64_t n, const int* joint, const double* v_0, const double* a_0,
                     double* q, double* t_rel3, double* dir) {
  Exposed* L = static_cast<Exposed*>(h);
  for (int64_t i = 0; i < n; ++i) {
    T7 t = {};
    L->optBraking(joint ? joint[i] : 0, v_0[i], a_0[i], q[i], t, dir[i]);
    t_rel3[3 * i] = t[0]; t_rel3[3 * i + 1] = t[1]; t_rel3[3 * i + 2] = t[2];
  }
}

// t is pre-zeroed per item: the reference leaves it unwritten on the cc:340-344 path.
void ref_opt_switch_times(void* h, int64_t n, const int* joint, const double* q_goal,
                          const double* q_0, const double* v_0, const double* a_0,
                          const double* v_drive, double* t7, double* dir, unsigned char* mod,
                          unsigned char* ok, int threads) {
  Exposed* L = static_cast<Exposed*>(h);
  parallel_for(n, threads, [=](int64_t lo, int64_t hi, int) {
    for (int64_t i = lo; i < hi; ++i) {
      T7 t = {};
      char m = 0;
      bool s = L->optSwitchTimes(joint ? joint[i] : 0, q_goal[i], q_0[i], v_0[i], a_0[i], v_drive[i], t, dir[i], m);
      std::memcpy(t7 + 7 * i, t.data(), 56);
      mod[i] = (unsigned char)m;
      ok[i] = s;
    }
  });
}

void ref_time_scaling(void* h, int64_t n, const int* joint, const double* q_goal,
                      const double* q_0, const double* v_0, const double* a_0, const double* dir,
                      const double* t_required, double* t7, double* v_drive, unsigned char* mod,
                      unsigned char* ok, int threads) {
  Exposed* L = static_cast<Exposed*>(h);
  parallel_for(n, threads, [=](int64_t lo, int64_t hi, int) {
    for (int64_t i = lo; i < hi; ++i) {
      T7 t = {};
      char m = 0;
      bool s = L->timeScaling(joint ? joint[i] : 0, q_goal[i], q_0[i], v_0[i], a_0[i], dir[i], t_required[i], t, v_drive[i], m);
      std::memcpy(t7 + 7 * i, t.data(), 56);
      mod[i] = (unsigned char)m;
      ok[i] = s;
    }
  });
}

// ---- problem-major batched legs: x[p*dof + joint] -------------------------------------
// stages 1-3 only. reached[p] = 1 if the reference would call getTrajectory.
void ref_solve_batch(void* h, int64_t n, const double* q_goal, const double* q_0,
                     const double* v_0, const double* a_0, double* t_opt, double* t_scaled,
                     double* dir, double* v_drive, unsigned char* mod, int* slowest, int* ts_ok,
                     unsigned char* reached, int threads) {
  Exposed* L = static_cast<Exposed*>(h);
  const int dof = L->dof;
  parallel_for(n, threads, [=](int64_t lo, int64_t hi, int) {
    for (int64_t p = lo; p < hi; ++p) {
      const int64_t o = p * dof;
      reached[p] = (unsigned char)solve_one(L, q_goal + o, q_0 + o, v_0 + o, a_0 + o, t_opt + 7 * o,
                                            t_scaled + 7 * o, dir + o, v_drive + o, mod + o,
                                            slowest + p, ts_ok + o);
    }
  });
}

// getTrajectory on given switching times (reference cc:706-841). Output rows are
// [joint][sample] with `stride` doubles per row; returns traj.length, or -(needed) if
// stride is too small.
int ref_get_trajectory(void* h, const double* t7, const double* dir, const unsigned char* mod,
                       const double* q_0, const double* v_0, const double* a_0,
                       const double* v_drive, int64_t stride, double* q, double* v, double* a, double* j) {
  Exposed* L = static_cast<Exposed*>(h);
  const int dof = L->dof;
  std::vector<T7> t(dof);
  for (int i = 0; i < dof; ++i) std::memcpy(t[i].data(), t7 + 7 * i, 56);
  std::vector<double> d(dir, dir + dof), q0(q_0, q_0 + dof), v0(v_0, v_0 + dof), a0(a_0, a_0 + dof), vd(v_drive, v_drive + dof);
  std::vector<char> m(mod, mod + dof);
  long_term_planner::Trajectory tr = L->getTrajectory(t, d, m, q0, v0, a0, vd);
  if (tr.length > stride) return -tr.length;
  for (int i = 0; i < dof; ++i) {
    std::memcpy(q + i * stride, tr.q[i].data(), sizeof(double) * tr.length);
    std::memcpy(v + i * stride, tr.v[i].data(), sizeof(double) * tr.length);
    std::memcpy(a + i * stride, tr.a[i].data(), sizeof(double) * tr.length);
    std::memcpy(j + i * stride, tr.j[i].data(), sizeof(double) * tr.length);
  }
  return tr.length;
}

// full planTrajectory (reference cc:7-63), one problem. length_out = -1 if traj untouched.
int ref_plan(void* h, const double* q_goal, const double* q_0, const double* v_0, const double* a_0,
             int64_t stride, double* q, double* v, double* a, double* j, int* length_out) {
  Exposed* L = static_cast<Exposed*>(h);
  const int dof = L->dof;
  std::vector<double> g(q_goal, q_goal + dof), q0(q_0, q_0 + dof), v0(v_0, v_0 + dof), a0(a_0, a_0 + dof);
  long_term_planner::Trajectory tr;
  tr.length = -1;
  bool ok = L->planTrajectory(g, q0, v0, a0, tr);
  *length_out = tr.length;
  if (tr.length > 0 && q && tr.length <= stride) {
    for (int i = 0; i < dof; ++i) {
      std::memcpy(q + i * stride, tr.q[i].data(), sizeof(double) * tr.length);
      std::memcpy(v + i * stride, tr.v[i].data(), sizeof(double) * tr.length);
      std::memcpy(a + i * stride, tr.a[i].data(), sizeof(double) * tr.length);
      std::memcpy(j + i * stride, tr.j[i].data(), sizeof(double) * tr.length);
    }
  }
  return ok ? 1 : 0;
}

// Timing leg: full planTrajectory over a batch, nothing kept except a checksum of the
// final positions, the success flags and trajectory lengths. Problem-major inputs.
double ref_plan_batch(void* h, int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                      const double* a_0, unsigned char* success, int* length, int threads) {
  Exposed* L = static_cast<Exposed*>(h);
  const int dof = L->dof;
  std::vector<double> sums((size_t)(threads > 1 ? threads : 1), 0.0);
  parallel_for(n, threads, [&, dof](int64_t lo, int64_t hi, int tid) {
    double acc = 0;
    for (int64_t p = lo; p < hi; ++p) {
      const int64_t o = p * dof;
      std::vector<double> g(q_goal + o, q_goal + o + dof), q0(q_0 + o, q_0 + o + dof), v0(v_0 + o, v_0 + o + dof), a0(a_0 + o, a_0 + o + dof);
      long_term_planner::Trajectory tr;
      tr.length = -1;
      bool ok = L->planTrajectory(g, q0, v0, a0, tr);
      success[p] = ok;
      length[p] = tr.length;
      if (tr.length > 0) for (int i = 0; i < dof; ++i) acc += tr.q[i][tr.length - 1];
    }
    sums[tid] = acc;
  });
  double s = 0;
  for (double x : sums) s += x;
  return s;
}

// roots<T>() + getSmallestPositiveNonComplexRoot<T>() (reference roots.h:22-50).
// coeffs highest power first, deg+1 of them. re/im receive deg values in Eigen's order.
double ref_roots_f64(const double* coeffs, int deg, double* re, double* im) {
  Eigen::VectorXd p(deg + 1);
  for (int i = 0; i <= deg; ++i) p[i] = coeffs[i];
  auto r = long_term_planner::roots<double>(p);
  for (int i = 0; i < deg; ++i) { re[i] = r(i, 0).real(); im[i] = r(i, 0).imag(); }
  return long_term_planner::getSmallestPositiveNonComplexRoot<double>(r);
}
float ref_roots_f32(const float* coeffs, int deg, float* re, float* im) {
  Eigen::VectorXf p(deg + 1);
  for (int i = 0; i <= deg; ++i) p[i] = coeffs[i];
  auto r = long_term_planner::roots<float>(p);
  for (int i = 0; i < deg; ++i) { re[i] = r(i, 0).real(); im[i] = r(i, 0).imag(); }
  return long_term_planner::getSmallestPositiveNonComplexRoot<float>(r);
}

const char* ref_build_info(void) {
  return "reference long_term_planner.cc (unmodified) + oracle/eigen_shim EigenSolver; "
         "g++ -std=c++17 -O2 -ffp-contract=off";
}

}  // extern "C"
