// long_term_planner.h -- drop-in host class for the B200-native planner.
//
// Same namespace, type names, constructor, setters, planTrajectory and protected per-joint
// methods (name, argument order, meaning, bool-only error reporting) as the reference class
// yannickBurkhardt/LongTermPlanner include/long_term_planner/long_term_planner.h:37-45
// (Trajectory), :54-56 (sign), :103-131 (constructors), :144-205 (public methods),
// :223-307 (protected methods) -- so that code written against the reference, including its
// own test fixture (tests/include/long_term_planner_fixture.h:34-57, which re-exports the
// protected methods with using-declarations), compiles unchanged. Every method forwards to
// the C ABI of include/ltp_b200.h, i.e. to the CUDA kernels; there is no CPU implementation
// behind this class. planTrajectories() is the new batched entry point over
// structure-of-arrays device buffers.
//
// Not part of the reference's interface: the Eigen dependency (gone), device(), the batched
// types. Link with liblong_term_planner.so + libltp_b200.so (longtermplanner_b200/lib/).
#ifndef long_term_planner_H
#define long_term_planner_H

// <stdlib.h>/<math.h> (not only the <c...> forms) on purpose: they put the floating-point
// overloads of abs() into the global namespace, which code written against the reference
// relies on (the reference got them through Eigen's headers)
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <complex>   // the reference header drags these in through roots.h; kept so that
#include <cstdint>   // sources relying on the transitive includes still compile
#include <functional>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <numeric>
#include <tuple>
#include <vector>

#include "../ltp_b200.h"

namespace long_term_planner {

/// Trajectory structure (reference long_term_planner.h:37-45).
struct Trajectory {
  int dof;
  double t_sample;
  int length;
  std::vector<std::vector<double>> q;
  std::vector<std::vector<double>> v;
  std::vector<std::vector<double>> a;
  std::vector<std::vector<double>> j;
};

/// -1 / 0 / +1 (reference long_term_planner.h:54-56).
template <typename T>
int sign(T val) {
  return (T(0) < val) - (val < T(0));
}

/// Device buffers of a batched plan. All pointers are DEVICE pointers on the planner's GPU,
/// laid out as described in ltp_b200.h (inputs joint-major [dof][n]; trajectories time-major
/// (samples, n, dof) or one row per (problem, joint)).
struct BatchPlan {
  ltp_solution solution;   ///< 64-byte records [dof][n][8] (seven switching times + v_drive), directions,
                           ///< lengths, flags -- see ltp_b200.h; v_drive as a separate array is optional
  double* q = nullptr;     ///< sampled positions     (may all four be null: solve only)
  double* v = nullptr;     ///< sampled velocities
  double* a = nullptr;     ///< sampled accelerations
  double* j = nullptr;     ///< sampled jerks
  uint8_t* success = nullptr;  ///< [n] planTrajectory's return value per problem
  int32_t horizon = 0;     ///< 0: exact length per problem; > 0: fixed number of samples
  int32_t layout = LTP_LAYOUT_TIME_MAJOR;
  int64_t stride = 0;      ///< sample capacity (time-major) or doubles per row (rows)
};

class LongTermPlanner {
 private:
  int dof_;
  double t_sample_;
  std::vector<double> q_min_, q_max_, v_max_, a_max_, j_max_;
  int device_;
  // the device-side planner is created on first use and re-synchronised after a setter ran;
  // copies of this object get their own (copy = value semantics, like the reference)
  mutable std::shared_ptr<ltp_planner> handle_;
  mutable bool dirty_;
  ltp_planner* handle() const;

 public:
  /// Dummy planner (reference long_term_planner.h:103-105).
  LongTermPlanner() : dof_(0), t_sample_(0.001), device_(0), dirty_(true) {}

  /// reference long_term_planner.h:118-131
  LongTermPlanner(int dof, double t_sample, std::vector<double> q_min, std::vector<double> q_max,
                  std::vector<double> v_max, std::vector<double> a_max, std::vector<double> j_max)
      : dof_(dof), t_sample_(t_sample), q_min_(q_min), q_max_(q_max), v_max_(v_max), a_max_(a_max),
        j_max_(j_max), device_(0), dirty_(true) {}

  LongTermPlanner(const LongTermPlanner& o)
      : dof_(o.dof_), t_sample_(o.t_sample_), q_min_(o.q_min_), q_max_(o.q_max_), v_max_(o.v_max_),
        a_max_(o.a_max_), j_max_(o.j_max_), device_(o.device_), dirty_(true) {}
  LongTermPlanner& operator=(const LongTermPlanner& o) {
    if (this != &o) {
      dof_ = o.dof_; t_sample_ = o.t_sample_; q_min_ = o.q_min_; q_max_ = o.q_max_; v_max_ = o.v_max_;
      a_max_ = o.a_max_; j_max_ = o.j_max_; device_ = o.device_;
      handle_.reset();
      dirty_ = true;
    }
    return *this;
  }

  /// Plan one trajectory (reference long_term_planner.h:144-150, long_term_planner.cc:7-63).
  /// Returns false, leaving traj untouched, when the inputs are rejected or a joint has no
  /// solution; returns false with traj populated when a joint ends outside its limits.
  bool planTrajectory(const std::vector<double>& q_goal, const std::vector<double>& q_0,
                      const std::vector<double>& v_0, const std::vector<double>& a_0, Trajectory& traj);

  /// reference long_term_planner.h:161-165, long_term_planner.cc:68-77
  bool checkInputs(const std::vector<double>& q_0, const std::vector<double>& v_0,
                   const std::vector<double>& a_0);

  /// reference long_term_planner.h:176-187
  inline void setLimits(std::vector<double> q_min, std::vector<double> q_max, std::vector<double> v_max,
                        std::vector<double> a_max, std::vector<double> j_max) {
    q_min_ = q_min; q_max_ = q_max; v_max_ = v_max; a_max_ = a_max; j_max_ = j_max;
    dirty_ = true;
  }
  /// reference long_term_planner.h:194-196
  inline void setSampleTime(double t_sample) { t_sample_ = t_sample; dirty_ = true; }
  /// reference long_term_planner.h:203-205 (takes a double there, too)
  inline void setDoF(double dof) { dof_ = dof; dirty_ = true; }

  /// GPU the planner runs on (default 0). Not in the reference.
  inline void setDevice(int device) { device_ = device; handle_.reset(); dirty_ = true; }
  inline int device() const { return device_; }

  /// NEW: n independent problems in one call. Inputs are device pointers, joint-major
  /// [dof][n]. Enqueues the solve (and, when plan.q is set, the sampler) on `stream`
  /// (cudaStream_t as void*) and returns an ltp_status without synchronising.
  int planTrajectories(int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                       const double* a_0, const BatchPlan& plan, void* stream = nullptr);

  /// NEW: planTrajectories for more trajectories than fit in memory (ltp_plan_stream): the run
  /// is cut into chunks of `chunk` problems that are solved and sampled (time-major) into a
  /// two-slot ring and handed to `consume` (see ltp_b200.h). Synchronises before returning.
  /// sorted_slots (exact-length mode): every chunk's problems are ordered by trajectory length
  /// on the device and trajectory slot k holds problem ltp_chunk.order[k] (ltp_set_stream_sorted).
  int planStream(int64_t n, const double* q_goal, const double* q_0, const double* v_0, const double* a_0,
                 int64_t chunk, int32_t horizon, int64_t capacity, ltp_chunk_consumer consume, void* user,
                 ltp_stream_stats* stats = nullptr, bool sorted_slots = false, void* input_stream = nullptr);

  /// NEW: receding-horizon step (ltp_advance_batch): the state `tick` samples into time-major
  /// trajectories becomes the next start state, clamped to what checkInputs accepts.
  /// capacity: sample capacity (first extent) of q, v, a; valid: pass the solution's `reached`.
  int advance(int64_t n, int32_t tick, int64_t capacity, const int32_t* traj_len, const uint8_t* valid,
              const double* q, const double* v, const double* a, double* q_0, double* v_0, double* a_0,
              void* stream = nullptr);

 protected:
  /// reference long_term_planner.h:223-231, long_term_planner.cc:82-353
  bool optSwitchTimes(int joint, double q_goal, double q_0, double v_0, double a_0, double v_drive,
                      std::array<double, 7>& t, double& dir, char& mod_jerk_profile);
  /// reference long_term_planner.h:249-259, long_term_planner.cc:358-645
  bool timeScaling(int joint, double q_goal, double q_0, double v_0, double a_0, double dir,
                   double t_required, std::array<double, 7>& scaled_t, double& v_drive,
                   char& mod_jerk_profile);
  /// reference long_term_planner.h:279-285, long_term_planner.cc:650-701
  bool optBraking(int joint, double v_0, double a_0, double& q, std::array<double, 7>& t_rel, double& dir);
  /// reference long_term_planner.h:299-307, long_term_planner.cc:706-841
  Trajectory getTrajectory(const std::vector<std::array<double, 7>>& t, const std::vector<double>& dir,
                           const std::vector<char>& mod_jerk_profile, const std::vector<double>& q_0,
                           const std::vector<double>& v_0, const std::vector<double>& a_0,
                           const std::vector<double>& v_drive);
};

}  // namespace long_term_planner

#endif  // long_term_planner_H
