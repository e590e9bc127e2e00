// Host side of the drop-in class: every method is a thin call into the C ABI
// (include/ltp_b200.h -> CUDA kernels). See include/long_term_planner/long_term_planner.h.
#include "long_term_planner/long_term_planner.h"

#include <cstdio>
#include <initializer_list>
#include <stdexcept>

namespace long_term_planner {

namespace {
const unsigned char kCaseFailUntouched = 13;  // cc:340-344: returns false WITHOUT writing t
}

ltp_planner* LongTermPlanner::handle() const {
  // The reference has no joint-count limit and reports errors through bool only; here a planner
  // with more than LTP_MAX_DOF joints or limit vectors shorter than dof cannot be put on the
  // device, and silently keeping the previous configuration would index past the staging block.
  if (dof_ < 0 || dof_ > LTP_MAX_DOF)
    throw std::invalid_argument("long_term_planner: dof outside [0, LTP_MAX_DOF = 32]");
  for (const std::vector<double>* v : {&q_min_, &q_max_, &v_max_, &a_max_, &j_max_})
    if ((int)v->size() < dof_) throw std::invalid_argument("long_term_planner: a limit vector is shorter than dof");
  if (!handle_) {
    ltp_planner* h = nullptr;
    int rc = ltp_create(&h, device_, dof_, t_sample_, q_min_.data(), q_max_.data(), v_max_.data(),
                        a_max_.data(), j_max_.data());
    if (rc == LTP_ERR_ARG) throw std::invalid_argument("long_term_planner: ltp_create rejected the configuration");
    if (rc != LTP_OK) {
      // a missing GPU is not something a caller can recover from by looking at `false`, so it is loud
      std::fprintf(stderr, "long_term_planner: ltp_create failed: %s %s\n", ltp_status_string(rc),
                   ltp_last_cuda_error());
      throw std::runtime_error("long_term_planner: no CUDA device / ltp_create failed (no CPU fallback)");
    }
    handle_ = std::shared_ptr<ltp_planner>(h, [](ltp_planner* p) { ltp_destroy(p); });
    dirty_ = false;
  } else if (dirty_) {
    int rc = ltp_set_dof(handle_.get(), dof_);
    if (rc == LTP_OK) rc = ltp_set_sample_time(handle_.get(), t_sample_);
    if (rc == LTP_OK && dof_ > 0)
      rc = ltp_set_limits(handle_.get(), q_min_.data(), q_max_.data(), v_max_.data(), a_max_.data(), j_max_.data());
    if (rc != LTP_OK) throw std::invalid_argument("long_term_planner: the device planner rejected the new configuration");
    dirty_ = false;
  }
  return handle_.get();
}

bool LongTermPlanner::checkInputs(const std::vector<double>& q_0, const std::vector<double>& v_0,
                                  const std::vector<double>& a_0) {
  for (int i = 0; i < dof_; i++) {
    if (q_0[i] < q_min_[i] || q_0[i] > q_max_[i] || std::fabs(v_0[i]) > v_max_[i] ||
        std::fabs(a_0[i]) > a_max_[i])
      return false;
    if (std::fabs(v_0[i] + 0.5 * a_0[i] * std::fabs(a_0[i]) / j_max_[i]) > v_max_[i]) return false;
  }
  return true;
}

bool LongTermPlanner::planTrajectory(const std::vector<double>& q_goal, const std::vector<double>& q_0,
                                     const std::vector<double>& v_0, const std::vector<double>& a_0,
                                     Trajectory& traj) {
  if (dof_ < 1) return false;
  ltp_planner* h = handle();
  {
    // latency path: the sampled rows are read straight out of the planner's staging block
    const double* view[4];
    int64_t pitch = 0;
    int32_t vlen = 0;
    uint8_t vok = 0;
    const int rc = ltp_plan_one_view(h, q_goal.data(), q_0.data(), v_0.data(), a_0.data(), view, &pitch, &vlen, &vok);
    if (rc == LTP_OK) {
      if (vlen <= 0) return false;  // early `return false` of the reference: traj untouched
      traj.dof = dof_;
      traj.length = vlen;
      traj.t_sample = t_sample_;
      std::vector<std::vector<double>>* out[4] = {&traj.q, &traj.v, &traj.a, &traj.j};
      for (int f = 0; f < 4; ++f) {
        out[f]->resize(dof_);
        for (int i = 0; i < dof_; ++i) (*out[f])[i].assign(view[f] + (size_t)i * pitch, view[f] + (size_t)i * pitch + vlen);
      }
      return vok != 0;
    }
    if (rc != LTP_ERR_CAPACITY) return false;
  }
  int64_t cap = 4096, needed = 0;
  // receive buffer, kept between calls (every sample that is read back below was written by
  // the call; nothing relies on a fill)
  thread_local std::vector<double> rows;
  int32_t len = 0;
  uint8_t ok = 0;
  for (;;) {
    if (rows.size() < (size_t)4 * dof_ * cap) rows.resize((size_t)4 * dof_ * cap);
    double* q = rows.data();
    double* v = q + (size_t)dof_ * cap;
    double* a = v + (size_t)dof_ * cap;
    double* j = a + (size_t)dof_ * cap;
    int rc = ltp_plan_host(h, 1, q_goal.data(), q_0.data(), v_0.data(), a_0.data(), 0, cap, q, v, a, j, &len,
                           &ok, &needed);
    if (rc == LTP_ERR_CAPACITY) {
      cap = needed;
      continue;
    }
    if (rc != LTP_OK) return false;
    break;
  }
  if (len <= 0) return false;  // early `return false` of the reference: traj untouched
  traj.dof = dof_;
  traj.length = len;
  traj.t_sample = t_sample_;
  std::vector<std::vector<double>>* out[4] = {&traj.q, &traj.v, &traj.a, &traj.j};
  for (int f = 0; f < 4; ++f) {
    out[f]->assign(dof_, std::vector<double>());
    for (int i = 0; i < dof_; ++i) {
      const double* src = rows.data() + ((size_t)f * dof_ + i) * cap;
      (*out[f])[i].assign(src, src + len);
    }
  }
  return ok != 0;
}

int LongTermPlanner::planTrajectories(int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                                      const double* a_0, const BatchPlan& plan, void* stream) {
  ltp_planner* h = handle();
  int rc = ltp_solve_batch(h, n, q_goal, q_0, v_0, a_0, &plan.solution, stream);
  if (rc != LTP_OK || !plan.q) return rc;
  return ltp_sample_batch(h, n, q_0, v_0, a_0, &plan.solution, plan.horizon, plan.layout, plan.stride, plan.q,
                          plan.v, plan.a, plan.j, plan.success, stream);
}

int LongTermPlanner::planStream(int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                                const double* a_0, int64_t chunk, int32_t horizon, int64_t capacity,
                                ltp_chunk_consumer consume, void* user, ltp_stream_stats* stats,
                                bool sorted_slots, void* input_stream) {
  ltp_planner* h = handle();
  ltp_set_stream_sorted(h, sorted_slots ? 1 : 0);
  return ltp_plan_stream(h, n, q_goal, q_0, v_0, a_0, chunk, horizon, capacity, consume, user, stats,
                         input_stream);
}

int LongTermPlanner::advance(int64_t n, int32_t tick, int64_t capacity, const int32_t* traj_len,
                             const uint8_t* valid, const double* q, const double* v, const double* a,
                             double* q_0, double* v_0, double* a_0, void* stream) {
  return ltp_advance_batch(handle(), n, tick, 1, capacity, traj_len, valid, q, v, a, q_0, v_0, a_0, stream);
}

bool LongTermPlanner::optSwitchTimes(int joint, double q_goal, double q_0, double v_0, double a_0,
                                     double v_drive, std::array<double, 7>& t, double& dir,
                                     char& mod_jerk_profile) {
  double tt[7];
  uint8_t mod = 0, kase = 0, ok = 0;
  double d = 0;
  if (ltp_opt_switch_times_host(handle(), joint, q_goal, q_0, v_0, a_0, v_drive, tt, &d, &mod, &kase, &ok) !=
      LTP_OK)
    return false;
  dir = d;
  mod_jerk_profile = (char)mod;
  if (!(ok == 0 && (kase & 15) == kCaseFailUntouched))
    for (int k = 0; k < 7; ++k) t[k] = tt[k];
  return ok != 0;
}

bool LongTermPlanner::timeScaling(int joint, double q_goal, double q_0, double v_0, double a_0, double dir,
                                  double t_required, std::array<double, 7>& scaled_t, double& v_drive,
                                  char& mod_jerk_profile) {
  double tt[7], vd = 0;
  uint8_t mod = 0, tsc = 0, ok = 0;
  if (ltp_time_scaling_host(handle(), joint, q_goal, q_0, v_0, a_0, dir, t_required, tt, &vd, &mod, &tsc,
                            &ok) != LTP_OK)
    return false;
  for (int k = 0; k < 7; ++k) scaled_t[k] = tt[k];
  v_drive = vd;
  mod_jerk_profile = (char)mod;
  return ok != 0;
}

bool LongTermPlanner::optBraking(int joint, double v_0, double a_0, double& q, std::array<double, 7>& t_rel,
                                 double& dir) {
  double t3[3];
  if (ltp_opt_braking_host(handle(), joint, v_0, a_0, &q, t3, &dir) != LTP_OK) return false;
  t_rel[0] = t3[0];  // only the first three entries are written (cc:679-688)
  t_rel[1] = t3[1];
  t_rel[2] = t3[2];
  return true;
}

Trajectory LongTermPlanner::getTrajectory(const std::vector<std::array<double, 7>>& t,
                                          const std::vector<double>& dir,
                                          const std::vector<char>& mod_jerk_profile,
                                          const std::vector<double>& q_0, const std::vector<double>& v_0,
                                          const std::vector<double>& a_0, const std::vector<double>& v_drive) {
  Trajectory traj;
  traj.dof = dof_;
  traj.t_sample = t_sample_;
  traj.length = 0;
  if (dof_ < 1) return traj;
  std::vector<double> t7((size_t)7 * dof_);
  std::vector<uint8_t> mod(dof_);
  for (int i = 0; i < dof_; ++i) {
    for (int k = 0; k < 7; ++k) t7[7 * i + k] = t[i][k];
    mod[i] = (uint8_t)mod_jerk_profile[i];
  }
  int64_t cap = 4096, needed = 0;
  std::vector<double> rows;
  int32_t len = 0;
  for (;;) {
    rows.assign((size_t)4 * dof_ * cap, 0.0);
    double* q = rows.data();
    double* v = q + (size_t)dof_ * cap;
    double* a = v + (size_t)dof_ * cap;
    double* j = a + (size_t)dof_ * cap;
    int rc = ltp_get_trajectory_host(handle(), t7.data(), dir.data(), mod.data(), q_0.data(), v_0.data(),
                                     a_0.data(), v_drive.data(), cap, q, v, a, j, &len, &needed);
    if (rc == LTP_ERR_CAPACITY) {
      cap = needed;
      continue;
    }
    if (rc != LTP_OK) return traj;
    break;
  }
  traj.length = len;
  std::vector<std::vector<double>>* out[4] = {&traj.q, &traj.v, &traj.a, &traj.j};
  for (int f = 0; f < 4; ++f) {
    out[f]->assign(dof_, std::vector<double>());
    for (int i = 0; i < dof_; ++i) {
      const double* src = rows.data() + ((size_t)f * dof_ + i) * cap;
      (*out[f])[i].assign(src, src + len);
    }
  }
  return traj;
}

}  // namespace long_term_planner
