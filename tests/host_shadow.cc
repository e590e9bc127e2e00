// TEST ARTEFACT ONLY. Compiles the product's per-thread device math
// (longtermplanner_b200/csrc/ltp_math.cuh) for the host so that its arithmetic can be
// compared with the oracle in a container that has no GPU (tests/test_devmath_host.py).
// The product library never links, loads or falls back to this: it exists because every
// GPU run costs minutes, and a formula typo should be caught before spending them.
// Build: g++ -O2 -ffp-contract=off (no FMA contraction; explicit fma() stays exact).
#include <cstdint>
#include <cmath>
#include <cstring>
#include <vector>

#include "../longtermplanner_b200/csrc/ltp_math.cuh"
#include "../longtermplanner_b200/csrc/ltp_pipeline.cuh"

using namespace ltp;

namespace {
struct Shadow {
  int dof;
  double ts;
  std::vector<JointLimits> lim;
};
}  // namespace

extern "C" {

void* shadow_create(int dof, double ts, const double* q_min, const double* q_max, const double* v_max,
                    const double* a_max, const double* j_max) {
  Shadow* s = new Shadow;
  s->dof = dof;
  s->ts = ts;
  s->lim.resize(dof);
  for (int i = 0; i < dof; ++i) {
    s->lim[i] = JointLimits{q_min[i], q_max[i], v_max[i], a_max[i], j_max[i], 0, 0, 0};
    derive_limits(s->lim[i]);
  }
  return s;
}
void shadow_destroy(void* h) { delete static_cast<Shadow*>(h); }

void shadow_opt_braking_items(void* h, int64_t n, const int* joint, const double* v_0, const double* a_0,
                              double* q, double* t_rel3, double* dir) {
  Shadow* s = static_cast<Shadow*>(h);
  for (int64_t i = 0; i < n; ++i) {
    const JointLimits& L = s->lim[joint ? joint[i] : 0];
    q[i] = brake_profile(L, s->ts, v_0[i], a_0[i], t_rel3[3 * i], t_rel3[3 * i + 1],
                         t_rel3[3 * i + 2], dir[i]);
  }
}

void shadow_opt_switch_times_items(void* h, int64_t n, const int* joint, const double* q_goal,
                                   const double* q_0, const double* v_0, const double* a_0,
                                   const double* v_drive, double* t7, double* dir, unsigned char* mod,
                                   unsigned char* kase, unsigned char* ok, int) {
  Shadow* s = static_cast<Shadow*>(h);
  for (int64_t i = 0; i < n; ++i) {
    const JointLimits& L = s->lim[joint ? joint[i] : 0];
    Prologue P = ost_prologue(L, s->ts, q_goal[i], q_0[i], v_0[i], a_0[i]);
    double t[7];
    zero7(t);
    unsigned char m = 0, c = 255;
    ok[i] = ost_body(L, s->ts, P, q_goal[i], q_0[i], v_drive[i], t, m, c);
    std::memcpy(t7 + 7 * i, t, 56);
    dir[i] = P.dir;
    mod[i] = m;
    kase[i] = c;
  }
}

void shadow_time_scaling_items(void* h, int64_t n, const int* joint, const double* q_goal,
                               const double* q_0, const double* v_0, const double* a_0, const double* dir,
                               const double* t_required, double* t7, double* v_drive, unsigned char* mod,
                               unsigned char* ts_case, unsigned char* final_case, unsigned char* ok, int) {
  Shadow* s = static_cast<Shadow*>(h);
  for (int64_t i = 0; i < n; ++i) {
    const JointLimits& L = s->lim[joint ? joint[i] : 0];
    TsInput I = make_ts_input(q_goal[i], q_0[i], v_0[i], a_0[i], dir[i], t_required[i]);
    Prologue P = ost_prologue(L, s->ts, q_goal[i], q_0[i], dir[i] * I.v_0, dir[i] * I.a_0);
    double t[7];
    zero7(t);
    unsigned char m = 0, fc = 255;
    int c = time_scaling_from(1, L, s->ts, P, I, t, v_drive[i], m, fc);
    std::memcpy(t7 + 7 * i, t, 56);
    mod[i] = m;
    ts_case[i] = (unsigned char)c;
    final_case[i] = fc;
    ok[i] = c != 9;
  }
}

// same sequence of device-function calls as ltp_solve_kernel, problem-major arrays
void shadow_solve_batch(void* h, int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                        const double* a_0, double* t_opt, double* t_scaled, double* dir, double* v_drive,
                        unsigned char* mod, unsigned char* opt_case, unsigned char* ts_case,
                        unsigned char* final_case, int* slowest, int* traj_len, unsigned char* reached, int) {
  Shadow* s = static_cast<Shadow*>(h);
  const int dof = s->dof;
  std::vector<Prologue> pro(dof);
  for (int64_t p = 0; p < n; ++p) {
    const int64_t o = p * dof;
    bool any_fail = false;
    for (int j = 0; j < dof; ++j) {
      const JointLimits& L = s->lim[j];
      bool in_ok = check_joint_input(L, q_0[o + j], v_0[o + j], a_0[o + j]);
      pro[j] = ost_prologue(L, s->ts, q_goal[o + j], q_0[o + j], v_0[o + j], a_0[o + j]);
      double* t = t_opt + 7 * (o + j);
      zero7(t);
      mod[o + j] = 0;
      opt_case[o + j] = 255;
      bool ok = ost_body(L, s->ts, pro[j], q_goal[o + j], q_0[o + j], L.v_max, t, mod[o + j], opt_case[o + j]);
      any_fail |= !(in_ok && ok);
      dir[o + j] = pro[j].dir;
    }
    double t_req = -1;
    int sl = -1;
    for (int j = 0; j < dof; ++j)
      if (t_opt[7 * (o + j) + 6] > t_req) { t_req = t_opt[7 * (o + j) + 6]; sl = j; }
    const bool rch = !any_fail && sl != -1;
    int len = 0;
    bool bad = false;
    for (int j = 0; j < dof; ++j) {
      const JointLimits& L = s->lim[j];
      double* t = t_scaled + 7 * (o + j);
      zero7(t);
      v_drive[o + j] = L.v_max;
      ts_case[o + j] = 255;
      final_case[o + j] = 255;
      if (rch) {
        if (j == sl) {
          ts_case[o + j] = 0;
          final_case[o + j] = opt_case[o + j];
        } else {
          TsInput I = make_ts_input(q_goal[o + j], q_0[o + j], v_0[o + j], a_0[o + j], pro[j].dir, t_req);
          ts_case[o + j] = (unsigned char)time_scaling_from(1, L, s->ts, pro[j], I, t, v_drive[o + j],
                                                            mod[o + j], final_case[o + j]);
          if (ts_case[o + j] == 9) final_case[o + j] = opt_case[o + j];
        }
        double m = t[0];
        for (int k = 1; k < 7; ++k)
          if (m < t[k]) m = t[k];
        if (m <= 0.0) std::memcpy(t, t_opt + 7 * (o + j), 56);
        bool fin = true;
        for (int k = 0; k < 7; ++k) fin &= (bool)std::isfinite(t[k]);
        int li = (fin && t[6] / s->ts <= 2.0e9) ? samples_for(t[6], s->ts) : -1;
        bad |= li < 0;
        len = li > len ? li : len;
      }
    }
    slowest[p] = sl;
    traj_len[p] = (rch && !bad) ? len : 0;
    reached[p] = rch;
  }
}

// Mirrors ltp_solve_fast_kernel + the work-list hand-over to the generic kernel: the
// closed-form pass decides per problem whether it is complete or deferred; deferred problems
// are recomputed by shadow_solve_batch (the generic sequence). Returns the number deferred.
int64_t shadow_solve_batch_auto(void* h, int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                                const double* a_0, double* t_opt, double* t_scaled, double* dir, double* v_drive,
                                unsigned char* mod, unsigned char* opt_case, unsigned char* ts_case,
                                unsigned char* final_case, int* slowest, int* traj_len, unsigned char* reached, int) {
  Shadow* s = static_cast<Shadow*>(h);
  const int dof = s->dof;
  std::vector<Prologue> pro(dof);
  int64_t deferred = 0;
  for (int64_t p = 0; p < n; ++p) {
    const int64_t o = p * dof;
    bool any_fail = false, defer = false;
    for (int j = 0; j < dof; ++j) {
      const JointLimits& L = s->lim[j];
      bool in_ok = check_joint_input(L, q_0[o + j], v_0[o + j], a_0[o + j]);
      pro[j] = ost_prologue(L, s->ts, q_goal[o + j], q_0[o + j], v_0[o + j], a_0[o + j]);
      double* t = t_opt + 7 * (o + j);
      zero7(t);
      mod[o + j] = 0;
      opt_case[o + j] = 255;
      int st = ost_body_t<false>(L, s->ts, pro[j], q_goal[o + j], q_0[o + j], L.v_max, t, mod[o + j], opt_case[o + j]);
      any_fail |= !(in_ok && st != OST_FAIL);
      defer |= st == OST_DEFER;
      dir[o + j] = pro[j].dir;
    }
    double t_req = -1;
    int sl = -1;
    for (int j = 0; j < dof; ++j)
      if (t_opt[7 * (o + j) + 6] > t_req) { t_req = t_opt[7 * (o + j) + 6]; sl = j; }
    const bool rch = !any_fail && sl != -1;
    int len = 0;
    bool bad = false;
    for (int j = 0; j < dof && !defer; ++j) {
      const JointLimits& L = s->lim[j];
      double* t = t_scaled + 7 * (o + j);
      zero7(t);
      v_drive[o + j] = L.v_max;
      ts_case[o + j] = 255;
      final_case[o + j] = 255;
      if (rch) {
        if (j == sl) {
          ts_case[o + j] = 0;
          final_case[o + j] = opt_case[o + j];
        } else {
          TsInput I = make_ts_input(q_goal[o + j], q_0[o + j], v_0[o + j], a_0[o + j], pro[j].dir, t_req);
          int c = time_scaling_closed_form(L, s->ts, pro[j], I, t, v_drive[o + j], mod[o + j], final_case[o + j]);
          if (c == 0) { defer = true; break; }
          ts_case[o + j] = (unsigned char)c;
          if (c == 9) final_case[o + j] = opt_case[o + j];
        }
        double m = t[0];
        for (int k = 1; k < 7; ++k)
          if (m < t[k]) m = t[k];
        if (m <= 0.0) std::memcpy(t, t_opt + 7 * (o + j), 56);
        bool fin = true;
        for (int k = 0; k < 7; ++k) fin &= (bool)std::isfinite(t[k]);
        int li = (fin && t[6] / s->ts <= 2.0e9) ? samples_for(t[6], s->ts) : -1;
        bad |= li < 0;
        len = li > len ? li : len;
      }
    }
    if (defer) {
      ++deferred;
      shadow_solve_batch(h, 1, q_goal + o, q_0 + o, v_0 + o, a_0 + o, t_opt + 7 * o, t_scaled + 7 * o, dir + o,
                         v_drive + o, mod + o, opt_case + o, ts_case + o, final_case + o, slowest + p,
                         traj_len + p, reached + p, 0);
      continue;
    }
    slowest[p] = sl;
    traj_len[p] = (rch && !bad) ? len : 0;
    reached[p] = rch;
  }
  return deferred;
}


// Mirrors the regrouped solve of ltp_b200.cu (ltp_solve_stage1_kernel -> ltp_solve_scale_kernel with
// its class-A / class-B / second-candidate slots -> every-branch kernel for what is deferred): the
// same per-item functions (csrc/ltp_pipeline.cuh) and the same hand-overs, with std::vector lists
// in place of the shared-memory lists. stats[0..4] = joints settled without a search, class A,
// class B, second-candidate items, problems deferred.
int64_t shadow_solve_batch_pipeline(void* h, int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                                    const double* a_0, double* t_opt, double* t_scaled, double* dir, double* v_drive,
                                    unsigned char* mod, unsigned char* opt_case, unsigned char* ts_case,
                                    unsigned char* final_case, int* slowest, int* traj_len, unsigned char* reached,
                                    int64_t* stats) {
  Shadow* s = static_cast<Shadow*>(h);
  const int dof = s->dof;
  const int kMark = 0x7fffffff;
  struct Queued { ScaleItem it; int64_t p; };
  std::vector<std::vector<Queued>> qa(dof), qb(dof), qc(dof);
  std::vector<unsigned char> jflag((size_t)n * dof);
  std::vector<int64_t> work;
  int64_t settled = 0;
  auto put = [&](int64_t p, int j, const JointResult& R) {
    const int64_t o = p * dof + j;
    std::memcpy(t_scaled + 7 * o, R.t, 56);
    v_drive[o] = R.v_drive;
    mod[o] = R.mod;
    ts_case[o] = R.ts_case;
    final_case[o] = R.final_case;
  };
  // kernel 1: every joint's record as if it were the slowest one of its problem
  for (int64_t p = 0; p < n; ++p)
    for (int j = 0; j < dof; ++j) {
      const int64_t o = p * dof + j;
      Stage1Out S1;
      stage1_joint(s->lim[j], s->ts, q_goal[o], q_0[o], v_0[o], a_0[o], S1);
      std::memcpy(t_opt + 7 * o, S1.t_opt, 56);
      std::memcpy(t_scaled + 7 * o, S1.t_opt, 56);
      dir[o] = S1.dir;
      v_drive[o] = s->lim[j].v_max;
      mod[o] = S1.mod;
      opt_case[o] = S1.opt_case;
      ts_case[o] = 0;
      final_case[o] = S1.opt_case;
      jflag[o] = S1.flags;
    }
  auto defer_problem = [&](int64_t p) {
    if (traj_len[p] != kMark) {
      traj_len[p] = kMark;
      work.push_back(p);
    }
  };
  auto join_len = [&](int64_t p, const double* t) {
    const int li = joint_sample_count(t, s->ts);
    if (li < 0) return defer_problem(p);
    if (traj_len[p] < li) traj_len[p] = li;  // atomicMax
  };
  // kernel 2, first part: slowest joint, then every joint decides what happens to it
  for (int64_t p = 0; p < n; ++p) {
    const int64_t o = p * dof;
    double t_req = -1;
    int sl = -1;
    unsigned char any = 0;
    for (int j = 0; j < dof; ++j) {
      any |= jflag[o + j];
      if (t_scaled[7 * (o + j) + 6] > t_req) { t_req = t_scaled[7 * (o + j) + 6]; sl = j; }
    }
    const bool rch = !(any & JF_FAIL) && sl != -1;
    slowest[p] = sl;
    reached[p] = rch;
    traj_len[p] = 0;
    if (any & JF_DEFER) {
      defer_problem(p);
      continue;
    }
    for (int j = 0; j < dof; ++j) {
      const int64_t oj = o + j;
      if (!rch) {  // aborted before time scaling (cc:15,29,39): the record says so
        JointResult R;
        zero7(R.t);
        R.v_drive = s->lim[j].v_max;
        R.mod = mod[oj];
        R.ts_case = 255;
        R.final_case = 255;
        put(p, j, R);
        continue;
      }
      if (j == sl) {
        join_len(p, t_scaled + 7 * oj);
        ++settled;
      } else if (jflag[oj] & JF_BRAKE_ONLY) {
        const Prologue pro = ost_prologue(s->lim[j], s->ts, q_goal[oj], q_0[oj], v_0[oj], a_0[oj]);
        JointResult R;
        const int act = stage3_brake_only(s->lim[j], s->ts, pro, q_goal[oj], q_0[oj], v_0[oj], a_0[oj], t_req,
                                          t_opt + 7 * oj, opt_case[oj], R);
        if (act == S3_SETTLED) {
          put(p, j, R);
          join_len(p, R.t);
          ++settled;
        } else {
          defer_problem(p);
        }
      } else {
        Queued Q;
        Q.p = p;
        const int act = stage3_classify(s->lim[j], q_goal[oj], q_0[oj], v_0[oj], a_0[oj], dir[oj], t_req, Q.it);
        (act == S3_QUEUE_A ? qa : qb)[j].push_back(Q);
      }
    }
  }
  auto finish = [&](int j, const Queued& Q, const JointResult& R) {
    if (joint_sample_count(R.t, s->ts) < 0) return defer_problem(Q.p);
    put(Q.p, j, R);
    join_len(Q.p, R.t);
  };
  int64_t na = 0, nb = 0, nc = 0;
  for (int j = 0; j < dof; ++j) {
    na += (int64_t)qa[j].size();
    nb += (int64_t)qb[j].size();
    for (const Queued& Q : qa[j]) {
      JointResult R;
      const int r = scale_attempt1_class_a(s->lim[j], s->ts, Q.it, R);
      if (r == SA_ACCEPT) finish(j, Q, R);
      else if (r == SA_DEFER) defer_problem(Q.p);
      else qc[j].push_back(Q);
    }
    for (const Queued& Q : qb[j]) {
      JointResult R;
      const int r = scale_attempt1_class_b(s->lim[j], s->ts, Q.it, R);
      if (r == SA_ACCEPT) finish(j, Q, R);
      else if (r == SA_DEFER) defer_problem(Q.p);
      else qc[j].push_back(Q);
    }
    nc += (int64_t)qc[j].size();
    for (const Queued& Q : qc[j]) {
      JointResult R;
      const int r = scale_attempt2(s->lim[j], s->ts, Q.it, R);
      if (r == SA_ACCEPT) finish(j, Q, R);
      else defer_problem(Q.p);
    }
  }
  for (int64_t p : work) {
    const int64_t o = p * dof;
    shadow_solve_batch(h, 1, q_goal + o, q_0 + o, v_0 + o, a_0 + o, t_opt + 7 * o, t_scaled + 7 * o, dir + o,
                       v_drive + o, mod + o, opt_case + o, ts_case + o, final_case + o, slowest + p, traj_len + p,
                       reached + p, 0);
  }
  if (stats) {
    stats[0] = settled; stats[1] = na; stats[2] = nb; stats[3] = nc; stats[4] = (int64_t)work.size();
  }
  return (int64_t)work.size();
}

int shadow_get_trajectory(void* h, const double* t7, const double* dir, const unsigned char* mod,
                          const double* q_0, const double* v_0, const double* a_0, const double* v_drive,
                          int64_t stride, double* q, double* v, double* a, double* j) {
  Shadow* s = static_cast<Shadow*>(h);
  const int dof = s->dof;
  int len = 0;
  for (int i = 0; i < dof; ++i) {
    int li = samples_for(t7[7 * i + 6], s->ts);
    len = li > len ? li : len;
  }
  if (len > stride) return -len;
  for (int jt = 0; jt < dof; ++jt) {
    RowSampler R;
    R.init(s->ts, s->lim[jt].j_max, t7 + 7 * jt, dir[jt], mod[jt], q_0[jt], v_0[jt], a_0[jt], v_drive[jt], len);
    alignas(16) double table[2 * kMaxSeg];
    SegTableT<2> T{table};
    T.build(R, len);
    SegCursorT<2> C;
    C.begin(R);
    for (int i = 0; i < len; ++i)
      C.step(T, i, j[jt * stride + i], a[jt * stride + i], v[jt * stride + i], q[jt * stride + i]);
  }
  return len;
}

// largest |peek_position(i -> len) - q[len-1]| over all joints and a set of start samples i:
// the closed-form jump the time-major kernel uses for the limit check of clipped rows
double shadow_peek_error(void* h, const double* t7, const double* dir, const unsigned char* mod,
                         const double* q_0, const double* v_0, const double* a_0, const double* v_drive) {
  Shadow* s = static_cast<Shadow*>(h);
  const int dof = s->dof;
  int len = 0;
  for (int i = 0; i < dof; ++i) {
    int li = samples_for(t7[7 * i + 6], s->ts);
    len = li > len ? li : len;
  }
  double worst = 0.0;
  for (int jt = 0; jt < dof; ++jt) {
    RowSampler R;
    R.init(s->ts, s->lim[jt].j_max, t7 + 7 * jt, dir[jt], mod[jt], q_0[jt], v_0[jt], a_0[jt], v_drive[jt], len);
    alignas(16) double table[2 * kMaxSeg];
    SegTableT<2> T{table};
    T.build(R, len);
    SegCursorT<2> C;
    C.begin(R);
    std::vector<double> peek(len + 1);
    double jj, aa, vv, qq = q_0[jt];
    for (int i = 0; i < len; ++i) {
      if (i % 7 == 0 || i + 3 >= len) peek[i] = C.peek_position(T, i, len); else peek[i] = NAN;
      C.step(T, i, jj, aa, vv, qq);
    }
    for (int i = 0; i < len; ++i)
      if (peek[i] == peek[i]) {
        const double e = std::fabs(peek[i] - qq);
        worst = e > worst ? e : worst;
      }
  }
  return worst;
}

// same rows by the sample-by-sample general rules (RowSampler::step); the segment machinery
// above must reproduce this bit for bit
int shadow_get_trajectory_general(void* h, const double* t7, const double* dir, const unsigned char* mod,
                                  const double* q_0, const double* v_0, const double* a_0, const double* v_drive,
                                  int64_t stride, double* q, double* v, double* a, double* j) {
  Shadow* s = static_cast<Shadow*>(h);
  const int dof = s->dof;
  int len = 0;
  for (int i = 0; i < dof; ++i) {
    int li = samples_for(t7[7 * i + 6], s->ts);
    len = li > len ? li : len;
  }
  if (len > stride) return -len;
  for (int jt = 0; jt < dof; ++jt) {
    RowSampler R;
    R.init(s->ts, s->lim[jt].j_max, t7 + 7 * jt, dir[jt], mod[jt], q_0[jt], v_0[jt], a_0[jt], v_drive[jt], len);
    for (int i = 0; i < len; ++i)
      R.step(i, j[jt * stride + i], a[jt * stride + i], v[jt * stride + i], q[jt * stride + i]);
  }
  return len;
}

}  // extern "C"
