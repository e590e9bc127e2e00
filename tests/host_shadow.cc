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

// div_by(x, d, RN(1/d)) against the plain division x / d on n numerators: returns the number of
// results whose bits differ (NaNs compare equal to NaNs) and leaves the first offender in *bad
int64_t shadow_div_by_mismatches(double d, int64_t n, const double* x, double* bad) {
  const double rd = 1.0 / d;
  int64_t miss = 0;
  for (int64_t i = 0; i < n; ++i) {
    const double a = div_by(x[i], d, rd), b = x[i] / d;
    if (std::memcmp(&a, &b, 8) != 0 && !(a != a && b != b)) {
      if (miss == 0 && bad) *bad = x[i];
      ++miss;
    }
  }
  return miss;
}

// Stage 1 and attempt 1 of the closed-form kernel with the range test of the divisions deferred
// (DivDeferred) against the same functions with every quotient tested in place (DivChecked): an
// item whose deferred run is not flagged must have the bits of the checked run (a flagged item is
// recomputed by the checked functions on the device, so there is nothing to compare). Returns the
// number of unflagged items that differ; flagged[0] / flagged[1] count the flagged stage-1 / attempt-1 runs.
static bool same_bits(const void* a, const void* b, size_t n) { return std::memcmp(a, b, n) == 0; }

int64_t shadow_deferred_vs_checked(void* h, int64_t n, const int* joint, const double* q_goal, const double* q_0,
                                   const double* v_0, const double* a_0, const double* t_req, int64_t* flagged) {
  Shadow* s = static_cast<Shadow*>(h);
  int64_t miss = 0, flg = 0, flg2 = 0;
  for (int64_t i = 0; i < n; ++i) {
    const JointLimits& L = s->lim[joint ? joint[i] : 0];
    // stage 1 (the policy of the closed-form kernel for up to 8 joints: limit-only factors read
    // from JointLimits instead of being formed per item)
    DivDeferredWide dv;
    const bool in_d = check_joint_input(L, q_0[i], v_0[i], a_0[i], dv);
    const Prologue pd = ost_prologue(L, s->ts, q_goal[i], q_0[i], v_0[i], a_0[i], dv);
    double td[7];
    zero7(td);
    unsigned char md = 0, cd = 255;
    const int sd = ost_body_dv<false, true>(L, s->ts, pd, q_goal[i], q_0[i], L.v_max, td, md, cd, dv);
    const bool in_c = check_joint_input(L, q_0[i], v_0[i], a_0[i]);
    const Prologue pc = ost_prologue(L, s->ts, q_goal[i], q_0[i], v_0[i], a_0[i]);
    double tc[7];
    zero7(tc);
    unsigned char mc = 0, cc = 255;
    const int sc = ost_body_t<false, true>(L, s->ts, pc, q_goal[i], q_0[i], L.v_max, tc, mc, cc);
    if (dv.bad) {
      ++flg;
    } else {
      const double fd[7] = {pd.v0m, pd.a0m, pd.dir, pd.dist, pd.b0, pd.b1, pd.b2};
      const double fc[7] = {pc.v0m, pc.a0m, pc.dir, pc.dist, pc.b0, pc.b1, pc.b2};
      if (in_d != in_c || sd != sc || md != mc || cd != cc || pd.brake_only != pc.brake_only ||
          !same_bits(fd, fc, 56) || !same_bits(td, tc, 56)) {
        ++miss;
        continue;
      }
    }
    // attempt 1 from the checked prologue (what the device has at that point either way)
    const TsInput I = make_ts_input(q_goal[i], q_0[i], v_0[i], a_0[i], pc.dir, t_req[i]);
    DivDeferredWide dv2;
    double ad[7], ac[7];
    zero7(ad);
    zero7(ac);
    double vd = L.v_max, vc = L.v_max;
    unsigned char m2d = mc, m2c = mc, fcd = 255, fcc = 255;
    const int rd = time_scaling_attempt1(L, s->ts, pc, I, ad, vd, m2d, fcd, dv2);
    const int rc = time_scaling_attempt1(L, s->ts, pc, I, ac, vc, m2c, fcc);
    if (dv2.bad) {
      ++flg2;
    } else if (rd != rc || m2d != m2c || fcd != fcc || !same_bits(&vd, &vc, 8) || !same_bits(ad, ac, 56)) {
      ++miss;
    }
    // attempt 2 the way its kernel runs it: prologue and nested solve under one flag
    DivDeferred dv3;
    const Prologue p3 = ost_prologue(L, s->ts, q_goal[i], q_0[i], v_0[i], a_0[i], dv3);
    const TsInput I3 = make_ts_input(q_goal[i], q_0[i], v_0[i], a_0[i], p3.dir, t_req[i]);
    double bd[7], bc[7];
    zero7(bd);
    zero7(bc);
    double wd = L.v_max, wc = L.v_max;
    unsigned char m3d = 0, m3c = 0, gcd = 255, gcc = 255;
    const int qd = time_scaling_attempt2(L, s->ts, p3, I3, bd, wd, m3d, gcd, dv3);
    if (dv3.bad) {
      ++flg2;
    } else {
      const int qc = time_scaling_attempt2(L, s->ts, pc, I, bc, wc, m3c, gcc);
      if (qd != qc || m3d != m3c || gcd != gcc || !same_bits(&wd, &wc, 8) || !same_bits(bd, bc, 56)) ++miss;
    }
  }
  if (flagged) {
    flagged[0] = flg;
    flagged[1] = flg2;
  }
  return miss;
}

// ts_candidate2 through the prepared reciprocals (checked and deferred) against the expression
// with the reference's divisions written out; returns the number of items whose bits differ
int64_t shadow_candidate2_mismatches(void* h, int64_t n, const int* joint, const double* q_goal, const double* q_0,
                                     const double* v_0, const double* a_0, const double* dir, const double* t_req) {
  Shadow* s = static_cast<Shadow*>(h);
  int64_t miss = 0;
  for (int64_t i = 0; i < n; ++i) {
    const JointLimits& L = s->lim[joint ? joint[i] : 0];
    const TsInput I = make_ts_input(q_goal[i], q_0[i], v_0[i], a_0[i], dir[i], t_req[i]);
    const double a = ts_candidate2_plain(L, I), b = ts_candidate2(L, I);
    DivDeferred dv;
    const double c = ts_candidate2(L, I, dv);
    const bool nan_ok = a != a && b != b;
    if (!same_bits(&a, &b, 8) && !nan_ok) ++miss;
    else if (!dv.bad && !same_bits(&a, &c, 8) && !(a != a && c != c)) ++miss;
  }
  return miss;
}

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
      int st = ost_body_t<false, true>(L, s->ts, pro[j], q_goal[o + j], q_0[o + j], L.v_max, t, mod[o + j], opt_case[o + j]);
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

// Mirrors ITEM MODE of ltp_b200.cu (batches of 8192 problems and more): what the closed forms do
// not settle is handed on per (problem, joint) -- tail items (ltp_solve_tail_kernel), pending
// problems (ltp_solve_pending_kernel), search items (ltp_solve_search_kernel) -- with the same
// per-joint device functions and the same hand-over rules, lists instead of device queues.
// stats[0..3] = tail items, pending problems, search items, whole-problem deferrals.
int64_t shadow_solve_batch_items(void* h, int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                                 const double* a_0, double* t_opt, double* t_scaled, double* dir, double* v_drive,
                                 unsigned char* mod, unsigned char* opt_case, unsigned char* ts_case,
                                 unsigned char* final_case, int* slowest, int* traj_len, unsigned char* reached,
                                 int64_t* stats) {
  Shadow* s = static_cast<Shadow*>(h);
  const int dof = s->dof;
  const double Ts = s->ts;
  struct Item { int64_t p; int j; double t_req; };
  std::vector<Item> tails, searches;
  std::vector<int64_t> pending, whole;
  std::vector<unsigned char> tail_ok((size_t)n * dof, 0);
  auto sample_count = [&](const double* t) {
    bool fin = true;
    for (int k = 0; k < 7; ++k) fin &= (bool)std::isfinite(t[k]);
    return (fin && t[6] / Ts <= 2.0e9) ? samples_for(t[6], Ts) : -1;
  };
  auto fallback = [&](double* t, const double* topt) {
    double m = t[0];
    for (int k = 1; k < 7; ++k)
      if (m < t[k]) m = t[k];
    if (m <= 0.0) std::memcpy(t, topt, 56);
  };
  // stages 2-3 of one problem with closed forms only; joints that stay open become search items.
  // from_tail: the problem waited for tail items (their t_opt / flags are in place already)
  auto finish_problem = [&](int64_t p, bool from_tail) {
    const int64_t o = p * dof;
    std::vector<Prologue> pro(dof);
    bool any_fail = false, need_tail = false;
    for (int j = 0; j < dof; ++j) {
      const JointLimits& L = s->lim[j];
      const bool in_ok = check_joint_input(L, q_0[o + j], v_0[o + j], a_0[o + j]);
      pro[j] = ost_prologue(L, Ts, q_goal[o + j], q_0[o + j], v_0[o + j], a_0[o + j]);
      double t[7];
      zero7(t);
      unsigned char m = 0, oc = 255;
      const int st = ost_body_t<false, true>(L, Ts, pro[j], q_goal[o + j], q_0[o + j], L.v_max, t, m, oc);
      dir[o + j] = pro[j].dir;
      bool ok = st == OST_OK;
      if (st == OST_DEFER) {
        if (!from_tail) {
          need_tail = true;
          tails.push_back({p, j, 0.0});
          continue;
        }
        ok = tail_ok[o + j] != 0;  // t_opt, mod, opt_case were written by the tail step
      } else {
        std::memcpy(t_opt + 7 * (o + j), t, 56);
        mod[o + j] = m;
        opt_case[o + j] = oc;
      }
      any_fail |= !(in_ok && ok);
    }
    if (need_tail) {
      pending.push_back(p);
      return;
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
      if (!rch) continue;
      bool open = false;
      if (j == sl) {
        ts_case[o + j] = 0;
        final_case[o + j] = opt_case[o + j];
      } else {
        const TsInput I = make_ts_input(q_goal[o + j], q_0[o + j], v_0[o + j], a_0[o + j], pro[j].dir, t_req);
        const int c = time_scaling_closed_form(L, Ts, pro[j], I, t, v_drive[o + j], mod[o + j], final_case[o + j]);
        open = c == 0;
        ts_case[o + j] = (unsigned char)c;
        if (c == 9) final_case[o + j] = opt_case[o + j];
      }
      if (open) {
        searches.push_back({p, j, t_req});
        continue;
      }
      fallback(t, t_opt + 7 * (o + j));
      const int li = sample_count(t);
      bad |= li < 0;
      len = li > len ? li : len;
    }
    slowest[p] = sl;
    reached[p] = rch;
    traj_len[p] = (rch && !bad) ? len : 0;
    if (bad) whole.push_back(p);
  };
  for (int64_t p = 0; p < n; ++p) finish_problem(p, false);
  for (const Item& it : tails) {  // ltp_solve_tail_kernel
    const int64_t o = it.p * dof + it.j;
    const JointLimits& L = s->lim[it.j];
    const Prologue pro = ost_prologue(L, Ts, q_goal[o], q_0[o], v_0[o], a_0[o]);
    double* t = t_opt + 7 * o;
    zero7(t);
    mod[o] = 0;
    opt_case[o] = 255;
    tail_ok[o] = ost_body(L, Ts, pro, q_goal[o], q_0[o], L.v_max, t, mod[o], opt_case[o]);
  }
  const std::vector<int64_t> waiting = pending;
  for (int64_t p : waiting) finish_problem(p, true);  // ltp_solve_pending_kernel
  for (const Item& it : searches) {  // ltp_solve_search_kernel
    const int64_t o = it.p * dof + it.j;
    const JointLimits& L = s->lim[it.j];
    const Prologue pro = ost_prologue(L, Ts, q_goal[o], q_0[o], v_0[o], a_0[o]);
    double topt[7];
    zero7(topt);
    unsigned char om = 0, oc = 255;
    ost_body(L, Ts, pro, q_goal[o], q_0[o], L.v_max, topt, om, oc);
    const TsInput I = make_ts_input(q_goal[o], q_0[o], v_0[o], a_0[o], pro.dir, it.t_req);
    double* t = t_scaled + 7 * o;
    zero7(t);
    v_drive[o] = L.v_max;
    unsigned char m = om, fc = 255;
    const int c = time_scaling_from(1, L, Ts, pro, I, t, v_drive[o], m, fc);
    if (c == 9) fc = oc;
    mod[o] = m;
    ts_case[o] = (unsigned char)c;
    final_case[o] = fc;
    fallback(t, topt);
    const int li = sample_count(t);
    if (li < 0) whole.push_back(it.p);
    else if (traj_len[it.p] < li) traj_len[it.p] = li;  // atomicMax
  }
  for (int64_t p : whole) {
    const int64_t o = p * dof;
    shadow_solve_batch(h, 1, q_goal + o, q_0 + o, v_0 + o, a_0 + o, t_opt + 7 * o, t_scaled + 7 * o, dir + o,
                       v_drive + o, mod + o, opt_case + o, ts_case + o, final_case + o, slowest + p, traj_len + p,
                       reached + p, 0);
  }
  if (stats) {
    stats[0] = (int64_t)tails.size(); stats[1] = (int64_t)waiting.size();
    stats[2] = (int64_t)searches.size(); stats[3] = (int64_t)whole.size();
  }
  return (int64_t)whole.size();
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
