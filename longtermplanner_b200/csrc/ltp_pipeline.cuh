// Per-item logic of the regrouped solve (stages 1-3, reference cc:14-55), shared by the kernels
// of ltp_b200.cu and -- compiled for the host -- by tests/host_shadow.cc, which replays the same
// hand-overs on the CPU against the oracle.
//
// Why regroup. One thread per (problem, joint) that walks the reference's control flow from top
// to bottom keeps 21.7 of 32 lanes busy on random Franka problems (ncu, round 1): the slowest
// joint of every problem idles through the whole cruise-speed search, a quarter of the others
// take the modified-profile branch of the nested solve while the rest wait, and an eighth goes
// on to the second candidate. Here the search is cut where its control flow forks, and the
// joints that take the same side travel together:
//
//   stage1_joint   (kernel 1, every joint) input check, braking solution, time-optimal solve;
//                  the joint's record is written as if it were the slowest one of its problem
//   stage3_classify(kernel 2, every joint that has to be stretched to the slowest one's time)
//                  first cruise-speed candidate (cc:378-396) and the cc:119 test at that speed:
//                  class A (normal profile) or class B (has to slow down first / candidate
//                  unusable)
//   scale_attempt  the nested phase solve + acceptance test of one candidate (cc:398-405); the
//                  class-A warps compile only the normal branch, the class-B warps only the
//                  modified one
//   scale_attempt2 second candidate (cc:408-436) for what is left (class C), same nested solve
//
// What travels between the steps is five doubles and one int per joint (ScaleItem): the nested
// solve of a joint that is not on the brake-only exit depends on the start state only through
// the mapped velocity and acceleration, the direction and the distance (Prologue).
#pragma once

#include "ltp_math.cuh"

namespace ltp {

// per-joint flags handed from kernel 1 to kernel 2
enum : unsigned char { JF_FAIL = 1, JF_DEFER = 2, JF_BRAKE_ONLY = 4 };

struct Stage1Out {
  double t_opt[7];
  double dir;
  unsigned char mod, opt_case, flags;
};

// cc:14-30 for one joint (checkInputs cc:68-77, optSwitchTimes at v_max without the quartic
// tail). JF_DEFER: the joint needs the quartic tail, its problem goes to the every-branch kernel.
LTP_HD void stage1_joint(const JointLimits& L, double Ts, double q_goal, double q_0, double v_0, double a_0,
                         Stage1Out& o) {
  const bool in_ok = check_joint_input(L, q_0, v_0, a_0);
  const Prologue pro = ost_prologue(L, Ts, q_goal, q_0, v_0, a_0);
  zero7(o.t_opt);
  o.mod = 0;
  o.opt_case = 255;
  const int st = ost_body_t<false>(L, Ts, pro, q_goal, q_0, L.v_max, o.t_opt, o.mod, o.opt_case);
  o.dir = pro.dir;
  o.flags = (unsigned char)((!(in_ok && st != OST_FAIL) ? JF_FAIL : 0) | (st == OST_DEFER ? JF_DEFER : 0) |
                            (pro.brake_only ? JF_BRAKE_ONLY : 0));
}

// What a joint carries from the classification to the warps that finish it.
struct ScaleItem {
  double t_req;  // required end time (cc:31-39)
  double V;      // first cruise-speed candidate (cc:378-396)
  double v0m, a0m, dist;  // Prologue: start state mapped to the positive direction, distance
  int meta;      // bit 8: dir < 0; the rest is the caller's (origin of the item)
};

LTP_HD Prologue item_prologue(const ScaleItem& it) {
  Prologue P;
  P.v0m = it.v0m;
  P.a0m = it.a0m;
  P.dir = (it.meta & 0x100) ? -1.0 : 1.0;
  P.dist = it.dist;
  P.b0 = P.b1 = P.b2 = 0.0;
  P.brake_only = false;
  return P;
}

struct JointResult {
  double t[7];
  double v_drive;
  unsigned char mod, ts_case, final_case;
};

LTP_HD double max7(const double* t) {
  double m = t[0];
#pragma unroll
  for (int k = 1; k < 7; ++k)
    if (m < t[k]) m = t[k];
  return m;
}

// cc:718 for one joint, -1 when a switching time is not finite / not representable
LTP_HD int joint_sample_count(const double* t, double Ts) {
  bool fin = true;
#pragma unroll
  for (int k = 0; k < 7; ++k) fin &= (bool)isfinite(t[k]);
  return (fin && t[6] / Ts <= 2.0e9) ? samples_for(t[6], Ts) : -1;
}

enum { S3_SETTLED = 1, S3_QUEUE_A = 2, S3_QUEUE_B = 3, S3_DEFER = 4 };

// A joint that is not the slowest one of its problem and not on the brake-only exit: the first
// candidate and the class. dir is what stage 1 stored for the joint (sign of the distance to go,
// cc:110); mapping the start state with it gives the operands of both cc:372-375 and cc:110-113.
LTP_HD int stage3_classify(const JointLimits& L, double q_goal, double q_0, double v_0, double a_0, double dir,
                           double t_req, ScaleItem& item) {
  const TsInput I = make_ts_input(q_goal, q_0, v_0, a_0, dir, t_req);
  const double V = ts_candidate1(L, I);
  item.t_req = t_req;
  item.V = V;
  item.v0m = I.v_0;
  item.a0m = I.a_0;
  item.dist = (q_goal - q_0) * dir;
  item.meta = dir < 0 ? 0x100 : 0;
  const bool v_ok = !isnan(V) && V > 0;
  return (v_ok && !ost_needs_mod_profile(L, I.v_0, I.a_0, V)) ? S3_QUEUE_A : S3_QUEUE_B;
}

// A joint on the brake-only exit (cc:102-107; ~0.2 % of random Franka joints): settled on the
// spot. pro is the joint's recomputed prologue. S3_SETTLED: R is final. S3_DEFER: every-branch
// kernel.
LTP_HD int stage3_brake_only(const JointLimits& L, double Ts, const Prologue& pro, double q_goal, double q_0,
                             double v_0, double a_0, double t_req, const double* t_opt, unsigned char opt_case,
                             JointResult& R) {
  R.v_drive = L.v_max;
  R.mod = 0;
  if (ts_brake_only_rejects(pro, t_req)) {  // cc:641-644 without evaluating a candidate, then cc:50-55
#pragma unroll
    for (int k = 0; k < 7; ++k) R.t[k] = t_opt[k];
    R.ts_case = 9;
    R.final_case = opt_case;
    return S3_SETTLED;
  }
  // the nested solve takes the cc:102-107 exit at any cruise speed and is accepted (the
  // acceptance test is what ts_brake_only_rejects evaluated); without a usable first candidate
  // the search goes on to the others
  const TsInput I = make_ts_input(q_goal, q_0, v_0, a_0, pro.dir, t_req);
  const double V = ts_candidate1(L, I);
  if (!(!isnan(V) && V > 0)) return S3_DEFER;
  const double T[7] = {pro.b0, pro.b1, pro.b2, 0, 0, 0, 0};
  cumsum7(T, R.t);
  R.v_drive = V;
  R.ts_case = 1;
  R.final_case = CASE_BRAKE_ONLY;
  if (max7(R.t) <= 0.0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) R.t[k] = t_opt[k];
  }
  return S3_SETTLED;
}

enum { SA_REJECT = 0, SA_ACCEPT = 1, SA_DEFER = 2 };

// Nested phase solve at cruise speed V + the acceptance test (cc:398-405 and its repeats).
// SA_DEFER: the joint needs the quartic tail, or was accepted with no positive time (the
// cc:50-55 fallback needs the time-optimal times, which do not travel): every-branch kernel.
template <int MODE>
LTP_HD int scale_attempt(const JointLimits& L, double Ts, const ScaleItem& it, double V, JointResult& R) {
  const Prologue P = item_prologue(it);
  zero7(R.t);
  R.mod = 0;
  R.final_case = 255;
  R.v_drive = V;
  const int st = ost_body_t<false, MODE>(L, Ts, P, 0.0, 0.0, V, R.t, R.mod, R.final_case);
  if (st == OST_DEFER) return SA_DEFER;
  if (st == OST_OK && it.t_req - R.t[6] < kTol && it.t_req - R.t[6] > -kTol / 10)
    return max7(R.t) <= 0.0 ? SA_DEFER : SA_ACCEPT;
  return SA_REJECT;
}

// class B: the first candidate again, for the joints that have to slow down first (or whose
// candidate is unusable, in which case the attempt is skipped like in cc:398)
LTP_HD int scale_attempt1_class_b(const JointLimits& L, double Ts, const ScaleItem& it, JointResult& R) {
  const bool v_ok = !isnan(it.V) && it.V > 0;
  if (!v_ok) return SA_REJECT;
  const int r = scale_attempt<OST_MODIFIED>(L, Ts, it, it.V, R);
  R.ts_case = 1;
  return r;
}

LTP_HD int scale_attempt1_class_a(const JointLimits& L, double Ts, const ScaleItem& it, JointResult& R) {
  const int r = scale_attempt<OST_NORMAL>(L, Ts, it, it.V, R);
  R.ts_case = 1;
  return r;
}

// the second candidate (cc:408-446). A joint it does not settle needs the root-solver
// candidates 3..8: SA_DEFER.
LTP_HD int scale_attempt2(const JointLimits& L, double Ts, const ScaleItem& it, JointResult& R) {
  // dir * (q_0 - q_goal) == -((q_goal - q_0) * dir): negation and the product with +-1 are exact
  const double V = ts_candidate2_core(L, it.a0m, it.v0m, it.t_req, -it.dist);
  R.v_drive = V;
  if (!isnan(V) && V > 0) {
    const int r = scale_attempt<OST_ANY>(L, Ts, it, V, R);
    R.ts_case = 2;
    if (r != SA_REJECT) return r;
  }
  return SA_DEFER;
}

}  // namespace ltp
