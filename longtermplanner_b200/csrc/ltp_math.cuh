// Per-joint FP64 math of the planning hot path, written for one CUDA thread per
// (problem, joint). Everything here is a __device__ function used by the kernels in
// ltp_kernels.cu; the same source can be compiled for the host (LTP_HD empty) by
// tests/host_shadow.cc so that the device arithmetic can be checked against the oracle
// in a container without a GPU. That host build is a test artefact only: the product
// library never links or dispatches to it.
//
// What is computed (reference = yannickBurkhardt/LongTermPlanner, src/long_term_planner.cc
// = "cc", include/long_term_planner/roots.h = "roots.h"):
//   brake_profile      fastest stop of a velocity                       cc:650-701
//   ost_prologue/body  seven switching times at a cruise speed V        cc:82-353
//   ts_candidate       the eight cruise-speed candidates of the search  cc:378-629
//   time_scaling       ordered search, first accepted candidate wins    cc:358-645
//   smallest_root      companion-matrix eigenvalues, smallest real > 1e-7   roots.h:22-50
//   RowSampler         jerk impulses + forward-Euler q/v/a/j recurrence cc:729-831
//
// Numerical contract. Results must match the reference binary (g++ -O2, no FMA
// contraction, glibc libm): every expression keeps the reference's operand order; the
// translation unit is compiled with -fmad=false so nvcc never fuses a*b+c; FP64 division
// and sqrt are IEEE-correct on the device. The reference's pow(x,3|4|6) calls go to
// glibc, whose result is the correctly rounded power in >99.9% of cases (measured); here
// they are computed as error-free products (explicit fma) rounded once, i.e. the
// correctly rounded power. pow(x,2) is a plain product on both sides, pow(x,0.5) is sqrt.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define LTP_HD __host__ __device__ __forceinline__
#define LTP_HD_NOINLINE __host__ __device__ __noinline__
#else
#define LTP_HD inline
#define LTP_HD_NOINLINE
#endif

namespace ltp {

// case byte, identical to oracle/ltp_oracle.h (the reference itself emits no case id)
enum : unsigned char {
  CASE_BRAKE_ONLY = 0, CASE_NOP4 = 5, CASE_Q1 = 6, CASE_Q1_P2 = 7, CASE_Q2 = 8,
  CASE_FAIL_UNTOUCHED = 13, CASE_DEGENERATE = 14, CASE_FAIL = 15,
  F_MOD = 0x10, F_BOTH = 0x20, F_NOP2 = 0x40, F_NOP6 = 0x80
};

struct JointLimits {
  double q_min, q_max, v_max, a_max, j_max;
  // derived once on the host (derive_limits): correctly rounded reciprocals for div_by()
  double r_a, r_j, a_over_j, r_v;
  // for the second cruise-speed candidate (ts_candidate2): reciprocals of J^2, J^3, 6 J^3, A J
  double r_j2, r_j3, r_6j3, r_aj;
  // Powers of A/J. The jerk phases 3, 5 and 7 last A/J unless a fix-up shortens them (cc:129-143,
  // 153-165, 682-687), so pow(T, 3) and pow(T, 4) of those durations are per-joint constants for
  // most joints: pow3(a_over_j), pow4(a_over_j), the same bits as evaluated per thread.
  double aoj3, aoj4;
  // The braking half of the time-optimal solve (cc:147-165 and the second half of cc:168-190 at
  // V = v_max) depends on the limits only: T5 and part2 as the body computes them, part2v = NaN
  // when that half needs the cc:153 fix-up (v_max / a_max < a_max / j_max) and is evaluated per thread.
  double t5v, part2v;
  // Limit-only factors of the first cruise-speed candidate (cc:378-396), each the prefix of a
  // product as C++ evaluates it left to right: A J, A^2 / 2, 36 A^2 J^2, 72 A^3 J, 144 A,
  // 72 A J^2, A^3, 36 A^4, 36 J^2
  double c_aj, c_a2h, c_36a2j2, c_72a3j, c_144a, c_72aj2, c_a3, c_36a4, c_36j2;
  // Terms of the radicand of the solve without a cruise phase (cc:202-223) that hold only the
  // durations T2 = T4 = T6 = A/J (no fix-up taken): (J^2 T^4)/4, (J^2 T^2 T^2)/2, (J^2 T^4)/2,
  // (2 J A T^3)/3, 2 J A T^2 T, 2 A^2 T^2
  double rk_q, rk_m, rk_h, rk_3, rk_7, rk_8;
};

#ifndef LTP_OST_MERGE
#define LTP_OST_MERGE 1
#endif
#ifndef LTP_OST_CONST
#define LTP_OST_CONST 2
#endif

constexpr double kEps = 4e-3;     // cc:96
constexpr double kTol = 0.1;      // cc:370
constexpr double kDblMin = 2.2250738585072014e-308;
constexpr double kDblEps = 2.220446049250313e-16;

LTP_HD double sq(double x) { return x * x; }

// correctly rounded x^3, x^4, x^6 via double-double products (see file header)
LTP_HD double pow3(double x) {
  double p = x * x, e = fma(x, x, -p);
  double h = p * x, l = fma(p, x, -h) + e * x;
  return h + l;
}
LTP_HD double pow4(double x) {
  double p = x * x, e = fma(x, x, -p);
  double h = p * p, l = fma(p, p, -h) + 2.0 * (p * e);
  return h + l;
}
LTP_HD double pow6(double x) {
  double p = x * x, e = fma(x, x, -p);
  double h = p * x, l = fma(p, x, -h) + e * x;  // x^3 = h + l
  double s = h + l, sl = (h - s) + l;
  double H = s * s, L = fma(s, s, -H) + 2.0 * (s * sl);
  return H + L;
}

// x / d for a divisor whose correctly rounded reciprocal rd = RN(1/d) is at hand: quotient
// estimate, exact remainder (fma), one correction -- three FP64 operations instead of the
// dozen of a division, and the SAME bits as the IEEE division x / d (Markstein's theorem for
// rd = RN(1/d); checked here on 1.2e9 random numerators against the divisors the limit sets
// produce). Outside the range where that argument holds (zero, subnormal, huge, inf, NaN
// results) the plain division is evaluated, so special values behave exactly as before.
LTP_HD_NOINLINE double div_slow(double x, double d) { return x / d; }

LTP_HD double div_by(double x, double d, double rd) {
  const double q = x * rd;
  const double r = fma(-q, d, x);
  const double f = fma(r, rd, q);
  // biased exponent of the result in [64, 1982] (|f| in ~[1e-289, 1e289]): integer test on
  // the exponent field of the high word, in place (mask, subtract, one unsigned compare)
#ifdef __CUDA_ARCH__
  const unsigned hi = (unsigned)__double2hiint(f);
#else
  unsigned long long bits;
  memcpy(&bits, &f, 8);
  const unsigned hi = (unsigned)(bits >> 32);
#endif
  if (((hi & 0x7ff00000u) - (64u << 20)) > (1918u << 20)) return div_slow(x, d);
  return f;
}
LTP_HD double div3(double x) { return div_by(x, 3.0, 1.0 / 3.0); }
LTP_HD double div12(double x) { return div_by(x, 12.0, 1.0 / 12.0); }

// Two ways to run the range test of div_by, as a policy object that the closed-form functions
// below take as their last argument (the signatures without it use DivChecked):
//   DivChecked   tests every quotient where it is computed and takes the plain division at once
//   DivDeferred  only records that some quotient left the range; the caller looks at `bad` once
//                the function is through and, if set, discards everything and calls the function
//                again with DivChecked. Same results by construction -- a run with no quotient out
//                of range executes the same operations either way -- but the three FP64
//                operations of a division are then followed by two integer instructions instead
//                of a divergent call (six), and straight-line code is not cut into a basic block
//                per division. A zero numerator is not out of range here (see below).
struct DivChecked {
  static constexpr bool kWideLimits = false;  // see DivDeferredWide
  LTP_HD double by(double x, double d, double rd) const { return div_by(x, d, rd); }
  LTP_HD double by3(double x) const { return div3(x); }
  LTP_HD double by12(double x) const { return div12(x); }
  // x / (2 d): half the quotient (exact) unless the quotient is outside the window, where the
  // halving could round a second time -- then the plain division by 2 d
  LTP_HD double half_by(double x, double d, double rd) const {
    const double q = x * rd;
    const double r = fma(-q, d, x);
    const double f = fma(r, rd, q);
#ifdef __CUDA_ARCH__
    const unsigned hi = (unsigned)__double2hiint(f);
#else
    unsigned long long bits;
    memcpy(&bits, &f, 8);
    const unsigned hi = (unsigned)(bits >> 32);
#endif
    if (((hi & 0x7ff00000u) - (64u << 20)) > (1918u << 20)) return div_slow(x, 2.0 * d);
    return 0.5 * f;
  }
};
struct DivDeferred {
  static constexpr bool kWideLimits = false;
  bool bad = false;
  LTP_HD double by(double x, double d, double rd) {
    const double q = x * rd;
    // The remainder is exact, so its rounding mode only decides the sign of an exact zero; rounded
    // down, a zero remainder is -0 and the correction step leaves the sign of q alone. That makes
    // the three operations return x / d = +-0 with the right sign for a zero numerator (rounded
    // to nearest, -0 / d came out as +0), so a zero numerator needs no other path.
#ifdef __CUDA_ARCH__
    const double r = __fma_rd(-q, d, x);
    const double f = fma(r, rd, q);
    const unsigned hi = (unsigned)__double2hiint(f);
#else
    double r = fma(-q, d, x);
    if (r == 0.0) r = -0.0;  // what round-down gives: the two addends are never both +0
    const double f = fma(r, rd, q);
    unsigned long long bits;
    memcpy(&bits, &f, 8);
    const unsigned hi = (unsigned)(bits >> 32);
#endif
    // the same window as div_by, on the high word shifted left by one (drops the sign); a zero
    // numerator is exempt
    bad |= (((hi << 1) - (64u << 21)) > (1918u << 21)) & (x != 0.0);
    return f;
  }
  LTP_HD double by3(double x) { return by(x, 3.0, 1.0 / 3.0); }
  LTP_HD double by12(double x) { return by(x, 12.0, 1.0 / 12.0); }
  LTP_HD double half_by(double x, double d, double rd) { return 0.5 * by(x, d, rd); }
};

// DivDeferred for a caller whose joints' limits stay resident in the constant cache when every
// field is used: the closed-form functions then also read the limit-only factors c_* and rk_* of
// JointLimits instead of forming them per thread (same values). With the 256 bytes per joint this
// makes, 7 joints fit and 12 do not: the 12-joint kernel went from 0.83 to 1.20 ms with them, the
// 7-joint one from 0.486 to 0.471 ms (profiles/r02_ab_limit_constants.log).
struct DivDeferredWide : DivDeferred {
  static constexpr bool kWideLimits = LTP_OST_CONST >= 2;
};

LTP_HD void derive_limits(JointLimits& L) {
  L.r_a = 1.0 / L.a_max;
  L.r_j = 1.0 / L.j_max;
  L.a_over_j = L.a_max / L.j_max;
  L.r_v = 1.0 / L.v_max;
  L.r_j2 = 1.0 / sq(L.j_max);
  L.r_j3 = 1.0 / pow3(L.j_max);
  L.r_6j3 = 1.0 / (6 * pow3(L.j_max));
  L.r_aj = 1.0 / (L.a_max * L.j_max);
  L.aoj3 = pow3(L.a_over_j);
  L.aoj4 = pow4(L.a_over_j);
  {  // cc:147-151 and the part2 of cc:168-190 at V = v_max, the expressions of ost_body_dv
    const double A = L.a_max, J = L.j_max, T4 = L.a_over_j, T6 = T4;
    const double T5 = L.v_max / A - 1.0 / 2.0 * (T4 + T6);
    L.t5v = T5;
    L.part2v = J * (1.0 / 6.0 * pow3(T6) + 1.0 / 2.0 * sq(T6) * (T5 + T4) - 1.0 / 6.0 * pow3(T4) +
                    1.0 / 2.0 * T6 * sq(T4)) +
               A * (1.0 / 2.0 * sq(T5) + T5 * T4);
    if (!(T5 >= -kEps)) L.part2v = NAN;  // cc:153: the fix-up path, evaluated per thread
  }
  {
    const double A = L.a_max, J = L.j_max, T = L.a_over_j;
    L.c_aj = A * J;
    L.c_a2h = sq(A) / 2;
    L.c_36a2j2 = 36 * sq(A) * sq(J);
    L.c_72a3j = 72.0 * pow3(A) * J;
    L.c_144a = 144 * A;
    L.c_72aj2 = 72.0 * A * sq(J);
    L.c_a3 = pow3(A);
    L.c_36a4 = 36 * pow4(A);
    L.c_36j2 = 36 * sq(J);
    L.rk_q = (sq(J) * L.aoj4) / 4;
    L.rk_m = (sq(J) * sq(T) * sq(T)) / 2;
    L.rk_h = (sq(J) * L.aoj4) / 2;
    L.rk_3 = (2.0 * J * A * L.aoj3) / 3;
    L.rk_7 = 2.0 * J * A * sq(T) * T;
    L.rk_8 = 2.0 * sq(A) * sq(T);
  }
}

// h:54-56: (double)((0 < x) - (x < 0)), written as two selects (no int -> double conversion);
// +1, -1, and +0.0 for zeros and NaN, like the reference's expression
LTP_HD double sgn(double x) { return x > 0.0 ? 1.0 : (x < 0.0 ? -1.0 : 0.0); }
// -sign(x) of cc:660-667: the reference negates the INT, so a zero stays +0.0
LTP_HD double neg_sgn(double x) { return x > 0.0 ? -1.0 : (x < 0.0 ? 1.0 : 0.0); }

// ------------------------------------------------------------------------------------
// cc:650-701. Returns the signed stop displacement; T[0..2] are the three durations.
// ------------------------------------------------------------------------------------
template <class DIV>
LTP_HD double brake_profile(const JointLimits& L, double Ts, double v_0, double a_0,
                            double& T0, double& T1, double& T2, double& dir, DIV& dv) {
  const double A = L.a_max, J = L.j_max;
  // cc:658-664 without the three-way branch: both tests are evaluated, one select picks the
  // operand whose sign decides (same values; a warp does not split three ways here)
  const bool same_sign = v_0 * a_0 > 0;
  const bool fast_enough = fabs(v_0) > dv.by(1.0 / 2.0 * sq(a_0), J, L.r_j);
  dir = neg_sgn((same_sign | fast_enough) ? v_0 : a_0);
  if (dir < 0) {
    a_0 = -a_0;
    v_0 = -v_0;
  }
  T0 = dv.by(A - a_0, J, L.r_j);
  T2 = L.a_over_j;
  double p3_2 = L.aoj3;  // pow3(T2): the per-joint constant unless the fix-up below changes T2
  T1 = dv.by(-v_0 - 1.0 / 2.0 * T0 * a_0, A, L.r_a) - 1.0 / 2.0 * (T0 + T2);
  if (T1 < -Ts) {
    T0 = -a_0 / J + sqrt(sq(a_0) / (2 * sq(J)) - v_0 / J);
    T2 = T0 + a_0 / J;
    T1 = 0;
    p3_2 = pow3(T2);
  }
#if !LTP_OST_CONST
  p3_2 = pow3(T2);
#endif
  double s = v_0 * (T0 + T1 + T2) +
             a_0 * (1.0 / 2.0 * sq(T0) + T0 * (T1 + T2) + 1.0 / 2.0 * sq(T2)) +
             J * (1.0 / 6.0 * pow3(T0) + 1.0 / 2.0 * sq(T0) * (T1 + T2) - 1.0 / 6.0 * p3_2 +
                  1.0 / 2.0 * T0 * sq(T2)) +
             A * (1.0 / 2.0 * sq(T1) + T1 * T2);
  return dir * s;
}

LTP_HD double brake_profile(const JointLimits& L, double Ts, double v_0, double a_0,
                            double& T0, double& T1, double& T2, double& dir) {
  DivChecked dv;
  return brake_profile(L, Ts, v_0, a_0, T0, T1, T2, dir, dv);
}

// ------------------------------------------------------------------------------------
// roots.h:22-50 over the eigenvalue algorithm of Eigen 3.4's EigenSolver (real Schur form
// of the companion matrix by Francis double-shift QR, no balancing).
// p: coefficients, highest power first, N+1 of them. Returns the smallest real root that
// is > 1e-7, +inf if there is none (also when the iteration fails or p is not finite).
// ------------------------------------------------------------------------------------
template <int N>
struct SmallSchur {
  double a[N][N];

  LTP_HD static void householder3(double v0, double v1, double v2, double& e0, double& e1,
                                  double& tau, double& beta) {
    double tail = v1 * v1 + v2 * v2;
    if (tail <= kDblMin) {
      tau = 0.0; beta = v0; e0 = 0.0; e1 = 0.0;
    } else {
      double b = sqrt(v0 * v0 + tail);
      if (v0 >= 0.0) b = -b;
      e0 = v1 / (v0 - b);
      e1 = v2 / (v0 - b);
      tau = (b - v0) / b;
      beta = b;
    }
  }
  LTP_HD static void householder2(double v0, double v1, double& e0, double& tau, double& beta) {
    double tail = v1 * v1;
    if (tail <= kDblMin) {
      tau = 0.0; beta = v0; e0 = 0.0;
    } else {
      double b = sqrt(v0 * v0 + tail);
      if (v0 >= 0.0) b = -b;
      e0 = v1 / (v0 - b);
      tau = (b - v0) / b;
      beta = b;
    }
  }

  // returns true on convergence; a holds the quasi-triangular factor afterwards
  LTP_HD_NOINLINE bool run() {
    double scale = 0.0;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) {
        double v = fabs(a[i][j]);
        if (v > scale) scale = v;
      }
    if (scale < kDblMin) {
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) a[i][j] = 0.0;
      return true;
    }
    // Hessenberg reduction of C/scale is the identity for a companion matrix
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) a[i][j] = a[i][j] / scale;
    const int max_iters = 40 * N;
    int iu = N - 1, iter = 0, total = 0;
    double exshift = 0.0, norm = 0.0;
    for (int j = 0; j < N; ++j) {
      int lim = (j + 2 < N) ? j + 2 : N;
      double cs = 0.0;
      for (int i = 0; i < lim; ++i) cs += fabs(a[i][j]);
      norm += cs;
    }
    double caz = norm * (kDblEps * kDblEps);
    if (!(caz > kDblMin)) caz = kDblMin;
    if (norm != 0.0) {
      while (iu >= 0) {
        int il = iu;
        while (il > 0) {
          double s = fabs(a[il - 1][il - 1]) + fabs(a[il][il]);
          s = s * kDblEps;
          if (!(s > caz)) s = caz;
          if (fabs(a[il][il - 1]) <= s) break;
          il--;
        }
        if (il == iu) {
          a[iu][iu] = a[iu][iu] + exshift;
          if (iu > 0) a[iu][iu - 1] = 0.0;
          iu--;
          iter = 0;
        } else if (il == iu - 1) {
          double p = 0.5 * (a[iu - 1][iu - 1] - a[iu][iu]);
          double q = p * p + a[iu][iu - 1] * a[iu - 1][iu];
          a[iu][iu] += exshift;
          a[iu - 1][iu - 1] += exshift;
          if (q >= 0.0) {
            double z = sqrt(fabs(q));
            double gp = (p >= 0.0) ? p + z : p - z;
            double gq = a[iu][iu - 1];
            double c, s;
            if (gq == 0.0) {
              c = gp < 0.0 ? -1.0 : 1.0;
              s = 0.0;
            } else if (gp == 0.0) {
              c = 0.0;
              s = gq < 0.0 ? 1.0 : -1.0;
            } else if (fabs(gp) > fabs(gq)) {
              double t = gq / gp;
              double u = sqrt(1.0 + t * t);
              if (gp < 0.0) u = -u;
              c = 1.0 / u;
              s = -t * c;
            } else {
              double t = gp / gq;
              double u = sqrt(1.0 + t * t);
              if (gq < 0.0) u = -u;
              s = -1.0 / u;
              c = -t * s;
            }
            for (int col = iu - 1; col < N; ++col) {
              double x = a[iu - 1][col], y = a[iu][col];
              a[iu - 1][col] = c * x - s * y;
              a[iu][col] = s * x + c * y;
            }
            for (int row = 0; row <= iu; ++row) {
              double x = a[row][iu - 1], y = a[row][iu];
              a[row][iu - 1] = c * x - s * y;
              a[row][iu] = s * x + c * y;
            }
            a[iu][iu - 1] = 0.0;
          }
          if (iu > 1) a[iu - 1][iu - 2] = 0.0;
          iu -= 2;
          iter = 0;
        } else {
          double sh0 = a[iu][iu], sh1 = a[iu - 1][iu - 1], sh2 = a[iu][iu - 1] * a[iu - 1][iu];
          if (iter == 10) {
            exshift += sh0;
            for (int i = 0; i <= iu; ++i) a[i][i] -= sh0;
            double s = fabs(a[iu][iu - 1]) + fabs(a[iu - 1][iu - 2]);
            sh0 = 0.75 * s;
            sh1 = 0.75 * s;
            sh2 = -0.4375 * s * s;
          }
          if (iter == 30) {
            double s = (sh1 - sh0) / 2.0;
            s = s * s + sh2;
            if (s > 0.0) {
              s = sqrt(s);
              if (sh1 < sh0) s = -s;
              s = s + (sh1 - sh0) / 2.0;
              s = sh0 - sh2 / s;
              exshift += s;
              for (int i = 0; i <= iu; ++i) a[i][i] -= s;
              sh0 = sh1 = sh2 = 0.964;
            }
          }
          iter++;
          total++;
          if (total > max_iters) break;
          int im;
          double v0 = 0.0, v1 = 0.0, v2 = 0.0;
          for (im = iu - 2; im >= il; --im) {
            double tmm = a[im][im];
            double r = sh0 - tmm;
            double s = sh1 - tmm;
            v0 = (r * s - sh2) / a[im + 1][im] + a[im][im + 1];
            v1 = a[im + 1][im + 1] - tmm - r - s;
            v2 = a[im + 2][im + 1];
            if (im == il) break;
            double lhs = a[im][im - 1] * (fabs(v1) + fabs(v2));
            double rhs = v0 * (fabs(a[im - 1][im - 1]) + fabs(tmm) + fabs(a[im + 1][im + 1]));
            if (fabs(lhs) < kDblEps * rhs) break;
          }
          for (int k = im; k <= iu - 2; ++k) {
            const bool first = (k == im);
            double w0, w1, w2, e0, e1, tau, beta;
            if (first) {
              w0 = v0; w1 = v1; w2 = v2;
            } else {
              w0 = a[k][k - 1]; w1 = a[k + 1][k - 1]; w2 = a[k + 2][k - 1];
            }
            householder3(w0, w1, w2, e0, e1, tau, beta);
            if (beta != 0.0) {
              if (first && k > il)
                a[k][k - 1] = -a[k][k - 1];
              else if (!first)
                a[k][k - 1] = beta;
              if (tau != 0.0) {
                for (int c = k; c < N; ++c) {  // rows k..k+2 from the left
                  double tmp = e0 * a[k + 1][c];
                  tmp += e1 * a[k + 2][c];
                  tmp += a[k][c];
                  a[k][c] -= tau * tmp;
                  a[k + 1][c] -= (tau * e0) * tmp;
                  a[k + 2][c] -= (tau * e1) * tmp;
                }
                const int r1 = (iu < k + 3) ? iu : k + 3;
                for (int r = 0; r <= r1; ++r) {  // columns k..k+2 from the right
                  double tmp = a[r][k + 1] * e0;
                  tmp += a[r][k + 2] * e1;
                  tmp += a[r][k];
                  a[r][k] -= tau * tmp;
                  a[r][k + 1] -= (tau * tmp) * e0;
                  a[r][k + 2] -= (tau * tmp) * e1;
                }
              }
            }
          }
          {
            double e0, tau, beta;
            householder2(a[iu - 1][iu - 2], a[iu][iu - 2], e0, tau, beta);
            if (beta != 0.0) {
              a[iu - 1][iu - 2] = beta;
              if (tau != 0.0) {
                for (int c = iu - 1; c < N; ++c) {
                  double tmp = e0 * a[iu][c];
                  tmp += a[iu - 1][c];
                  a[iu - 1][c] -= tau * tmp;
                  a[iu][c] -= (tau * e0) * tmp;
                }
                for (int r = 0; r <= iu; ++r) {
                  double tmp = a[r][iu] * e0;
                  tmp += a[r][iu - 1];
                  a[r][iu - 1] -= tau * tmp;
                  a[r][iu] -= (tau * tmp) * e0;
                }
              }
            }
          }
          for (int i = im + 2; i <= iu; ++i) {
            a[i][i - 2] = 0.0;
            if (i > im + 2) a[i][i - 3] = 0.0;
          }
        }
      }
    }
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j) a[i][j] = a[i][j] * scale;
    return total <= max_iters;
  }
};

template <int N>
LTP_HD_NOINLINE double smallest_root(const double* p) {
  SmallSchur<N> S;
  // roots.h:28-31
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) S.a[i][j] = 0.0;
  for (int i = 0; i + 1 < N; ++i) S.a[i + 1][i] = 1.0;
  for (int i = 0; i < N; ++i) S.a[i][N - 1] = (-1.0 * p[N - i]) / p[0];
  double best = INFINITY;
  if (!S.run()) return best;
  int i = 0;
  while (i < N) {
    if (i == N - 1 || S.a[i + 1][i] == 0.0) {
      double r = S.a[i][i];
      if (!isfinite(r)) break;
      if (r > 1e-7) best = fmin(best, r);  // roots.h:47 (imag == 0 exactly for a 1x1 block)
      ++i;
    } else {
      // complex pair unless z == 0 exactly; Eigen reports (re, +-z) and roots.h:47 tests
      // imag() == 0, so a 2x2 block whose discriminant rounds to 0 still counts as real
      double pp = 0.5 * (S.a[i][i] - S.a[i + 1][i + 1]);
      double t0 = S.a[i + 1][i], t1 = S.a[i][i + 1];
      double mx = fabs(pp);
      if (fabs(t0) > mx) mx = fabs(t0);
      if (fabs(t1) > mx) mx = fabs(t1);
      t0 /= mx;
      t1 /= mx;
      double p0 = pp / mx;
      double z = mx * sqrt(fabs(p0 * p0 + t0 * t1));
      double rr = S.a[i + 1][i + 1] + pp;
      if (!(isfinite(rr) && isfinite(z))) break;
      if (z == 0 && rr > 1e-7) best = fmin(best, rr);
      i += 2;
    }
  }
  return best;
}

// ------------------------------------------------------------------------------------
// cc:82-353 split into the part that does not depend on the cruise speed (prologue:
// braking solution, goal direction, BRAKE_ONLY exit) and the part that does (body).
// timeScaling re-enters optSwitchTimes up to eight times with the same start state
// (cc:400 restores the original signs), so the prologue is evaluated once per joint.
// ------------------------------------------------------------------------------------
struct Prologue {
  double v0m, a0m;      // start state mapped to the positive direction (cc:110-113)
  double dir;           // braking direction on the BRAKE_ONLY exit, else sign(q_diff)
  double dist;          // (q_goal - q_0) * dir                        (cc:190)
  double b0, b1, b2;    // braking durations (cc:100)
  bool brake_only;      // cc:102
};

template <class DIV>
LTP_HD Prologue ost_prologue(const JointLimits& L, double Ts, double q_goal, double q_0,
                             double v_0, double a_0, DIV& dv) {
  Prologue P;
  double dirb;
  double q_stop = brake_profile(L, Ts, v_0, a_0, P.b0, P.b1, P.b2, dirb, dv);
  double q_diff = q_goal - (q_0 + q_stop);
  P.brake_only = fabs(q_diff) < kEps;
  if (P.brake_only) {
    P.dir = dirb;
    P.v0m = v_0;
    P.a0m = a_0;
    P.dist = 0.0;
    return P;
  }
  P.dir = sgn(q_diff);
  if (P.dir < 0) {
    v_0 = -v_0;
    a_0 = -a_0;
  }
  P.v0m = v_0;
  P.a0m = a_0;
  P.dist = (q_goal - q_0) * P.dir;
  return P;
}

LTP_HD Prologue ost_prologue(const JointLimits& L, double Ts, double q_goal, double q_0,
                             double v_0, double a_0) {
  DivChecked dv;
  return ost_prologue(L, Ts, q_goal, q_0, v_0, a_0, dv);
}

LTP_HD void cumsum7(const double* T, double* t) {
  double acc = T[0];
  t[0] = acc;
#pragma unroll
  for (int i = 1; i < 7; ++i) {
    acc = acc + T[i];
    t[i] = acc;
  }
}

// any of T[0..6] < lim (a NaN is not). On the device this is spelled as a chain of seven
// compare-and-accumulate instructions: left to itself the compiler turns the OR of the seven
// compares into a minimum reduction, and an FP64 minimum is eight instructions here.
LTP_HD bool any_below7(const double* T, double lim) {
#ifdef __CUDA_ARCH__
  unsigned r;
  asm("{\n\t.reg .pred p;\n\t"
      "setp.lt.f64 p, %1, %8;\n\t"
      "setp.lt.or.f64 p, %2, %8, p;\n\t"
      "setp.lt.or.f64 p, %3, %8, p;\n\t"
      "setp.lt.or.f64 p, %4, %8, p;\n\t"
      "setp.lt.or.f64 p, %5, %8, p;\n\t"
      "setp.lt.or.f64 p, %6, %8, p;\n\t"
      "setp.lt.or.f64 p, %7, %8, p;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "d"(T[0]), "d"(T[1]), "d"(T[2]), "d"(T[3]), "d"(T[4]), "d"(T[5]), "d"(T[6]), "d"(lim));
  return r != 0;
#else
  bool below = false;
  for (int i = 0; i < 7; ++i) below |= T[i] < lim;
  return below;
#endif
}

LTP_HD bool any_below2(double a, double b, double lim) {
#ifdef __CUDA_ARCH__
  unsigned r;
  asm("{\n\t.reg .pred p;\n\t"
      "setp.lt.f64 p, %1, %3;\n\t"
      "setp.lt.or.f64 p, %2, %3, p;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(r)
      : "d"(a), "d"(b), "d"(lim));
  return r != 0;
#else
  return a < lim || b < lim;
#endif
}

LTP_HD void zero7(double* t) {
#pragma unroll
  for (int i = 0; i < 7; ++i) t[i] = 0.0;
}

// The rare tail of cc:245-337 (quartic root solves). Kept out of line: it carries the
// 4x4 Schur workspace and is taken by well under 1% of joints for realistic limits.
LTP_HD_NOINLINE unsigned char ost_quartic_tail(const JointLimits& L, const Prologue& P,
                                               double q_goal, double q_0, double* T,
                                               unsigned char& flags) {
  const double A = L.a_max, J = L.j_max, v_0 = P.v0m, a_0 = P.a0m, dir = P.dir;
  unsigned char base;
  double c[5];
  c[0] = 12;
  c[1] = 0;
  c[2] = -24 * sq(a_0) + 48 * J * v_0;
  c[3] = 48 * dir * sq(J) * q_0 - 48 * dir * sq(J) * q_goal + 16 * pow3(a_0) - 48 * a_0 * J * v_0;
  c[4] = -3 * pow4(a_0) + 12.0 * sq(a_0) * J * v_0 - 12.0 * sq(J) * sq(v_0);
  double r = smallest_root<4>(c);
  T[0] = (2.0 * sq(r) - 4 * a_0 * r + sq(a_0) - 2.0 * v_0 * J) / (4 * J * r);
  T[6] = sqrt(4 * sq(J) * sq(T[0]) + 8 * a_0 * J * T[0] + 2.0 * sq(a_0) + 4 * J * v_0) / (2.0 * J);
  T[4] = a_0 / J + T[0] + T[6];
  T[1] = 0;
  T[5] = 0;
  base = CASE_Q1;
  if (a_0 + T[0] * J > A) {  // cc:273-296
    T[0] = (A - a_0) / J;
    T[6] = 1.0 / J *
           (A / 2 +
            sqrt(9 * sq(A) +
                 6 * sqrt(-12.0 * A * pow3(J) * pow3(T[0]) + 9 * sq(a_0) * sq(J) * sq(T[0]) -
                          18 * a_0 * A * sq(J) * sq(T[0]) + 9 * sq(A) * sq(J) * sq(T[0]) +
                          36 * a_0 * sq(J) * T[0] * v_0 - 72.0 * A * dir * sq(J) * q_0 +
                          72.0 * A * dir * sq(J) * q_goal - 36 * A * sq(J) * T[0] * v_0 +
                          3 * pow4(A) + 36 * sq(J) * sq(v_0))) /
                6.0 -
            A);
    T[4] = T[6] + A / J;
    T[1] = -(-J * sq(T[4]) - 2.0 * J * T[4] * T[6] + J * sq(T[6]) + a_0 * T[0] + A * T[0] +
             2.0 * A * T[4] + 2.0 * A * T[6] + 2.0 * v_0) /
           (2.0 * A);
    T[5] = 0;
    base = CASE_Q1_P2;
  }
  if (T[6] * J > A) {  // cc:299-333
    T[6] = A / J;
    c[0] = 12;
    c[1] = -24 * A;
    c[2] = -12.0 * sq(a_0) + 12.0 * sq(A) + 24 * J * v_0;
    c[3] = 0;
    c[4] = 24 * dir * sq(J) * q_0 * A - 24 * dir * sq(J) * q_goal * A + 3 * pow4(a_0) +
           8 * pow3(a_0) * A + 6 * sq(a_0) * sq(A) - 12.0 * sq(a_0) * J * v_0 -
           24 * a_0 * J * v_0 * A - 12.0 * sq(A) * J * v_0 + 12.0 * sq(J) * sq(v_0);
    r = smallest_root<4>(c);
    T[0] = (r - a_0 - A) / J;
    T[4] = (a_0 + A) / J + T[0];
    T[5] = (sq(J) * sq(T[0]) + 2.0 * sq(J) * T[0] * T[4] - sq(J) * sq(T[4]) + 2.0 * a_0 * J * T[0] +
            2.0 * a_0 * J * T[4] - sq(A) + 2.0 * J * v_0) /
           (2.0 * J * A);
    T[1] = 0;
    if (base == CASE_Q1_P2) flags |= F_BOTH;
    base = CASE_Q2;
  }
  T[2] = 0;  // cc:335-336
  T[3] = 0;
  return base;
}

// Body of optSwitchTimes for cruise speed V. t receives the cumulative switching times;
// on the cc:340-344 failure it is left untouched, exactly like the reference.
// Returns OST_FAIL (reference returns false), OST_OK (true) or -- only when ALLOW_TAIL is
// false -- OST_DEFER: the joint needs the quartic tail (cc:245-337), nothing was written,
// and the caller must hand the problem to the generic kernel.
enum { OST_FAIL = 0, OST_OK = 1, OST_DEFER = 2 };

// V_IS_VMAX: the caller passes V = L.v_max (the time-optimal solve of stage 1), so the one
// division by the cruise speed can use the reciprocal prepared on the host like the others.
template <bool ALLOW_TAIL, bool V_IS_VMAX, class DIV>
LTP_HD int ost_body_dv(const JointLimits& L, double Ts, const Prologue& P, double q_goal, double q_0,
                       double V, double* t, unsigned char& mod, unsigned char& kase, DIV& dv) {
  const double A = L.a_max, J = L.j_max;
  const double eps = kEps;
  double T[7];
  unsigned char flags = 0;
  mod = 0;
  if (P.brake_only) {  // cc:102-107
    T[0] = P.b0; T[1] = P.b1; T[2] = P.b2; T[3] = 0; T[4] = 0; T[5] = 0; T[6] = 0;
    cumsum7(T, t);
    kase = CASE_BRAKE_ONLY;
    return OST_OK;
  }
  const double v_0 = P.v0m, a_0 = P.a0m;
#if LTP_OST_MERGE
  // cc:119-143 and cc:168-190 with the two jerk profiles on ONE instruction stream. The modified
  // profile brakes from v_0 - V (optBraking, cc:650-701) and the normal one accelerates to V, and
  // the reference writes both with the same expressions: T0 = (A - a)/J, T2 = A/J,
  // T1 = (x - T0 a/2)/A - (T0 + T2)/2 and the same displacement polynomial in (v, a, T0, T1, T2),
  // on (v, a, x) = (+-(v_0 - V), +-a_0, -v) there and (v_0, a_0, V - v_0) here. Selecting the
  // operands first lets a warp that holds both kinds of joints (nearly every warp: a fifth of the
  // searching joints is on the modified profile) evaluate them once instead of one after the
  // other. Same operations on the same operands per lane, so the same bits; only the fix-up for
  // a missing phase 2 differs between the two (different expressions in the reference) and
  // stays a branch.
  const bool m = v_0 + dv.by(0.5 * a_0 * fabs(a_0), J, L.r_j) > V;  // cc:119
  double ve = v_0, ae = a_0, x = V - v_0, dirb = 1.0;
  if (m) {
    mod = 1;
    flags |= F_MOD;
    ve = v_0 - V;
    // cc:658-667 as in brake_profile
    const bool same_sign = ve * ae > 0;
    const bool fast_enough = fabs(ve) > dv.by(1.0 / 2.0 * sq(ae), J, L.r_j);
    dirb = neg_sgn((same_sign | fast_enough) ? ve : ae);
    if (dirb < 0) {
      ae = -ae;
      ve = -ve;
    }
    x = -ve;
  }
  T[0] = dv.by(A - ae, J, L.r_j);
  T[2] = L.a_over_j;
  double p3_2 = L.aoj3;  // pow3(T[2]): the per-joint constant unless a fix-up below changes T[2]
  T[1] = dv.by(x - 1.0 / 2.0 * T[0] * ae, A, L.r_a) - 1.0 / 2.0 * (T[0] + T[2]);
  if (m) {
    if (T[1] < -Ts) {  // cc:682-687
      T[0] = -ae / J + sqrt(sq(ae) / (2 * sq(J)) - ve / J);
      T[2] = T[0] + ae / J;
      T[1] = 0;
      p3_2 = pow3(T[2]);
    }
  } else if (T[1] < -eps) {  // cc:129-143
    double rad = J * (V - v_0) + 0.5 * sq(a_0);
    if (rad > 0) {
      T[2] = dv.by(sqrt(rad), J, L.r_j);
      T[0] = T[2] - dv.by(a_0, J, L.r_j);
      T[1] = 0;
      flags |= F_NOP2;
      p3_2 = pow3(T[2]);
    } else {
      zero7(t);
      kase = CASE_DEGENERATE | flags;
      return OST_OK;
    }
  }
#else
  double q_brake = 0.0;
  if (v_0 + dv.by(0.5 * a_0 * fabs(a_0), J, L.r_j) > V) {  // cc:119-122
    mod = 1;
    flags |= F_MOD;
    double unused;
    q_brake = brake_profile(L, Ts, v_0 - V, a_0, T[0], T[1], T[2], unused, dv);
  } else {  // cc:125-143
    T[0] = dv.by(A - a_0, J, L.r_j);
    T[2] = L.a_over_j;
    T[1] = dv.by(V - v_0 - 0.5 * T[0] * a_0, A, L.r_a) - 0.5 * (T[0] + T[2]);
    if (T[1] < -eps) {
      double rad = J * (V - v_0) + 0.5 * sq(a_0);
      if (rad > 0) {
        T[2] = dv.by(sqrt(rad), J, L.r_j);
        T[0] = T[2] - dv.by(a_0, J, L.r_j);
        T[1] = 0;
        flags |= F_NOP2;
      } else {
        zero7(t);
        kase = CASE_DEGENERATE | flags;
        return OST_OK;
      }
    }
  }
#endif
  // cc:147-165
  T[4] = L.a_over_j;
  T[6] = T[4];
#if LTP_OST_MERGE && LTP_OST_CONST
  // the time-optimal solve (V = v_max): this half is a function of the limits, prepared on the host
  const bool half_prepared = V_IS_VMAX && L.part2v == L.part2v;
  double p3_4 = L.aoj3;  // pow3(T[4]) = pow3(T[6])
  if (half_prepared) {
    T[5] = L.t5v;
  } else
#endif
  {
    T[5] = dv.by(V, A, L.r_a) - 1.0 / 2.0 * (T[4] + T[6]);
    if (T[5] < -eps) {
      double rad = dv.by(V, J, L.r_j);
      if (rad > 0) {
        T[4] = sqrt(rad);
        T[6] = T[4];
        T[5] = 0;
        flags |= F_NOP6;
#if LTP_OST_MERGE && LTP_OST_CONST
        p3_4 = pow3(T[4]);
#endif
      } else {
        zero7(t);
        kase = CASE_DEGENERATE | flags;
        return OST_OK;
      }
    }
  }
  // cc:168-190
  double part1;
#if LTP_OST_MERGE
  {
    const double T012 = T[0] + T[1] + T[2];
    const double s = ve * T012 +
                     ae * (1.0 / 2.0 * sq(T[0]) + T[0] * (T[1] + T[2]) + 1.0 / 2.0 * sq(T[2])) +
                     J * (1.0 / 6.0 * pow3(T[0]) + 1.0 / 2.0 * sq(T[0]) * (T[1] + T[2]) -
                          1.0 / 6.0 * (LTP_OST_CONST ? p3_2 : pow3(T[2])) + 1.0 / 2.0 * T[0] * sq(T[2])) +
                     A * (1.0 / 2.0 * sq(T[1]) + T[1] * T[2]);
    part1 = m ? dirb * s + V * T012 : s;
  }
#else
  if (mod == 1) {
    part1 = q_brake + V * (T[0] + T[1] + T[2]);
  } else {
    part1 = v_0 * (T[0] + T[1] + T[2]) +
            a_0 * (1.0 / 2.0 * sq(T[0]) + T[0] * (T[1] + T[2]) + 1.0 / 2.0 * sq(T[2])) +
            J * (1.0 / 6.0 * pow3(T[0]) + 1.0 / 2.0 * sq(T[0]) * (T[1] + T[2]) -
                 1.0 / 6.0 * pow3(T[2]) + 1.0 / 2.0 * T[0] * sq(T[2])) +
            A * (1.0 / 2.0 * sq(T[1]) + T[1] * T[2]);
  }
#endif
#if LTP_OST_MERGE && LTP_OST_CONST
  double part2;
  if (half_prepared) {
    part2 = L.part2v;
  } else {
    part2 = J * (1.0 / 6.0 * p3_4 + 1.0 / 2.0 * sq(T[6]) * (T[5] + T[4]) - 1.0 / 6.0 * p3_4 +
                 1.0 / 2.0 * T[6] * sq(T[4])) +
            A * (1.0 / 2.0 * sq(T[5]) + T[5] * T[4]);
  }
#else
  double part2 = J * (1.0 / 6.0 * pow3(T[6]) + 1.0 / 2.0 * sq(T[6]) * (T[5] + T[4]) -
                      1.0 / 6.0 * pow3(T[4]) + 1.0 / 2.0 * T[6] * sq(T[4])) +
                 A * (1.0 / 2.0 * sq(T[5]) + T[5] * T[4]);
#endif
  T[3] = V_IS_VMAX ? dv.by(P.dist - part1 - part2, L.v_max, L.r_v) : (P.dist - part1 - part2) / V;

  unsigned char base = (unsigned char)(1 + ((flags & F_NOP2) ? 1 : 0) + ((flags & F_NOP6) ? 2 : 0));

  if (T[3] < -eps) {  // cc:194
    if (mod == 1) {   // cc:195-199
      zero7(t);
      kase = CASE_FAIL | flags;
      return OST_FAIL;
    }
    // cc:202-223; the powers of T[2], T[4] = T[6] are the per-joint constants unless a fix-up
    // changed those durations (the modified profile, which sets no flag, left above)
#if LTP_OST_MERGE && LTP_OST_CONST
    // the terms that hold only T[2], T[4] = T[6]
    double k_q2, k_m, k_q4, k_h4, k_32, k_34, k_7, k_82, k_84;
    if (DIV::kWideLimits && !(flags & (F_NOP2 | F_NOP6))) {
      k_q2 = L.rk_q; k_m = L.rk_m; k_q4 = L.rk_q; k_h4 = L.rk_h; k_32 = L.rk_3; k_34 = L.rk_3;
      k_7 = L.rk_7; k_82 = L.rk_8; k_84 = L.rk_8;
    } else {
      double c3_2 = L.aoj3, c4_2 = L.aoj4, c3_4 = L.aoj3, c4_4 = L.aoj4;
      if (flags & F_NOP2) {
        c3_2 = pow3(T[2]);
        c4_2 = pow4(T[2]);
      }
      if (flags & F_NOP6) {
        c3_4 = pow3(T[4]);
        c4_4 = pow4(T[4]);
      }
      k_q2 = (sq(J) * c4_2) / 4;
      k_m = (sq(J) * sq(T[2]) * sq(T[4])) / 2;
      k_q4 = (sq(J) * c4_4) / 4;
      k_h4 = (sq(J) * c4_4) / 2;
      k_32 = dv.by3(2.0 * J * A * c3_2);
      k_34 = dv.by3(2.0 * J * A * c3_4);
      k_7 = 2.0 * J * A * sq(T[4]) * T[6];
      k_82 = 2.0 * sq(A) * sq(T[2]);
      k_84 = 2.0 * sq(A) * sq(T[4]);
    }
#else
    const double k_q2 = (sq(J) * pow4(T[2])) / 4, k_m = (sq(J) * sq(T[2]) * sq(T[4])) / 2,
                 k_q4 = (sq(J) * pow4(T[4])) / 4, k_h4 = (sq(J) * pow4(T[6])) / 2,
                 k_32 = dv.by3(2.0 * J * A * pow3(T[2])), k_34 = dv.by3(2.0 * J * A * pow3(T[4])),
                 k_7 = 2.0 * J * A * sq(T[4]) * T[6], k_82 = 2.0 * sq(A) * sq(T[2]),
                 k_84 = 2.0 * sq(A) * sq(T[4]);
#endif
    double rad = (sq(J) * pow4(T[0])) / 2 - k_q2 + k_m - k_q4 + k_h4 + 2.0 * J * a_0 * pow3(T[0]) -
                 dv.by3(2.0 * J * A * pow3(T[0])) - 2.0 * J * A * T[0] * sq(T[2]) + k_32 + k_34 - k_7 - k_34 +
                 2.0 * J * v_0 * sq(T[0]) + 2.0 * sq(a_0) * sq(T[0]) - 2.0 * a_0 * A * sq(T[0]) -
                 2.0 * a_0 * A * sq(T[2]) + 4 * a_0 * v_0 * T[0] + k_82 + k_84 - 4 * A * v_0 * T[0] +
                 4 * P.dist * A + 2.0 * sq(v_0);
    if (rad > 0) {  // cc:224-236
      // x / (4 A) == (x / A) / 4 bit for bit (scaling by 4 is exact)
      T[5] = 0.25 * dv.by(-(4 * A * T[4] - 2.0 * sqrt(rad) + J * sq(T[2]) - J * sq(T[4]) + 2.0 * J * sq(T[6])),
                           A, L.r_a);
      T[1] = dv.by(-v_0 - a_0 * T[0] - 1.0 / 2.0 * J * sq(T[0]) + 1.0 / 2.0 * J * sq(T[2]) +
                        1.0 / 2.0 * J * sq(T[6]) - 1.0 / 2.0 * J * sq(T[4]),
                    A, L.r_a) -
             T[2] + T[5] + T[4];
      T[3] = 0;
      base = CASE_NOP4;
    } else {
      zero7(t);
      kase = CASE_DEGENERATE | flags;
      return OST_OK;
    }
    if (any_below2(T[5], T[1], -eps)) {
      if (!ALLOW_TAIL) return OST_DEFER;
      base = ost_quartic_tail(L, P, q_goal, q_0, T, flags);
    }
  }
  // cc:340-348
  // (the reference's second test, T < 0 && T >= -eps, is T < 0 once T < -eps is excluded;
  // a NaN fails both and stays)
  if (any_below7(T, -eps)) {
    kase = CASE_FAIL_UNTOUCHED | flags;
    return OST_FAIL;
  }
#pragma unroll
  for (int i = 0; i < 7; ++i)
    if (T[i] < 0.0) T[i] = 0.0;
  cumsum7(T, t);  // cc:351
  kase = base | flags;
  return OST_OK;
}

template <bool ALLOW_TAIL, bool V_IS_VMAX = false>
LTP_HD int ost_body_t(const JointLimits& L, double Ts, const Prologue& P, double q_goal, double q_0,
                      double V, double* t, unsigned char& mod, unsigned char& kase) {
  DivChecked dv;
  return ost_body_dv<ALLOW_TAIL, V_IS_VMAX>(L, Ts, P, q_goal, q_0, V, t, mod, kase, dv);
}

LTP_HD bool ost_body(const JointLimits& L, double Ts, const Prologue& P, double q_goal, double q_0,
                     double V, double* t, unsigned char& mod, unsigned char& kase) {
  return ost_body_t<true>(L, Ts, P, q_goal, q_0, V, t, mod, kase) == OST_OK;
}

// ------------------------------------------------------------------------------------
// cc:378-629: the k-th cruise-speed candidate (k = 1..8) of the time-scaling search.
// v_0, a_0 are in the mapped frame (cc:372-375); tr is the required end time.
// ------------------------------------------------------------------------------------
struct TsInput {
  double q_goal, q_0, v_0, a_0, dir, tr;
};

template <class DIV>
LTP_HD double ts_candidate1(const JointLimits& L, const TsInput& I, DIV& dv) {
  const double A = L.a_max, J = L.j_max, a_0 = I.a_0, v_0 = I.v_0, tr = I.tr, dir = I.dir;
  if constexpr (DIV::kWideLimits) {
    // the same expression with the limit-only prefixes of its products read from the limits
    // (derive_limits forms them with the same operations in the same order)
    return dv.by(L.c_aj * tr / 2 - sq(a_0) / 4 + a_0 * A / 2 - L.c_a2h + v_0 * J / 2 -
                      dv.by12(sqrt(L.c_36a2j2 * sq(tr) - 36 * sq(a_0) * A * J * tr +
                                 72.0 * a_0 * sq(A) * J * tr - L.c_72a3j * tr +
                                 L.c_144a * dir * sq(J) * I.q_0 - L.c_144a * dir * sq(J) * I.q_goal +
                                 L.c_72aj2 * v_0 * tr - 9 * pow4(a_0) + 12.0 * pow3(a_0) * A +
                                 36 * sq(a_0) * sq(A) + 36 * sq(a_0) * J * v_0 - 72.0 * a_0 * L.c_a3 -
                                 72.0 * a_0 * A * J * v_0 + L.c_36a4 - L.c_36j2 * sq(v_0))),
                  J, L.r_j);
  } else {
    return dv.by(A * J * tr / 2 - sq(a_0) / 4 + a_0 * A / 2 - sq(A) / 2 + v_0 * J / 2 -
                      dv.by12(sqrt(36 * sq(A) * sq(J) * sq(tr) - 36 * sq(a_0) * A * J * tr +
                                 72.0 * a_0 * sq(A) * J * tr - 72.0 * pow3(A) * J * tr +
                                 144 * A * dir * sq(J) * I.q_0 - 144 * A * dir * sq(J) * I.q_goal +
                                 72.0 * A * sq(J) * v_0 * tr - 9 * pow4(a_0) + 12.0 * pow3(a_0) * A +
                                 36 * sq(a_0) * sq(A) + 36 * sq(a_0) * J * v_0 - 72.0 * a_0 * pow3(A) -
                                 72.0 * a_0 * A * J * v_0 + 36 * pow4(A) - 36 * sq(J) * sq(v_0))),
                  J, L.r_j);
  }
}

LTP_HD double ts_candidate1(const JointLimits& L, const TsInput& I) {
  DivChecked dv;
  return ts_candidate1(L, I, dv);
}

// cc:408-436. The reference's 17 divisions by 2J, A, 6J^3, 2J^3, 2J^2, J, 2AJ and AJ go through the
// reciprocals prepared on the host (x / (2 d) as half of x / d: scaling by two is exact; 1.0 / J is
// the reciprocal itself); the last one, by a sum, is a plain division. Same bits as the expression
// written with '/' (ts_candidate2_plain below; tests/test_devmath_host.py compares the two).
template <class DIV>
LTP_HD double ts_candidate2(const JointLimits& L, const TsInput& I, DIV& dv) {
  const double A = L.a_max, J = L.j_max, a_0 = I.a_0, v_0 = I.v_0, tr = I.tr, dir = I.dir;
  // w, h, g: sub-expressions the reference writes out repeatedly (cc:413-431)
  const double w = dv.by(v_0 + dv.half_by(a_0 * (a_0 - A), J, L.r_j), A, L.r_a);
  const double h = 0.5 * L.a_over_j;
  const double g = dv.half_by(a_0 - A, J, L.r_j);
  const double sA = a_0 + A;
  const double J2 = sq(J), J3 = pow3(J), J3x6 = 6 * J3, AJ = A * J;
  return -(dir * (I.q_0 - I.q_goal) -
           J * (dv.by(pow3(sA), J3x6, L.r_6j3) - dv.by(pow3(A), J3x6, L.r_6j3) +
                dv.half_by(sq(A) * sA, J3, L.r_j3) + dv.half_by(sq(sA) * (w + h + g), J2, L.r_j2)) +
           a_0 * (dv.half_by(sq(sA), J2, L.r_j2) + dv.half_by(sq(A), J2, L.r_j2) +
                  dv.by(sA * (w + h + g), J, L.r_j)) -
           A * (sq(w - h + g) / 2 + dv.by(A * (w - h + g), J, L.r_j)) + v_0 * (w + dv.by(sA, J, L.r_j) + h + g)) /
         (h - dv.by(v_0, A, L.r_a) + A * (dv.by(w - h + g, A, L.r_a) + L.r_j) -
          dv.half_by(sq(a_0) + 2.0 * a_0 * A + 4 * sq(A) - 2.0 * J * tr * A + 2.0 * J * v_0, AJ, L.r_aj) +
          dv.half_by(sq(sA), AJ, L.r_aj) - dv.by(a_0 * sA, AJ, L.r_aj));
}

LTP_HD double ts_candidate2(const JointLimits& L, const TsInput& I) {
  DivChecked dv;
  return ts_candidate2(L, I, dv);
}

// the same expression with the reference's divisions written out (test oracle for the above)
LTP_HD double ts_candidate2_plain(const JointLimits& L, const TsInput& I) {
  const double A = L.a_max, J = L.j_max, a_0 = I.a_0, v_0 = I.v_0, tr = I.tr, dir = I.dir;
  // w, h, g: sub-expressions the reference writes out repeatedly (cc:413-431)
  const double w = (v_0 + (a_0 * (a_0 - A)) / (2.0 * J)) / A;
  const double h = A / (2.0 * J);
  const double g = (a_0 - A) / (2.0 * J);
  const double sA = a_0 + A;
  const double J3 = pow3(J);
  return -(dir * (I.q_0 - I.q_goal) -
           J * (pow3(sA) / (6 * J3) - pow3(A) / (6 * J3) + (sq(A) * sA) / (2.0 * J3) +
                (sq(sA) * (w + h + g)) / (2.0 * sq(J))) +
           a_0 * (sq(sA) / (2.0 * sq(J)) + sq(A) / (2.0 * sq(J)) + (sA * (w + h + g)) / J) -
           A * (sq(w - h + g) / 2 + (A * (w - h + g)) / J) + v_0 * (w + sA / J + h + g)) /
         (h - v_0 / A + A * ((w - h + g) / A + 1.0 / J) -
          (sq(a_0) + 2.0 * a_0 * A + 4 * sq(A) - 2.0 * J * tr * A + 2.0 * J * v_0) / (2.0 * A * J) +
          sq(sA) / (2.0 * A * J) - (a_0 * sA) / (A * J));
}

// candidates 3..8 need a polynomial root (quartic, quartic, quintic, quartic, quartic, sextic)
LTP_HD_NOINLINE double ts_candidate_root(int k, const JointLimits& L, const TsInput& I) {
  const double A = L.a_max, J = L.j_max, a_0 = I.a_0, v_0 = I.v_0, tr = I.tr, dir = I.dir;
  const double q_0 = I.q_0, q_goal = I.q_goal;
  double p[7], r;
  switch (k) {
    case 3:  // cc:449-473
      p[0] = 3;
      p[1] = 12.0 * A;
      p[2] = -24 * A * J * tr - 12.0 * sq(a_0) - 24 * a_0 * A + 12.0 * sq(A) + 24 * J * v_0;
      p[3] = 0;
      p[4] = 48 * sq(a_0) * A * J * tr - 96 * dir * sq(J) * A * q_0 + 96 * dir * sq(J) * A * q_goal -
             96 * A * sq(J) * v_0 * tr + 12.0 * pow4(a_0) + 16 * pow3(a_0) * A -
             24 * sq(a_0) * sq(A) - 48 * sq(a_0) * J * v_0 + 48 * sq(A) * J * v_0 +
             48 * sq(J) * sq(v_0);
      r = smallest_root<4>(p);
      return (-2.0 * sq(a_0) + 4 * J * v_0 + sq(r)) / (4 * J);
    case 4:  // cc:485-514
      p[0] = 12;
      p[1] = 24 * A;
      p[2] = -24 * A * J * tr + 24 * sq(a_0) - 48 * a_0 * A + 24 * sq(A) - 24 * J * v_0 + 12.0 * a_0 -
             12.0 * A;
      p[3] = 0;
      p[4] = -24 * dir * sq(J) * A * q_0 + 24 * dir * sq(J) * A * q_goal + 9 * pow4(a_0) -
             12.0 * pow3(a_0) * A - 24 * sq(a_0) * J * v_0 + 48 * a_0 * A * J * v_0 + 4 * pow4(A) -
             24 * sq(A) * J * v_0 + 12.0 * sq(J) * sq(v_0) + 6 * pow3(a_0) + 6 * sq(a_0) * A -
             12.0 * a_0 * sq(A) - 12.0 * a_0 * J * v_0 + 12.0 * A * J * v_0 + 4 * a_0 * A - 4 * sq(A);
      r = smallest_root<4>(p);
      return sq(r) / J;
    case 5: {  // cc:526-541
      const double J3 = pow3(J), J4 = pow4(J), a3 = pow3(a_0), a4 = pow4(a_0);
      p[0] = (144 * J * tr + 144 * a_0);
      p[1] = (-72.0 * sq(J) * sq(tr) - 144 * a_0 * J * tr + 36 * sq(a_0) - 216 * J * v_0);
      p[2] = (144 * dir * sq(J) * q_0 - 144 * dir * sq(J) * q_goal + 48 * a3 - 144 * a_0 * J * v_0);
      p[3] = (-144 * dir * J3 * q_0 * tr + 144 * dir * J3 * q_goal * tr - 48 * a3 * J * tr -
              144 * a_0 * dir * sq(J) * q_0 + 144 * a_0 * dir * sq(J) * q_goal +
              144 * a_0 * sq(J) * v_0 * tr + 6 * a4 - 72.0 * sq(a_0) * J * v_0 + 216 * sq(J) * sq(v_0));
      p[4] = 0;
      p[5] = -72.0 * sq(dir) * J4 * sq(q_0) + 144 * sq(dir) * J4 * q_0 * q_goal -
             72.0 * sq(dir) * J4 * sq(q_goal) - 48 * a3 * dir * sq(J) * q_0 +
             48 * a3 * dir * sq(J) * q_goal + 144 * a_0 * dir * J3 * q_0 * v_0 -
             144 * a_0 * dir * J3 * q_goal * v_0 + pow6(a_0) - 6 * a4 * J * v_0 +
             36 * sq(a_0) * sq(J) * sq(v_0) - 72.0 * J3 * pow3(v_0);
      r = smallest_root<5>(p);
      return sq(r) / J;
    }
    case 6:  // cc:553-567
      p[0] = 3;
      p[1] = -6 * 1.4142135623730951 * A;  // -6*sqrt(2)*a_max
      p[2] = (12.0 * A * J * tr - 6 * sq(a_0) - 12.0 * a_0 * A - 6 * sq(A) - 12.0 * J * v_0);
      p[3] = 0;
      p[4] = -12.0 * sq(a_0) * A * J * tr - 24 * dir * sq(J) * A * q_0 + 24 * dir * sq(J) * A * q_goal -
             24 * A * sq(J) * v_0 * tr + 3 * pow4(a_0) + 4 * pow3(a_0) * A + 6 * sq(a_0) * sq(A) +
             12.0 * sq(a_0) * J * v_0 + 12.0 * sq(A) * J * v_0 + 12.0 * sq(J) * sq(v_0);
      r = smallest_root<4>(p);
      return -(sq(r) - sq(a_0) - 2.0 * J * v_0) / (2.0 * J);
    case 7:  // cc:579-593
      p[0] = 12;
      p[1] = -24 * A;
      p[2] = (24 * A * J * tr - 12.0 * sq(a_0) - 24 * a_0 * A - 12.0 * sq(A) - 24 * J * v_0);
      p[3] = 0;
      p[4] = 24 * dir * sq(J) * A * q_0 - 24 * dir * sq(J) * A * q_goal + 3 * pow4(a_0) +
             8 * pow3(a_0) * A + 6 * sq(a_0) * sq(A) + 12.0 * sq(a_0) * J * v_0 +
             24 * a_0 * A * J * v_0 + 12.0 * sq(A) * J * v_0 + 12.0 * sq(J) * sq(v_0);
      r = smallest_root<4>(p);
      return sq(r) / J;
    default: {  // 8, cc:606-629
      const double J3 = pow3(J), J4 = pow4(J), a3 = pow3(a_0), a4 = pow4(a_0);
      p[0] = 144;
      p[1] = (-144 * J * tr + 144 * a_0);
      p[2] = (72.0 * sq(J) * sq(tr) - 144 * a_0 * J * tr - 36 * sq(a_0) - 216 * J * v_0);
      p[3] = (-144 * dir * sq(J) * q_0 + 144 * dir * sq(J) * q_goal - 48 * a3 - 144 * a_0 * J * v_0);
      p[4] = (144 * dir * J3 * q_0 * tr - 144 * dir * J3 * q_goal * tr + 48 * a3 * J * tr -
              144 * a_0 * dir * sq(J) * q_0 + 144 * a_0 * dir * sq(J) * q_goal +
              144 * a_0 * sq(J) * v_0 * tr + 6 * a4 + 72.0 * sq(a_0) * J * v_0 + 216 * sq(J) * sq(v_0));
      p[5] = 0;
      p[6] = 72.0 * sq(dir) * J4 * sq(q_0) - 144 * sq(dir) * J4 * q_0 * q_goal +
             72.0 * sq(dir) * J4 * sq(q_goal) + 48 * a3 * dir * sq(J) * q_0 -
             48 * a3 * dir * sq(J) * q_goal + 144 * a_0 * dir * J3 * q_0 * v_0 -
             144 * a_0 * dir * J3 * q_goal * v_0 - pow6(a_0) - 6 * a4 * J * v_0 -
             36 * sq(a_0) * sq(J) * sq(v_0) - 72.0 * J3 * pow3(v_0);
      r = smallest_root<6>(p);
      return sq(r) / J;
    }
  }
}

// One attempt of the search (the block repeated at cc:398-405, 439-446, ...): if the
// candidate is usable, re-solve the switching times at that cruise speed and accept when
// the end time falls inside (t_req - 0.1, t_req + 0.01).
LTP_HD bool ts_try(const JointLimits& L, double Ts, const Prologue& P, const TsInput& I, double V,
                   double* scaled_t, unsigned char& mod, unsigned char& kase) {
  if (!isnan(V) && V > 0) {
    bool ok = ost_body(L, Ts, P, I.q_goal, I.q_0, V, scaled_t, mod, kase);
    if (ok && I.tr - scaled_t[6] < kTol && I.tr - scaled_t[6] > -kTol / 10) return true;
  }
  return false;
}

// cc:358-645 from attempt `first` on (1 = whole search). P must be the prologue of this
// joint's start state. Returns the accepted attempt (1..8) or 9 after the cc:641-644 reset.
// A joint on the BRAKE_ONLY exit gets the same switching times from every nested solve,
// whatever the cruise speed (cc:102-107), so the acceptance test of cc:402 has the same
// outcome in all eight attempts. True when that outcome is "rejected": the search ends in
// the cc:641-644 state without evaluating a single candidate (the reference grinds through
// six eigen-solves to get there).
LTP_HD bool ts_brake_only_rejects(const Prologue& P, double tr) {
  if (!P.brake_only) return false;
  const double t6 = ((((((P.b0 + P.b1) + P.b2) + 0.0) + 0.0) + 0.0) + 0.0);
  return !(tr - t6 < kTol && tr - t6 > -kTol / 10);
}

LTP_HD void ts_fail_state(const JointLimits& L, double* scaled_t, double& v_drive, unsigned char& mod,
                          unsigned char& final_case) {
  mod = 0;
  zero7(scaled_t);
  v_drive = L.v_max;
  final_case = CASE_FAIL;
}

// Closed-form part of the search (attempts 1 and 2) for the fast kernel, one function per
// attempt so that the joints attempt 1 does not settle can be regrouped in between.
// attempt1 returns 1 (accepted), 9 (the outcome is already known to be failure), 0 (a nested
// solve needs the quartic tail: the caller defers the whole problem to the generic kernel) or
// -1 (rejected: go on with attempt 2). attempt2 returns 2 (accepted) or 0 (the joint needs the
// root-solver attempts 3..8 or a quartic tail: defer).
template <class DIV>
LTP_HD int time_scaling_attempt1(const JointLimits& L, double Ts, const Prologue& P, const TsInput& I,
                                 double* scaled_t, double& v_drive, unsigned char& mod,
                                 unsigned char& final_case, DIV& dv) {
  if (ts_brake_only_rejects(P, I.tr)) {
    ts_fail_state(L, scaled_t, v_drive, mod, final_case);
    return 9;
  }
  const double V = ts_candidate1(L, I, dv);
  v_drive = V;
  if (!isnan(V) && V > 0) {
    const int st = ost_body_dv<false, false>(L, Ts, P, I.q_goal, I.q_0, V, scaled_t, mod, final_case, dv);
    if (st == OST_DEFER) return 0;
    if (st == OST_OK && I.tr - scaled_t[6] < kTol && I.tr - scaled_t[6] > -kTol / 10) return 1;
  }
  return -1;
}

LTP_HD int time_scaling_attempt1(const JointLimits& L, double Ts, const Prologue& P, const TsInput& I,
                                 double* scaled_t, double& v_drive, unsigned char& mod,
                                 unsigned char& final_case) {
  DivChecked dv;
  return time_scaling_attempt1(L, Ts, P, I, scaled_t, v_drive, mod, final_case, dv);
}

template <class DIV>
LTP_HD int time_scaling_attempt2(const JointLimits& L, double Ts, const Prologue& P, const TsInput& I,
                                 double* scaled_t, double& v_drive, unsigned char& mod,
                                 unsigned char& final_case, DIV& dv) {
  const double V = ts_candidate2(L, I, dv);
  v_drive = V;
  if (!isnan(V) && V > 0) {
    const int st = ost_body_dv<false, false>(L, Ts, P, I.q_goal, I.q_0, V, scaled_t, mod, final_case, dv);
    if (st == OST_DEFER) return 0;
    if (st == OST_OK && I.tr - scaled_t[6] < kTol && I.tr - scaled_t[6] > -kTol / 10) return 2;
  }
  return 0;
}

LTP_HD int time_scaling_attempt2(const JointLimits& L, double Ts, const Prologue& P, const TsInput& I,
                                 double* scaled_t, double& v_drive, unsigned char& mod,
                                 unsigned char& final_case) {
  DivChecked dv;
  return time_scaling_attempt2(L, Ts, P, I, scaled_t, v_drive, mod, final_case, dv);
}

LTP_HD int time_scaling_closed_form(const JointLimits& L, double Ts, const Prologue& P, const TsInput& I,
                                    double* scaled_t, double& v_drive, unsigned char& mod,
                                    unsigned char& final_case) {
  const int c = time_scaling_attempt1(L, Ts, P, I, scaled_t, v_drive, mod, final_case);
  return c >= 0 ? c : time_scaling_attempt2(L, Ts, P, I, scaled_t, v_drive, mod, final_case);
}

LTP_HD int time_scaling_from(int first, const JointLimits& L, double Ts, const Prologue& P,
                             const TsInput& I, double* scaled_t, double& v_drive,
                             unsigned char& mod, unsigned char& final_case) {
  double V;
  if (ts_brake_only_rejects(P, I.tr)) {
    ts_fail_state(L, scaled_t, v_drive, mod, final_case);
    return 9;
  }
  if (first <= 1) {
    V = ts_candidate1(L, I);
    v_drive = V;
    if (ts_try(L, Ts, P, I, V, scaled_t, mod, final_case)) return 1;
  }
  if (first <= 2) {
    V = ts_candidate2(L, I);
    v_drive = V;
    if (ts_try(L, Ts, P, I, V, scaled_t, mod, final_case)) return 2;
  }
  for (int k = (first > 3 ? first : 3); k <= 8; ++k) {
    V = ts_candidate_root(k, L, I);
    v_drive = V;
    if (ts_try(L, Ts, P, I, V, scaled_t, mod, final_case)) return k;
  }
  ts_fail_state(L, scaled_t, v_drive, mod, final_case);
  return 9;
}

LTP_HD TsInput make_ts_input(double q_goal, double q_0, double v_0, double a_0, double dir,
                             double tr) {
  TsInput I;
  I.q_goal = q_goal;
  I.q_0 = q_0;
  if (dir < 0) {  // cc:372-375
    v_0 = -v_0;
    a_0 = -a_0;
  }
  I.v_0 = v_0;
  I.a_0 = a_0;
  I.dir = dir;
  I.tr = tr;
  return I;
}

// cc:68-77 for one joint
template <class DIV>
LTP_HD bool check_joint_input(const JointLimits& L, double q_0, double v_0, double a_0, DIV& dv) {
  if (q_0 < L.q_min || q_0 > L.q_max || fabs(v_0) > L.v_max || fabs(a_0) > L.a_max) return false;
  if (fabs(v_0 + dv.by(0.5 * a_0 * fabs(a_0), L.j_max, L.r_j)) > L.v_max) return false;
  return true;
}

LTP_HD bool check_joint_input(const JointLimits& L, double q_0, double v_0, double a_0) {
  DivChecked dv;
  return check_joint_input(L, q_0, v_0, a_0, dv);
}

// cc:718 for one joint: (int)ceil(t6/Ts) + 1, or 0 when the time is not representable
// (the reference would invoke undefined behaviour there).
LTP_HD int samples_for(double t6, double Ts) {
  double x = ceil(t6 / Ts);
  if (!(x >= -1.0e9 && x <= 2.0e9)) return 0;
  return (int)x + 1;
}

// ------------------------------------------------------------------------------------
// cc:729-831 for one (problem, joint) row.
//
// The reference materialises a jerk array (piecewise-constant fills, cc:759-766, plus up to
// eight fractional impulses, cc:768-807) and then runs the forward-Euler recurrence
// (cc:810-831). Here the jerk array is never stored: RowSampler::init derives the sample
// indices and the impulses, build_segments cuts the row at every index where the jerk or
// the update rule can change (at most 26 pieces), and the hot loop advances the reference's
// own recurrence inside a piece with five FP64 operations per sample:
//     a += Ts*j;  v = cruise ? v_drive*dir : v + Ts*a;  q += Ts*v
// (multiply and add stay unfused; Ts*j is the same product the reference forms per sample).
// Sample 0 is t = Ts (cc:810-812). Impulses whose index falls outside [0, limit) are dropped
// (the reference writes them out of bounds).
// ------------------------------------------------------------------------------------
constexpr int kMaxSeg = 27;

struct RowSampler {
  double Ts, jp0, jp2, jp4, jp6;  // jerk of the four non-zero phases (phases 2, 4, 6 are 0)
  double vcruise;                 // v_drive * dir (cc:823)
  int s[7];
  int imp_idx[7];                 // -1 = slot unused
  double imp_v[7];                // value added at imp_idx (first addend)
  double imp0b, imp4b, imp4c;     // further addends of the two combined impulses (cc:781, 798)
  unsigned char n0, n4;           // number of addends of slots 0 and 4
  bool phase4;
  double a, v, q;

  LTP_HD void set_imp(int slot, int idx, int limit, double x) {
    imp_idx[slot] = (idx >= 0 && idx < limit) ? idx : -1;
    imp_v[slot] = x;
  }

  LTP_HD void init(double Ts_, double J, const double* t, double dir, unsigned char mod,
                   double q_0, double v_0, double a_0, double v_drive, int limit) {
    Ts = Ts_;
    // cc:734-744: profile {+1,0,-1,0,-1,0,+1} or modified {-1,0,+1,0,-1,0,+1}
    const double dj = dir * J;
    jp0 = dj * (mod == 1 ? -1 : 1);
    jp2 = dj * (mod == 1 ? 1 : -1);
    jp4 = dj * -1;
    jp6 = dj * 1;
    vcruise = v_drive * dir;
    double fr[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      double r = t[k] / Ts;
      double fl = floor(r);
      fr[k] = t[k] - Ts * fl;               // cc:747
      double idx = (k & 1) ? ceil(r) : fl;  // cc:751-757
      // clamp before the cast: NaN / huge values are undefined in the reference
      if (!(idx >= -2.0e9)) idx = -2.0e9;
      if (idx > 2.0e9) idx = 2.0e9;
      s[k] = (int)idx;
    }
    n0 = 1;
    n4 = 1;
    imp0b = imp4b = imp4c = 0.0;
    // cc:768-807, source order
    if (s[2] >= s[1]) {
      set_imp(0, s[0] + 1, limit, fr[0] / Ts * jp0);
      set_imp(1, s[1], (s[1] > 0) ? limit : 0, (1 - fr[1] / Ts) * jp2);
      set_imp(2, s[2] + 1, limit, fr[2] / Ts * jp2);
    } else {
      set_imp(0, s[1], (s[1] > 0) ? limit : 0, fr[0] / Ts * jp0);
      imp0b = (fr[2] - fr[0]) / Ts * jp2;
      n0 = 2;
      set_imp(1, -1, 0, 0.0);
      set_imp(2, -1, 0, 0.0);
    }
    set_imp(3, s[3], (s[3] > 0) ? limit : 0, (1 - fr[3] / Ts) * jp4);
    if (s[2] - s[0] > 0) {
      set_imp(4, s[4] + 1, limit, fr[4] / Ts * jp4);
    } else {
      set_imp(4, s[4], (s[4] > 0) ? limit : 0, fr[4] / Ts * jp4);
      imp4b = fr[0] / Ts * jp0;
      imp4c = (fr[2] - fr[0]) / Ts * jp2;
      n4 = 3;
    }
    set_imp(5, s[5], (s[5] > 0) ? limit : 0, (1 - fr[5] / Ts) * jp6);
    set_imp(6, s[6] + 1, limit, fr[6] / Ts * jp6);
    phase4 = s[3] - s[2] > 2;  // cc:813
    a = a_0; v = v_0; q = q_0;
  }

  // piecewise-constant part of the jerk at sample i: the reference fills ranges in phase
  // order and later fills overwrite earlier ones (cc:759-766) -> last writer wins
  LTP_HD double base_jerk(int i) const {
    if (s[6] - s[5] > 0 && i >= s[5] && i < s[6]) return jp6;
    if (s[5] - s[4] > 0 && i >= s[4] && i < s[5]) return 0.0;
    if (s[4] - s[3] > 0 && i >= s[3] && i < s[4]) return jp4;
    if (s[3] - s[2] > 0 && i >= s[2] && i < s[3]) return 0.0;
    if (s[2] - s[1] > 0 && i >= s[1] && i < s[2]) return jp2;
    if (s[1] - s[0] > 0 && i >= s[0] && i < s[1]) return 0.0;
    if (s[0] > 0 && i < s[0]) return jp0;
    return 0.0;
  }

  LTP_HD double jerk_at(int i) const {
    double j = base_jerk(i);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      if (imp_idx[k] == i) {
        j = j + imp_v[k];
        if (k == 0 && n0 > 1) j = j + imp0b;
        if (k == 4 && n4 > 1) {
          j = j + imp4b;
          j = j + imp4c;
        }
      }
    }
    return j;
  }

  // update rules of cc:815-829 at sample i (sample 0 uses the plain formulas, cc:810-812)
  LTP_HD bool a_zero(int i) const { return i > 0 && i > s[6]; }
  LTP_HD bool v_cruise(int i) const { return i > 0 && phase4 && i >= s[2] + 1 && i < s[3] - 1; }

  // smallest index > i at which the jerk or an update rule can change
  LTP_HD int next_break(int i) const {
    int best = 0x7fffffff;
#define LTP_CAND(c)                        \
  do {                                     \
    const int c_ = (c);                    \
    if (c_ > i && c_ < best) best = c_;    \
  } while (0)
#pragma unroll
    for (int k = 0; k < 7; ++k) LTP_CAND(s[k]);
    LTP_CAND(1);
    LTP_CAND(s[6] + 1);
    LTP_CAND(s[2] + 1);
    LTP_CAND(s[3] - 1);
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      LTP_CAND(imp_idx[k]);
      LTP_CAND(imp_idx[k] + 1);
    }
#undef LTP_CAND
    return best;
  }

  // one sample by the general rules; used by the host shadow and as the definition the
  // segment machinery below must reproduce
  LTP_HD void step(int i, double& jo, double& ao, double& vo, double& qo) {
    const double j = jerk_at(i);
    if (!a_zero(i)) a = a + Ts * j; else a = 0.0;
    if (v_cruise(i)) v = vcruise;
    else if (!a_zero(i)) v = v + Ts * a;
    else v = 0.0;
    q = q + Ts * v;
    jo = j; ao = a; vo = v; qo = q;
  }
};

// The row cut into pieces of constant jerk and constant update rule. Entry m of a row holds
// two 64-bit words, adjacent and 16-byte aligned (one 128-bit load per sample); consecutive
// entries of a row are ENTRY_STRIDE doubles apart, so the kernels keep one column per lane in
// shared memory ([entry][lane][2], conflict-free per quarter-warp for any mix of entry
// indices) while the host build uses a plain array (ENTRY_STRIDE 2).
//   word 0  jerk of the piece (the value that is emitted)
//   word 1  low 32 bits: first sample index of the NEXT piece; bit 32: the velocity is
//           overridden with v_drive*dir (cc:822-823); bit 33: "live", cleared from the sample
//           after the last switching time on, where the reference pins a and v to exactly 0
//           (cc:817-829)
LTP_HD double seg_pack(int next, int cruise, int live) {
  const unsigned hi = (unsigned)cruise | ((unsigned)live << 1);
  const long long b = (long long)(((unsigned long long)hi << 32) | (unsigned)next);
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double d;
  memcpy(&d, &b, 8);
  return d;
#endif
}
LTP_HD void seg_unpack(double d, int& next, bool& cruise, bool& live) {
#ifdef __CUDA_ARCH__
  next = __double2loint(d);
  const int hi = __double2hiint(d);
#else
  long long b;
  memcpy(&b, &d, 8);
  next = (int)(unsigned)(b & 0xffffffffll);
  const int hi = (int)((unsigned long long)b >> 32);
#endif
  cruise = (hi & 1) != 0;
  live = (hi & 2) != 0;
}

template <int ENTRY_STRIDE>
struct SegTableT {
  double* base;  // entry 0 of this row

  // stop_at_end: leave the entries behind the last piece unwritten (they are never entered);
  // for a caller that builds one table per thread and pays for every iteration
  LTP_HD void build(const RowSampler& R, int limit, bool stop_at_end = false) const {
    int cur = 0;
#pragma unroll 1
    for (int m = 0; m < kMaxSeg; ++m) {
      const int at = cur < limit ? cur : 0;  // entries past the end are never entered
      const double j = R.jerk_at(at);
      const int live = R.a_zero(at) ? 0 : 1;
      const int vc = R.v_cruise(at) ? 1 : 0;
      if (cur < limit) {
        cur = R.next_break(cur);
        if (cur >= limit) cur = 0x7fffffff;
      }
      double* e = base + m * ENTRY_STRIDE;
      e[0] = j;
      e[1] = seg_pack(cur, vc, live);
      if (stop_at_end && cur == 0x7fffffff) break;
    }
  }
};

// Streaming state of one row. step() is branch-free: the piece index advances by a compare
// and the two table words of the current piece are re-read every sample. Arithmetic per
// sample is exactly the reference's (cc:810-831): a + Ts*j, v + Ts*a, q + Ts*v with multiply
// and add unfused, and the pinned values (0, 0 and v_drive*dir) selected, not computed.
template <int ENTRY_STRIDE>
struct SegCursorT {
  double Ts, vcruise, a, v, q;
  int m, next;

  LTP_HD void begin(const RowSampler& R) {
    Ts = R.Ts; vcruise = R.vcruise; a = R.a; v = R.v; q = R.q;
    m = 0;
    next = 0x7fffffff;  // replaced by entry 0 on the first step (i = 0 never equals it)
  }

  LTP_HD void step(const SegTableT<ENTRY_STRIDE>& T, int i, double& jo, double& ao, double& vo, double& qo) {
    m += (i == next) ? 1 : 0;
    const double* e = T.base + m * ENTRY_STRIDE;
#ifdef __CUDA_ARCH__
    const double2 w = *reinterpret_cast<const double2*>(e);
    const double jv = w.x, pk = w.y;
#else
    const double jv = e[0], pk = e[1];
#endif
    bool vc, live;
    seg_unpack(pk, next, vc, live);
    const double a1 = a + Ts * jv;
    a = live ? a1 : 0.0;
    const double v1 = v + Ts * a;
    v = vc ? vcruise : (live ? v1 : 0.0);
    q = q + Ts * v;
    jo = jv; ao = a; vo = v; qo = q;
  }

  // Position after sample `end - 1`, starting from the state after sample `i - 1`, WITHOUT
  // stepping: inside a piece the recurrence has the closed form
  //   a_k = a + c k,  v_k = v + Ts (a k + c k(k+1)/2),  q_n = q + Ts (v n + Ts (a n(n+1)/2 + c n(n+1)(n+2)/6))
  // (c = Ts * jerk; discrete sums, not the continuous polynomials). The result differs from
  // the sequentially rounded one by ~1e-12, so it may only be used for decisions with a guard
  // band (the final joint-limit check of a clipped row); the cursor itself is not advanced.
  LTP_HD double peek_position(const SegTableT<ENTRY_STRIDE>& T, int i, int end) const {
    double pa = a, pv = v, pq = q;
    int pm = m, pnext = next;
    while (i < end) {
      pm += (i == pnext) ? 1 : 0;
      const double* e = T.base + pm * ENTRY_STRIDE;
      bool vc, live;
      seg_unpack(e[1], pnext, vc, live);
      const int stop = pnext < end ? pnext : end;
      const double k = (double)(stop - i);
      const double c = live ? Ts * e[0] : 0.0;
      if (!live) pa = 0.0;
      if (vc) {
        pv = vcruise;
        pq = pq + Ts * pv * k;
      } else if (live) {
        pq = pq + Ts * (pv * k + Ts * (pa * (k * (k + 1.0) * 0.5) + c * (k * (k + 1.0) * (k + 2.0) / 6.0)));
        pv = pv + Ts * (pa * k + c * (k * (k + 1.0) * 0.5));
      } else {
        pv = 0.0;
      }
      pa = pa + c * k;
      i = stop;
    }
    return pq;
  }
};

}  // namespace ltp
