/* TEST INFRASTRUCTURE ONLY -- see ltp_oracle.h for the contract and the pinning status.
 *
 * Plain-C restatement of the reference's planning hot path. "cc:N" cites
 * /root/reference/src/long_term_planner.cc line N, "roots.h:N" cites
 * /root/reference/include/long_term_planner/roots.h.
 *
 * Arithmetic contract (what makes this agree bit-for-bit with the reference compiled by
 * g++ -O2 -ffp-contract=off): every floating-point expression keeps the reference's
 * operand order and association; a second power is a plain product (what g++ makes of
 * pow(x,2)); third/fourth/sixth powers and the 0.5 power of cc:226 stay calls into libm's
 * pow(), exactly as in the reference binary; no fused multiply-add anywhere.
 */
#include "ltp_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define SQ(x) ((x) * (x))
#define EPS_T 4e-3 /* cc:96 */

struct ltpo_planner {
  int dof;
  double ts;
  double *q_min, *q_max, *v_max, *a_max, *j_max;
};

static double* dup_vec(const double* p, int n) {
  double* r = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (n > 0) memcpy(r, p, sizeof(double) * (size_t)n);
  return r;
}

ltpo_planner* ltpo_create(int dof, double t_sample, const double* q_min, const double* q_max,
                          const double* v_max, const double* a_max, const double* j_max) {
  ltpo_planner* L = (ltpo_planner*)malloc(sizeof(*L));
  L->dof = dof;
  L->ts = t_sample;
  L->q_min = dup_vec(q_min, dof);
  L->q_max = dup_vec(q_max, dof);
  L->v_max = dup_vec(v_max, dof);
  L->a_max = dup_vec(a_max, dof);
  L->j_max = dup_vec(j_max, dof);
  return L;
}

void ltpo_destroy(ltpo_planner* L) {
  if (!L) return;
  free(L->q_min); free(L->q_max); free(L->v_max); free(L->a_max); free(L->j_max);
  free(L);
}

/* long_term_planner.h:54-56 */
static int sgn(double x) { return (0.0 < x) - (x < 0.0); }

/* ------------------------------------------------------------------------------------ */
/* roots.h:22-50 on top of the eigenvalue algorithm of Eigen 3.4's EigenSolver            */
/* (RealSchur on the companion matrix). Double precision only.      */
/* ------------------------------------------------------------------------------------ */
#define RN 6

static void householder(const double* v, int n, double* ess, double* tau, double* beta) {
  double tail = 0.0;
  for (int i = 1; i < n; ++i) tail += v[i] * v[i];
  double c0 = v[0];
  if (tail <= 2.2250738585072014e-308) {
    *tau = 0.0;
    *beta = c0;
    for (int i = 0; i < n - 1; ++i) ess[i] = 0.0;
  } else {
    double b = sqrt(c0 * c0 + tail);
    if (c0 >= 0.0) b = -b;
    for (int i = 0; i < n - 1; ++i) ess[i] = v[i + 1] / (c0 - b);
    *tau = (b - c0) / b;
    *beta = b;
  }
}

static void refl_left(double a[RN][RN], int n, int k, int m, int c0, const double* ess, double tau) {
  if (tau == 0.0) return;
  for (int c = c0; c < n; ++c) {
    double tmp = ess[0] * a[k + 1][c];
    if (m == 3) tmp += ess[1] * a[k + 2][c];
    tmp += a[k][c];
    a[k][c] -= tau * tmp;
    a[k + 1][c] -= (tau * ess[0]) * tmp;
    if (m == 3) a[k + 2][c] -= (tau * ess[1]) * tmp;
  }
}

static void refl_right(double a[RN][RN], int k, int m, int r1, const double* ess, double tau) {
  if (tau == 0.0) return;
  for (int r = 0; r <= r1; ++r) {
    double tmp = a[r][k + 1] * ess[0];
    if (m == 3) tmp += a[r][k + 2] * ess[1];
    tmp += a[r][k];
    a[r][k] -= tau * tmp;
    a[r][k + 1] -= (tau * tmp) * ess[0];
    if (m == 3) a[r][k + 2] -= (tau * tmp) * ess[1];
  }
}

/* real Schur form of an upper-Hessenberg n x n matrix, in place; returns 1 on convergence */
static int real_schur(double a[RN][RN], int n) {
  const double eps = 2.220446049250313e-16, tiny = 2.2250738585072014e-308;
  double scale = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      double v = fabs(a[i][j]);
      if (v > scale) scale = v;
    }
  if (scale < tiny) {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) a[i][j] = 0.0;
    return 1;
  }
  /* Hessenberg reduction of C/scale is the identity for a companion matrix */
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[i][j] = a[i][j] / scale;
  for (int i = 2; i < n; ++i)
    for (int j = 0; j < i - 1; ++j) a[i][j] = 0.0;

  const int max_iters = 40 * n;
  int iu = n - 1, iter = 0, total = 0;
  double exshift = 0.0, norm = 0.0;
  for (int j = 0; j < n; ++j) {
    int lim = (j + 2 < n) ? j + 2 : n;
    double cs = 0.0;
    for (int i = 0; i < lim; ++i) cs += fabs(a[i][j]);
    norm += cs;
  }
  double caz = norm * (eps * eps);
  if (!(caz > tiny)) caz = tiny;
  if (norm != 0.0) {
    while (iu >= 0) {
      int il = iu;
      while (il > 0) {
        double s = fabs(a[il - 1][il - 1]) + fabs(a[il][il]);
        s = s * eps;
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
          for (int col = iu - 1; col < n; ++col) {
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
        double v[3] = {0.0, 0.0, 0.0};
        for (im = iu - 2; im >= il; --im) {
          double tmm = a[im][im];
          double r = sh0 - tmm;
          double s = sh1 - tmm;
          v[0] = (r * s - sh2) / a[im + 1][im] + a[im][im + 1];
          v[1] = a[im + 1][im + 1] - tmm - r - s;
          v[2] = a[im + 2][im + 1];
          if (im == il) break;
          double lhs = a[im][im - 1] * (fabs(v[1]) + fabs(v[2]));
          double rhs = v[0] * (fabs(a[im - 1][im - 1]) + fabs(tmm) + fabs(a[im + 1][im + 1]));
          if (fabs(lhs) < eps * rhs) break;
        }
        for (int k = im; k <= iu - 2; ++k) {
          int first = (k == im);
          double w[3], ess[2], tau, beta;
          if (first) {
            w[0] = v[0]; w[1] = v[1]; w[2] = v[2];
          } else {
            w[0] = a[k][k - 1]; w[1] = a[k + 1][k - 1]; w[2] = a[k + 2][k - 1];
          }
          householder(w, 3, ess, &tau, &beta);
          if (beta != 0.0) {
            if (first && k > il)
              a[k][k - 1] = -a[k][k - 1];
            else if (!first)
              a[k][k - 1] = beta;
            refl_left(a, n, k, 3, k, ess, tau);
            refl_right(a, k, 3, (iu < k + 3) ? iu : k + 3, ess, tau);
          }
        }
        {
          double w[2] = {a[iu - 1][iu - 2], a[iu][iu - 2]}, ess[1], tau, beta;
          householder(w, 2, ess, &tau, &beta);
          if (beta != 0.0) {
            a[iu - 1][iu - 2] = beta;
            refl_left(a, n, iu - 1, 2, iu - 1, ess, tau);
            refl_right(a, iu - 1, 2, iu, ess, tau);
          }
        }
        for (int i = im + 2; i <= iu; ++i) {
          a[i][i - 2] = 0.0;
          if (i > im + 2) a[i][i - 3] = 0.0;
        }
      }
    }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) a[i][j] = a[i][j] * scale;
  return total <= max_iters;
}

/* Trace of the polynomials handed to the root finder (single-threaded use: tests and
 * tools/root_crosscheck.py compare every one of them with LAPACK). Record layout: 8 doubles =
 * degree, up to 7 coefficients (highest power first, the rest 0). */
static double* g_trace = 0;
static int64_t g_trace_cap = 0, g_trace_len = 0;
void ltpo_trace_roots(double* buf, int64_t capacity) {
  g_trace = buf;
  g_trace_cap = capacity;
  g_trace_len = 0;
}
int64_t ltpo_trace_count(void) { return g_trace_len; }

double ltpo_roots(const double* coeffs, int deg, double* re, double* im) {
  const int n = deg;
  if (g_trace) {
    if (g_trace_len < g_trace_cap && deg >= 1 && deg <= 6) {
      double* rec = g_trace + 8 * g_trace_len;
      rec[0] = (double)deg;
      for (int i = 0; i < 7; ++i) rec[1 + i] = i <= deg ? coeffs[i] : 0.0;
    }
    g_trace_len++;
  }
  double a[RN][RN];
  double out_re[RN], out_im[RN];
  for (int i = 0; i < n; ++i) { out_re[i] = NAN; out_im[i] = NAN; }
  if (n >= 1 && n <= RN) {
    /* roots.h:28-31: zero matrix, ones below the diagonal, last column = -p_k/p_0 in
     * ascending power */
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) a[i][j] = 0.0;
    for (int i = 0; i + 1 < n; ++i) a[i + 1][i] = 1.0;
    for (int i = 0; i < n; ++i) a[i][n - 1] = (-1.0 * coeffs[n - i]) / coeffs[0];
    if (real_schur(a, n)) {
      int i = 0;
      while (i < n) {
        if (i == n - 1 || a[i + 1][i] == 0.0) {
          if (!isfinite(a[i][i])) break;
          out_re[i] = a[i][i];
          out_im[i] = 0.0;
          ++i;
        } else {
          double p = 0.5 * (a[i][i] - a[i + 1][i + 1]);
          double t0 = a[i + 1][i], t1 = a[i][i + 1];
          double mx = fabs(p);
          if (fabs(t0) > mx) mx = fabs(t0);
          if (fabs(t1) > mx) mx = fabs(t1);
          t0 /= mx;
          t1 /= mx;
          double p0 = p / mx;
          double z = mx * sqrt(fabs(p0 * p0 + t0 * t1));
          double rr = a[i + 1][i + 1] + p;
          if (!(isfinite(rr) && isfinite(z))) break;
          out_re[i] = rr; out_im[i] = z;
          out_re[i + 1] = rr; out_im[i + 1] = -z;
          i += 2;
        }
      }
    }
  }
  /* roots.h:43-50 */
  double best = INFINITY;
  for (int i = 0; i < n; ++i) {
    if (out_im[i] == 0 && out_re[i] > 1e-7) best = fmin(best, out_re[i]);
    if (re) re[i] = out_re[i];
    if (im) im[i] = out_im[i];
  }
  return best;
}

/* ------------------------------------------------------------------------------------ */
/* cc:68-77                                                                              */
/* ------------------------------------------------------------------------------------ */
int ltpo_check_inputs(const ltpo_planner* L, const double* q_0, const double* v_0, const double* a_0) {
  for (int i = 0; i < L->dof; ++i) {
    if (q_0[i] < L->q_min[i] || q_0[i] > L->q_max[i] || fabs(v_0[i]) > L->v_max[i] ||
        fabs(a_0[i]) > L->a_max[i])
      return 0;
    if (fabs(v_0[i] + 0.5 * a_0[i] * fabs(a_0[i]) / L->j_max[i]) > L->v_max[i]) return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------------------------ */
/* cc:650-701                                                                            */
/* ------------------------------------------------------------------------------------ */
void ltpo_opt_braking(const ltpo_planner* L, int joint, double v_0, double a_0, double* q,
                      double T[7], double* dir) {
  const double A = L->a_max[joint], J = L->j_max[joint];
  /* cc:658-670 */
  if (v_0 * a_0 > 0) {
    *dir = -sgn(v_0);
  } else if (fabs(v_0) > 1.0 / 2.0 * SQ(a_0) / J) {
    *dir = -sgn(v_0);
  } else {
    *dir = -sgn(a_0);
  }
  if (*dir < 0) { /* cc:673-676 */
    a_0 = -a_0;
    v_0 = -v_0;
  }
  /* cc:679-681 */
  T[0] = (A - a_0) / J;
  T[2] = A / J;
  T[1] = (-v_0 - 1.0 / 2.0 * T[0] * a_0) / A - 1.0 / 2.0 * (T[0] + T[2]);
  if (T[1] < -L->ts) { /* cc:685-689 */
    T[0] = -a_0 / J + sqrt(SQ(a_0) / (2 * SQ(J)) - v_0 / J);
    T[2] = T[0] + a_0 / J;
    T[1] = 0;
  }
  /* cc:692-696 */
  double s = v_0 * (T[0] + T[1] + T[2]) +
             a_0 * (1.0 / 2.0 * SQ(T[0]) + T[0] * (T[1] + T[2]) + 1.0 / 2.0 * SQ(T[2])) +
             J * (1.0 / 6.0 * pow(T[0], 3) + 1.0 / 2.0 * SQ(T[0]) * (T[1] + T[2]) -
                  1.0 / 6.0 * pow(T[2], 3) + 1.0 / 2.0 * T[0] * SQ(T[2])) +
             A * (1.0 / 2.0 * SQ(T[1]) + T[1] * T[2]);
  *q = *dir * s; /* cc:699 */
}

static void cumsum7(const double* r, double* t) {
  double acc = r[0];
  t[0] = acc;
  for (int i = 1; i < 7; ++i) {
    acc = acc + r[i];
    t[i] = acc;
  }
}

/* ------------------------------------------------------------------------------------ */
/* cc:82-353                                                                             */
/* ------------------------------------------------------------------------------------ */
int ltpo_opt_switch_times(const ltpo_planner* L, int joint, double q_goal, double q_0, double v_0,
                          double a_0, double V, double t[7], double* dir, unsigned char* mod,
                          unsigned char* kase) {
  const double A = L->a_max[joint], J = L->j_max[joint];
  const double eps = EPS_T;
  double T[7] = {0, 0, 0, 0, 0, 0, 0};
  unsigned char flags = 0;
  *mod = 0; /* cc:95 */

  double q_stop = 0;
  ltpo_opt_braking(L, joint, v_0, a_0, &q_stop, T, dir); /* cc:100 */
  double q_diff = q_goal - (q_0 + q_stop);
  if (fabs(q_diff) < eps) { /* cc:102-107 */
    cumsum7(T, t);
    *kase = LTPO_CASE_BRAKE_ONLY;
    return 1;
  }
  *dir = sgn(q_diff); /* cc:108 */
  if (*dir < 0) {
    v_0 = -v_0;
    a_0 = -a_0;
  }

  double q_brake = 0.0, unused;
  if (v_0 + 0.5 * a_0 * fabs(a_0) / J > V) { /* cc:119-122 */
    *mod = 1;
    flags |= LTPO_F_MOD;
    ltpo_opt_braking(L, joint, v_0 - V, a_0, &q_brake, T, &unused);
  } else { /* cc:125-143 */
    T[0] = (A - a_0) / J;
    T[2] = A / J;
    T[1] = (V - v_0 - 0.5 * T[0] * a_0) / A - 0.5 * (T[0] + T[2]);
    if (T[1] < -eps) {
      double rad = J * (V - v_0) + 0.5 * SQ(a_0);
      if (rad > 0) {
        T[2] = sqrt(rad) / J;
        T[0] = T[2] - a_0 / J;
        T[1] = 0;
        flags |= LTPO_F_NOP2;
      } else {
        for (int i = 0; i < 7; ++i) t[i] = 0.0;
        *kase = LTPO_CASE_DEGENERATE | flags;
        return 1;
      }
    }
  }

  /* cc:147-165 */
  T[4] = A / J;
  T[6] = T[4];
  T[5] = V / A - 1.0 / 2.0 * (T[4] + T[6]);
  if (T[5] < -eps) {
    double rad = V / J;
    if (rad > 0) {
      T[4] = sqrt(rad);
      T[6] = T[4];
      T[5] = 0;
      flags |= LTPO_F_NOP6;
    } else {
      for (int i = 0; i < 7; ++i) t[i] = 0.0;
      *kase = LTPO_CASE_DEGENERATE | flags;
      return 1;
    }
  }

  /* cc:168-190 */
  double part1;
  if (*mod == 1) {
    part1 = q_brake + V * (T[0] + T[1] + T[2]);
  } else {
    part1 = v_0 * (T[0] + T[1] + T[2]) +
            a_0 * (1.0 / 2.0 * SQ(T[0]) + T[0] * (T[1] + T[2]) + 1.0 / 2.0 * SQ(T[2])) +
            J * (1.0 / 6.0 * pow(T[0], 3) + 1.0 / 2.0 * SQ(T[0]) * (T[1] + T[2]) -
                 1.0 / 6.0 * pow(T[2], 3) + 1.0 / 2.0 * T[0] * SQ(T[2])) +
            A * (1.0 / 2.0 * SQ(T[1]) + T[1] * T[2]);
  }
  double part2 = J * (1.0 / 6.0 * pow(T[6], 3) + 1.0 / 2.0 * SQ(T[6]) * (T[5] + T[4]) -
                      1.0 / 6.0 * pow(T[4], 3) + 1.0 / 2.0 * T[6] * SQ(T[4])) +
                 A * (1.0 / 2.0 * SQ(T[5]) + T[5] * T[4]);
  T[3] = ((q_goal - q_0) * *dir - part1 - part2) / V;

  unsigned char base = (unsigned char)(1 + ((flags & LTPO_F_NOP2) ? 1 : 0) + ((flags & LTPO_F_NOP6) ? 2 : 0));

  if (T[3] < -eps) { /* cc:194 */
    if (*mod == 1) { /* cc:195-199 */
      for (int i = 0; i < 7; ++i) t[i] = 0.0;
      *kase = LTPO_CASE_FAIL | flags;
      return 0;
    }
    /* cc:202-223 */
    double rad = (SQ(J) * pow(T[0], 4)) / 2 - (SQ(J) * pow(T[2], 4)) / 4 +
                 (SQ(J) * SQ(T[2]) * SQ(T[4])) / 2 - (SQ(J) * pow(T[4], 4)) / 4 +
                 (SQ(J) * pow(T[6], 4)) / 2 + 2.0 * J * a_0 * pow(T[0], 3) -
                 (2.0 * J * A * pow(T[0], 3)) / 3 - 2.0 * J * A * T[0] * SQ(T[2]) +
                 (2.0 * J * A * pow(T[2], 3)) / 3 + (2.0 * J * A * pow(T[4], 3)) / 3 -
                 2.0 * J * A * SQ(T[4]) * T[6] - (2.0 * J * A * pow(T[6], 3)) / 3 +
                 2.0 * J * v_0 * SQ(T[0]) + 2.0 * SQ(a_0) * SQ(T[0]) - 2.0 * a_0 * A * SQ(T[0]) -
                 2.0 * a_0 * A * SQ(T[2]) + 4 * a_0 * v_0 * T[0] + 2.0 * SQ(A) * SQ(T[2]) +
                 2.0 * SQ(A) * SQ(T[4]) - 4 * A * v_0 * T[0] + 4 * *dir * (q_goal - q_0) * A +
                 2.0 * SQ(v_0);
    if (rad > 0) { /* cc:224-236 */
      T[5] = -(4 * A * T[4] - 2.0 * pow(rad, (1.0 / 2)) + J * SQ(T[2]) - J * SQ(T[4]) +
               2.0 * J * SQ(T[6])) /
             (4 * A);
      T[1] = (-v_0 - a_0 * T[0] - 1.0 / 2.0 * J * SQ(T[0]) + 1.0 / 2.0 * J * SQ(T[2]) +
              1.0 / 2.0 * J * SQ(T[6]) - 1.0 / 2.0 * J * SQ(T[4])) /
                 A -
             T[2] + T[5] + T[4];
      T[3] = 0;
      base = LTPO_CASE_NOP4;
    } else {
      for (int i = 0; i < 7; ++i) t[i] = 0.0;
      *kase = LTPO_CASE_DEGENERATE | flags;
      return 1;
    }

    if (T[5] < -eps || T[1] < -eps) { /* cc:245 */
      double c[5];
      c[0] = 12;
      c[1] = 0;
      c[2] = -24 * SQ(a_0) + 48 * J * v_0;
      c[3] = 48 * *dir * SQ(J) * q_0 - 48 * *dir * SQ(J) * q_goal + 16 * pow(a_0, 3) -
             48 * a_0 * J * v_0;
      c[4] = -3 * pow(a_0, 4) + 12.0 * SQ(a_0) * J * v_0 - 12.0 * SQ(J) * SQ(v_0);
      double r = ltpo_roots(c, 4, 0, 0); /* cc:256-261 */
      T[0] = (2.0 * SQ(r) - 4 * a_0 * r + SQ(a_0) - 2.0 * v_0 * J) / (4 * J * r);
      T[6] = sqrt(4 * SQ(J) * SQ(T[0]) + 8 * a_0 * J * T[0] + 2.0 * SQ(a_0) + 4 * J * v_0) /
             (2.0 * J);
      T[4] = a_0 / J + T[0] + T[6];
      T[1] = 0;
      T[5] = 0;
      base = LTPO_CASE_Q1;

      if (a_0 + T[0] * J > A) { /* cc:273-296 */
        T[0] = (A - a_0) / J;
        T[6] = 1.0 / J *
               (A / 2 +
                sqrt(9 * SQ(A) +
                     6 * sqrt(-12.0 * A * pow(J, 3) * pow(T[0], 3) + 9 * SQ(a_0) * SQ(J) * SQ(T[0]) -
                              18 * a_0 * A * SQ(J) * SQ(T[0]) + 9 * SQ(A) * SQ(J) * SQ(T[0]) +
                              36 * a_0 * SQ(J) * T[0] * v_0 - 72.0 * A * *dir * SQ(J) * q_0 +
                              72.0 * A * *dir * SQ(J) * q_goal - 36 * A * SQ(J) * T[0] * v_0 +
                              3 * pow(A, 4) + 36 * SQ(J) * SQ(v_0))) /
                    6.0 -
                A);
        T[4] = T[6] + A / J;
        T[1] = -(-J * SQ(T[4]) - 2.0 * J * T[4] * T[6] + J * SQ(T[6]) + a_0 * T[0] + A * T[0] +
                 2.0 * A * T[4] + 2.0 * A * T[6] + 2.0 * v_0) /
               (2.0 * A);
        T[5] = 0;
        base = LTPO_CASE_Q1_P2;
      }

      if (T[6] * J > A) { /* cc:299-333 */
        T[6] = A / J;
        c[0] = 12;
        c[1] = -24 * A;
        c[2] = -12.0 * SQ(a_0) + 12.0 * SQ(A) + 24 * J * v_0;
        c[3] = 0;
        c[4] = 24 * *dir * SQ(J) * q_0 * A - 24 * *dir * SQ(J) * q_goal * A + 3 * pow(a_0, 4) +
               8 * pow(a_0, 3) * A + 6 * SQ(a_0) * SQ(A) - 12.0 * SQ(a_0) * J * v_0 -
               24 * a_0 * J * v_0 * A - 12.0 * SQ(A) * J * v_0 + 12.0 * SQ(J) * SQ(v_0);
        r = ltpo_roots(c, 4, 0, 0); /* cc:316-321 */
        T[0] = (r - a_0 - A) / J;
        T[4] = (a_0 + A) / J + T[0];
        T[5] = (SQ(J) * SQ(T[0]) + 2.0 * SQ(J) * T[0] * T[4] - SQ(J) * SQ(T[4]) +
                2.0 * a_0 * J * T[0] + 2.0 * a_0 * J * T[4] - SQ(A) + 2.0 * J * v_0) /
               (2.0 * J * A);
        T[1] = 0;
        if (base == LTPO_CASE_Q1_P2) flags |= LTPO_F_BOTH;
        base = LTPO_CASE_Q2;
      }
      T[2] = 0; /* cc:335-336 */
      T[3] = 0;
    }
  }
  /* cc:340-348 */
  for (int i = 0; i < 7; ++i) {
    if (T[i] < -eps) {
      *kase = LTPO_CASE_FAIL_UNTOUCHED | flags;
      return 0; /* t is NOT written (cc:344) */
    } else if (T[i] < 0.0 && T[i] >= -eps) {
      T[i] = 0.0;
    }
  }
  cumsum7(T, t); /* cc:351 */
  *kase = base | flags;
  return 1;
}

/* ------------------------------------------------------------------------------------ */
/* cc:358-645                                                                            */
/* ------------------------------------------------------------------------------------ */
typedef struct {
  const ltpo_planner* L;
  int joint;
  double q_goal, q_0, v_0, a_0, dir, t_req;
} ts_ctx;

/* the block repeated after every candidate: cc:398-405, 439-446, 475-482, ... */
static int ts_try(const ts_ctx* c, double V, double scaled_t[7], unsigned char* mod,
                  unsigned char* kase) {
  const double tol = 0.1; /* cc:370 */
  if (!isnan(V) && V > 0) {
    double trash;
    int ok = ltpo_opt_switch_times(c->L, c->joint, c->q_goal, c->q_0, c->dir * c->v_0,
                                   c->dir * c->a_0, V, scaled_t, &trash, mod, kase);
    if (ok && c->t_req - scaled_t[6] < tol && c->t_req - scaled_t[6] > -tol / 10) return 1;
  }
  return 0;
}

int ltpo_time_scaling(const ltpo_planner* L, int joint, double q_goal, double q_0, double v_0,
                      double a_0, double dir, double tr, double scaled_t[7], double* v_drive,
                      unsigned char* mod, unsigned char* ts_case, unsigned char* final_case) {
  const double A = L->a_max[joint], J = L->j_max[joint];
  if (dir < 0) { /* cc:372-375 */
    v_0 = -v_0;
    a_0 = -a_0;
  }
  ts_ctx c = {L, joint, q_goal, q_0, v_0, a_0, dir, tr};
  double V, r, p[7];
  *final_case = LTPO_CASE_FAIL;

  /* attempt 1, cc:378-396 */
  V = (A * J * tr / 2 - SQ(a_0) / 4 + a_0 * A / 2 - SQ(A) / 2 + v_0 * J / 2 -
       sqrt(36 * SQ(A) * SQ(J) * SQ(tr) - 36 * SQ(a_0) * A * J * tr + 72.0 * a_0 * SQ(A) * J * tr -
            72.0 * pow(A, 3) * J * tr + 144 * A * dir * SQ(J) * q_0 -
            144 * A * dir * SQ(J) * q_goal + 72.0 * A * SQ(J) * v_0 * tr - 9 * pow(a_0, 4) +
            12.0 * pow(a_0, 3) * A + 36 * SQ(a_0) * SQ(A) + 36 * SQ(a_0) * J * v_0 -
            72.0 * a_0 * pow(A, 3) - 72.0 * a_0 * A * J * v_0 + 36 * pow(A, 4) -
            36 * SQ(J) * SQ(v_0)) /
           12) /
      J;
  *v_drive = V;
  if (ts_try(&c, V, scaled_t, mod, final_case)) { *ts_case = 1; return 1; }

  /* attempt 2, cc:408-436. w, h, g are sub-expressions that the reference spells out
   * several times; each is evaluated with the reference's own operation order. */
  {
    const double w = (v_0 + (a_0 * (a_0 - A)) / (2.0 * J)) / A;
    const double h = A / (2.0 * J);
    const double g = (a_0 - A) / (2.0 * J);
    const double sA = a_0 + A;
    V = -(dir * (q_0 - q_goal) -
          J * (pow(sA, 3) / (6 * pow(J, 3)) - pow(A, 3) / (6 * pow(J, 3)) +
               (SQ(A) * sA) / (2.0 * pow(J, 3)) + (SQ(sA) * (w + h + g)) / (2.0 * SQ(J))) +
          a_0 * (SQ(sA) / (2.0 * SQ(J)) + SQ(A) / (2.0 * SQ(J)) + (sA * (w + h + g)) / J) -
          A * (SQ(w - h + g) / 2 + (A * (w - h + g)) / J) + v_0 * (w + sA / J + h + g)) /
        (h - v_0 / A + A * ((w - h + g) / A + 1.0 / J) -
         (SQ(a_0) + 2.0 * a_0 * A + 4 * SQ(A) - 2.0 * J * tr * A + 2.0 * J * v_0) / (2.0 * A * J) +
         SQ(sA) / (2.0 * A * J) - (a_0 * sA) / (A * J));
  }
  *v_drive = V;
  if (ts_try(&c, V, scaled_t, mod, final_case)) { *ts_case = 2; return 1; }

  /* attempt 3, cc:449-473 */
  p[0] = 3;
  p[1] = 12.0 * A;
  p[2] = -24 * A * J * tr - 12.0 * SQ(a_0) - 24 * a_0 * A + 12.0 * SQ(A) + 24 * J * v_0;
  p[3] = 0;
  p[4] = 48 * SQ(a_0) * A * J * tr - 96 * dir * SQ(J) * A * q_0 + 96 * dir * SQ(J) * A * q_goal -
         96 * A * SQ(J) * v_0 * tr + 12.0 * pow(a_0, 4) + 16 * pow(a_0, 3) * A -
         24 * SQ(a_0) * SQ(A) - 48 * SQ(a_0) * J * v_0 + 48 * SQ(A) * J * v_0 +
         48 * SQ(J) * SQ(v_0);
  r = ltpo_roots(p, 4, 0, 0);
  V = (-2.0 * SQ(a_0) + 4 * J * v_0 + SQ(r)) / (4 * J);
  *v_drive = V;
  if (ts_try(&c, V, scaled_t, mod, final_case)) { *ts_case = 3; return 1; }

  /* attempt 4, cc:485-514 (the dimensionally odd terms are the reference's) */
  p[0] = 12;
  p[1] = 24 * A;
  p[2] = -24 * A * J * tr + 24 * SQ(a_0) - 48 * a_0 * A + 24 * SQ(A) - 24 * J * v_0 + 12.0 * a_0 -
         12.0 * A;
  p[3] = 0;
  p[4] = -24 * dir * SQ(J) * A * q_0 + 24 * dir * SQ(J) * A * q_goal + 9 * pow(a_0, 4) -
         12.0 * pow(a_0, 3) * A - 24 * SQ(a_0) * J * v_0 + 48 * a_0 * A * J * v_0 +
         4 * pow(A, 4) - 24 * SQ(A) * J * v_0 + 12.0 * SQ(J) * SQ(v_0) + 6 * pow(a_0, 3) +
         6 * SQ(a_0) * A - 12.0 * a_0 * SQ(A) - 12.0 * a_0 * J * v_0 + 12.0 * A * J * v_0 +
         4 * a_0 * A - 4 * SQ(A);
  r = ltpo_roots(p, 4, 0, 0);
  V = SQ(r) / J;
  *v_drive = V;
  if (ts_try(&c, V, scaled_t, mod, final_case)) { *ts_case = 4; return 1; }

  /* attempt 5, cc:526-541 */
  p[0] = (144 * J * tr + 144 * a_0);
  p[1] = (-72.0 * SQ(J) * SQ(tr) - 144 * a_0 * J * tr + 36 * SQ(a_0) - 216 * J * v_0);
  p[2] = (144 * dir * SQ(J) * q_0 - 144 * dir * SQ(J) * q_goal + 48 * pow(a_0, 3) -
          144 * a_0 * J * v_0);
  p[3] = (-144 * dir * pow(J, 3) * q_0 * tr + 144 * dir * pow(J, 3) * q_goal * tr -
          48 * pow(a_0, 3) * J * tr - 144 * a_0 * dir * SQ(J) * q_0 +
          144 * a_0 * dir * SQ(J) * q_goal + 144 * a_0 * SQ(J) * v_0 * tr + 6 * pow(a_0, 4) -
          72.0 * SQ(a_0) * J * v_0 + 216 * SQ(J) * SQ(v_0));
  p[4] = 0;
  p[5] = -72.0 * SQ(dir) * pow(J, 4) * SQ(q_0) + 144 * SQ(dir) * pow(J, 4) * q_0 * q_goal -
         72.0 * SQ(dir) * pow(J, 4) * SQ(q_goal) - 48 * pow(a_0, 3) * dir * SQ(J) * q_0 +
         48 * pow(a_0, 3) * dir * SQ(J) * q_goal + 144 * a_0 * dir * pow(J, 3) * q_0 * v_0 -
         144 * a_0 * dir * pow(J, 3) * q_goal * v_0 + pow(a_0, 6) - 6 * pow(a_0, 4) * J * v_0 +
         36 * SQ(a_0) * SQ(J) * SQ(v_0) - 72.0 * pow(J, 3) * pow(v_0, 3);
  r = ltpo_roots(p, 5, 0, 0);
  V = SQ(r) / J;
  *v_drive = V;
  if (ts_try(&c, V, scaled_t, mod, final_case)) { *ts_case = 5; return 1; }

  /* attempt 6, cc:553-567 */
  p[0] = 3;
  p[1] = -6 * sqrt(2) * A;
  p[2] = (12.0 * A * J * tr - 6 * SQ(a_0) - 12.0 * a_0 * A - 6 * SQ(A) - 12.0 * J * v_0);
  p[3] = 0;
  p[4] = -12.0 * SQ(a_0) * A * J * tr - 24 * dir * SQ(J) * A * q_0 + 24 * dir * SQ(J) * A * q_goal -
         24 * A * SQ(J) * v_0 * tr + 3 * pow(a_0, 4) + 4 * pow(a_0, 3) * A + 6 * SQ(a_0) * SQ(A) +
         12.0 * SQ(a_0) * J * v_0 + 12.0 * SQ(A) * J * v_0 + 12.0 * SQ(J) * SQ(v_0);
  r = ltpo_roots(p, 4, 0, 0);
  V = -(SQ(r) - SQ(a_0) - 2.0 * J * v_0) / (2.0 * J);
  *v_drive = V;
  if (ts_try(&c, V, scaled_t, mod, final_case)) { *ts_case = 6; return 1; }

  /* attempt 7, cc:579-593 */
  p[0] = 12;
  p[1] = -24 * A;
  p[2] = (24 * A * J * tr - 12.0 * SQ(a_0) - 24 * a_0 * A - 12.0 * SQ(A) - 24 * J * v_0);
  p[3] = 0;
  p[4] = 24 * dir * SQ(J) * A * q_0 - 24 * dir * SQ(J) * A * q_goal + 3 * pow(a_0, 4) +
         8 * pow(a_0, 3) * A + 6 * SQ(a_0) * SQ(A) + 12.0 * SQ(a_0) * J * v_0 +
         24 * a_0 * A * J * v_0 + 12.0 * SQ(A) * J * v_0 + 12.0 * SQ(J) * SQ(v_0);
  r = ltpo_roots(p, 4, 0, 0);
  V = SQ(r) / J;
  *v_drive = V;
  if (ts_try(&c, V, scaled_t, mod, final_case)) { *ts_case = 7; return 1; }

  /* attempt 8, cc:606-629 */
  p[0] = 144;
  p[1] = (-144 * J * tr + 144 * a_0);
  p[2] = (72.0 * SQ(J) * SQ(tr) - 144 * a_0 * J * tr - 36 * SQ(a_0) - 216 * J * v_0);
  p[3] = (-144 * dir * SQ(J) * q_0 + 144 * dir * SQ(J) * q_goal - 48 * pow(a_0, 3) -
          144 * a_0 * J * v_0);
  p[4] = (144 * dir * pow(J, 3) * q_0 * tr - 144 * dir * pow(J, 3) * q_goal * tr +
          48 * pow(a_0, 3) * J * tr - 144 * a_0 * dir * SQ(J) * q_0 +
          144 * a_0 * dir * SQ(J) * q_goal + 144 * a_0 * SQ(J) * v_0 * tr + 6 * pow(a_0, 4) +
          72.0 * SQ(a_0) * J * v_0 + 216 * SQ(J) * SQ(v_0));
  p[5] = 0;
  p[6] = 72.0 * SQ(dir) * pow(J, 4) * SQ(q_0) - 144 * SQ(dir) * pow(J, 4) * q_0 * q_goal +
         72.0 * SQ(dir) * pow(J, 4) * SQ(q_goal) + 48 * pow(a_0, 3) * dir * SQ(J) * q_0 -
         48 * pow(a_0, 3) * dir * SQ(J) * q_goal + 144 * a_0 * dir * pow(J, 3) * q_0 * v_0 -
         144 * a_0 * dir * pow(J, 3) * q_goal * v_0 - pow(a_0, 6) - 6 * pow(a_0, 4) * J * v_0 -
         36 * SQ(a_0) * SQ(J) * SQ(v_0) - 72.0 * pow(J, 3) * pow(v_0, 3);
  r = ltpo_roots(p, 6, 0, 0);
  V = SQ(r) / J;
  *v_drive = V;
  if (ts_try(&c, V, scaled_t, mod, final_case)) { *ts_case = 8; return 1; }

  /* cc:641-644 */
  *mod = 0;
  for (int i = 0; i < 7; ++i) scaled_t[i] = 0.0;
  *v_drive = L->v_max[joint];
  *ts_case = 9;
  *final_case = LTPO_CASE_FAIL;
  return 0;
}

/* ------------------------------------------------------------------------------------ */
/* cc:14-55 for one problem (joint-contiguous arrays)                                    */
/* ------------------------------------------------------------------------------------ */
static int finite7(const double* t) {
  for (int i = 0; i < 7; ++i)
    if (!isfinite(t[i])) return 0;
  return 1;
}

static int solve_one(const ltpo_planner* L, const double* q_goal, const double* q_0,
                     const double* v_0, const double* a_0, double* t_opt, double* t_scaled,
                     double* dir, double* v_drive, unsigned char* mod, unsigned char* opt_case,
                     unsigned char* ts_case, unsigned char* final_case, int* slowest, int* traj_len) {
  const int dof = L->dof;
  *slowest = -1;
  *traj_len = 0;
  for (int i = 0; i < dof; ++i) {
    v_drive[i] = L->v_max[i]; /* cc:42 */
    mod[i] = 0;
    dir[i] = 0;
    opt_case[i] = 255;
    ts_case[i] = 255;
    final_case[i] = 255;
  }
  memset(t_opt, 0, sizeof(double) * 7 * (size_t)dof); /* value-initialised, cc:18-20 */
  memset(t_scaled, 0, sizeof(double) * 7 * (size_t)dof);
  if (!ltpo_check_inputs(L, q_0, v_0, a_0)) return 0; /* cc:14-15 */
  for (int i = 0; i < dof; ++i) { /* cc:27-30 */
    int ok = ltpo_opt_switch_times(L, i, q_goal[i], q_0[i], v_0[i], a_0[i], L->v_max[i],
                                   t_opt + 7 * i, dir + i, mod + i, opt_case + i);
    if (!ok) return 0;
  }
  double t_required = -1; /* cc:31-39 */
  for (int i = 0; i < dof; ++i) {
    if (t_opt[7 * i + 6] > t_required) {
      t_required = t_opt[7 * i + 6];
      *slowest = i;
    }
  }
  if (*slowest == -1) return 0;
  for (int i = 0; i < dof; ++i) { /* cc:42-48 */
    if (i == *slowest) {
      ts_case[i] = 0;
      final_case[i] = opt_case[i];
      continue;
    }
    ltpo_time_scaling(L, i, q_goal[i], q_0[i], v_0[i], a_0[i], dir[i], t_required,
                      t_scaled + 7 * i, v_drive + i, mod + i, ts_case + i, final_case + i);
    if (ts_case[i] == 9) final_case[i] = opt_case[i];
  }
  for (int i = 0; i < dof; ++i) { /* cc:50-55 */
    double m = t_scaled[7 * i];
    for (int k = 1; k < 7; ++k)
      if (m < t_scaled[7 * i + k]) m = t_scaled[7 * i + k]; /* std::max_element */
    if (m <= 0.0) memcpy(t_scaled + 7 * i, t_opt + 7 * i, 56);
  }
  /* cc:716-719; a non-finite time would be undefined behaviour in the reference */
  int len = 0;
  for (int i = 0; i < dof; ++i) {
    if (!finite7(t_scaled + 7 * i) || t_scaled[7 * i + 6] / L->ts > 2.0e9) return 1;
    int li = (int)ceil(t_scaled[7 * i + 6] / L->ts) + 1;
    if (li > len) len = li;
  }
  *traj_len = len;
  return 1;
}

/* ------------------------------------------------------------------------------------ */
/* cc:706-841                                                                            */
/* ------------------------------------------------------------------------------------ */
int ltpo_get_trajectory(const ltpo_planner* L, const double* t7, const double* dir,
                        const unsigned char* mod, const double* q_0, const double* v_0,
                        const double* a_0, const double* v_drive, int64_t stride, double* q,
                        double* v, double* a, double* j) {
  const int dof = L->dof;
  const double Ts = L->ts;
  int len = 0; /* cc:716-719 */
  for (int i = 0; i < dof; ++i) {
    int li = (int)ceil(t7[7 * i + 6] / Ts) + 1;
    if (li > len) len = li;
  }
  if (len > stride) return -len;
  for (int jt = 0; jt < dof; ++jt) {
    const double* t = t7 + 7 * jt;
    double* jj = j + jt * stride;
    double* aa = a + jt * stride;
    double* vv = v + jt * stride;
    double* qq = q + jt * stride;
    for (int i = 0; i < len; ++i) jj[i] = aa[i] = vv[i] = qq[i] = 0.0; /* cc:725-728 */
    /* cc:734-744 */
    static const int prof_std[7] = {1, 0, -1, 0, -1, 0, 1};
    static const int prof_mod[7] = {-1, 0, 1, 0, -1, 0, 1};
    const int* prof = (mod[jt] == 1) ? prof_mod : prof_std;
    double jp[7], fr[7];
    int s[7];
    for (int k = 0; k < 7; ++k) jp[k] = dir[jt] * L->j_max[jt] * prof[k];
    for (int k = 0; k < 7; ++k) fr[k] = t[k] - Ts * floor(t[k] / Ts); /* cc:746-748 */
    for (int k = 0; k < 7; ++k) /* cc:751-757: floor for even k, ceil for odd k */
      s[k] = (k & 1) ? (int)ceil(t[k] / Ts) : (int)floor(t[k] / Ts);
    /* cc:759-766, clipped to the array (the reference would run past the end) */
#define FILL(lo, hi, val)                                  \
  for (int i_ = ((lo) < 0 ? 0 : (lo)); i_ < (hi) && i_ < len; ++i_) jj[i_] = (val)
#define ADDJ(idx, val)                                     \
  do {                                                     \
    int k_ = (idx);                                        \
    if (k_ >= 0 && k_ < len) jj[k_] = jj[k_] + (val);      \
  } while (0)
    if (s[0] > 0) FILL(0, s[0], jp[0]);
    for (int k = 1; k < 7; ++k)
      if (s[k] - s[k - 1] > 0) FILL(s[k - 1], s[k], jp[k]);
    /* cc:768-783 */
    if (s[2] >= s[1]) {
      ADDJ(s[0] + 1, fr[0] / Ts * jp[0]);
      if (s[1] > 0) ADDJ(s[1], (1 - fr[1] / Ts) * jp[2]);
      ADDJ(s[2] + 1, fr[2] / Ts * jp[2]);
    } else {
      if (s[1] > 0) {
        int k_ = s[1];
        if (k_ < len) jj[k_] = jj[k_] + fr[0] / Ts * jp[0] + (fr[2] - fr[0]) / Ts * jp[2];
      }
    }
    if (s[3] > 0) ADDJ(s[3], (1 - fr[3] / Ts) * jp[4]); /* cc:786-788 */
    if (s[2] - s[0] > 0) {                              /* cc:790-800 */
      ADDJ(s[4] + 1, fr[4] / Ts * jp[4]);
    } else {
      if (s[4] > 0) {
        int k_ = s[4];
        if (k_ < len)
          jj[k_] = jj[k_] + fr[4] / Ts * jp[4] + fr[0] / Ts * jp[0] + (fr[2] - fr[0]) / Ts * jp[2];
      }
    }
    if (s[5] > 0) ADDJ(s[5], (1 - fr[5] / Ts) * jp[6]); /* cc:803-805 */
    ADDJ(s[6] + 1, fr[6] / Ts * jp[6]);                 /* cc:807 */
#undef FILL
#undef ADDJ
    /* cc:810-831 */
    aa[0] = a_0[jt] + Ts * jj[0];
    vv[0] = v_0[jt] + Ts * aa[0];
    qq[0] = q_0[jt] + Ts * vv[0];
    int phase4 = s[3] - s[2] > 2;
    for (int i = 1; i < len; ++i) {
      if (i <= s[6])
        aa[i] = aa[i - 1] + Ts * jj[i];
      else
        aa[i] = 0.0;
      if (phase4 && i >= s[2] + 1 && i < s[3] - 1)
        vv[i] = v_drive[jt] * dir[jt];
      else if (i <= s[6])
        vv[i] = vv[i - 1] + Ts * aa[i];
      else
        vv[i] = 0.0;
      qq[i] = qq[i - 1] + Ts * vv[i];
    }
  }
  return len;
}

/* ------------------------------------------------------------------------------------ */
/* cc:7-63                                                                               */
/* ------------------------------------------------------------------------------------ */
int ltpo_plan(const ltpo_planner* L, const double* q_goal, const double* q_0, const double* v_0,
              const double* a_0, int64_t stride, double* q, double* v, double* a, double* j,
              int* length) {
  const int dof = L->dof;
  double* buf = (double*)malloc(sizeof(double) * (size_t)dof * 16);
  unsigned char* cb = (unsigned char*)malloc((size_t)dof * 4);
  double *t_opt = buf, *t_scaled = buf + 7 * dof, *dir = buf + 14 * dof, *vd = buf + 15 * dof;
  int slowest, len;
  *length = -1;
  int reached = solve_one(L, q_goal, q_0, v_0, a_0, t_opt, t_scaled, dir, vd, cb, cb + dof,
                          cb + 2 * dof, cb + 3 * dof, &slowest, &len);
  int ok = 0;
  if (reached && len > 0) {
    int own = 0;
    if (!q || len > stride) { /* caller only wants flags: sample into scratch */
      stride = len;
      q = (double*)malloc(sizeof(double) * (size_t)dof * (size_t)len * 4);
      v = q + (size_t)dof * len;
      a = v + (size_t)dof * len;
      j = a + (size_t)dof * len;
      own = 1;
    }
    len = ltpo_get_trajectory(L, t_scaled, dir, cb, q_0, v_0, a_0, vd, stride, q, v, a, j);
    *length = len;
    ok = 1;
    for (int i = 0; i < dof; ++i) /* cc:59-61 */
      if (q[i * stride + len - 1] < L->q_min[i] || q[i * stride + len - 1] > L->q_max[i]) ok = 0;
    if (own) free(q);
  }
  free(buf);
  free(cb);
  return ok;
}

/* ------------------------------------------------------------------------------------ */
/* batched legs                                                                          */
/* ------------------------------------------------------------------------------------ */
typedef struct job {
  void (*fn)(struct job*, int64_t, int64_t);
  int64_t lo, hi;
  const ltpo_planner* L;
  const int* joint;
  const double *q_goal, *q_0, *v_0, *a_0, *v_in, *dir_in, *t_req;
  double *t_opt, *t_scaled, *dir, *v_drive, *q;
  unsigned char *mod, *c0, *c1, *c2, *ok;
  int *slowest, *traj_len;
  double acc;
} job;

static void* job_main(void* p) {
  job* j = (job*)p;
  j->fn(j, j->lo, j->hi);
  return 0;
}

static double run_jobs(job* proto, int64_t n, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if (n < threads) threads = (int)(n > 0 ? n : 1);
  job* js = (job*)malloc(sizeof(job) * (size_t)threads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  for (int t = 0; t < threads; ++t) {
    js[t] = *proto;
    js[t].lo = n * t / threads;
    js[t].hi = n * (t + 1) / threads;
    js[t].acc = 0;
    if (threads > 1) pthread_create(&th[t], 0, job_main, &js[t]);
    else job_main(&js[t]);
  }
  double acc = 0;
  for (int t = 0; t < threads; ++t) {
    if (threads > 1) pthread_join(th[t], 0);
    acc += js[t].acc;
  }
  free(js);
  free(th);
  return acc;
}

void ltpo_opt_braking_items(const ltpo_planner* L, int64_t n, const int* joint, const double* v_0,
                            const double* a_0, double* q, double* t_rel3, double* dir) {
  for (int64_t i = 0; i < n; ++i) {
    double T[7] = {0, 0, 0, 0, 0, 0, 0};
    ltpo_opt_braking(L, joint ? joint[i] : 0, v_0[i], a_0[i], q + i, T, dir + i);
    t_rel3[3 * i] = T[0];
    t_rel3[3 * i + 1] = T[1];
    t_rel3[3 * i + 2] = T[2];
  }
}

static void ost_items(job* j, int64_t lo, int64_t hi) {
  for (int64_t i = lo; i < hi; ++i) {
    double t[7] = {0, 0, 0, 0, 0, 0, 0};
    j->ok[i] = (unsigned char)ltpo_opt_switch_times(j->L, j->joint ? j->joint[i] : 0, j->q_goal[i],
                                                    j->q_0[i], j->v_0[i], j->a_0[i], j->v_in[i], t,
                                                    j->dir + i, j->mod + i, j->c0 + i);
    memcpy(j->t_opt + 7 * i, t, 56);
  }
}

void ltpo_opt_switch_times_items(const ltpo_planner* L, int64_t n, const int* joint,
                                 const double* q_goal, const double* q_0, const double* v_0,
                                 const double* a_0, const double* v_drive, double* t7, double* dir,
                                 unsigned char* mod, unsigned char* kase, unsigned char* ok,
                                 int threads) {
  job p;
  memset(&p, 0, sizeof p);
  p.fn = ost_items; p.L = L; p.joint = joint; p.q_goal = q_goal; p.q_0 = q_0; p.v_0 = v_0;
  p.a_0 = a_0; p.v_in = v_drive; p.t_opt = t7; p.dir = dir; p.mod = mod; p.c0 = kase; p.ok = ok;
  run_jobs(&p, n, threads);
}

static void ts_items(job* j, int64_t lo, int64_t hi) {
  for (int64_t i = lo; i < hi; ++i) {
    double t[7] = {0, 0, 0, 0, 0, 0, 0};
    unsigned char m = 0;
    j->ok[i] = (unsigned char)ltpo_time_scaling(j->L, j->joint ? j->joint[i] : 0, j->q_goal[i],
                                                j->q_0[i], j->v_0[i], j->a_0[i], j->dir_in[i],
                                                j->t_req[i], t, j->v_drive + i, &m, j->c1 + i,
                                                j->c2 + i);
    j->mod[i] = m;
    memcpy(j->t_scaled + 7 * i, t, 56);
  }
}

void ltpo_time_scaling_items(const ltpo_planner* L, int64_t n, const int* joint,
                             const double* q_goal, const double* q_0, const double* v_0,
                             const double* a_0, const double* dir, const double* t_required,
                             double* t7, double* v_drive, unsigned char* mod,
                             unsigned char* ts_case, unsigned char* final_case,
                             unsigned char* ok, int threads) {
  job p;
  memset(&p, 0, sizeof p);
  p.fn = ts_items; p.L = L; p.joint = joint; p.q_goal = q_goal; p.q_0 = q_0; p.v_0 = v_0;
  p.a_0 = a_0; p.dir_in = dir; p.t_req = t_required; p.t_scaled = t7; p.v_drive = v_drive;
  p.mod = mod; p.c1 = ts_case; p.c2 = final_case; p.ok = ok;
  run_jobs(&p, n, threads);
}

static void solve_items(job* j, int64_t lo, int64_t hi) {
  const int dof = j->L->dof;
  for (int64_t p = lo; p < hi; ++p) {
    const int64_t o = p * dof;
    j->ok[p] = (unsigned char)solve_one(j->L, j->q_goal + o, j->q_0 + o, j->v_0 + o, j->a_0 + o,
                                        j->t_opt + 7 * o, j->t_scaled + 7 * o, j->dir + o,
                                        j->v_drive + o, j->mod + o, j->c0 + o, j->c1 + o,
                                        j->c2 + o, j->slowest + p, j->traj_len + p);
  }
}

void ltpo_solve_batch(const ltpo_planner* L, int64_t n, const double* q_goal, const double* q_0,
                      const double* v_0, const double* a_0, double* t_opt, double* t_scaled,
                      double* dir, double* v_drive, unsigned char* mod, unsigned char* opt_case,
                      unsigned char* ts_case, unsigned char* final_case, int* slowest,
                      int* traj_len, unsigned char* reached, int threads) {
  job p;
  memset(&p, 0, sizeof p);
  p.fn = solve_items; p.L = L; p.q_goal = q_goal; p.q_0 = q_0; p.v_0 = v_0; p.a_0 = a_0;
  p.t_opt = t_opt; p.t_scaled = t_scaled; p.dir = dir; p.v_drive = v_drive; p.mod = mod;
  p.c0 = opt_case; p.c1 = ts_case; p.c2 = final_case; p.slowest = slowest; p.traj_len = traj_len;
  p.ok = reached;
  run_jobs(&p, n, threads);
}

static void plan_items(job* j, int64_t lo, int64_t hi) {
  const int dof = j->L->dof;
  double acc = 0;
  double* scratch = 0;
  int64_t cap = 0;
  for (int64_t p = lo; p < hi; ++p) {
    const int64_t o = p * dof;
    int len = -1;
    /* first try with the scratch we have; grow on demand */
    int ok = 0;
    for (;;) {
      ok = ltpo_plan(j->L, j->q_goal + o, j->q_0 + o, j->v_0 + o, j->a_0 + o, cap,
                     cap ? scratch : 0, cap ? scratch + dof * cap : 0,
                     cap ? scratch + 2 * dof * cap : 0, cap ? scratch + 3 * dof * cap : 0, &len);
      if (len > cap) { /* sampled into a private buffer inside ltpo_plan; keep ours big enough */
        cap = (int64_t)len + 256;
        free(scratch);
        scratch = (double*)malloc(sizeof(double) * (size_t)dof * (size_t)cap * 4);
        continue;
      }
      break;
    }
    j->ok[p] = (unsigned char)ok;
    j->traj_len[p] = len;
    if (len > 0)
      for (int i = 0; i < dof; ++i) acc += scratch[i * cap + len - 1];
  }
  free(scratch);
  j->acc = acc;
}

double ltpo_plan_batch(const ltpo_planner* L, int64_t n, const double* q_goal, const double* q_0,
                       const double* v_0, const double* a_0, unsigned char* success, int* length,
                       int threads) {
  job p;
  memset(&p, 0, sizeof p);
  p.fn = plan_items; p.L = L; p.q_goal = q_goal; p.q_0 = q_0; p.v_0 = v_0; p.a_0 = a_0;
  p.ok = success; p.traj_len = length;
  return run_jobs(&p, n, threads);
}
