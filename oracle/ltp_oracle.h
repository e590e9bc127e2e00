/* TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement, in plain C, of the reference's
 * planning hot path (yannickBurkhardt/LongTermPlanner, src/long_term_planner.cc and
 * include/long_term_planner/roots.h). Each function cites the reference lines it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this. The product library (longtermplanner_b200/csrc) never includes, links or
 * calls anything in this directory.
 *
 * Pinning: the restatement is checked bit-for-bit (times, dir, mod, v_drive, success,
 * trajectory length and samples) against oracle/_ref/libltp_ref.so, which is the
 * reference's own UNMODIFIED .cc compiled by g++ with the same flags, and against the
 * golden tables of the reference's tests (tests/golden/). The one part that cannot be
 * pinned is the Eigen 3.4 eigenvalue solver behind roots.h:32 (Eigen is absent from this
 * container): for the root-solver branches parity is "unpinned" beyond the reference's
 * own 1e-5 golden vector (tests/src/roots_tests.cc:10-31).
 */
#ifndef LTP_ORACLE_H
#define LTP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Case byte shared by oracle and GPU (the reference emits no case id):
 *   low nibble: 0 BRAKE_ONLY cc:102-107 | 1..4 cruise phase exists (1: P2,P6; 2: no P2;
 *               3: no P6; 4: neither) | 5 no cruise phase, closed form cc:202-236 |
 *               6 quartic #1 cc:246-270 | 7 quartic #1 then P2 re-inserted cc:273-296 |
 *               8 quartic #2 cc:299-333 | 13 failure at the final safety check, t NOT written
 *               (false, cc:340-344) | 14 "should never occur" zero return (true) |
 *               15 failure with t zeroed (false, cc:195-199; also the time-scaling reset)
 *   0x10 modified jerk profile (cc:119-122)   0x20 both cc:273 and cc:299 fired
 *   0x40 no-P2 branch cc:131-137 was taken    0x80 no-P6 branch cc:153-159 was taken */
enum {
  LTPO_CASE_BRAKE_ONLY = 0, LTPO_CASE_NOP4 = 5, LTPO_CASE_Q1 = 6, LTPO_CASE_Q1_P2 = 7,
  LTPO_CASE_Q2 = 8, LTPO_CASE_FAIL_UNTOUCHED = 13, LTPO_CASE_DEGENERATE = 14, LTPO_CASE_FAIL = 15,
  LTPO_F_MOD = 0x10, LTPO_F_BOTH = 0x20, LTPO_F_NOP2 = 0x40, LTPO_F_NOP6 = 0x80
};
/* ts_case: 0 slowest joint (not scaled, cc:44-46); 1..8 accepted attempt; 9 all failed
 * (cc:641-644, then the cc:50-55 fallback); 255 plan aborted before time scaling. */

typedef struct ltpo_planner ltpo_planner;

ltpo_planner* ltpo_create(int dof, double t_sample, const double* q_min, const double* q_max,
                          const double* v_max, const double* a_max, const double* j_max);
void ltpo_destroy(ltpo_planner* L);

int ltpo_check_inputs(const ltpo_planner* L, const double* q_0, const double* v_0, const double* a_0);
void ltpo_opt_braking(const ltpo_planner* L, int joint, double v_0, double a_0, double* q,
                      double t_rel[7], double* dir);
int ltpo_opt_switch_times(const ltpo_planner* L, int joint, double q_goal, double q_0, double v_0,
                          double a_0, double v_drive, double t[7], double* dir,
                          unsigned char* mod, unsigned char* kase);
int ltpo_time_scaling(const ltpo_planner* L, int joint, double q_goal, double q_0, double v_0,
                      double a_0, double dir, double t_required, double scaled_t[7],
                      double* v_drive, unsigned char* mod, unsigned char* ts_case,
                      unsigned char* final_case);
/* roots.h:22-50; coeffs highest power first; returns the smallest real root > 1e-7 or +inf */
double ltpo_roots(const double* coeffs, int deg, double* re, double* im);
/* record every polynomial handed to ltpo_roots (8 doubles each: degree, coefficients);
 * buf = NULL stops. Not thread-safe: trace single-threaded runs only. */
void ltpo_trace_roots(double* buf, int64_t capacity);
int64_t ltpo_trace_count(void);

/* item-array forms of the per-joint primitives (joint may be NULL -> joint 0) */
void ltpo_opt_braking_items(const ltpo_planner* L, int64_t n, const int* joint, const double* v_0,
                            const double* a_0, double* q, double* t_rel3, double* dir);
void ltpo_opt_switch_times_items(const ltpo_planner* L, int64_t n, const int* joint,
                                 const double* q_goal, const double* q_0, const double* v_0,
                                 const double* a_0, const double* v_drive, double* t7, double* dir,
                                 unsigned char* mod, unsigned char* kase, unsigned char* ok,
                                 int threads);
void ltpo_time_scaling_items(const ltpo_planner* L, int64_t n, const int* joint,
                             const double* q_goal, const double* q_0, const double* v_0,
                             const double* a_0, const double* dir, const double* t_required,
                             double* t7, double* v_drive, unsigned char* mod,
                             unsigned char* ts_case, unsigned char* final_case,
                             unsigned char* ok, int threads);

/* stages 1-3 (cc:14-55), problem-major x[p*dof + joint], times [p][joint][7].
 * reached[p] = 1 if the reference would go on to getTrajectory. traj_len per cc:716-719
 * (0 when not reached or when a switching time is not finite). */
void ltpo_solve_batch(const ltpo_planner* L, int64_t n, const double* q_goal, const double* q_0,
                      const double* v_0, const double* a_0, double* t_opt, double* t_scaled,
                      double* dir, double* v_drive, unsigned char* mod, unsigned char* opt_case,
                      unsigned char* ts_case, unsigned char* final_case, int* slowest,
                      int* traj_len, unsigned char* reached, int threads);

/* cc:706-841. rows [joint][stride]; returns traj_len (or -needed if stride too small).
 * Writes that the reference performs out of bounds are dropped. */
int ltpo_get_trajectory(const ltpo_planner* L, const double* t7, const double* dir,
                        const unsigned char* mod, const double* q_0, const double* v_0,
                        const double* a_0, const double* v_drive, int64_t stride, double* q,
                        double* v, double* a, double* j);

/* cc:7-63 for one problem; returns success; *length = -1 if the trajectory is untouched */
int ltpo_plan(const ltpo_planner* L, const double* q_goal, const double* q_0, const double* v_0,
              const double* a_0, int64_t stride, double* q, double* v, double* a, double* j,
              int* length);

/* timing / checksum leg: full plans, nothing kept but flags, lengths and the sum of the
 * final positions. */
double ltpo_plan_batch(const ltpo_planner* L, int64_t n, const double* q_goal, const double* q_0,
                       const double* v_0, const double* a_0, unsigned char* success, int* length,
                       int threads);

#ifdef __cplusplus
}
#endif
#endif
