/* ltp_b200.h -- C ABI of the B200-native batched planning hot path.
 *
 * This is the drop-in boundary. The reference (yannickBurkhardt/LongTermPlanner) has no
 * FFI layer of its own: its boundary is the C++ class LongTermPlanner
 * (include/long_term_planner/long_term_planner.h:61-308). Each entry point below names
 * the reference interface it replaces; include/long_term_planner/long_term_planner.h in
 * THIS repository re-creates that class on top of these calls.
 *
 * Conventions
 *  - plain C types only; every pointer is either a HOST pointer ("_host" entry points and
 *    the limit vectors) or a DEVICE pointer on the planner's device (everything else).
 *  - batched device buffers are structure-of-arrays, JOINT-MAJOR:  x[joint * n + problem];
 *    switching times are t[(k * dof + joint) * n + problem], k = 0..6.
 *  - trajectories are written sample-contiguous, one row per (problem, joint):
 *    q[(problem * dof + joint) * stride + sample], mirroring Trajectory::q[joint][sample]
 *    (reference long_term_planner.h:41-44).
 *  - all device work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = the
 *    legacy default stream); no entry point synchronises unless it says so.
 *  - return value: LTP_OK or a negative ltp_status. Nothing is printed, nothing throws.
 *  - there is no CPU fallback: without a CUDA device ltp_create fails with LTP_ERR_CUDA.
 */
#ifndef LTP_B200_H
#define LTP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LTP_MAX_DOF 32

typedef enum {
  LTP_OK = 0,
  LTP_ERR_ARG = -1,      /* null pointer, dof out of range, n < 0, bad stride ... */
  LTP_ERR_CUDA = -2,     /* a CUDA runtime call failed; see ltp_last_cuda_error() */
  LTP_ERR_CAPACITY = -3  /* caller-provided row capacity too small; see ltp_plan_host */
} ltp_status;

/* case byte (the reference emits no case id; encoding documented in DESIGN.md):
 *  low nibble 0 brake-only | 1..4 cruise phase exists (1: P2,P6; 2: no P2; 3: no P6; 4: none)
 *  | 5 no cruise phase (closed form) | 6 quartic #1 | 7 quartic #1 + P2 | 8 quartic #2
 *  | 13 failure at the final safety check (reference leaves t unwritten) | 14 degenerate
 *  zero return | 15 failure (t zeroed);  0x10 modified jerk profile; 0x20 both
 *  re-insertions fired; 0x40 / 0x80 no-P2 / no-P6 branch taken.
 * ts_case: 0 slowest joint, 1..8 accepted attempt, 9 search failed (optimal times kept),
 *  255 plan aborted before time scaling. */

typedef struct ltp_planner ltp_planner;

/* replaces LongTermPlanner::LongTermPlanner(dof, t_sample, q_min, q_max, v_max, a_max,
 * j_max) (reference long_term_planner.h:118-131). Limit vectors: host, dof doubles. */
int ltp_create(ltp_planner** out, int device, int dof, double t_sample, const double* q_min,
               const double* q_max, const double* v_max, const double* a_max,
               const double* j_max);
/* replaces setLimits / setSampleTime / setDoF (reference long_term_planner.h:176-205) */
int ltp_set_limits(ltp_planner* p, const double* q_min, const double* q_max, const double* v_max,
                   const double* a_max, const double* j_max);
int ltp_set_sample_time(ltp_planner* p, double t_sample);
int ltp_set_dof(ltp_planner* p, int dof);
/* how ltp_solve_batch runs: AUTO = closed-form kernel for every problem + generic kernel for
 * the problems that need a polynomial root solve (default); GENERIC = the generic kernel for
 * every problem. Results are identical; the switch exists for validation and profiling. */
#define LTP_SOLVE_AUTO 0
#define LTP_SOLVE_GENERIC 1
int ltp_set_solve_mode(ltp_planner* p, int mode);
/* Per-kernel timing for benchmarks. While on, every launch of the hot kernels is
 * bracketed by a CUDA event pair on the launching stream. ltp_profile_read waits for the
 * recorded launches of one kernel, returns the sum of their durations in ms and their number
 * since the last reset. Off by default; results do not depend on it. */
#define LTP_PROFILE_SOLVE_FAST 0      /* closed-form kernel: stages 1-2 + first candidate (every problem) */
#define LTP_PROFILE_SOLVE_GENERIC 1   /* every-branch kernel (work list, or all problems in GENERIC mode) */
#define LTP_PROFILE_SAMPLE_TIME_MAJOR 2
#define LTP_PROFILE_SAMPLE_ROWS 3
#define LTP_PROFILE_SOLVE_ATTEMPT2 4  /* second candidate for the queued joints */
#define LTP_PROFILE_SOLVE_ITEMS 5     /* tail + pending + search kernels (item mode, large batches) */
#define LTP_PROFILE_KERNELS 6
int ltp_set_profiling(ltp_planner* p, int on);
int ltp_profile_read(ltp_planner* p, int kernel, double* ms_sum, int64_t* launches, int reset);
int ltp_get_dof(const ltp_planner* p);
int ltp_get_device(const ltp_planner* p);
void ltp_destroy(ltp_planner* p);
const char* ltp_status_string(int status);
const char* ltp_last_cuda_error(void);
/* number of kernel launches issued through this planner since creation */
int64_t ltp_launch_count(const ltp_planner* p);

/* ---- per-joint primitives, batched (device pointers, joint-major [dof][n]) ----------- */

/* replaces LongTermPlanner::optBraking (reference long_term_planner.cc:650-701).
 * t_rel: [3][dof][n]. */
int ltp_opt_braking_batch(ltp_planner* p, int64_t n, const double* v_0, const double* a_0,
                          double* q_stop, double* t_rel, double* dir, void* stream);

/* replaces LongTermPlanner::optSwitchTimes (reference long_term_planner.cc:82-353).
 * t: [7][dof][n] (left at 0 where the reference leaves it unwritten); kase may be NULL. */
int ltp_opt_switch_times_batch(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0,
                               const double* v_0, const double* a_0, const double* v_drive,
                               double* t, double* dir, uint8_t* mod, uint8_t* kase, uint8_t* ok,
                               void* stream);

/* replaces LongTermPlanner::timeScaling (reference long_term_planner.cc:358-645).
 * dir, t_required: [dof][n] inputs. ts_case / final_case may be NULL. */
int ltp_time_scaling_batch(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0,
                           const double* v_0, const double* a_0, const double* dir,
                           const double* t_required, double* t, double* v_drive, uint8_t* mod,
                           uint8_t* ts_case, uint8_t* final_case, uint8_t* ok, void* stream);

/* ---- the batched planner: planTrajectories ------------------------------------------ */

/* Result of stages 1-3 (phase times, synchronisation, time scaling). Device pointers.
 * Pointers marked optional may be NULL. */
typedef struct {
  double* t_scaled;    /* [dof][n][8]  one 64-byte record per (joint, problem), 32-byte aligned:
                        *      t_scaled[(joint * n + problem) * 8 + k], k = 0..6 the final cumulative
                        *      switching times (reference t_scaled, cc:50-55), k = 7 v_drive.
                        *      A record is two whole 32-byte sectors written once by whichever
                        *      kernel settles the joint: no store of the solve is ever partial */
  double* dir;         /* [dof][n] */
  double* v_drive;     /* optional [dof][n]  copy of slot 7 of the records */
  uint8_t* mod;        /* [dof][n]  modified jerk profile flag */
  int32_t* slowest;    /* [n]  index of the slowest joint, -1 if none */
  int32_t* traj_len;   /* [n]  samples of the trajectory (cc:716-719); 0 if not reached */
  uint8_t* reached;    /* [n]  1 if the reference would go on to getTrajectory (cc:58). With
                        *      reached = 0 the reference has returned false before producing
                        *      anything (cc:15,29,39): traj_len is 0 and the other fields of that
                        *      problem are unspecified (all joints are evaluated in parallel here,
                        *      the reference stops at the first joint that fails) */
  double* t_opt;       /* optional [7][dof][n]  time-optimal switching times (cc:27-30) */
  uint8_t* opt_case;   /* optional [dof][n] */
  uint8_t* ts_case;    /* optional [dof][n] */
  uint8_t* final_case; /* optional [dof][n] */
} ltp_solution;

/* replaces the solve part of LongTermPlanner::planTrajectory
 * (reference long_term_planner.cc:14-55) for n independent problems.
 * ONE STREAM PER PLANNER AT A TIME: the work list of the solve, the sort bins of
 * ltp_sample_batch_sorted and the staging block of the host calls belong to the planner, not to
 * the stream. Calls on one planner must be stream-ordered with respect to each other (same
 * stream, or separated by events); for concurrent streams or host threads create one planner
 * per stream -- a planner is a few hundred bytes plus its scratch. (The reference's
 * planTrajectory is re-entrant on one object; the drop-in class is not, for the same reason.)
 * The solve scratch (about n * (8 + 72 * dof) bytes: work lists, the per-joint queues of the second
 * cruise-speed candidate and the item lists) is allocated on the first call and whenever n grows,
 * which must not happen inside a CUDA-graph capture: call ltp_reserve first. */
int ltp_reserve(ltp_planner* p, int64_t n); /* scratch for solves of up to n problems, now */
int ltp_solve_batch(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0,
                    const double* v_0, const double* a_0, const ltp_solution* sol, void* stream);

/* replaces LongTermPlanner::getTrajectory + the final joint-limit check
 * (reference long_term_planner.cc:58-61, 706-841) for n problems.
 *   horizon == 0: every problem writes exactly traj_len[p] samples, clipped to the sample
 *                 capacity `stride` (traj_len[p] > stride tells the caller a row was clipped;
 *                 the success flag still refers to the complete trajectory);
 *   horizon  > 0: every problem writes exactly `horizon` samples (clipped, or continued
 *                 with the recurrence's own steady state q_last, 0, 0, 0). A problem that was
 *                 not planned (reached = 0 or traj_len <= 0) holds its start position:
 *                 (q_0, 0, 0, 0) in every sample, success = 0. With horizon == 0 nothing is
 *                 stored for such a problem.
 * layout:
 *   LTP_LAYOUT_ROWS        q[(problem * dof + joint) * stride + sample]  -- one row per
 *                          (problem, joint) like Trajectory::q[joint][sample]; stride =
 *                          doubles per row, >= the samples written. 32-byte vector stores
 *                          when stride % 4 == 0 and the base pointers are 32-byte aligned.
 *   LTP_LAYOUT_TIME_MAJOR  q[(sample * n + problem) * dof + joint]  -- a (stride, n, dof)
 *                          tensor; stride = sample capacity. This is the layout a batched
 *                          consumer steps through (all environments read sample k together)
 *                          and the one that streams to HBM at full bandwidth -- provided a
 *                          sample plane (n * dof doubles) is a whole number of 256-byte
 *                          pieces, i.e. n * dof is a multiple of 32: pad the batch to a
 *                          multiple of 32 problems (repeat the last one). Measured on B200,
 *                          7 joints: n = 4096 5.8 TB/s, n = 4104 4.8 TB/s, n = 4097 3.1 TB/s.
 *                          Larger batches do better still (n = 16384: 6.5 TB/s).
 * success: [n], 1 iff reached and every joint ends inside [q_min, q_max]. */
#define LTP_LAYOUT_ROWS 0
#define LTP_LAYOUT_TIME_MAJOR 1
int ltp_sample_batch(ltp_planner* p, int64_t n, const double* q_0, const double* v_0,
                     const double* a_0, const ltp_solution* sol, int32_t horizon, int32_t layout,
                     int64_t stride, double* q, double* v, double* a, double* j, uint8_t* success,
                     void* stream);

/* ltp_sample_batch in time-major layout and exact-length mode with the SLOTS ordered by
 * trajectory length (longest first): q[(sample * n + k) * dof + joint] belongs to problem
 * order[k]; `order` (device, n int32) is written by the call, `success` stays indexed by
 * problem. The lanes of a sampler warp run until the longest of their 32 rows ends; with
 * problems of mixed length in index order a sixth of the store slots of random problems is
 * idle, with equal neighbours none (5.4 -> 6.2 TB/s on B200). Same samples, permuted slots. */
int ltp_sample_batch_sorted(ltp_planner* p, int64_t n, const double* q_0, const double* v_0,
                            const double* a_0, const ltp_solution* sol, int64_t capacity, double* q,
                            double* v, double* a, double* j, uint8_t* success, int32_t* order,
                            void* stream);

/* ---- streaming: more trajectories than fit in memory -------------------------------- */

/* One chunk of a streamed run, as seen by the consumer. Everything is a DEVICE pointer into
 * a slot of the planner's ring and stays valid until work enqueued on `stream` by the
 * consumer has run (the slot is reused by the chunk after next, on the same stream). Inputs
 * and solution are joint-major over `count` problems; q, v, a, j are time-major
 * (capacity, count, dof). */
typedef struct {
  int64_t first, count;        /* problems first .. first + count - 1 of the run */
  int64_t capacity;            /* sample capacity of the trajectory tensors */
  int32_t horizon;
  ltp_solution solution;
  const double *q_goal, *q_0, *v_0, *a_0;
  const double *q, *v, *a, *j;
  const uint8_t* success;
  /* NULL: trajectory slot k of the chunk holds problem k of the chunk. Otherwise (sorted-slot
   * mode, ltp_set_stream_sorted): slot k holds problem order[k], i.e.
   * q[(sample * count + k) * dof + joint] belongs to problem first + order[k]. The solution,
   * the inputs and `success` stay indexed by problem. */
  const int32_t* order;
} ltp_chunk;

/* called on the host right after a chunk's kernels were enqueued; enqueue the consuming work
 * on `stream` (do not synchronise). Non-zero return aborts the run. */
typedef int (*ltp_chunk_consumer)(void* user, const ltp_chunk* chunk, void* stream);

typedef struct {
  int64_t problems, chunks;
  int64_t reached, success;    /* problems that were planned / ended inside the joint limits */
  int64_t clipped;             /* problems whose trajectory is longer than what was written */
  int64_t samples;             /* (problem, joint, sample) triples written */
  int64_t bytes;               /* samples * 32 (q, v, a, j as f64) */
  int64_t max_traj_len;
} ltp_stream_stats;

/* planTrajectory for n problems whose trajectories do not fit in memory at once (reference
 * long_term_planner.cc:7-63 per problem). Inputs: device, joint-major [dof][n]. The run is cut
 * into chunks of `chunk` problems; each chunk is solved and sampled (time-major, `horizon`
 * and `capacity` as in ltp_sample_batch) into one of two ring slots on its own stream, handed
 * to `consume` (may be NULL), and the slot is recycled (`chunk` a multiple of 32 keeps the
 * sample planes aligned, see LTP_LAYOUT_TIME_MAJOR). Synchronises before returning;
 * `stats` (may be NULL) receives the totals, accumulated on the device. The run uses the
 * planner's own streams; input_stream (a cudaStream_t, NULL = the default stream) is the stream
 * on which the caller produced the inputs: the run waits for the work enqueued on it so far. */
int ltp_plan_stream(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0,
                    const double* v_0, const double* a_0, int64_t chunk, int32_t horizon,
                    int64_t capacity, ltp_chunk_consumer consume, void* user,
                    ltp_stream_stats* stats, void* input_stream);
/* Sorted-slot mode of ltp_plan_stream for exact-length sampling (horizon = 0); off by default.
 * The lanes of a sampler warp run until the longest of their 32 rows ends, so with problems of
 * mixed length a sixth of the store slots of random problems is idle. When on, every chunk's
 * problems are ordered by trajectory length on the device (longest first) and trajectory slot k
 * holds problem ltp_chunk.order[k]: same samples, permuted slots, full warps. */
int ltp_set_stream_sorted(ltp_planner* p, int on);

/* Receding-horizon replanning (the reference's stated use, README.md:10-13: a new target
 * arrives before the previous one is reached): the state `tick` samples into the current
 * time-major trajectories (sample index tick, i.e. time (tick + 1) * t_sample, reference
 * cc:810-812) becomes the next start state, joint-major [dof][n]. traj_len (may be NULL for
 * fixed-horizon trajectories, which hold every sample up to the horizon): [n], the tick is
 * limited to traj_len - 1 per problem. valid (may be NULL): [n], problems with 0 keep their
 * q_0/v_0/a_0 (pass the solution's `reached`: the samplers store nothing for a problem that
 * was not planned unless a fixed horizon is used). capacity: the sample capacity of q, v, a
 * (their first extent); the sample read is limited to capacity - 1 (a trajectory clipped by the
 * capacity), and with traj_len == NULL a tick >= capacity is LTP_ERR_ARG. clamp != 0 pulls a
 * state that overshoots a limit by the recurrence's rounding back inside what checkInputs
 * accepts (reference cc:68-77). */
int ltp_advance_batch(ltp_planner* p, int64_t n, int32_t tick, int32_t clamp, int64_t capacity,
                      const int32_t* traj_len, const uint8_t* valid, const double* q, const double* v,
                      const double* a, double* q_0, double* v_0, double* a_0, void* stream);

/* Layout bridge: dst[c][r] = src[r][c] for a rows x cols matrix of doubles on the device
 * (shared-memory tiled, both sides coalesced). A vectorised environment keeps its state
 * problem-major, x[problem][joint]; ltp_transpose(p, n, dof, x_pm, x_jm, stream) gives the
 * joint-major [dof][n] layout the entry points above take, and (p, dof, n, ...) goes back.
 * (The time-major trajectories already are (samples, n, dof), i.e. problem-major per sample.) */
int ltp_transpose(ltp_planner* p, int64_t rows, int64_t cols, const double* src, double* dst, void* stream);

/* ---- host-buffer entry points (what a caller without device buffers uses) ------------ */

/* Stages 1-3 with HOST buffers: copies the four inputs in, solves, copies the requested
 * outputs back, and synchronises. Same layouts as ltp_solve_batch; the ltp_solution
 * holds HOST pointers here. Output mask: every field except traj_len and reached may be NULL
 * and is then not copied back -- the device->host transfer is what bounds this call (520 B per
 * 7-DoF plan for the full solution: records incl. v_drive, dir, mod, slowest, traj_len, reached;
 * 453 B for records + traj_len + reached, 5 B for the durations alone). Pinned host memory makes the copies asynchronous and lets chunks overlap. */
int ltp_solve_host(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0,
                   const double* v_0, const double* a_0, const ltp_solution* host_sol);

/* Full planTrajectory for n problems with HOST buffers (reference long_term_planner.cc:7-63).
 * Rows are [n][dof][capacity]. If a trajectory needs more than `capacity` samples nothing is
 * written for any problem, *needed receives the required capacity and LTP_ERR_CAPACITY is
 * returned. traj_len / success: [n] host. horizon as in ltp_sample_batch. */
int ltp_plan_host(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0,
                  const double* v_0, const double* a_0, int32_t horizon, int64_t capacity,
                  double* q, double* v, double* a, double* j, int32_t* traj_len,
                  uint8_t* success, int64_t* needed);

/* planTrajectory (reference long_term_planner.cc:7-63) for ONE problem, result left in the
 * planner's pinned staging block instead of being copied out: rows4[f] (f = 0..3: q, v, a, j)
 * points at dof rows of *row_stride doubles each, of which the first *length are samples.
 * The pointers stay valid until the next call on this planner. This is what the drop-in
 * class's planTrajectory uses (it copies straight into Trajectory's vectors). *length <= 0:
 * the reference's early `return false`. LTP_ERR_CAPACITY when the trajectory is longer than a
 * staging row (*length tells how long): use ltp_plan_host. */
int ltp_plan_one_view(ltp_planner* p, const double* q_goal, const double* q_0, const double* v_0,
                      const double* a_0, const double** rows4, int64_t* row_stride,
                      int32_t* length, uint8_t* success);

/* single-item host forms of the protected per-joint methods (used by the C++ drop-in) */
int ltp_opt_braking_host(ltp_planner* p, int joint, double v_0, double a_0, double* q_stop,
                         double* t_rel3, double* dir);
int ltp_opt_switch_times_host(ltp_planner* p, int joint, double q_goal, double q_0, double v_0,
                              double a_0, double v_drive, double* t7, double* dir, uint8_t* mod,
                              uint8_t* kase, uint8_t* ok);
int ltp_time_scaling_host(ltp_planner* p, int joint, double q_goal, double q_0, double v_0,
                          double a_0, double dir, double t_required, double* t7, double* v_drive,
                          uint8_t* mod, uint8_t* ts_case, uint8_t* ok);
/* getTrajectory on given switching times; t7 [dof][7] host, rows [dof][capacity] host */
int ltp_get_trajectory_host(ltp_planner* p, const double* t7, const double* dir,
                            const uint8_t* mod, const double* q_0, const double* v_0,
                            const double* a_0, const double* v_drive, int64_t capacity, double* q,
                            double* v, double* a, double* j, int32_t* length, int64_t* needed);

#ifdef __cplusplus
}
#endif
#endif /* LTP_B200_H */
