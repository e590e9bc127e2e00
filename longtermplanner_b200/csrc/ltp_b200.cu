// B200 (sm_100a) kernels and C ABI of the batched planning hot path. See include/ltp_b200.h
// for the boundary and DESIGN.md for the data layout and the per-kernel rooflines.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false ...
// (-fmad=false is part of the numerical contract, see ltp_math.cuh).
//
// Thread mapping of the solve / sample kernels: a CTA is (32 problems) x (dof joints);
// warp w works on joint w of 32 consecutive problems. All lanes of a warp therefore share
// one limit set, every lane is busy for any dof (no padding to a power of two), and the
// joint-major buffers x[joint * n + problem] are read and written fully coalesced. The
// per-problem reductions (slowest joint, trajectory length, final limit check) go through
// a few bytes of shared memory per problem.
#include <cuda_runtime.h>
#include <type_traits>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <initializer_list>
#include <new>

#include "../../include/ltp_b200.h"
#include "ltp_math.cuh"

namespace {

using namespace ltp;

// The per-joint limits travel in two arrays: the fields every kernel reads (136 bytes per joint,
// the stride they had before the limit-only factors were added) and the factors only the
// closed-form kernel for up to 8 joints reads (DivDeferredWide). One 256-byte record per joint --
// a power-of-two stride -- cost the 12-joint kernel 27 % (0.83 -> 1.05 ms) even though it never
// touches the second half: its joints' lines then meet in the same sets of the constant cache
// (profiles/r02_ab_limit_constants.log, padded strides of 264 / 280 / 296 bytes: 0.94 / 0.85 / 0.85 ms).
#define LTP_HOT_FIELDS(X)                                                                              \
  X(q_min) X(q_max) X(v_max) X(a_max) X(j_max) X(r_a) X(r_j) X(a_over_j) X(r_v) X(r_j2) X(r_j3) X(r_6j3) \
  X(r_aj) X(aoj3) X(aoj4) X(t5v) X(part2v)
#define LTP_WIDE_FIELDS(X)                                                                             \
  X(c_aj) X(c_a2h) X(c_36a2j2) X(c_72a3j) X(c_144a) X(c_72aj2) X(c_a3) X(c_36a4) X(c_36j2) X(rk_q) X(rk_m) \
  X(rk_h) X(rk_3) X(rk_7) X(rk_8)
#define LTP_DECL(f) double f;
struct JointHot { LTP_HOT_FIELDS(LTP_DECL) };
struct JointWide { LTP_WIDE_FIELDS(LTP_DECL) };
#undef LTP_DECL
static_assert(sizeof(JointHot) + sizeof(JointWide) == sizeof(JointLimits), "a field of JointLimits is in neither list");

struct PlannerParams {
  int dof;
  double ts;
  double r_ts;  // RN(1 / ts), for the sample counts (div_by)
  JointHot hot[LTP_MAX_DOF];
  JointWide wide[LTP_MAX_DOF];
  // the joint's limits as the closed-form functions take them; a field is loaded where it is used
  __host__ __device__ __forceinline__ JointLimits joint(int jt) const {
    JointLimits L;
#define LTP_GET_HOT(f) L.f = hot[jt].f;
#define LTP_GET_WIDE(f) L.f = wide[jt].f;
    LTP_HOT_FIELDS(LTP_GET_HOT)
    LTP_WIDE_FIELDS(LTP_GET_WIDE)
#undef LTP_GET_HOT
#undef LTP_GET_WIDE
    return L;
  }
  void set_joint(int jt, const JointLimits& L) {
#define LTP_PUT_HOT(f) hot[jt].f = L.f;
#define LTP_PUT_WIDE(f) wide[jt].f = L.f;
    LTP_HOT_FIELDS(LTP_PUT_HOT)
    LTP_WIDE_FIELDS(LTP_PUT_WIDE)
#undef LTP_PUT_HOT
#undef LTP_PUT_WIDE
  }
};

struct DeviceSolution {  // ltp_solution, by value
  double* t_scaled;
  double* dir;
  double* v_drive;
  uint8_t* mod;
  int32_t* slowest;
  int32_t* traj_len;
  uint8_t* reached;
  double* t_opt;
  uint8_t* opt_case;
  uint8_t* ts_case;
  uint8_t* final_case;
};

constexpr int kTile = 32;  // problems per CTA

// ------------------------------------------------------------------------------------
// Stages 1-3 (reference cc:14-55) run as two kernels on one stream:
//
//   ltp_solve_fast_kernel    every problem. Closed-form work only: the phase solve without
//                            the quartic tail, the slowest-joint reduction, and attempts 1
//                            and 2 of the time-scaling search (both closed form). A problem
//                            in which any joint needs a polynomial root solve is appended to
//                            a work list and left for the second kernel.
//   ltp_solve_generic_kernel the work list (or, in generic-only mode, every problem): every
//                            branch of the reference evaluated in-thread, including the
//                            Francis-QR root finder. It overwrites whatever the fast kernel
//                            stored for a deferred problem.
//
// Keeping the root solver out of the first kernel keeps its code small and its warps
// converged; grouping the root-solve problems keeps the lanes of the second kernel busy.
// The split changes no result: a deferred problem is recomputed from its inputs by exactly
// the code that handles it in generic-only mode.
// ------------------------------------------------------------------------------------
// Queue of the joints whose first cruise-speed candidate was rejected: filled by the
// closed-form kernel, drained (regrouped into full warps) by ltp_solve_attempt2_kernel.
// Structure of arrays, so that both sides access it fully coalesced; the start state travels
// with the entry (the producer has it in registers, a gather in the consumer costs more).
struct Attempt2Queue {
  double *t_req, *q_goal, *q_0, *v_0, *a_0;
  int2* where;  // (problem, joint)
};

constexpr int kDeferredMark = 0x7fffffff;  // traj_len of a problem that sits in the work list
constexpr int64_t kItemModeMin = 8192;     // problems from which the solve runs in item mode

struct SolveShared {
  double* t6;            // [dof][32]
  int* len;              // [dof][32]
  int* arrived;          // [32]
  int* warp_items;       // [dof + 2]   (spare)
  int* tail_items;       // [dof + 2]   the same for the tail items
  unsigned char* flag;   // [dof][32]  bit0 fail, bit1 defer
};

__host__ __device__ inline size_t solve_smem_bytes(int dof) {
  return (size_t)dof * kTile * (sizeof(double) + sizeof(int) + 1) + (kTile + 2 * (dof + 2)) * sizeof(int);
}

__device__ __forceinline__ SolveShared carve_shared(unsigned char* raw, int dof) {
  SolveShared s;
  s.t6 = reinterpret_cast<double*>(raw);
  s.len = reinterpret_cast<int*>(s.t6 + dof * kTile);
  s.arrived = s.len + dof * kTile;
  s.warp_items = s.arrived + kTile;
  s.tail_items = s.warp_items + dof + 2;
  s.flag = reinterpret_cast<unsigned char*>(s.tail_items + dof + 2);
  return s;
}

// Device scratch of one solve. counters: [0] whole-problem work list, [2] tail items,
// [3] tail-pending problems, [4] search items, [8 + joint] that joint's part of the attempt-2
// queue. Lists:
//   work_list   problems the every-branch kernel recomputes as a whole (n ints)
//   queue       joints whose first cruise-speed candidate was rejected (dof * n entries)
//   tail_items  (problem, joint) pairs whose time-optimal solve needs the quartic tail (dof * n)
//   pending     problems with at least one tail item: their stages 2-3 wait for it (n ints)
//   search_*    (problem, joint) pairs whose search needs a polynomial root, with the required
//               end time (dof * n)
struct SolveScratch {
  int* counters;
  int* work_list;
  Attempt2Queue queue;
  int2* tail_items;
  int* pending;
  int2* search_items;
  double* search_t_req;
};
enum { kCntWork = 0, kCntTail = 2, kCntPending = 3, kCntSearch = 4, kCounters = 8 };
// The attempt-2 queue is one sub-queue per joint (entries [joint * n, joint * n + count[joint]),
// counts behind the counters above): a warp of the closed-form kernel holds 32 problems of ONE
// joint, so it reserves its slots with one atomic of its own and needs no word from the other
// warps of its CTA -- the CTA-wide slot assignment it replaces cost two barriers after attempt 1.
constexpr int kCounterInts = kCounters + LTP_MAX_DOF;

inline size_t round16(size_t x) { return (x + 15) / 16 * 16; }

inline size_t solve_scratch_bytes(int dof, int64_t n) {
  const size_t cap = (size_t)dof * (size_t)n;
  const size_t dn = (size_t)dof * (size_t)n;
  return round16(kCounterInts * sizeof(int)) + 2 * round16((size_t)n * sizeof(int)) + cap * (5 * sizeof(double) + sizeof(int2)) +
         dn * (2 * sizeof(int2) + sizeof(double));
}

inline SolveScratch carve_scratch(void* base, int dof, int64_t n) {
  SolveScratch s;
  unsigned char* b = static_cast<unsigned char*>(base);
  const size_t cap = (size_t)dof * (size_t)n;
  const size_t dn = (size_t)dof * (size_t)n;
  s.counters = reinterpret_cast<int*>(b);
  b += round16(kCounterInts * sizeof(int));
  s.work_list = reinterpret_cast<int*>(b);
  b += round16((size_t)n * sizeof(int));
  s.pending = reinterpret_cast<int*>(b);
  b += round16((size_t)n * sizeof(int));
  double* q = reinterpret_cast<double*>(b);
  s.queue.t_req = q;
  s.queue.q_goal = q + cap;
  s.queue.q_0 = q + 2 * cap;
  s.queue.v_0 = q + 3 * cap;
  s.queue.a_0 = q + 4 * cap;
  s.search_t_req = q + 5 * cap;
  int2* w = reinterpret_cast<int2*>(q + 5 * cap + dn);
  s.queue.where = w;
  s.tail_items = w + cap;
  s.search_items = w + cap + dn;
  return s;
}

// per-joint stores shared by both kernels: the part known after the time-optimal solve ...
__device__ __forceinline__ void store_joint_opt(const DeviceSolution& S, int dof, int jt, int64_t n, int64_t p,
                                                const double* t_opt, double dir, unsigned char opt_case) {
  const int64_t at = (int64_t)jt * n + p;
  S.dir[at] = dir;
  if (S.t_opt) {
#pragma unroll
    for (int k = 0; k < 7; ++k) S.t_opt[((int64_t)k * dof + jt) * n + p] = t_opt[k];
  }
  if (S.opt_case) S.opt_case[at] = opt_case;
}

// The switching times of a (problem, joint) live in one 64-byte record -- t[0..6] and v_drive,
// t_scaled[(joint * n + problem) * 8 + k] -- i.e. two whole 32-byte sectors that belong to nobody
// else. Whichever kernel settles the joint writes the record with two 256-bit stores, so no sector
// of the solution is ever written partially, however many kernels hand the joint on before it is
// settled (with the round-1 layout [7][dof][n] a sector held the same time of four problems, and
// every joint settled by a later kernel cost the memory system a read-modify-write per time).
__device__ __forceinline__ double* record_of(const DeviceSolution& S, int jt, int64_t n, int64_t p) {
  return S.t_scaled + ((int64_t)jt * n + p) * 8;
}

__device__ __forceinline__ void store_record(double* rec, const double* t, double v_drive) {
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(rec), "d"(t[0]), "d"(t[1]), "d"(t[2]), "d"(t[3])
               : "memory");
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(rec + 4), "d"(t[4]), "d"(t[5]), "d"(t[6]), "d"(v_drive)
               : "memory");
}

__device__ __forceinline__ void load_record(const double* rec, double* t, double& v_drive) {
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(t[0]), "=d"(t[1]), "=d"(t[2]), "=d"(t[3]) : "l"(rec));
  asm volatile("ld.global.v4.f64 {%0, %1, %2, %3}, [%4];"
               : "=d"(t[4]), "=d"(t[5]), "=d"(t[6]), "=d"(v_drive)
               : "l"(rec + 4));
}

// ... and the part that the time-scaling search decides
__device__ __forceinline__ void store_joint_scaled(const DeviceSolution& S, int dof, int jt, int64_t n, int64_t p,
                                                   const double* t_sc, double v_drive, unsigned char mod,
                                                   unsigned char ts_case, unsigned char final_case) {
  const int64_t at = (int64_t)jt * n + p;
  store_record(record_of(S, jt, n, p), t_sc, v_drive);
  if (S.v_drive) S.v_drive[at] = v_drive;
  S.mod[at] = mod;
  if (S.ts_case) S.ts_case[at] = ts_case;
  if (S.final_case) S.final_case[at] = final_case;
  (void)dof;
}

__device__ __forceinline__ void store_joint(const DeviceSolution& S, int dof, int jt, int64_t n, int64_t p,
                                            const double* t_sc, const double* t_opt, double dir,
                                            double v_drive, unsigned char mod, unsigned char opt_case,
                                            unsigned char ts_case, unsigned char final_case) {
  store_joint_scaled(S, dof, jt, n, p, t_sc, v_drive, mod, ts_case, final_case);
  store_joint_opt(S, dof, jt, n, p, t_opt, dir, opt_case);
}

// cc:718 for one joint, -1 when a switching time is not finite / not representable. The seven
// times are a running sum (cumsum7) or all zero wherever this is called, so a time that is not
// finite leaves the last one not finite, and the last one is the only one looked at. t6 / Ts goes
// through the reciprocal of the sample time (same bits, see DivDeferred: zero -- a joint that does
// not move -- included); a quotient outside the window takes the plain division.
__device__ __forceinline__ int joint_samples(const double* t_sc, double Ts, double r_ts) {
  const double t6 = t_sc[6];
  DivDeferred dv;
  double q = dv.by(t6, Ts, r_ts);
  if (dv.bad) q = t6 / Ts;
  if (!(isfinite(t6) && q <= 2.0e9)) return -1;
  const double x = ceil(q);  // samples_for(), with the quotient at hand
  if (!(x >= -1.0e9 && x <= 2.0e9)) return 0;
  return (int)x + 1;
}

// The last of a problem's dof threads to get here reduces the per-joint lengths and writes
// the per-problem outputs; no CTA-wide barrier, so a warp that is still busy with a slow
// joint does not hold the others back. Returns true in that last thread if the problem
// must be appended to the whole-problem work list.
// tail_pending (item mode): a joint of the problem waits for its quartic tail; stages 2-3 of the
// whole problem are redone by ltp_solve_pending_kernel, nothing is decided here.
__device__ __forceinline__ bool finish_problem(const SolveShared& sh, const DeviceSolution& S, int dof,
                                               int lane, int jt, int64_t p, int my_len, bool my_defer,
                                               bool reached, int slowest, bool tail_pending = false) {
  sh.len[jt * kTile + lane] = my_len;
  if (my_defer) sh.flag[jt * kTile + lane] |= 4;
  __threadfence_block();
  const int prev = atomicAdd(&sh.arrived[lane], 1);
  if (prev != dof - 1) return false;
  __threadfence_block();
  int len = 0;
  bool bad = false, defer = false;
  for (int i = 0; i < dof; ++i) {
    const int li = sh.len[i * kTile + lane];
    bad |= li < 0;
    len = li > len ? li : len;
    defer |= (sh.flag[i * kTile + lane] & 4) != 0;
  }
  S.slowest[p] = slowest;
  S.reached[p] = (uint8_t)reached;
  if (tail_pending) {
    S.traj_len[p] = 0;
    return false;
  }
  S.traj_len[p] = defer ? kDeferredMark : ((reached && !bad) ? len : 0);
  return defer;
}

// Item mode (large batches): what cannot be settled in closed form is handed on per (problem,
// joint), so that the kernels that run the root finder do so on full warps of joints that all
// need it -- the every-branch kernel run on whole problems keeps 4 of 32 lanes busy on the
// reference's test limits (ncu: 4.1 active threads per instruction, 9 ms for the 22 % of 2^20
// problems that reach it).
//   tail item    the joint's time-optimal solve needs the quartic tail (cc:245-337)
//                -> ltp_solve_tail_kernel computes it; the problem waits in `pending`
//   pending      -> ltp_solve_pending_kernel: stages 2-3 of those problems, closed form
//   search item  the joint's cruise-speed search needs candidates 3..8 (cc:449-644) or a nested
//                solve with a tail -> ltp_solve_search_kernel, one thread per joint
// Only a problem with a non-finite switching time still goes to the whole-problem list.
__device__ __forceinline__ void push_search_item(const SolveScratch& X, int64_t p, int jt, double t_req) {
  const int e = atomicAdd(X.counters + kCntSearch, 1);
  X.search_items[e] = make_int2((int)p, jt);
  X.search_t_req[e] = t_req;
}

__device__ __forceinline__ void push_whole_problem(const DeviceSolution& S, const SolveScratch& X, int64_t p) {
  if (atomicExch(S.traj_len + p, kDeferredMark) != kDeferredMark) {
    const int slot = atomicAdd(X.counters + kCntWork, 1);
    X.work_list[slot] = (int)p;
  }
}

// what the tail kernel leaves for the pending kernel in v_drive: mod, opt_case, "solve succeeded"
__device__ __forceinline__ double pack_tail_flags(unsigned mod, unsigned opt_case, unsigned ok) {
  return __longlong_as_double((long long)((mod & 255u) | ((opt_case & 255u) << 8) | ((ok & 1u) << 16)));
}

// Register budget: the kernel is bound by FP64 dependency latency, so it is compiled for
// ~28 resident warps per SM. Measured on B200, 2^20 Franka problems, round 1: 7 joints 1.12 ms at
// 128 registers / 14 warps, 1.04 ms at 80 / 21, 0.95 ms at 72 / 28. Round 2, after the instruction
// diet: 7 joints 0.593 ms at 80 / 21, 0.537 ms at 72 / 28, 0.558 ms at 56 / 35; 12 joints 0.985 ms
// at 80 / 24, 0.933 ms at 56 / 36 -- hence three CTAs per SM for 12 joints; 6 joints 0.476 ms at
// 64 / 30, 0.447 ms at 80 / 24 -- hence four CTAs for 6.
#ifndef LTP_FAST_WARPS
#define LTP_FAST_WARPS 28
#endif
#ifndef LTP_FAST_BLOCKS_12
#define LTP_FAST_BLOCKS_12 3
#endif
#ifndef LTP_FAST_BLOCKS_6
#define LTP_FAST_BLOCKS_6 4
#endif
constexpr int fast_min_blocks(int maxw) {
  return maxw == 12 ? LTP_FAST_BLOCKS_12 : maxw == 6 ? LTP_FAST_BLOCKS_6 : ((LTP_FAST_WARPS + maxw / 2) / maxw > 0 ? (LTP_FAST_WARPS + maxw / 2) / maxw : 1);
}
// The closed-form kernel runs stage 1 and attempt 1 with the range test of the prepared-reciprocal
// divisions deferred (DivDeferred, ltp_math.cuh): one look at a flag per stage instead of a
// divergent call site per division. The thread whose flag is set -- a zero quotient, e.g. a start at
// rest, or a quotient near the ends of the exponent range -- discards the stage and repeats it here
// with every quotient tested where it is computed, out of line; results come back through memory.
#ifndef LTP_FAST_DEFER
#define LTP_FAST_DEFER 1
#endif
#ifndef LTP_FAST_RV
#define LTP_FAST_RV 1
#endif
#ifndef LTP_FAST_PREFETCH
#define LTP_FAST_PREFETCH 592  // tiles ahead; 148 SMs x 4 resident CTAs (0.530 -> 0.524 ms for 7 joints)
#endif
struct Stage1Redo {
  Prologue pro;
  double t_opt[7];
  int st;
  unsigned char mod, opt_case, in_ok;
};
__device__ __noinline__ void stage1_checked(const PlannerParams& P, int jt, double qg, double q0, double v0,
                                            double a0, Stage1Redo* o) {
  const JointLimits L = P.joint(jt);
  o->in_ok = check_joint_input(L, q0, v0, a0);
  o->pro = ost_prologue(L, P.ts, qg, q0, v0, a0);
  zero7(o->t_opt);
  o->mod = 0;
  o->opt_case = 255;
  o->st = ost_body_t<false, true>(L, P.ts, o->pro, qg, q0, L.v_max, o->t_opt, o->mod, o->opt_case);
}
struct Attempt1Redo {
  double t[7];
  double v_drive;
  int c;
  unsigned char mod, final_case;
};
__device__ __noinline__ void attempt1_checked(const PlannerParams& P, int jt, double qg, double q0, double v0,
                                              double a0, double t_req, unsigned char mod, unsigned char final_case,
                                              Attempt1Redo* o) {
  const JointLimits L = P.joint(jt);
  const Prologue pro = ost_prologue(L, P.ts, qg, q0, v0, a0);
  const TsInput I = make_ts_input(qg, q0, v0, a0, pro.dir, t_req);
  zero7(o->t);
  o->v_drive = L.v_max;
  o->mod = mod;
  o->final_case = final_case;
  o->c = time_scaling_attempt1(L, P.ts, pro, I, o->t, o->v_drive, o->mod, o->final_case);
}

struct Attempt2Redo {
  double t[7];
  double v_drive;
  int c;
  unsigned char mod, final_case;
};
__device__ __noinline__ void attempt2_checked(const PlannerParams& P, int jt, double qg, double q0, double v0,
                                              double a0, double t_req, Attempt2Redo* o) {
  const JointLimits L = P.joint(jt);
  const Prologue pro = ost_prologue(L, P.ts, qg, q0, v0, a0);
  const TsInput I = make_ts_input(qg, q0, v0, a0, pro.dir, t_req);
  zero7(o->t);
  o->v_drive = L.v_max;
  o->mod = 0;
  o->final_case = 255;
  o->c = time_scaling_attempt2(L, P.ts, pro, I, o->t, o->v_drive, o->mod, o->final_case);
}

#ifndef LTP_FAST_EXACT
#define LTP_FAST_EXACT 1
#endif
#ifndef LTP_FAST_UNIFORM_JT
#define LTP_FAST_UNIFORM_JT 0
#endif
// EXACT: the CTA has exactly MAXW warps (the arm sizes the kernel is specialised for), so the
// shared-memory offsets are constants and the loops over the joints unroll
template <int MAXW, bool EXACT = false>
__global__ void __launch_bounds__(kTile * MAXW, fast_min_blocks(MAXW))
ltp_solve_fast_kernel(const __grid_constant__ PlannerParams P, int64_t n, const double* __restrict__ q_goal,
                      const double* __restrict__ q_0, const double* __restrict__ v_0,
                      const double* __restrict__ a_0, DeviceSolution S, SolveScratch X, int items) {
  extern __shared__ unsigned char smem_raw[];
  const int dof = EXACT ? MAXW : P.dof;
  const SolveShared sh = carve_shared(smem_raw, dof);
  int* const work_list = X.work_list;
  int* const work_count = X.counters + kCntWork;
#if LTP_FAST_UNIFORM_JT
  // a warp is one joint: the warp-wide reduction returns the joint index in a uniform register,
  // so the limits are addressed through the uniform datapath
  const int lane = threadIdx.x, jt = (int)__reduce_max_sync(0xffffffffu, (unsigned)threadIdx.y);
#else
  const int lane = threadIdx.x, jt = threadIdx.y;
#endif
  const int64_t p = (int64_t)blockIdx.x * kTile + lane;
  const bool valid = p < n;
  const JointLimits L = P.joint(jt);
  const double Ts = P.ts;
  const int64_t at = (int64_t)jt * n + p;

  double qg = 0, q0 = 0, v0 = 0, a0 = 0;
  if (valid) {
    qg = q_goal[at]; q0 = q_0[at]; v0 = v_0[at]; a0 = a_0[at];
  }
#if LTP_FAST_PREFETCH
  // The inputs come from DRAM (235 MB per 2^20 problems, more than the L2 holds) and a tile's
  // first dependent instruction waits for them. Ask the L2 for the lines of the tile that takes
  // this CTA's place when it retires (LTP_FAST_PREFETCH tiles ahead: about one CTA lifetime), one
  // request per 128-byte line.
  {
    const int64_t pf = p + (int64_t)LTP_FAST_PREFETCH * kTile;
    if ((lane & 15) == 0 && pf < n) {
      const int64_t fa = (int64_t)jt * n + pf;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q_goal + fa));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(q_0 + fa));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(v_0 + fa));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a_0 + fa));
    }
  }
#endif
  if (jt == 0) sh.arrived[lane] = 0;
  // stage 1 (cc:14-30)
  double t_opt[7];
  zero7(t_opt);
  unsigned char mod = 0, opt_case = 255;
#if LTP_FAST_DEFER
  // up to 8 joints: the whole JointLimits block of the CTA's joints stays in the constant cache,
  // so the limit-only factors are read from it (DivDeferredWide, ltp_math.cuh)
  using Dv = typename std::conditional<(MAXW <= 8), DivDeferredWide, DivDeferred>::type;
  Dv dv;
  bool in_ok = check_joint_input(L, q0, v0, a0, dv);
  Prologue pro = ost_prologue(L, Ts, qg, q0, v0, a0, dv);
  int st1 = ost_body_dv<false, LTP_FAST_RV != 0>(L, Ts, pro, qg, q0, L.v_max, t_opt, mod, opt_case, dv);
  if (dv.bad) {
    Stage1Redo r;
    stage1_checked(P, jt, qg, q0, v0, a0, &r);
    in_ok = r.in_ok != 0;
    pro = r.pro;
#pragma unroll
    for (int k = 0; k < 7; ++k) t_opt[k] = r.t_opt[k];
    st1 = r.st;
    mod = r.mod;
    opt_case = r.opt_case;
  }
#else
  const bool in_ok = check_joint_input(L, q0, v0, a0);
  const Prologue pro = ost_prologue(L, Ts, qg, q0, v0, a0);
  const int st1 = ost_body_t<false, true>(L, Ts, pro, qg, q0, L.v_max, t_opt, mod, opt_case);
#endif
  sh.t6[jt * kTile + lane] = t_opt[6];
  sh.flag[jt * kTile + lane] = (unsigned char)((!(in_ok && st1 != OST_FAIL) ? 1 : 0) | (st1 == OST_DEFER ? 2 : 0));
  // item mode: the joints that need the quartic tail are listed below; the barrier that stage 2
  // needs anyway tells every thread whether the CTA has any (none for realistic limits)
  const bool tail_item = items && valid && st1 == OST_DEFER;
  const int cta_has_tail = __syncthreads_or(tail_item);
  // stage 2 (cc:31-39): strict '>' so the lowest joint index wins ties, NaN never wins
  double t_req = -1;
  int slowest = -1;
  unsigned char any = 0;
  for (int i = 0; i < dof; ++i) {
    any |= sh.flag[i * kTile + lane];
    const double ti = sh.t6[i * kTile + lane];
    if (ti > t_req) {
      t_req = ti;
      slowest = i;
    }
  }
  // a joint that needs the quartic tail has no t_opt yet: the whole problem is deferred
  const bool defer1 = (any & 2) != 0;
  const bool reached = !(any & 1) && slowest != -1;
  if (cta_has_tail) {  // uniform over the CTA
    // tail items: slot = base of this CTA (one atomic per CTA) + items of the warps before mine +
    // those of the lanes before mine; the problems that wait for them (every warp sees the same
    // set; warp 0 lists them)
    const unsigned tail_mask = __ballot_sync(0xffffffffu, tail_item);
    const unsigned pend = __ballot_sync(0xffffffffu, defer1 && valid);
    if (lane == 0) sh.tail_items[jt] = __popc(tail_mask);
    __syncthreads();
    int t_before = 0, t_total = 0;
    for (int w = 0; w < dof; ++w) {
      const int c = sh.tail_items[w];
      t_before += (w < jt) ? c : 0;
      t_total += c;
    }
    if (jt == 0 && lane == 0) {
      sh.tail_items[dof] = atomicAdd(X.counters + kCntTail, t_total);
      sh.tail_items[dof + 1] = atomicAdd(X.counters + kCntPending, __popc(pend));
    }
    __syncthreads();
    const unsigned below = (1u << lane) - 1u;
    if (tail_item) X.tail_items[sh.tail_items[dof] + t_before + __popc(tail_mask & below)] = make_int2((int)p, jt);
    if (jt == 0 && ((pend >> lane) & 1u)) X.pending[sh.tail_items[dof + 1] + __popc(pend & below)] = (int)p;
  }
  // stage 3 (cc:42-55), closed-form attempts only. Attempt 1 runs here; the joints it does
  // not settle (about a third) are queued for ltp_solve_attempt2_kernel, which runs attempt 2
  // on full warps of such joints ("grouped by case").
  double t_sc[7];
  zero7(t_sc);
  double v_drive = L.v_max;
  unsigned char ts_case = 255, final_case = 255;
  bool my_defer = false, need2 = false;
  if (reached && !defer1) {
    if (jt == slowest) {
      ts_case = 0;
      final_case = opt_case;
    } else {
      const TsInput I = make_ts_input(qg, q0, v0, a0, pro.dir, t_req);
#if LTP_FAST_DEFER
      const unsigned char mod1 = mod;
      Dv dv2;
      int c = time_scaling_attempt1(L, Ts, pro, I, t_sc, v_drive, mod, final_case, dv2);
      if (dv2.bad) {
        Attempt1Redo r;
        attempt1_checked(P, jt, qg, q0, v0, a0, t_req, mod1, 255, &r);
#pragma unroll
        for (int k = 0; k < 7; ++k) t_sc[k] = r.t[k];
        v_drive = r.v_drive;
        c = r.c;
        mod = r.mod;
        final_case = r.final_case;
      }
#else
      const int c = time_scaling_attempt1(L, Ts, pro, I, t_sc, v_drive, mod, final_case);
#endif
      my_defer = (c == 0);
      need2 = (c == -1) && valid;
      if (items && my_defer && valid) push_search_item(X, p, jt, t_req);  // rare: a nested solve with a tail
      ts_case = (unsigned char)c;
      if (c == 9) final_case = opt_case;
    }
    if (!need2) {
      double m = t_sc[0];
#pragma unroll
      for (int k = 1; k < 7; ++k)
        if (m < t_sc[k]) m = t_sc[k];
      if (m <= 0.0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) t_sc[k] = t_opt[k];
      }
    }
  }
  // queue slot = base of this warp in its joint's sub-queue (one atomic per warp) + the queued
  // lanes before mine
  const unsigned need_mask = __ballot_sync(0xffffffffu, need2);
  if (need_mask != 0) {  // uniform over the warp
    int base = 0;
    if (lane == 0) base = atomicAdd(X.counters + kCounters + jt, __popc(need_mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (need2) {
      const int64_t e = (int64_t)jt * n + base + __popc(need_mask & ((1u << lane) - 1u));
      X.queue.t_req[e] = t_req;
      X.queue.q_goal[e] = qg;
      X.queue.q_0[e] = q0;
      X.queue.v_0[e] = v0;
      X.queue.a_0[e] = a0;
      X.queue.where[e] = make_int2((int)p, jt);
    }
  }
  if (!valid) return;
  store_joint_opt(S, dof, jt, n, p, t_opt, pro.dir, opt_case);
  // a queued joint contributes length 0 here; its length arrives by atomicMax later
  const int my_len = (reached && !defer1 && !my_defer && !need2) ? joint_samples(t_sc, Ts, P.r_ts) : 0;
  // legacy mode: a joint that is not settled here sends its whole problem to the every-branch
  // kernel; item mode: only a non-finite time does (the joint itself is listed as an item)
  const bool whole = items ? (my_len < 0) : (my_defer || defer1);
  if (finish_problem(sh, S, dof, lane, jt, p, my_len, whole, reached, slowest, items && defer1)) {
    const int slot = atomicAdd(work_count, 1);
    work_list[slot] = (int)p;
  }
  if (!need2 && !(items && my_defer)) store_joint_scaled(S, dof, jt, n, p, t_sc, v_drive, mod, ts_case, final_case);
}

// Attempt 2 of the cruise-speed search (reference cc:408-446) for the queued joints: one
// thread per entry, consecutive entries in consecutive lanes, so every lane of every warp
// has the same work whatever the mix of joints that needed it. The start state comes with the
// entry and the prologue is recomputed; results go straight to the solution arrays, the
// sample count joins the problem's by atomicMax. A joint that attempt 2 does not settle
// either sends its problem to the work list of the generic kernel (once per problem: the
// exchange on traj_len elects the sender).
#ifndef LTP_TM_BUILD_EARLY_EXIT
#define LTP_TM_BUILD_EARLY_EXIT 1
#endif
#ifndef LTP_A2_MINB
#define LTP_A2_MINB 2
#endif
__global__ void __launch_bounds__(256, LTP_A2_MINB)
ltp_solve_attempt2_kernel(const __grid_constant__ PlannerParams P, int64_t n, DeviceSolution S, SolveScratch X,
                          int items) {
  const int dof = P.dof;
  const double Ts = P.ts;
  const int step = gridDim.x * blockDim.x;
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  // the sub-queues end to end: flat entry f lives at joint * n + (f - entries of the joints before)
  __shared__ int sub_count[LTP_MAX_DOF];
  if ((int)threadIdx.x < dof) sub_count[threadIdx.x] = X.counters[kCounters + threadIdx.x];
  __syncthreads();
  int count = 0;
  for (int j = 0; j < dof; ++j) count += sub_count[j];
  auto slot_of = [&](int f) -> int64_t {
    int j = 0;
    while (j < dof - 1 && f >= sub_count[j]) f -= sub_count[j++];
    return (int64_t)j * n + f;
  };
  if (e >= count) return;
  // software pipeline: the next entry's six loads are in flight while this one is evaluated
  // (the kernel was waiting on memory for half of its cycles, four warps per scheduler)
  int64_t at = slot_of(e);
  int2 where = X.queue.where[at];
  double t_req = X.queue.t_req[at], qg = X.queue.q_goal[at], q0 = X.queue.q_0[at], v0 = X.queue.v_0[at],
         a0 = X.queue.a_0[at];
  while (true) {
    const int en = e + step;
    const bool more = en < count;
    int2 where_n = where;
    double t_req_n = 0, qg_n = 0, q0_n = 0, v0_n = 0, a0_n = 0;
    if (more) {
      at = slot_of(en);
      where_n = X.queue.where[at];
      t_req_n = X.queue.t_req[at]; qg_n = X.queue.q_goal[at]; q0_n = X.queue.q_0[at];
      v0_n = X.queue.v_0[at]; a0_n = X.queue.a_0[at];
    }
    const int64_t p = where.x;
    const int jt = where.y;
    // (a problem that is already on its way to the generic kernel is not skipped: what is
    // stored here is overwritten there, and a look at traj_len first costs a second trip to
    // memory per entry)
    const JointLimits L = P.joint(jt);
    double t[7];
    zero7(t);
    double v_drive = L.v_max;
    unsigned char mod = 0, final_case = 255;
#if LTP_FAST_DEFER
    // deferred range test of the divisions, as in the closed-form kernel
    DivDeferred dv;
    const Prologue pro = ost_prologue(L, Ts, qg, q0, v0, a0, dv);
    const TsInput I = make_ts_input(qg, q0, v0, a0, pro.dir, t_req);
    int c = time_scaling_attempt2(L, Ts, pro, I, t, v_drive, mod, final_case, dv);
    if (dv.bad) {
      Attempt2Redo r;
      attempt2_checked(P, jt, qg, q0, v0, a0, t_req, &r);
#pragma unroll
      for (int k = 0; k < 7; ++k) t[k] = r.t[k];
      v_drive = r.v_drive;
      c = r.c;
      mod = r.mod;
      final_case = r.final_case;
    }
#else
    const Prologue pro = ost_prologue(L, Ts, qg, q0, v0, a0);
    const TsInput I = make_ts_input(qg, q0, v0, a0, pro.dir, t_req);
    const int c = time_scaling_attempt2(L, Ts, pro, I, t, v_drive, mod, final_case);
#endif
    double m = t[0];
#pragma unroll
    for (int k = 1; k < 7; ++k)
      if (m < t[k]) m = t[k];
    // an accepted solve with no positive time falls back to the time-optimal times (cc:50-55),
    // which are not at hand here: that (degenerate) problem goes to the generic kernel too
    const bool open = (c == 0 || m <= 0.0);
    const int len = open ? -1 : joint_samples(t, Ts, P.r_ts);
    if (items && open) {
      push_search_item(X, p, jt, t_req);  // candidates 3..8 (or the fallback) for this joint alone
    } else if (len < 0) {
      push_whole_problem(S, X, p);
    } else {
      store_joint_scaled(S, dof, jt, n, p, t, v_drive, mod, (unsigned char)c, final_case);
      atomicMax(S.traj_len + p, len);
    }
    if (!more) break;
    e = en;
    where = where_n;
    t_req = t_req_n; qg = qg_n; q0 = q0_n; v0 = v0_n; a0 = a0_n;
  }
}

// Tail items: the time-optimal solve of one (problem, joint) including the quartic tail
// (cc:245-337, one or two quartic root solves), one thread per item -- every lane of a warp runs
// the root finder. The result is parked in the joint's record (times in slots 0..6, the flags in
// the v_drive slot; the pending kernel overwrites it with the final values) and in dir / t_opt /
// opt_case.
__global__ void __launch_bounds__(128)
ltp_solve_tail_kernel(const __grid_constant__ PlannerParams P, int64_t n, const double* __restrict__ q_goal,
                      const double* __restrict__ q_0, const double* __restrict__ v_0,
                      const double* __restrict__ a_0, DeviceSolution S, SolveScratch X) {
  const int dof = P.dof;
  const int count = X.counters[kCntTail];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
    const int2 w = X.tail_items[e];
    const int64_t p = w.x;
    const int jt = w.y;
    const JointLimits L = P.joint(jt);
    const int64_t at = (int64_t)jt * n + p;
    const double qg = q_goal[at], q0 = q_0[at];
    const Prologue pro = ost_prologue(L, P.ts, qg, q0, v_0[at], a_0[at]);
    double t_opt[7];
    zero7(t_opt);
    unsigned char mod = 0, opt_case = 255;
    const bool ok = ost_body(L, P.ts, pro, qg, q0, L.v_max, t_opt, mod, opt_case);
    store_record(record_of(S, jt, n, p), t_opt, pack_tail_flags(mod, opt_case, ok ? 1u : 0u));
    store_joint_opt(S, dof, jt, n, p, t_opt, pro.dir, opt_case);
  }
}

// Pending problems: stages 2 and 3 of the problems that waited for a tail, same mapping as the
// closed-form kernel (32 problems x dof joints per CTA, grid-stride over the pending list). The
// tail joints read their time-optimal solve back; the search runs its two closed-form candidates
// (cc:378-446); a joint they do not settle becomes a search item.
template <int MAXW>
__global__ void __launch_bounds__(kTile * MAXW, MAXW <= 8 ? 2 : 1)
ltp_solve_pending_kernel(const __grid_constant__ PlannerParams P, int64_t n, const double* __restrict__ q_goal,
                         const double* __restrict__ q_0, const double* __restrict__ v_0,
                         const double* __restrict__ a_0, DeviceSolution S, SolveScratch X) {
  extern __shared__ unsigned char smem_raw[];
  const int dof = P.dof;
  const SolveShared sh = carve_shared(smem_raw, dof);
  const int lane = threadIdx.x, jt = threadIdx.y;
  const JointLimits L = P.joint(jt);
  const double Ts = P.ts;
  const int64_t count = X.counters[kCntPending];
  for (int64_t tile = blockIdx.x; tile * kTile < count; tile += gridDim.x) {
    const int64_t w = tile * kTile + lane;
    const bool valid = w < count;
    const int64_t p = valid ? (int64_t)X.pending[w] : 0;
    const int64_t at = (int64_t)jt * n + p;
    double qg = 0, q0 = 0, v0 = 0, a0 = 0;
    if (valid) {
      qg = q_goal[at]; q0 = q_0[at]; v0 = v_0[at]; a0 = a_0[at];
    }
    if (jt == 0) sh.arrived[lane] = 0;
    // stage 1 again (cc:14-30), the tail joints from what ltp_solve_tail_kernel left
    const bool in_ok = check_joint_input(L, q0, v0, a0);
    const Prologue pro = ost_prologue(L, Ts, qg, q0, v0, a0);
    double t_opt[7];
    zero7(t_opt);
    unsigned char mod = 0, opt_case = 255;
    const int st1 = ost_body_t<false>(L, Ts, pro, qg, q0, L.v_max, t_opt, mod, opt_case);
    bool ost_ok = st1 == OST_OK;
    if (st1 == OST_DEFER && valid) {
      double parked;
      load_record(record_of(S, jt, n, p), t_opt, parked);
      const unsigned f = (unsigned)__double2loint(parked);
      mod = (unsigned char)(f & 255u);
      opt_case = (unsigned char)((f >> 8) & 255u);
      ost_ok = ((f >> 16) & 1u) != 0;
    }
    sh.t6[jt * kTile + lane] = t_opt[6];
    sh.flag[jt * kTile + lane] = (unsigned char)(!(in_ok && ost_ok));
    __syncthreads();
    // stage 2 (cc:31-39)
    double t_req = -1;
    int slowest = -1;
    bool any_fail = false;
    for (int i = 0; i < dof; ++i) {
      any_fail |= (sh.flag[i * kTile + lane] & 1) != 0;
      const double ti = sh.t6[i * kTile + lane];
      if (ti > t_req) {
        t_req = ti;
        slowest = i;
      }
    }
    const bool reached = !any_fail && slowest != -1;
    // stage 3 (cc:42-55), closed-form candidates
    double t_sc[7];
    zero7(t_sc);
    double v_drive = L.v_max;
    unsigned char ts_case = 255, final_case = 255;
    bool open = false;
    if (reached) {
      if (jt == slowest) {
        ts_case = 0;
        final_case = opt_case;
      } else {
        const TsInput I = make_ts_input(qg, q0, v0, a0, pro.dir, t_req);
        const int c = time_scaling_closed_form(L, Ts, pro, I, t_sc, v_drive, mod, final_case);
        open = c == 0;
        ts_case = (unsigned char)c;
        if (c == 9) final_case = opt_case;
      }
      double m = t_sc[0];
#pragma unroll
      for (int k = 1; k < 7; ++k)
        if (m < t_sc[k]) m = t_sc[k];
      if (m <= 0.0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) t_sc[k] = t_opt[k];
      }
    }
    if (valid) {
      if (open) push_search_item(X, p, jt, t_req);
      const int my_len = (reached && !open) ? joint_samples(t_sc, Ts, P.r_ts) : 0;
      if (finish_problem(sh, S, dof, lane, jt, p, my_len, my_len < 0, reached, slowest)) {
        const int slot = atomicAdd(X.counters + kCntWork, 1);
        X.work_list[slot] = (int)p;
      }
      if (!open) store_joint_scaled(S, dof, jt, n, p, t_sc, v_drive, mod, ts_case, final_case);  // else: search kernel
    }
    __syncthreads();  // shared arrays are reused by the next tile
  }
}

// Search items: the whole cruise-speed search (cc:358-645) of one (problem, joint) whose two
// closed-form candidates did not settle it, one thread per item -- full warps of joints that all
// go on to the root-finder candidates. The result joins the problem like in the attempt-2 kernel.
__global__ void __launch_bounds__(128)
ltp_solve_search_kernel(const __grid_constant__ PlannerParams P, int64_t n, const double* __restrict__ q_goal,
                        const double* __restrict__ q_0, const double* __restrict__ v_0,
                        const double* __restrict__ a_0, DeviceSolution S, SolveScratch X) {
  const int dof = P.dof;
  const double Ts = P.ts;
  const int count = X.counters[kCntSearch];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < count; e += gridDim.x * blockDim.x) {
    const int2 w = X.search_items[e];
    const double t_req = X.search_t_req[e];
    const int64_t p = w.x;
    const int jt = w.y;
    const JointLimits L = P.joint(jt);
    const int64_t at = (int64_t)jt * n + p;
    const double qg = q_goal[at], q0 = q_0[at], v0 = v_0[at], a0 = a_0[at];
    const Prologue pro = ost_prologue(L, Ts, qg, q0, v0, a0);
    // the joint's time-optimal solve again: its case byte and times are what a failed search
    // (cc:641-644) and the cc:50-55 fallback report
    double t_opt[7];
    zero7(t_opt);
    unsigned char mod = 0, opt_case = 255;
    ost_body(L, Ts, pro, qg, q0, L.v_max, t_opt, mod, opt_case);
    const TsInput I = make_ts_input(qg, q0, v0, a0, pro.dir, t_req);
    double t_sc[7];
    zero7(t_sc);
    double v_drive = L.v_max;
    unsigned char final_case = 255;
    const unsigned char ts_case = (unsigned char)time_scaling_from(1, L, Ts, pro, I, t_sc, v_drive, mod, final_case);
    if (ts_case == 9) final_case = opt_case;
    double m = t_sc[0];
#pragma unroll
    for (int k = 1; k < 7; ++k)
      if (m < t_sc[k]) m = t_sc[k];
    if (m <= 0.0) {
#pragma unroll
      for (int k = 0; k < 7; ++k) t_sc[k] = t_opt[k];
    }
    const int len = joint_samples(t_sc, Ts, P.r_ts);
    if (len < 0) {
      push_whole_problem(S, X, p);
    } else {
      store_joint_scaled(S, dof, jt, n, p, t_sc, v_drive, mod, ts_case, final_case);
      atomicMax(S.traj_len + p, len);
    }
  }
}

// every branch evaluated in-thread. work_list == nullptr: problem = tile index (generic-only
// mode); otherwise tiles of 32 entries of the work list, grid-stride.
template <int MAXW>
__global__ void __launch_bounds__(kTile * MAXW)
ltp_solve_generic_kernel(const __grid_constant__ PlannerParams P, int64_t n, const double* __restrict__ q_goal,
                         const double* __restrict__ q_0, const double* __restrict__ v_0,
                         const double* __restrict__ a_0, DeviceSolution S, const int* __restrict__ work_list,
                         const int* __restrict__ work_count) {
  extern __shared__ unsigned char smem_raw[];
  const int dof = P.dof;
  const SolveShared sh = carve_shared(smem_raw, dof);
  const int lane = threadIdx.x, jt = threadIdx.y;
  const JointLimits L = P.joint(jt);
  const double Ts = P.ts;
  const int64_t count = work_list ? (int64_t)*work_count : n;
  for (int64_t tile = blockIdx.x; tile * kTile < count; tile += gridDim.x) {
    const int64_t w = tile * kTile + lane;
    const bool valid = w < count;
    const int64_t p = valid ? (work_list ? (int64_t)work_list[w] : w) : 0;
    const int64_t at = (int64_t)jt * n + p;
    double qg = 0, q0 = 0, v0 = 0, a0 = 0;
    if (valid) {
      qg = q_goal[at]; q0 = q_0[at]; v0 = v_0[at]; a0 = a_0[at];
    }
    if (jt == 0) sh.arrived[lane] = 0;
    // stage 1 (cc:14-30)
    const bool in_ok = check_joint_input(L, q0, v0, a0);
    const Prologue pro = ost_prologue(L, Ts, qg, q0, v0, a0);
    double t_opt[7];
    zero7(t_opt);
    unsigned char mod = 0, opt_case = 255;
    const bool ost_ok = ost_body(L, Ts, pro, qg, q0, L.v_max, t_opt, mod, opt_case);
    sh.t6[jt * kTile + lane] = t_opt[6];
    sh.flag[jt * kTile + lane] = (unsigned char)(!(in_ok && ost_ok));
    __syncthreads();
    // stage 2 (cc:31-39)
    double t_req = -1;
    int slowest = -1;
    bool any_fail = false;
    for (int i = 0; i < dof; ++i) {
      any_fail |= (sh.flag[i * kTile + lane] & 1) != 0;
      const double ti = sh.t6[i * kTile + lane];
      if (ti > t_req) {
        t_req = ti;
        slowest = i;
      }
    }
    const bool reached = !any_fail && slowest != -1;
    // stage 3 (cc:42-55)
    double t_sc[7];
    zero7(t_sc);
    double v_drive = L.v_max;
    unsigned char ts_case = 255, final_case = 255;
    if (reached) {
      if (jt == slowest) {
        ts_case = 0;
        final_case = opt_case;
      } else {
        const TsInput I = make_ts_input(qg, q0, v0, a0, pro.dir, t_req);
        ts_case = (unsigned char)time_scaling_from(1, L, Ts, pro, I, t_sc, v_drive, mod, final_case);
        if (ts_case == 9) final_case = opt_case;
      }
      double m = t_sc[0];
#pragma unroll
      for (int k = 1; k < 7; ++k)
        if (m < t_sc[k]) m = t_sc[k];
      if (m <= 0.0) {
#pragma unroll
        for (int k = 0; k < 7; ++k) t_sc[k] = t_opt[k];
      }
    }
    const int my_len = reached ? joint_samples(t_sc, Ts, P.r_ts) : 0;
    if (valid) {
      finish_problem(sh, S, dof, lane, jt, p, my_len, false, reached, slowest);
      store_joint(S, dof, jt, n, p, t_sc, t_opt, pro.dir, v_drive, mod, opt_case, ts_case, final_case);
    }
    __syncthreads();  // shared arrays are reused by the next tile
  }
}

// ------------------------------------------------------------------------------------
// per-joint primitives. grid.y = joint (or 1 with joint_fixed >= 0), 1 thread per item
// ------------------------------------------------------------------------------------
__global__ void ltp_opt_braking_kernel(const __grid_constant__ PlannerParams P, int joint_fixed,
                                       int64_t n, const double* __restrict__ v_0,
                                       const double* __restrict__ a_0, double* q_stop, double* t_rel,
                                       double* dir) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int jt = joint_fixed >= 0 ? joint_fixed : blockIdx.y;
  const int row = joint_fixed >= 0 ? 0 : jt, rows = joint_fixed >= 0 ? 1 : P.dof;
  const JointLimits L = P.joint(jt);
  const int64_t at = (int64_t)row * n + p;
  double T0, T1, T2, d;
  const double q = brake_profile(L, P.ts, v_0[at], a_0[at], T0, T1, T2, d);
  q_stop[at] = q;
  dir[at] = d;
  t_rel[((int64_t)0 * rows + row) * n + p] = T0;
  t_rel[((int64_t)1 * rows + row) * n + p] = T1;
  t_rel[((int64_t)2 * rows + row) * n + p] = T2;
}

__global__ void ltp_opt_switch_times_kernel(const __grid_constant__ PlannerParams P, int joint_fixed,
                                            int64_t n, const double* __restrict__ q_goal,
                                            const double* __restrict__ q_0,
                                            const double* __restrict__ v_0,
                                            const double* __restrict__ a_0,
                                            const double* __restrict__ v_drive, double* t, double* dir,
                                            uint8_t* mod, uint8_t* kase, uint8_t* ok) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int jt = joint_fixed >= 0 ? joint_fixed : blockIdx.y;
  const int row = joint_fixed >= 0 ? 0 : jt, rows = joint_fixed >= 0 ? 1 : P.dof;
  const JointLimits L = P.joint(jt);
  const int64_t at = (int64_t)row * n + p;
  const double qg = q_goal[at], q0 = q_0[at];
  const Prologue pro = ost_prologue(L, P.ts, qg, q0, v_0[at], a_0[at]);
  double tt[7];
  zero7(tt);
  unsigned char m = 0, c = 255;
  const bool good = ost_body(L, P.ts, pro, qg, q0, v_drive[at], tt, m, c);
#pragma unroll
  for (int k = 0; k < 7; ++k) t[((int64_t)k * rows + row) * n + p] = tt[k];
  dir[at] = pro.dir;
  mod[at] = m;
  if (kase) kase[at] = c;
  ok[at] = (uint8_t)good;
}

__global__ void ltp_time_scaling_kernel(const __grid_constant__ PlannerParams P, int joint_fixed,
                                        int64_t n, const double* __restrict__ q_goal,
                                        const double* __restrict__ q_0, const double* __restrict__ v_0,
                                        const double* __restrict__ a_0, const double* __restrict__ dir,
                                        const double* __restrict__ t_required, double* t,
                                        double* v_drive, uint8_t* mod, uint8_t* ts_case,
                                        uint8_t* final_case, uint8_t* ok) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int jt = joint_fixed >= 0 ? joint_fixed : blockIdx.y;
  const int row = joint_fixed >= 0 ? 0 : jt, rows = joint_fixed >= 0 ? 1 : P.dof;
  const JointLimits L = P.joint(jt);
  const int64_t at = (int64_t)row * n + p;
  const double qg = q_goal[at], q0 = q_0[at], d = dir[at];
  const TsInput I = make_ts_input(qg, q0, v_0[at], a_0[at], d, t_required[at]);
  // the reference re-enters optSwitchTimes with (dir * v_0, dir * a_0) (cc:400); for
  // dir = +-1 that is the caller's start state again
  const Prologue pro = ost_prologue(L, P.ts, qg, q0, d * I.v_0, d * I.a_0);
  double tt[7];
  zero7(tt);
  double vd = 0;
  unsigned char m = 0, fc = 255;
  const int c = time_scaling_from(1, L, P.ts, pro, I, tt, vd, m, fc);
#pragma unroll
  for (int k = 0; k < 7; ++k) t[((int64_t)k * rows + row) * n + p] = tt[k];
  v_drive[at] = vd;
  mod[at] = m;
  if (ts_case) ts_case[at] = (uint8_t)c;
  if (final_case) final_case[at] = fc;
  ok[at] = (uint8_t)(c != 9);
}

// ------------------------------------------------------------------------------------
// stage 4: dense q/v/a/j sampling (reference cc:706-841) + final limit check (cc:59-61)
//
// One thread per (problem, joint) row runs the reference's sequential recurrence (it is what
// defines the result: a closed form would differ in the last bits) and streams the row out
// four samples per field at a time: one 256-bit, sector-aligned, evict-first store per field
// (STG.E.EF.256), so every 32-byte sector is written exactly once and completely. The
// arithmetic is ~5 FP64 operations per 32 output bytes; the kernel is bound by HBM writes.
// A CTA is one warp: floor(32/dof) whole problems, joint index fastest, so the final
// per-problem limit check is a warp vote and the grid (n*dof/28 CTAs for 7 joints) spreads
// evenly over the SMs even for a few thousand problems.
// ------------------------------------------------------------------------------------
using RowCursor = SegCursorT<64>;
__device__ __forceinline__ void cursor_begin(RowCursor& C, const RowSampler& R, const SegTableT<64>&) { C.begin(R); }

__device__ __forceinline__ void store4(double* dst, double x0, double x1, double x2, double x3) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(x0), "d"(x1), "d"(x2), "d"(x3)
               : "memory");
}

template <bool VEC>
__global__ void __launch_bounds__(32)
ltp_sample_kernel(const __grid_constant__ PlannerParams P, int64_t n, int ppb, const double* __restrict__ q_0,
                  const double* __restrict__ v_0, const double* __restrict__ a_0, DeviceSolution S,
                  int horizon, int64_t stride, double* __restrict__ q, double* __restrict__ v,
                  double* __restrict__ a, double* __restrict__ j, uint8_t* success) {
  __shared__ __align__(16) double s_tab[kMaxSeg][32][2];  // per-lane segment table, [entry][lane][word]
  const int dof = P.dof;
  const int lane = threadIdx.x;
  const SegTableT<64> T{&s_tab[0][lane][0]};
  const int slot = lane / dof, jt = lane - slot * dof;   // problem slot within the CTA, joint
  const int64_t p = (int64_t)blockIdx.x * ppb + slot;
  const bool valid = slot < ppb && p < n;
  bool row_ok = false;
  if (valid && horizon > 0 && !(S.reached[p] && S.traj_len[p] > 0)) {
    // fixed horizon: a problem without a plan holds its start position (ltp_b200.h)
    const double hold = q_0[(int64_t)jt * n + p];
    const int64_t base = (p * dof + jt) * stride;
    for (int i = 0; i < horizon; ++i) {
      q[base + i] = hold; v[base + i] = 0.0; a[base + i] = 0.0; j[base + i] = 0.0;
    }
  }
  if (valid && S.reached[p]) {
    const int len = S.traj_len[p];
    if (len > 0) {
      const JointLimits L = P.joint(jt);
      const int64_t at = (int64_t)jt * n + p;
      double t[7], v_drive_row;
      load_record(record_of(S, jt, n, p), t, v_drive_row);
      // samples stored: the fixed horizon, or the exact length clipped to the row capacity
      const int n_out = horizon > 0 ? horizon : (len < stride ? len : (int)stride);
      const int n_run = n_out > len ? n_out : len;    // samples computed
      RowCursor C;
      {
        RowSampler R;
        R.init(P.ts, L.j_max, t, S.dir[at], S.mod[at], q_0[at], v_0[at], a_0[at], v_drive_row, n_run);
        T.build(R, n_run);
        cursor_begin(C, R, T);
      }
      const int64_t base = (p * dof + jt) * stride;
      double* qo = q + base; double* vo = v + base; double* ao = a + base; double* jo = j + base;
      double q_last = 0.0;
      int i = 0;
      if (VEC) {
        const int n_vec = n_out & ~3;
        for (; i < n_vec; i += 4) {
          double jj[4], aa[4], vv[4], qq[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) C.step(T, i + u, jj[u], aa[u], vv[u], qq[u]);
          const unsigned k = (unsigned)(len - 1 - i);
          if (k < 4u) q_last = k == 0 ? qq[0] : k == 1 ? qq[1] : k == 2 ? qq[2] : qq[3];
          store4(qo + i, qq[0], qq[1], qq[2], qq[3]);
          store4(vo + i, vv[0], vv[1], vv[2], vv[3]);
          store4(ao + i, aa[0], aa[1], aa[2], aa[3]);
          store4(jo + i, jj[0], jj[1], jj[2], jj[3]);
        }
      }
      for (; i < n_run; ++i) {
        double jj, aa, vv, qq;
        C.step(T, i, jj, aa, vv, qq);
        if (i == len - 1) q_last = qq;
        if (i < n_out) {
          qo[i] = qq; vo[i] = vv; ao[i] = aa; jo[i] = jj;
        }
      }
      row_ok = !(q_last < L.q_min || q_last > L.q_max);  // cc:60
    }
  }
  // all joints of a problem sit in adjacent lanes of this warp
  const unsigned ok_mask = __ballot_sync(0xffffffffu, row_ok);
  if (valid && jt == 0) {
    const unsigned want = ((dof >= 32) ? 0xffffffffu : ((1u << dof) - 1u)) << (slot * dof);
    success[p] = (uint8_t)((ok_mask & want) == want);
  }
}

// Latency variant of the row sampler for a handful of rows (the single-plan host calls): one
// row per one-warp CTA; the warp builds the piece table together, lane 0 then runs the row.
// The batch kernels above hide the latency of the per-sample chain
// (table look-up, then a -> v -> q) behind thousands of other rows; with seven rows in flight
// that chain, ~115 cycles per sample, IS the run time. Here the row is walked piece by piece:
// inside a piece the jerk and the update rule are constants, so the three recurrences
// a += Ts*j, v += Ts*a, q += Ts*v are three independent chains of one dependent add per sample
// and run overlapped, four samples per iteration. Same operations on the same operands as
// SegCursorT::step, hence the same bits. Writes row_ok[row] (1 = the row ends inside its joint
// limits, cc:59-61); the caller combines the rows of a problem.
template <bool LIVE, bool CRUISE>
__device__ __forceinline__ void run_piece(int& i, const int stop, const int n_out, const int len, const double Ts,
                                          const double jv, const double vcruise, double& a, double& v, double& q,
                                          double& q_last, double* qo, double* vo, double* ao, double* jo) {
  const double c = Ts * jv;
  if (LIVE && !CRUISE) {
    // The common piece (accelerating or braking): blocks of four samples, software-pipelined
    // one block deep per recurrence -- while q of block b-1 is summed up, v of block b and a of
    // block b+1 are already on their way, so an iteration costs one chain of four dependent
    // adds (~28 cycles each on this part) instead of the eight of a -> v -> q in sequence.
    while ((i & 3) != 0 && i < stop) {
      a = a + c;
      v = v + Ts * a;
      q = q + Ts * v;
      if (i == len - 1) q_last = q;
      if (i < n_out) {
        qo[i] = q; vo[i] = v; ao[i] = a; jo[i] = jv;
      }
      ++i;
    }
    const int lim = stop < n_out ? stop : n_out;
    const int nblk = (lim - i) >> 2;
    if (nblk >= 2) {
#define LTP_CHAIN_A(dst, from)            \
  dst[0] = (from) + c; dst[1] = dst[0] + c; dst[2] = dst[1] + c; dst[3] = dst[2] + c
#define LTP_CHAIN_S(dst, from, src)       \
  {                                       \
    const double t0_ = Ts * src[0], t1_ = Ts * src[1], t2_ = Ts * src[2], t3_ = Ts * src[3]; \
    dst[0] = (from) + t0_; dst[1] = dst[0] + t1_; dst[2] = dst[1] + t2_; dst[3] = dst[2] + t3_; \
  }
#define LTP_EMIT(at, ab, vb, qb)                                                     \
  {                                                                                  \
    const unsigned k_ = (unsigned)(len - 1 - (at));                                  \
    if (k_ < 4u) q_last = k_ == 0 ? qb[0] : k_ == 1 ? qb[1] : k_ == 2 ? qb[2] : qb[3]; \
    store4(qo + (at), qb[0], qb[1], qb[2], qb[3]);                                   \
    store4(vo + (at), vb[0], vb[1], vb[2], vb[3]);                                   \
    store4(ao + (at), ab[0], ab[1], ab[2], ab[3]);                                   \
    store4(jo + (at), jv, jv, jv, jv);                                               \
  }
      double a_cur[4], a_nxt[4], v_cur[4], qb[4];
      LTP_CHAIN_A(a_cur, a);                 // block 0
      LTP_CHAIN_A(a_nxt, a_cur[3]);          // block 1
      LTP_CHAIN_S(v_cur, v, a_cur);          // block 0
      for (int b = 1; b < nblk - 1; ++b) {
        double a_nn[4], v_nxt[4];
        LTP_CHAIN_A(a_nn, a_nxt[3]);         // a of block b + 1
        LTP_CHAIN_S(v_nxt, v_cur[3], a_nxt); // v of block b
        LTP_CHAIN_S(qb, q, v_cur);           // q of block b - 1
        q = qb[3];
        LTP_EMIT(i + 4 * (b - 1), a_cur, v_cur, qb);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a_cur[u] = a_nxt[u]; a_nxt[u] = a_nn[u]; v_cur[u] = v_nxt[u];
        }
      }
      double v_nxt[4];
      LTP_CHAIN_S(v_nxt, v_cur[3], a_nxt);   // v of the last block
      LTP_CHAIN_S(qb, q, v_cur);             // q of the last block but one
      q = qb[3];
      LTP_EMIT(i + 4 * (nblk - 2), a_cur, v_cur, qb);
      LTP_CHAIN_S(qb, q, v_nxt);             // q of the last block
      LTP_EMIT(i + 4 * (nblk - 1), a_nxt, v_nxt, qb);
      a = a_nxt[3]; v = v_nxt[3]; q = qb[3];
      i += 4 * nblk;
#undef LTP_CHAIN_A
#undef LTP_CHAIN_S
#undef LTP_EMIT
    }
  }
  while (i < stop) {
    if ((i & 3) == 0 && i + 4 <= stop && i + 4 <= n_out) {
      double aa[4], vv[4], qq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a = LIVE ? a + c : 0.0;
        v = CRUISE ? vcruise : (LIVE ? v + Ts * a : 0.0);
        q = q + Ts * v;
        aa[u] = a; vv[u] = v; qq[u] = q;
      }
      const unsigned k = (unsigned)(len - 1 - i);
      if (k < 4u) q_last = k == 0 ? qq[0] : k == 1 ? qq[1] : k == 2 ? qq[2] : qq[3];
      store4(qo + i, qq[0], qq[1], qq[2], qq[3]);
      store4(vo + i, vv[0], vv[1], vv[2], vv[3]);
      store4(ao + i, aa[0], aa[1], aa[2], aa[3]);
      store4(jo + i, jv, jv, jv, jv);
      i += 4;
    } else {
      a = LIVE ? a + c : 0.0;
      v = CRUISE ? vcruise : (LIVE ? v + Ts * a : 0.0);
      q = q + Ts * v;
      if (i == len - 1) q_last = q;
      if (i < n_out) {
        qo[i] = q; vo[i] = v; ao[i] = a; jo[i] = jv;
      }
      ++i;
    }
  }
}

__global__ void __launch_bounds__(32)
ltp_sample_row_latency_kernel(const __grid_constant__ PlannerParams P, int64_t n, const double* __restrict__ q_0,
                              const double* __restrict__ v_0, const double* __restrict__ a_0, DeviceSolution S,
                              int horizon, int64_t stride, double* __restrict__ q, double* __restrict__ v,
                              double* __restrict__ a, double* __restrict__ j, uint8_t* row_ok) {
  __shared__ __align__(16) double s_tab[kMaxSeg][2];
  const int dof = P.dof;
  const int lane = threadIdx.x;
  const int64_t row = blockIdx.x;
  const int64_t p = row / dof;
  const int jt = (int)(row - p * dof);
  bool ok = false;
  if (S.reached[p]) {  // uniform over the warp
    const int len = S.traj_len[p];
    if (len > 0) {
      const JointLimits L = P.joint(jt);
      const int64_t at = (int64_t)jt * n + p;
      double t[7], v_drive_row;
      load_record(record_of(S, jt, n, p), t, v_drive_row);
      const int n_out = horizon > 0 ? horizon : (len < stride ? len : (int)stride);
      const int n_run = n_out > len ? n_out : len;
      RowSampler R;
      R.init(P.ts, L.j_max, t, S.dir[at], S.mod[at], q_0[at], v_0[at], a_0[at], v_drive_row, n_run);
      {
        // The piece table, built by the whole warp instead of piece after piece by one thread:
        // every lane takes one of the 26 places where the jerk or the update rule can change
        // (the candidates of RowSampler::next_break, plus sample 0), duplicates drop out, the
        // rank of a lane's place among the distinct ones is its piece number, the next larger
        // place ends its piece. Same table as SegTableT::build, ~1.5 us instead of ~9.
        int key = 0x7fffffff;
        if (lane < 7) key = R.s[lane];
        else if (lane == 7) key = 1;
        else if (lane == 8) key = R.s[6] + 1;
        else if (lane == 9) key = R.s[2] + 1;
        else if (lane == 10) key = R.s[3] - 1;
        else if (lane < 18) key = R.imp_idx[lane - 11];
        else if (lane < 25) key = R.imp_idx[lane - 18] + 1;
        else if (lane == 25) key = 0;
        const bool in_range = (lane == 25) || (lane < 25 && key > 0 && key < n_run);
        if (!in_range) key = 0x7fffffff;
        const unsigned same = __match_any_sync(0xffffffffu, key);
        const bool mine = in_range && (lane == __ffs(same) - 1);  // first lane holding this place
        int rank = 0, nxt = 0x7fffffff;
#pragma unroll 8
        for (int o = 0; o < 32; ++o) {
          const int k2 = __shfl_sync(0xffffffffu, key, o);
          const bool f2 = __shfl_sync(0xffffffffu, (int)mine, o) != 0;
          rank += (f2 && k2 < key) ? 1 : 0;
          nxt = (f2 && k2 > key && k2 < nxt) ? k2 : nxt;
        }
        if (mine && rank < kMaxSeg) {
          s_tab[rank][0] = R.jerk_at(key);
          s_tab[rank][1] = seg_pack(nxt, R.v_cruise(key) ? 1 : 0, R.a_zero(key) ? 0 : 1);
        }
        __syncwarp();
      }
      if (lane != 0) return;
      double av = R.a, vv = R.v, qv = R.q, q_last = 0.0;
      const double Ts = R.Ts, vcruise = R.vcruise;
      const int64_t base = (p * dof + jt) * stride;
      double* qo = q + base; double* vo = v + base; double* ao = a + base; double* jo = j + base;
      int i = 0;
      for (int m = 0; m < kMaxSeg && i < n_run; ++m) {
        const double jv = s_tab[m][0];
        int next;
        bool vc, live;
        seg_unpack(s_tab[m][1], next, vc, live);
        const int stop = next < n_run ? next : n_run;
        if (!live && vc) run_piece<false, true>(i, stop, n_out, len, Ts, jv, vcruise, av, vv, qv, q_last, qo, vo, ao, jo);
        else if (!live) run_piece<false, false>(i, stop, n_out, len, Ts, jv, vcruise, av, vv, qv, q_last, qo, vo, ao, jo);
        else if (vc) run_piece<true, true>(i, stop, n_out, len, Ts, jv, vcruise, av, vv, qv, q_last, qo, vo, ao, jo);
        else run_piece<true, false>(i, stop, n_out, len, Ts, jv, vcruise, av, vv, qv, q_last, qo, vo, ao, jo);
      }
      ok = !(q_last < L.q_min || q_last > L.q_max);  // cc:60
    }
  }
  if (lane == 0) row_ok[row] = (uint8_t)ok;
}

// Time-major variant: q[(sample * n + problem) * dof + joint] (a torch tensor of shape
// (samples, n, dof)). Lane l of CTA b owns row r = 32 b + l of the flattened (problem, joint)
// index, so the 32 lanes of a warp write 32 consecutive doubles -- one aligned 256-byte
// piece, two complete 128-byte lines -- per field and sample, and all rows advance through
// the output in lockstep: HBM sees four long sequential write streams instead of n*dof*4
// interleaved ones. Measured on B200 (tools/experiments/store_pattern_probe.cu,
// store_parallelism_probe.cu): ~6.0 TB/s for this pattern against ~4.3 TB/s for one
// 32-byte sector per row per store at a 16 KB row stride.
//
// More warps per row tile (splitting the sample axis for the stores, with a compute-only
// run-up per chunk) do not help: the pure-store kernel of this pattern tops out at
// 5.9-6.05 TB/s with 1, 2, 4 or 8 warps per tile, and this kernel reaches 96 % of that with
// one (tools/experiments/sampler_variants.cu).
// Alignment matters (measured, tools/sampler_scaling_probe.py): a sample plane is n*dof*8 bytes,
// so the warp's 256-byte piece stays line-aligned from sample to sample only if n*dof is a
// multiple of 32 -- 5.8 TB/s at n = 4096 (7 joints), 4.8 TB/s at n = 4104 (sector-aligned,
// pieces straddle lines), 3.1 TB/s at n = 4097 (partial sectors: the neighbouring warp completes
// them later and the memory system reads them back to merge), whatever the cache hint of the
// store. Callers pad the batch to a multiple of 32 problems (ltp_b200.h).
// success[] must have been initialised with reached[]; a row that ends outside its joint
// limits clears its problem's flag.
__device__ __forceinline__ void clear_flag(uint8_t* flags, int64_t p) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(flags + p);
  unsigned* word = reinterpret_cast<unsigned*>(addr & ~uintptr_t(3));
  const unsigned shift = (unsigned)(addr & 3u) * 8u;
  atomicAnd(word, ~(0xffu << shift));
}

__device__ __forceinline__ void store1(double* dst, double x) {
  asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(dst), "d"(x) : "memory");
}

// OffT: byte offsets inside one field are kept in 32 bits when the field is smaller than
// 4 GiB, which lets every store use the [uniform 64-bit base + 32-bit offset] address form.
template <typename OffT>
__global__ void __launch_bounds__(32)
ltp_sample_tm_kernel(const __grid_constant__ PlannerParams P, int64_t n, const double* __restrict__ q_0,
                     const double* __restrict__ v_0, const double* __restrict__ a_0, DeviceSolution S,
                     int horizon, int64_t capacity, double* __restrict__ q, double* __restrict__ v,
                     double* __restrict__ a, double* __restrict__ j, uint8_t* success,
                     const int* __restrict__ order) {
  __shared__ __align__(16) double s_tab[kMaxSeg][32][2];  // per-lane segment table, [entry][lane][word]
  const int dof = P.dof;
  const int lane = threadIdx.x;
  const SegTableT<64> T{&s_tab[0][lane][0]};
  const int64_t rows = n * dof;
  const int64_t r = (int64_t)blockIdx.x * 32 + lane;
  if (r >= rows) return;
  // Sorted-slot mode (order != nullptr, exact-length streaming): the lanes of a warp run until
  // the longest of their rows ends, and with problems in index order only 84 % of the
  // lane-iterations of random Franka problems are live. Here slot k of the OUTPUT holds problem
  // order[k] (longest first; built by the ltp_order_* kernels): rows of equal length share a
  // warp (> 99 % live) and the stores stay one aligned 256-byte piece per field and sample.
  // (Keeping the rows where they were and permuting only the threads was measured too: the
  // scattered 96-byte fragments cost more than half of the bandwidth.)
  const int64_t k = r / dof;
  const int jt = (int)(r - k * dof);
  const int64_t p = order ? (int64_t)order[k] : k;
  const int64_t at = (int64_t)jt * n + p;
  const bool planned = S.reached[p] != 0;
  const int len = planned ? S.traj_len[p] : 0;
  if (len <= 0) {
    if (planned) clear_flag(success, p);
    // fixed horizon: a problem without a plan holds its start position (ltp_b200.h)
    if (horizon > 0) {
      const double hold = q_0[at];
      const int64_t step = rows;
      for (int64_t i = 0, off = r; i < horizon; ++i, off += step) {
        store1(q + off, hold);
        store1(v + off, 0.0);
        store1(a + off, 0.0);
        store1(j + off, 0.0);
      }
    }
    return;
  }
  const JointLimits L = P.joint(jt);
  double t[7], v_drive_row;
  load_record(record_of(S, jt, n, p), t, v_drive_row);
  // samples stored: the fixed horizon, or the exact length clipped to the sample capacity
  const int n_out = horizon > 0 ? horizon : (len < capacity ? len : (int)capacity);
  const int n_run = n_out > len ? n_out : len;    // samples computed
  RowCursor C;
  {
    RowSampler R;
    R.init(P.ts, L.j_max, t, S.dir[at], S.mod[at], q_0[at], v_0[at], a_0[at], v_drive_row, n_run);
    T.build(R, n_run, LTP_TM_BUILD_EARLY_EXIT != 0);
    cursor_begin(C, R, T);
  }
  char* qb = reinterpret_cast<char*>(q);
  char* vb = reinterpret_cast<char*>(v);
  char* ab = reinterpret_cast<char*>(a);
  char* jb = reinterpret_cast<char*>(j);
  OffT off = (OffT)r * 8;
  const OffT step = (OffT)rows * 8;
  double jj, aa, vv, qq;
  int i = 0;
#pragma unroll 4
  for (; i < n_out; ++i) {
    C.step(T, i, jj, aa, vv, qq);
    store1(reinterpret_cast<double*>(qb + off), qq);
    store1(reinterpret_cast<double*>(vb + off), vv);
    store1(reinterpret_cast<double*>(ab + off), aa);
    store1(reinterpret_cast<double*>(jb + off), jj);
    off += step;
  }
  // cc:60. The position is constant from the last switching sample on (v is pinned to 0), so
  // the value after the last computed sample (index >= traj_len - 1) is q[traj_len - 1].
  double q_end = C.q;
  if (i < n_run) {
    // Clipped row: the check still refers to the complete trajectory. Jump to the end with
    // the per-piece closed form; only if that lands within 1e-9 of a limit (it is accurate
    // to ~1e-12) step through the remaining samples to get the reference's exact value.
    q_end = C.peek_position(T, i, n_run);
    const double band = 1e-9;
    if (!(fabs(q_end - L.q_min) > band && fabs(q_end - L.q_max) > band)) {
      for (; i < n_run; ++i) C.step(T, i, jj, aa, vv, qq);
      q_end = C.q;
    }
  }
  if (q_end < L.q_min || q_end > L.q_max) clear_flag(success, p);
}

// ------------------------------------------------------------------------------------
// Problems ordered by trajectory length, longest first, for the exact-length time-major
// sampler: a counting sort over kOrderBins buckets of 2^shift samples (three small kernels,
// everything stays on the device). The order inside a bucket is whatever the atomics produce;
// the sampled output does not depend on it.
// ------------------------------------------------------------------------------------
constexpr int kOrderBins = 1024;

__device__ __forceinline__ int order_bucket(const int32_t* traj_len, const uint8_t* reached, int64_t p, int shift) {
  const int len = reached[p] ? traj_len[p] : 0;
  int k = len > 0 ? (len >> shift) : 0;
  k = k < kOrderBins ? k : kOrderBins - 1;
  return kOrderBins - 1 - k;
}

__global__ void __launch_bounds__(256)
ltp_order_hist_kernel(int64_t n, const int32_t* __restrict__ traj_len, const uint8_t* __restrict__ reached, int shift,
                      int* bins) {
  __shared__ int h[kOrderBins];
  for (int i = threadIdx.x; i < kOrderBins; i += blockDim.x) h[i] = 0;
  __syncthreads();
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    atomicAdd(&h[order_bucket(traj_len, reached, p, shift)], 1);
  __syncthreads();
  for (int i = threadIdx.x; i < kOrderBins; i += blockDim.x)
    if (h[i]) atomicAdd(&bins[i], h[i]);
}

__global__ void __launch_bounds__(kOrderBins)
ltp_order_scan_kernel(int* bins) {  // counts -> first position of each bucket
  __shared__ int s[kOrderBins];
  const int t = threadIdx.x;
  const int mine = bins[t];
  s[t] = mine;
  __syncthreads();
  for (int d = 1; d < kOrderBins; d <<= 1) {
    const int add = t >= d ? s[t - d] : 0;
    __syncthreads();
    s[t] += add;
    __syncthreads();
  }
  bins[t] = s[t] - mine;
}

__global__ void __launch_bounds__(256)
ltp_order_scatter_kernel(int64_t n, const int32_t* __restrict__ traj_len, const uint8_t* __restrict__ reached, int shift,
                         int* cursors, int* __restrict__ order) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    order[atomicAdd(&cursors[order_bucket(traj_len, reached, p, shift)], 1)] = (int)p;
}

// ------------------------------------------------------------------------------------
// streaming driver: totals of one chunk, accumulated in device memory across chunks
// ------------------------------------------------------------------------------------
struct StreamTotals {  // mirrors ltp_stream_stats
  long long problems, reached, success, clipped, samples, max_len;
};

__global__ void __launch_bounds__(256)
ltp_chunk_totals_kernel(int64_t n, int dof, int horizon, int64_t capacity, const int32_t* __restrict__ traj_len,
                        const uint8_t* __restrict__ reached, const uint8_t* __restrict__ success,
                        StreamTotals* totals) {
  long long rch = 0, suc = 0, clip = 0, smp = 0, mx = 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int len = traj_len[p];
    const bool r = reached[p] != 0 && len > 0;
    rch += reached[p] != 0;
    suc += success[p] != 0;
    if (r) {
      const long long out = horizon > 0 ? horizon : (len < capacity ? len : capacity);
      clip += (horizon > 0 ? len > horizon : len > capacity);
      smp += out * dof;
      mx = len > mx ? len : mx;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    rch += __shfl_down_sync(0xffffffffu, rch, o);
    suc += __shfl_down_sync(0xffffffffu, suc, o);
    clip += __shfl_down_sync(0xffffffffu, clip, o);
    smp += __shfl_down_sync(0xffffffffu, smp, o);
    const long long other = __shfl_down_sync(0xffffffffu, mx, o);
    mx = other > mx ? other : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd((unsigned long long*)&totals->reached, (unsigned long long)rch);
    atomicAdd((unsigned long long*)&totals->success, (unsigned long long)suc);
    atomicAdd((unsigned long long*)&totals->clipped, (unsigned long long)clip);
    atomicAdd((unsigned long long*)&totals->samples, (unsigned long long)smp);
    atomicMax(&totals->max_len, mx);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&totals->problems, (unsigned long long)n);
}

// ------------------------------------------------------------------------------------
// receding-horizon replanning: the state reached `tick` samples into the current plans
// becomes the start state of the next solve. Time-major trajectories (samples, n, dof) ->
// joint-major [dof][n]. The forward-Euler samples may overshoot a limit by a few ulps (or by
// the discretisation error of a sample); with clamp != 0 the state is pulled back inside what
// checkInputs accepts (reference cc:68-77) so that the next plan is not rejected for it.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ltp_advance_kernel(const __grid_constant__ PlannerParams P, int64_t n, int tick, int clamp, int64_t capacity,
                   const int32_t* __restrict__ traj_len, const uint8_t* __restrict__ valid,
                   const double* __restrict__ q, const double* __restrict__ v, const double* __restrict__ a,
                   double* __restrict__ q_0, double* __restrict__ v_0, double* __restrict__ a_0) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int jt = blockIdx.y;
  if (p >= n) return;
  const int64_t at = (int64_t)jt * n + p;
  if (valid && !valid[p]) return;  // no plan: the environment keeps its state
  const int dof = P.dof;
  const JointLimits L = P.joint(jt);
  // exact-length trajectories end at traj_len - 1; from there on the state is the final one
  int at_sample = tick;
  if (traj_len) {
    const int len = traj_len[p];
    if (len <= 0) return;
    at_sample = tick < len ? tick : len - 1;
  }
  // a trajectory clipped by the sample capacity holds capacity samples, whatever its length
  if (at_sample > capacity - 1) at_sample = (int)(capacity - 1);
  const int64_t src = ((int64_t)at_sample * n + p) * dof + jt;
  double qq = q[src], vv = v[src], aa = a[src];
  if (clamp) {
    aa = fmin(fmax(aa, -L.a_max), L.a_max);
    vv = fmin(fmax(vv, -L.v_max), L.v_max);
    qq = fmin(fmax(qq, L.q_min), L.q_max);
    // |v + a|a|/(2 j)| <= v_max (cc:74): shrink the acceleration towards the largest one
    // from which the joint can still be stopped below v_max
    const double room = L.v_max - fabs(vv);
    if (fabs(vv + 0.5 * aa * fabs(aa) / L.j_max) > L.v_max && vv * aa > 0) {
      double lim_a = sqrt(2.0 * L.j_max * (room > 0 ? room : 0.0));
      lim_a = lim_a * (1.0 - 4.0 * kDblEps);
      aa = aa > 0 ? fmin(aa, lim_a) : fmax(aa, -lim_a);
    }
  }
  q_0[at] = qq;
  v_0[at] = vv;
  a_0[at] = aa;
}


// ------------------------------------------------------------------------------------
// Layout bridge for callers that keep their state problem-major, x[problem][joint] -- what a
// vectorised environment holds (n_env x dof), while the kernels want joint-major x[joint][problem]
// (one joint of 32 consecutive problems per warp, every access a full 256-byte piece).
// dst[c][r] = src[r][c] for a rows x cols matrix of doubles, through a 32 x 33 shared tile so
// that both the reads and the writes are coalesced.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ltp_transpose_kernel(int64_t rows, int64_t cols, int64_t dst_pitch, const double* __restrict__ src,
                     double* __restrict__ dst) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8 threads
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int64_t r = r0 + ty + k, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + k][tx] = src[r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int64_t c = c0 + ty + k, r = r0 + tx;
    if (r < rows && c < cols) dst[c * dst_pitch + r] = tile[tx][ty + k];
  }
}

}  // namespace

// ======================================================================================
// C ABI
// ======================================================================================
constexpr int kProfRing = 64;  // timed launches kept between two ltp_profile_read calls

struct ltp_planner {
  int device;
  PlannerParams params;
  std::atomic<int64_t> launches;
  // scratch for the host-buffer entry points (grown on demand)
  void* d_scratch;
  size_t d_scratch_bytes;
  // pinned, device-mapped host staging of the small-batch host calls (ltp_plan_host with a
  // handful of problems): the kernels read the inputs from it and write the rows into it
  void* h_stage;
  int stream_sorted;  // ltp_set_stream_sorted
  int* d_bins;        // ltp_sample_batch_sorted: kOrderBins counters
  cudaStream_t stream;  // internal stream of the host entry points
  // work list of the two-kernel solve: [0] = count, [1..] = problem indices
  void* d_work;  // SolveScratch of ltp_solve_batch
  int64_t d_work_capacity;
  int d_work_dof;
  int solve_mode;  // LTP_SOLVE_AUTO / LTP_SOLVE_GENERIC
  int sm_count;
  // optional per-kernel timing (ltp_set_profiling): CUDA events recorded on the launching
  // stream directly around the hot kernels, read back by ltp_profile_read
  bool profiling;
  struct Timed {
    cudaEvent_t ev[kProfRing][2];
    int head, pending;
    double ms_sum;
    int64_t count;
  } timed[LTP_PROFILE_KERNELS];
  // chunk pipeline of ltp_solve_host: two slots, each with its own stream, device
  // buffers and work list, so that the copy-in of chunk k+1 and the copy-out of chunk k
  // run on the two DMA engines at the same time as the kernels of the chunk between them
  cudaStream_t pipe_stream[2];
  void* pipe_buf[2];
  size_t pipe_bytes[2];
  // ltp_plan_stream: ring of trajectory slots (grown on demand) and the device totals
  void* ring_buf[2];
  size_t ring_bytes[2];
  StreamTotals* d_totals;
};

namespace {

thread_local char g_cuda_err[256] = "";

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_cuda_err, sizeof g_cuda_err, "%s: %s", what, cudaGetErrorString(e));
  return LTP_ERR_CUDA;
}

#define LTP_CUDA(call)                                   \
  do {                                                   \
    cudaError_t e_ = (call);                             \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
  } while (0)

struct DeviceGuard {
  int prev;
  bool ok;
  explicit DeviceGuard(int dev) : prev(-1), ok(false) {
    if (cudaGetDevice(&prev) != cudaSuccess) return;
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) return;
    ok = true;
  }
  ~DeviceGuard() {
    if (ok && prev >= 0) cudaSetDevice(prev);
  }
};

int fill_limits(PlannerParams& P, const double* q_min, const double* q_max, const double* v_max,
                const double* a_max, const double* j_max) {
  if (!q_min || !q_max || !v_max || !a_max || !j_max) return LTP_ERR_ARG;
  for (int i = 0; i < P.dof; ++i) {
    JointLimits L{};
    L.q_min = q_min[i];
    L.q_max = q_max[i];
    L.v_max = v_max[i];
    L.a_max = a_max[i];
    L.j_max = j_max[i];
    derive_limits(L);
    P.set_joint(i, L);
  }
  return LTP_OK;
}

DeviceSolution to_dev(const ltp_solution* s) {
  DeviceSolution d;
  d.t_scaled = s->t_scaled; d.dir = s->dir; d.v_drive = s->v_drive; d.mod = s->mod;
  d.slowest = s->slowest; d.traj_len = s->traj_len; d.reached = s->reached; d.t_opt = s->t_opt;
  d.opt_case = s->opt_case; d.ts_case = s->ts_case; d.final_case = s->final_case;
  return d;
}

bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }

int ensure_scratch(ltp_planner* p, size_t bytes) {
  if (bytes <= p->d_scratch_bytes) return LTP_OK;
  if (p->d_scratch) LTP_CUDA(cudaFree(p->d_scratch));
  p->d_scratch = nullptr;
  p->d_scratch_bytes = 0;
  LTP_CUDA(cudaMalloc(&p->d_scratch, bytes));
  p->d_scratch_bytes = bytes;
  return LTP_OK;
}

size_t up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// fold the finished event pairs of one kernel slot into its sum (synchronises on them)
void prof_drain(ltp_planner::Timed& t) {
  for (int k = 0; k < t.pending; ++k) {
    const int slot = (t.head - t.pending + k + 2 * kProfRing) % kProfRing;
    float ms = 0.f;
    if (cudaEventSynchronize(t.ev[slot][1]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, t.ev[slot][0], t.ev[slot][1]) == cudaSuccess) {
      t.ms_sum += ms;
      t.count++;
    }
  }
  t.pending = 0;
}

// RAII bracket around one kernel launch: records an event pair on `st` when profiling is on
struct ProfScope {
  ltp_planner::Timed* t;
  cudaStream_t st;
  int slot;
  ProfScope(ltp_planner* p, int which, cudaStream_t s) : t(nullptr), st(s), slot(0) {
    if (!p->profiling) return;
    t = &p->timed[which];
    if (t->pending == kProfRing) prof_drain(*t);
    slot = t->head;
    if (!t->ev[slot][0]) {
      cudaEventCreate(&t->ev[slot][0]);
      cudaEventCreate(&t->ev[slot][1]);
    }
    cudaEventRecord(t->ev[slot][0], st);
  }
  ~ProfScope() {
    if (!t) return;
    cudaEventRecord(t->ev[slot][1], st);
    t->head = (t->head + 1) % kProfRing;
    t->pending++;
  }
};

// carve a solution (all fields) out of a scratch block; returns bytes used
size_t carve_solution(unsigned char* base, int dof, int64_t n, ltp_solution* s) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    unsigned char* r = base ? base + off : nullptr;
    off += up(bytes, 256);
    return r;
  };
  const size_t dn = (size_t)dof * (size_t)n;
  s->t_scaled = (double*)take(8 * dn * 8);  // 64-byte records: t[0..6], v_drive
  s->t_opt = (double*)take(7 * dn * 8);
  s->dir = (double*)take(dn * 8);
  s->v_drive = (double*)take(dn * 8);
  s->mod = (uint8_t*)take(dn);
  s->opt_case = (uint8_t*)take(dn);
  s->ts_case = (uint8_t*)take(dn);
  s->final_case = (uint8_t*)take(dn);
  s->slowest = (int32_t*)take((size_t)n * 4);
  s->traj_len = (int32_t*)take((size_t)n * 4);
  s->reached = (uint8_t*)take((size_t)n);
  return off;
}

}  // namespace

extern "C" {

const char* ltp_status_string(int status) {
  switch (status) {
    case LTP_OK: return "ok";
    case LTP_ERR_ARG: return "invalid argument";
    case LTP_ERR_CUDA: return "CUDA error";
    case LTP_ERR_CAPACITY: return "row capacity too small";
    default: return "unknown status";
  }
}

const char* ltp_last_cuda_error(void) { return g_cuda_err; }

int ltp_create(ltp_planner** out, int device, int dof, double t_sample, const double* q_min,
               const double* q_max, const double* v_max, const double* a_max, const double* j_max) {
  if (!out) return LTP_ERR_ARG;
  *out = nullptr;
  if (dof < 0 || dof > LTP_MAX_DOF) return LTP_ERR_ARG;
  int count = 0;
  LTP_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return cuda_fail(cudaErrorInvalidDevice, "ltp_create(device)");
  ltp_planner* p = new (std::nothrow) ltp_planner();
  if (!p) return LTP_ERR_ARG;
  p->device = device;
  p->params.dof = dof;
  p->params.ts = t_sample;
  p->params.r_ts = 1.0 / t_sample;
  p->launches = 0;
  p->d_scratch = nullptr;
  p->d_scratch_bytes = 0;
  p->h_stage = nullptr;
  p->stream_sorted = 0;
  p->d_bins = nullptr;
  p->stream = nullptr;
  p->d_work = nullptr;
  p->d_work_capacity = 0;
  p->d_work_dof = 0;
  p->solve_mode = LTP_SOLVE_AUTO;
  p->sm_count = 148;
  p->profiling = false;
  std::memset(p->timed, 0, sizeof p->timed);
  for (int i = 0; i < 2; ++i) {
    p->pipe_stream[i] = nullptr;
    p->pipe_buf[i] = nullptr;
    p->pipe_bytes[i] = 0;
    p->ring_buf[i] = nullptr;
    p->ring_bytes[i] = 0;
  }
  p->d_totals = nullptr;
  std::memset(p->params.hot, 0, sizeof p->params.hot);
  std::memset(p->params.wide, 0, sizeof p->params.wide);
  if (dof > 0) {
    int rc = fill_limits(p->params, q_min, q_max, v_max, a_max, j_max);
    if (rc != LTP_OK) { delete p; return rc; }
  }
  {
    DeviceGuard g(device);
    if (!g.ok) { delete p; return cuda_fail(cudaGetLastError(), "cudaSetDevice"); }
    cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaStreamCreate"); }
    cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, device);
  }
  *out = p;
  return LTP_OK;
}

int ltp_set_limits(ltp_planner* p, const double* q_min, const double* q_max, const double* v_max,
                   const double* a_max, const double* j_max) {
  if (!p) return LTP_ERR_ARG;
  return fill_limits(p->params, q_min, q_max, v_max, a_max, j_max);
}

int ltp_set_sample_time(ltp_planner* p, double t_sample) {
  if (!p) return LTP_ERR_ARG;
  p->params.ts = t_sample;
  p->params.r_ts = 1.0 / t_sample;
  return LTP_OK;
}

int ltp_set_dof(ltp_planner* p, int dof) {
  if (!p || dof < 0 || dof > LTP_MAX_DOF) return LTP_ERR_ARG;
  p->params.dof = dof;
  return LTP_OK;
}

int ltp_set_solve_mode(ltp_planner* p, int mode) {
  if (!p || (mode != LTP_SOLVE_AUTO && mode != LTP_SOLVE_GENERIC)) return LTP_ERR_ARG;
  p->solve_mode = mode;
  return LTP_OK;
}

int ltp_set_stream_sorted(ltp_planner* p, int on) {
  if (!p) return LTP_ERR_ARG;
  p->stream_sorted = on ? 1 : 0;
  return LTP_OK;
}

int ltp_set_profiling(ltp_planner* p, int on) {
  if (!p) return LTP_ERR_ARG;
  p->profiling = on != 0;
  return LTP_OK;
}

int ltp_profile_read(ltp_planner* p, int kernel, double* ms_sum, int64_t* launches, int reset) {
  if (!p || kernel < 0 || kernel >= LTP_PROFILE_KERNELS) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  ltp_planner::Timed& t = p->timed[kernel];
  prof_drain(t);
  if (ms_sum) *ms_sum = t.ms_sum;
  if (launches) *launches = t.count;
  if (reset) {
    t.ms_sum = 0.0;
    t.count = 0;
  }
  return LTP_OK;
}

int ltp_get_dof(const ltp_planner* p) { return p ? p->params.dof : LTP_ERR_ARG; }
int ltp_get_device(const ltp_planner* p) { return p ? p->device : LTP_ERR_ARG; }
int64_t ltp_launch_count(const ltp_planner* p) { return p ? p->launches.load() : 0; }

void ltp_destroy(ltp_planner* p) {
  if (!p) return;
  {
    DeviceGuard g(p->device);
    if (p->d_scratch) cudaFree(p->d_scratch);
    if (p->h_stage) cudaFreeHost(p->h_stage);
    if (p->d_bins) cudaFree(p->d_bins);
    if (p->d_work) cudaFree(p->d_work);
    if (p->d_totals) cudaFree(p->d_totals);
    if (p->stream) cudaStreamDestroy(p->stream);
    for (auto& t : p->timed)
      for (auto& e : t.ev)
        for (auto& x : e)
          if (x) cudaEventDestroy(x);
    for (int i = 0; i < 2; ++i) {
      if (p->pipe_buf[i]) cudaFree(p->pipe_buf[i]);
      if (p->ring_buf[i]) cudaFree(p->ring_buf[i]);
      if (p->pipe_stream[i]) cudaStreamDestroy(p->pipe_stream[i]);
    }
  }
  delete p;
}

int ltp_opt_braking_batch(ltp_planner* p, int64_t n, const double* v_0, const double* a_0,
                          double* q_stop, double* t_rel, double* dir, void* stream) {
  if (!p || n < 0 || p->params.dof < 1) return LTP_ERR_ARG;
  if (n == 0) return LTP_OK;
  if (!v_0 || !a_0 || !q_stop || !t_rel || !dir) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  dim3 block(128), grid((unsigned)((n + 127) / 128), p->params.dof);
  ltp_opt_braking_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(p->params, -1, n, v_0, a_0, q_stop, t_rel, dir);
  p->launches++;
  LTP_CUDA(cudaGetLastError());
  return LTP_OK;
}

int ltp_opt_switch_times_batch(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0,
                               const double* v_0, const double* a_0, const double* v_drive, double* t,
                               double* dir, uint8_t* mod, uint8_t* kase, uint8_t* ok, void* stream) {
  if (!p || n < 0 || p->params.dof < 1) return LTP_ERR_ARG;
  if (n == 0) return LTP_OK;
  if (!q_goal || !q_0 || !v_0 || !a_0 || !v_drive || !t || !dir || !mod || !ok) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  dim3 block(128), grid((unsigned)((n + 127) / 128), p->params.dof);
  ltp_opt_switch_times_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
      p->params, -1, n, q_goal, q_0, v_0, a_0, v_drive, t, dir, mod, kase, ok);
  p->launches++;
  LTP_CUDA(cudaGetLastError());
  return LTP_OK;
}

int ltp_time_scaling_batch(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0,
                           const double* v_0, const double* a_0, const double* dir,
                           const double* t_required, double* t, double* v_drive, uint8_t* mod,
                           uint8_t* ts_case, uint8_t* final_case, uint8_t* ok, void* stream) {
  if (!p || n < 0 || p->params.dof < 1) return LTP_ERR_ARG;
  if (n == 0) return LTP_OK;
  if (!q_goal || !q_0 || !v_0 || !a_0 || !dir || !t_required || !t || !v_drive || !mod || !ok)
    return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  dim3 block(128), grid((unsigned)((n + 127) / 128), p->params.dof);
  ltp_time_scaling_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
      p->params, -1, n, q_goal, q_0, v_0, a_0, dir, t_required, t, v_drive, mod, ts_case, final_case, ok);
  p->launches++;
  LTP_CUDA(cudaGetLastError());
  return LTP_OK;
}

// launches of stages 1-3 on `st`. work: device buffer of n + 1 ints ([0] = count), only
// touched in LTP_SOLVE_AUTO mode.
// launches of stages 1-3 on `st`. work: device buffer of n + 1 ints ([0] = count), only
// touched in LTP_SOLVE_AUTO mode.
static int solve_launch(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                        const double* a_0, const ltp_solution* sol, void* scratch, cudaStream_t st) {
  const int dof = p->params.dof;
  const dim3 block(kTile, dof);
  const unsigned tiles = (unsigned)((n + kTile - 1) / kTile);
  const size_t smem = solve_smem_bytes(dof);
  const DeviceSolution ds = to_dev(sol);
  const SolveScratch X = carve_scratch(scratch, dof, n);
#define LTP_DISPATCH_W(KERNEL, GRID, ...)                                                   \
  do {                                                                                      \
    if (dof <= 1) KERNEL<1><<<GRID, block, smem, st>>>(__VA_ARGS__);                        \
    else if (dof <= 2) KERNEL<2><<<GRID, block, smem, st>>>(__VA_ARGS__);                   \
    else if (dof <= 4) KERNEL<4><<<GRID, block, smem, st>>>(__VA_ARGS__);                   \
    else if (dof <= 8) KERNEL<8><<<GRID, block, smem, st>>>(__VA_ARGS__);                   \
    else if (dof <= 16) KERNEL<16><<<GRID, block, smem, st>>>(__VA_ARGS__);                 \
    else KERNEL<32><<<GRID, block, smem, st>>>(__VA_ARGS__);                                \
    p->launches++;                                                                          \
  } while (0)
  // the closed-form kernel is compiled for the exact CTA size of the common arm sizes so that
  // its register budget is the largest that still fits the intended number of CTAs per SM
#define LTP_DISPATCH_FAST(GRID, ...)                                                        \
  do {                                                                                      \
    if (dof == 6) ltp_solve_fast_kernel<6, LTP_FAST_EXACT><<<GRID, block, smem, st>>>(__VA_ARGS__);         \
    else if (dof == 7) ltp_solve_fast_kernel<7, LTP_FAST_EXACT><<<GRID, block, smem, st>>>(__VA_ARGS__);    \
    else if (dof == 12) ltp_solve_fast_kernel<12, LTP_FAST_EXACT><<<GRID, block, smem, st>>>(__VA_ARGS__);  \
    else { LTP_DISPATCH_W(ltp_solve_fast_kernel, GRID, __VA_ARGS__); break; }               \
    p->launches++;                                                                          \
  } while (0)
  // a single tile is one CTA whichever way it is run: the every-branch kernel alone (one launch
  // instead of a memset and three) gives the same results sooner
  if (p->solve_mode == LTP_SOLVE_GENERIC || n <= kTile) {
    ProfScope ps(p, LTP_PROFILE_SOLVE_GENERIC, st);
    LTP_DISPATCH_W(ltp_solve_generic_kernel, tiles, p->params, n, q_goal, q_0, v_0, a_0, ds,
                   (const int*)nullptr, (const int*)nullptr);
  } else {
    LTP_CUDA(cudaMemsetAsync(X.counters, 0, kCounterInts * sizeof(int), st));
    // large batches hand unsettled joints on one by one (item mode); a small batch is a handful of
    // CTAs whichever way it is run and keeps the three-launch sequence
    const int items = n >= kItemModeMin ? 1 : 0;
    {
      ProfScope ps(p, LTP_PROFILE_SOLVE_FAST, st);
      LTP_DISPATCH_FAST(tiles, p->params, n, q_goal, q_0, v_0, a_0, ds, X, items);
    }
    // the queue and the work list are drained by fixed-size grid-stride launches: their
    // lengths never leave the device
    if (dof > 1) {
      ProfScope ps(p, LTP_PROFILE_SOLVE_ATTEMPT2, st);
      const int64_t want = ((int64_t)(dof - 1) * n + 255) / 256;
      const int64_t cap = (int64_t)p->sm_count * 16;
      ltp_solve_attempt2_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(p->params, n, ds, X, items);
      p->launches++;
    }
    if (items) {
      // list lengths never leave the device: fixed-size grid-stride launches
      ProfScope ps(p, LTP_PROFILE_SOLVE_ITEMS, st);
      const unsigned gi = (unsigned)(p->sm_count * 3);
      ltp_solve_tail_kernel<<<gi, 128, 0, st>>>(p->params, n, q_goal, q_0, v_0, a_0, ds, X);
      const unsigned gp = tiles < (unsigned)(p->sm_count * 2) ? tiles : (unsigned)(p->sm_count * 2);
      LTP_DISPATCH_W(ltp_solve_pending_kernel, gp, p->params, n, q_goal, q_0, v_0, a_0, ds, X);
      ltp_solve_search_kernel<<<gi, 128, 0, st>>>(p->params, n, q_goal, q_0, v_0, a_0, ds, X);
      p->launches += 2;
    }
    const unsigned g2 = tiles < (unsigned)(p->sm_count * 4) ? tiles : (unsigned)(p->sm_count * 4);
    ProfScope ps(p, LTP_PROFILE_SOLVE_GENERIC, st);
    LTP_DISPATCH_W(ltp_solve_generic_kernel, g2, p->params, n, q_goal, q_0, v_0, a_0, ds,
                   (const int*)X.work_list, (const int*)(X.counters + kCntWork));
  }
#undef LTP_DISPATCH_FAST
#undef LTP_DISPATCH_W
  LTP_CUDA(cudaGetLastError());
  return LTP_OK;
}

static int reserve_solve_scratch(ltp_planner* p, int64_t n) {
  if (p->d_work_capacity >= n && p->d_work_dof >= p->params.dof) return LTP_OK;
  if (p->d_work) LTP_CUDA(cudaFree(p->d_work));
  p->d_work = nullptr;
  p->d_work_capacity = 0;
  LTP_CUDA(cudaMalloc(&p->d_work, solve_scratch_bytes(p->params.dof, n)));
  p->d_work_capacity = n;
  p->d_work_dof = p->params.dof;
  return LTP_OK;
}

int ltp_reserve(ltp_planner* p, int64_t n) {
  if (!p || n < 0 || n > 0x7fffffff || p->params.dof < 1) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  return reserve_solve_scratch(p, n);
}

int ltp_solve_batch(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0,
                    const double* v_0, const double* a_0, const ltp_solution* sol, void* stream) {
  if (!p || n < 0 || p->params.dof < 1) return LTP_ERR_ARG;
  if (n == 0) return LTP_OK;
  if (!q_goal || !q_0 || !v_0 || !a_0 || !sol) return LTP_ERR_ARG;
  if (!sol->t_scaled || !sol->dir || !sol->mod || !sol->slowest || !sol->traj_len || !sol->reached)
    return LTP_ERR_ARG;
  if (!aligned32(sol->t_scaled)) return LTP_ERR_ARG;  // the records are written with 256-bit stores
  if (n > 0x7fffffff) return LTP_ERR_ARG;  // problem indices travel as int32 in the work list
  DeviceGuard g(p->device);
  if (p->solve_mode != LTP_SOLVE_GENERIC && n > kTile) {
    const int rc = reserve_solve_scratch(p, n);  // allocates only when n grows (ltp_reserve avoids that)
    if (rc != LTP_OK) return rc;
  }
  return solve_launch(p, n, q_goal, q_0, v_0, a_0, sol, p->d_work, (cudaStream_t)stream);
}

static size_t order_scratch_bytes(int64_t n) { return (size_t)(kOrderBins + n) * sizeof(int); }

// order_scratch: kOrderBins + n ints on the device -> sorted-slot output, the order is left in
// order_scratch + kOrderBins; nullptr: slot = problem index
static int sample_tm_launch2(ltp_planner* p, int64_t n, const double* q_0, const double* v_0, const double* a_0,
                             const ltp_solution* sol, int32_t horizon, int64_t stride, double* q, double* v,
                             double* a, double* j, uint8_t* success, int* bins, int* ord, cudaStream_t st);

static int sample_tm_launch(ltp_planner* p, int64_t n, const double* q_0, const double* v_0, const double* a_0,
                            const ltp_solution* sol, int32_t horizon, int64_t stride, double* q, double* v,
                            double* a, double* j, uint8_t* success, int* order_scratch, cudaStream_t st) {
  return sample_tm_launch2(p, n, q_0, v_0, a_0, sol, horizon, stride, q, v, a, j, success, order_scratch,
                           order_scratch ? order_scratch + kOrderBins : nullptr, st);
}

// bins (kOrderBins ints) + ord (n ints), both on the device -> sorted-slot output; nullptr: slot = problem
static int sample_tm_launch2(ltp_planner* p, int64_t n, const double* q_0, const double* v_0, const double* a_0,
                             const ltp_solution* sol, int32_t horizon, int64_t stride, double* q, double* v,
                             double* a, double* j, uint8_t* success, int* bins, int* ord, cudaStream_t st) {
  const int dof = p->params.dof;
  LTP_CUDA(cudaMemcpyAsync(success, sol->reached, (size_t)n, cudaMemcpyDeviceToDevice, st));
  const int* order = nullptr;
  if (bins && ord) {
    int shift = 0;
    while ((stride >> shift) >= kOrderBins) ++shift;
    const unsigned g = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    LTP_CUDA(cudaMemsetAsync(bins, 0, kOrderBins * sizeof(int), st));
    ltp_order_hist_kernel<<<g, 256, 0, st>>>(n, sol->traj_len, sol->reached, shift, bins);
    ltp_order_scan_kernel<<<1, kOrderBins, 0, st>>>(bins);
    ltp_order_scatter_kernel<<<g, 256, 0, st>>>(n, sol->traj_len, sol->reached, shift, bins, ord);
    p->launches += 3;
    order = ord;
  }
  const int64_t rows = n * dof;
  const int64_t samples = horizon > 0 ? horizon : stride;
  const unsigned grid = (unsigned)((rows + 31) / 32);
  ProfScope ps(p, LTP_PROFILE_SAMPLE_TIME_MAJOR, st);
  if ((double)rows * 8.0 * (double)(samples + 1) < 4294967296.0)
    ltp_sample_tm_kernel<uint32_t><<<grid, 32, 0, st>>>(p->params, n, q_0, v_0, a_0, to_dev(sol), horizon, stride,
                                                           q, v, a, j, success, order);
  else
    ltp_sample_tm_kernel<uint64_t><<<grid, 32, 0, st>>>(p->params, n, q_0, v_0, a_0, to_dev(sol), horizon, stride,
                                                           q, v, a, j, success, order);
  p->launches++;
  LTP_CUDA(cudaGetLastError());
  return LTP_OK;
}

int ltp_sample_batch_sorted(ltp_planner* p, int64_t n, const double* q_0, const double* v_0, const double* a_0,
                            const ltp_solution* sol, int64_t capacity, double* q, double* v, double* a, double* j,
                            uint8_t* success, int32_t* order, void* stream) {
  if (!p || n < 0 || p->params.dof < 1 || capacity < 1) return LTP_ERR_ARG;
  if (n == 0) return LTP_OK;
  if (!q_0 || !v_0 || !a_0 || !sol || !q || !v || !a || !j || !success || !order) return LTP_ERR_ARG;
  if (!sol->t_scaled || !sol->dir || !sol->mod || !sol->traj_len || !sol->reached || !aligned32(sol->t_scaled))
    return LTP_ERR_ARG;
  if (n > 0x7fffffff) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  if (!p->d_bins) LTP_CUDA(cudaMalloc(&p->d_bins, kOrderBins * sizeof(int)));
  return sample_tm_launch2(p, n, q_0, v_0, a_0, sol, 0, capacity, q, v, a, j, success, p->d_bins, order,
                           (cudaStream_t)stream);
}

int ltp_sample_batch(ltp_planner* p, int64_t n, const double* q_0, const double* v_0, const double* a_0,
                     const ltp_solution* sol, int32_t horizon, int32_t layout, int64_t stride, double* q,
                     double* v, double* a, double* j, uint8_t* success, void* stream) {
  if (layout != LTP_LAYOUT_ROWS && layout != LTP_LAYOUT_TIME_MAJOR) return LTP_ERR_ARG;
  if (!p || n < 0 || p->params.dof < 1 || horizon < 0) return LTP_ERR_ARG;
  if (n == 0) return LTP_OK;
  if (!q_0 || !v_0 || !a_0 || !sol || !q || !v || !a || !j || !success || stride < 1 ||
      (horizon > 0 && stride < horizon))
    return LTP_ERR_ARG;
  if (!sol->t_scaled || !sol->dir || !sol->mod || !sol->traj_len || !sol->reached || !aligned32(sol->t_scaled))
    return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  const int dof = p->params.dof;
  if (layout == LTP_LAYOUT_TIME_MAJOR)
    return sample_tm_launch(p, n, q_0, v_0, a_0, sol, horizon, stride, q, v, a, j, success, nullptr,
                            (cudaStream_t)stream);
  const int ppb = 32 / dof > 0 ? 32 / dof : 1;  // whole problems per one-warp CTA (dof <= 32)
  const unsigned grid = (unsigned)((n + ppb - 1) / ppb);
  const bool vec = (stride % 4 == 0) && aligned32(q) && aligned32(v) && aligned32(a) && aligned32(j);
  ProfScope ps(p, LTP_PROFILE_SAMPLE_ROWS, (cudaStream_t)stream);
  if (vec)
    ltp_sample_kernel<true><<<grid, 32, 0, (cudaStream_t)stream>>>(p->params, n, ppb, q_0, v_0, a_0, to_dev(sol),
                                                                    horizon, stride, q, v, a, j, success);
  else
    ltp_sample_kernel<false><<<grid, 32, 0, (cudaStream_t)stream>>>(p->params, n, ppb, q_0, v_0, a_0, to_dev(sol),
                                                                     horizon, stride, q, v, a, j, success);
  p->launches++;
  LTP_CUDA(cudaGetLastError());
  return LTP_OK;
}

// ---- host-buffer entry points ---------------------------------------------------------

// Problems per chunk of the host-buffer pipeline: 2^16 problems are 15 MB in and 34 MB out
// for 7 joints -- long enough for full-rate DMA, short enough that the pipeline fills and
// drains in a few percent of a 2^20-problem call.
static const int64_t kHostChunk = 1 << 16;

int ltp_solve_host(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0,
                   const double* v_0, const double* a_0, const ltp_solution* hs) {
  if (!p || n < 0 || !q_goal || !q_0 || !v_0 || !a_0 || !hs || p->params.dof < 1) return LTP_ERR_ARG;
  // output mask: any field may be NULL and is then not copied back (the transfer out is what
  // bounds this call); traj_len and reached are always delivered
  if (!hs->traj_len || !hs->reached) return LTP_ERR_ARG;
  if (n == 0) return LTP_OK;
  if (n > 0x7fffffff) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  const int dof = p->params.dof;
  // Chunks of problems go through two pipeline slots (own stream, own device buffers, own
  // work list). Within a slot everything is stream-ordered: copy-in, the solve kernels,
  // copy-out, then the slot's next chunk. The two slots overlap each other, so the
  // host->device and device->host DMA engines and the SMs are busy at the same time and the
  // call approaches the slower PCIe direction instead of the sum of the three stages.
  // A joint-major host array x[row * n + problem] restricted to a chunk is `rows` pieces of
  // c values at a pitch of n values: one 2-D copy per array.
  const int64_t c_max = n < kHostChunk ? n : kHostChunk;
  const int slots = n > c_max ? 2 : 1;
  ltp_solution ds[2];
  double* d_in[2][4];
  void* d_work[2];
  const size_t in_bytes = up((size_t)dof * (size_t)c_max * 8, 256);
  const size_t sol_bytes = carve_solution(nullptr, dof, c_max, &ds[0]);
  const size_t work_bytes = up(solve_scratch_bytes(dof, c_max), 256);
  const size_t need = 4 * in_bytes + sol_bytes + work_bytes;
  for (int s = 0; s < slots; ++s) {
    if (!p->pipe_stream[s]) LTP_CUDA(cudaStreamCreateWithFlags(&p->pipe_stream[s], cudaStreamNonBlocking));
    if (p->pipe_bytes[s] < need) {
      if (p->pipe_buf[s]) LTP_CUDA(cudaFree(p->pipe_buf[s]));
      p->pipe_buf[s] = nullptr;
      p->pipe_bytes[s] = 0;
      LTP_CUDA(cudaMalloc(&p->pipe_buf[s], need));
      p->pipe_bytes[s] = need;
    }
    unsigned char* base = (unsigned char*)p->pipe_buf[s];
    for (int i = 0; i < 4; ++i) d_in[s][i] = (double*)(base + i * in_bytes);
    carve_solution(base + 4 * in_bytes, dof, c_max, &ds[s]);
    d_work[s] = base + 4 * in_bytes + sol_bytes;
    if (!hs->v_drive) ds[s].v_drive = nullptr;  // slot 7 of the records holds it anyway
    if (!hs->t_opt) ds[s].t_opt = nullptr;
    if (!hs->opt_case) ds[s].opt_case = nullptr;
    if (!hs->ts_case) ds[s].ts_case = nullptr;
    if (!hs->final_case) ds[s].final_case = nullptr;
  }
  const double* h_in[4] = {q_goal, q_0, v_0, a_0};
  int64_t k = 0;
  for (int64_t p0 = 0; p0 < n; p0 += c_max, ++k) {
    const int s = (int)(k % slots);
    const int64_t c = (n - p0) < c_max ? (n - p0) : c_max;
    cudaStream_t st = p->pipe_stream[s];
    const ltp_solution& d = ds[s];
    for (int i = 0; i < 4; ++i)
      LTP_CUDA(cudaMemcpy2DAsync(d_in[s][i], (size_t)c * 8, h_in[i] + p0, (size_t)n * 8, (size_t)c * 8, dof,
                                 cudaMemcpyHostToDevice, st));
    int rc = solve_launch(p, c, d_in[s][0], d_in[s][1], d_in[s][2], d_in[s][3], &d, d_work[s], st);
    if (rc != LTP_OK) return rc;
#define LTP_OUT2D(FIELD, ROWS, ELEM)                                                                   \
  LTP_CUDA(cudaMemcpy2DAsync(hs->FIELD + p0, (size_t)n * (ELEM), d.FIELD, (size_t)c * (ELEM),           \
                             (size_t)c * (ELEM), (ROWS), cudaMemcpyDeviceToHost, st))
    if (hs->t_scaled)  // [dof][n] records of 64 bytes: dof pieces of c records at a pitch of n records
      LTP_CUDA(cudaMemcpy2DAsync(hs->t_scaled + p0 * 8, (size_t)n * 64, d.t_scaled, (size_t)c * 64, (size_t)c * 64, dof,
                                 cudaMemcpyDeviceToHost, st));
    if (hs->dir) LTP_OUT2D(dir, dof, 8);
    if (hs->v_drive) LTP_OUT2D(v_drive, dof, 8);
    if (hs->mod) LTP_OUT2D(mod, dof, 1);
    if (hs->slowest) LTP_OUT2D(slowest, 1, 4);
    LTP_OUT2D(traj_len, 1, 4);
    LTP_OUT2D(reached, 1, 1);
    if (hs->t_opt) LTP_OUT2D(t_opt, 7 * dof, 8);
    if (hs->opt_case) LTP_OUT2D(opt_case, dof, 1);
    if (hs->ts_case) LTP_OUT2D(ts_case, dof, 1);
    if (hs->final_case) LTP_OUT2D(final_case, dof, 1);
#undef LTP_OUT2D
  }
  for (int s = 0; s < slots; ++s) LTP_CUDA(cudaStreamSynchronize(p->pipe_stream[s]));
  return LTP_OK;
}

static int grow(void** buf, size_t* have, size_t need) {
  if (*have >= need) return LTP_OK;
  if (*buf) LTP_CUDA(cudaFree(*buf));
  *buf = nullptr;
  *have = 0;
  LTP_CUDA(cudaMalloc(buf, need));
  *have = need;
  return LTP_OK;
}

int ltp_plan_stream(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                    const double* a_0, int64_t chunk, int32_t horizon, int64_t capacity,
                    ltp_chunk_consumer consume, void* user, ltp_stream_stats* stats, void* input_stream) {
  if (!p || n < 0 || chunk < 1 || chunk > 0x7fffffff || horizon < 0 || capacity < 1 ||
      (horizon > 0 && capacity < horizon) || p->params.dof < 1)
    return LTP_ERR_ARG;
  if (stats) std::memset(stats, 0, sizeof *stats);
  if (n == 0) return LTP_OK;
  if (!q_goal || !q_0 || !v_0 || !a_0) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  const int dof = p->params.dof;
  const int64_t c_max = n < chunk ? n : chunk;
  const int slots = n > c_max ? 2 : 1;
  const bool sorted = p->stream_sorted && horizon == 0;
  // per slot: compact chunk inputs, solution, work list (pipe_buf) and the four trajectory
  // fields + success flags (ring_buf), all reused by every second chunk
  ltp_solution ds[2];
  double* d_in[2][4];
  void* d_work[2];
  double* d_traj[2][4];
  uint8_t* d_succ[2];
  const size_t in_bytes = up((size_t)dof * (size_t)c_max * 8, 256);
  const size_t sol_bytes = carve_solution(nullptr, dof, c_max, &ds[0]);
  const size_t work_bytes = up(solve_scratch_bytes(dof, c_max), 256);
  const size_t order_bytes = up(order_scratch_bytes(c_max), 256);
  int* d_order[2];
  const size_t field_bytes = up((size_t)capacity * (size_t)c_max * (size_t)dof * 8, 256);
  const size_t succ_bytes = up((size_t)c_max, 256);
  for (int s = 0; s < slots; ++s) {
    if (!p->pipe_stream[s]) LTP_CUDA(cudaStreamCreateWithFlags(&p->pipe_stream[s], cudaStreamNonBlocking));
    int rc = grow(&p->pipe_buf[s], &p->pipe_bytes[s], 4 * in_bytes + sol_bytes + work_bytes + order_bytes);
    if (rc != LTP_OK) return rc;
    rc = grow(&p->ring_buf[s], &p->ring_bytes[s], 4 * field_bytes + succ_bytes);
    if (rc != LTP_OK) return rc;
    unsigned char* base = (unsigned char*)p->pipe_buf[s];
    for (int i = 0; i < 4; ++i) d_in[s][i] = (double*)(base + i * in_bytes);
    carve_solution(base + 4 * in_bytes, dof, c_max, &ds[s]);
    ds[s].t_opt = nullptr; ds[s].opt_case = nullptr; ds[s].ts_case = nullptr; ds[s].final_case = nullptr;
    ds[s].v_drive = nullptr;  // slot 7 of the records
    d_work[s] = base + 4 * in_bytes + sol_bytes;
    d_order[s] = (int*)(base + 4 * in_bytes + sol_bytes + work_bytes);
    unsigned char* ring = (unsigned char*)p->ring_buf[s];
    for (int i = 0; i < 4; ++i) d_traj[s][i] = (double*)(ring + i * field_bytes);
    d_succ[s] = ring + 4 * field_bytes;
  }
  if (!p->d_totals) LTP_CUDA(cudaMalloc(&p->d_totals, sizeof(StreamTotals)));
  LTP_CUDA(cudaMemsetAsync(p->d_totals, 0, sizeof(StreamTotals), p->pipe_stream[0]));
  LTP_CUDA(cudaStreamSynchronize(p->pipe_stream[0]));
  // The chunk copies run on the planner's own (non-blocking) streams: they must not start before
  // whatever the caller enqueued on input_stream to produce the inputs has run.
  {
    cudaEvent_t ready;
    LTP_CUDA(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ready, (cudaStream_t)input_stream);
    for (int s = 0; s < slots && e == cudaSuccess; ++s) e = cudaStreamWaitEvent(p->pipe_stream[s], ready, 0);
    cudaEventDestroy(ready);
    if (e != cudaSuccess) return cuda_fail(e, "ltp_plan_stream(input_stream)");
  }
  const double* src[4] = {q_goal, q_0, v_0, a_0};
  int64_t k = 0;
  for (int64_t p0 = 0; p0 < n; p0 += c_max, ++k) {
    const int s = (int)(k % slots);
    const int64_t c = (n - p0) < c_max ? (n - p0) : c_max;
    cudaStream_t st = p->pipe_stream[s];
    // a chunk of a joint-major array is dof pieces of c values at a pitch of n values
    for (int i = 0; i < 4; ++i)
      LTP_CUDA(cudaMemcpy2DAsync(d_in[s][i], (size_t)c * 8, src[i] + p0, (size_t)n * 8, (size_t)c * 8, dof,
                                 cudaMemcpyDeviceToDevice, st));
    int rc = solve_launch(p, c, d_in[s][0], d_in[s][1], d_in[s][2], d_in[s][3], &ds[s], d_work[s], st);
    if (rc != LTP_OK) return rc;
    rc = sample_tm_launch(p, c, d_in[s][1], d_in[s][2], d_in[s][3], &ds[s], horizon, capacity, d_traj[s][0],
                          d_traj[s][1], d_traj[s][2], d_traj[s][3], d_succ[s],
                          sorted ? d_order[s] : nullptr, st);
    if (rc != LTP_OK) return rc;
    const unsigned tg = (unsigned)((c + 255) / 256);
    ltp_chunk_totals_kernel<<<tg < 1024u ? tg : 1024u, 256, 0, st>>>(c, dof, horizon, capacity, ds[s].traj_len,
                                                                     ds[s].reached, d_succ[s], p->d_totals);
    p->launches++;
    LTP_CUDA(cudaGetLastError());
    if (consume) {
      ltp_chunk view;
      view.first = p0;
      view.count = c;
      view.capacity = capacity;
      view.horizon = horizon;
      view.solution = ds[s];
      view.q_goal = d_in[s][0]; view.q_0 = d_in[s][1]; view.v_0 = d_in[s][2]; view.a_0 = d_in[s][3];
      view.q = d_traj[s][0]; view.v = d_traj[s][1]; view.a = d_traj[s][2]; view.j = d_traj[s][3];
      view.success = d_succ[s];
      view.order = sorted ? d_order[s] + kOrderBins : nullptr;
      const int crc = consume(user, &view, (void*)st);
      if (crc != 0) {
        for (int t = 0; t < slots; ++t) cudaStreamSynchronize(p->pipe_stream[t]);
        return crc < 0 ? crc : LTP_ERR_ARG;
      }
    }
  }
  for (int s = 0; s < slots; ++s) LTP_CUDA(cudaStreamSynchronize(p->pipe_stream[s]));
  if (stats) {
    StreamTotals h;
    LTP_CUDA(cudaMemcpy(&h, p->d_totals, sizeof h, cudaMemcpyDeviceToHost));
    stats->problems = h.problems; stats->reached = h.reached; stats->success = h.success;
    stats->clipped = h.clipped; stats->samples = h.samples; stats->max_traj_len = h.max_len;
    stats->chunks = k;
    stats->bytes = h.samples * 32;
  }
  return LTP_OK;
}

int ltp_transpose(ltp_planner* p, int64_t rows, int64_t cols, const double* src, double* dst, void* stream) {
  if (!p || rows < 0 || cols < 0) return LTP_ERR_ARG;
  if (rows == 0 || cols == 0) return LTP_OK;
  if (!src || !dst || src == dst) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  const int64_t gx = (cols + 31) / 32, gy = (rows + 31) / 32;
  if (gx > 0x7fffffff) return LTP_ERR_ARG;
  // grid.y holds at most 65535 tiles: walk the rows in slabs (a slab of rows of src is contiguous,
  // its transpose is a block of columns of dst with pitch `rows`)
  for (int64_t y0 = 0; y0 < gy; y0 += 65535) {
    const int64_t slab = (gy - y0) < 65535 ? (gy - y0) : 65535;
    const int64_t r_off = y0 * 32;
    const int64_t r_cnt = (rows - r_off) < slab * 32 ? (rows - r_off) : slab * 32;
    ltp_transpose_kernel<<<dim3((unsigned)gx, (unsigned)slab), 256, 0, (cudaStream_t)stream>>>(
        r_cnt, cols, rows, src + r_off * cols, dst + r_off);
    p->launches++;
  }
  LTP_CUDA(cudaGetLastError());
  return LTP_OK;
}

int ltp_advance_batch(ltp_planner* p, int64_t n, int32_t tick, int32_t clamp, int64_t capacity,
                      const int32_t* traj_len, const uint8_t* valid, const double* q, const double* v,
                      const double* a, double* q_0, double* v_0, double* a_0, void* stream) {
  if (!p || n < 0 || tick < 0 || capacity < 1 || p->params.dof < 1) return LTP_ERR_ARG;
  // without lengths every problem is read at `tick`: that sample has to exist
  if (!traj_len && tick >= capacity) return LTP_ERR_ARG;
  if (n == 0) return LTP_OK;
  if (!q || !v || !a || !q_0 || !v_0 || !a_0) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  dim3 grid((unsigned)((n + 255) / 256), p->params.dof);
  ltp_advance_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p->params, n, tick, clamp, capacity, traj_len, valid, q,
                                                             v, a, q_0, v_0, a_0);
  p->launches++;
  LTP_CUDA(cudaGetLastError());
  return LTP_OK;
}

// Small batches (the drop-in planTrajectory is n = 1) and the single-item calls: latency, not
// bandwidth. Everything the host exchanges with the kernels sits in one block of pinned,
// device-mapped host memory: the kernels read their inputs from it and write their results
// (for a plan: the sampled rows, as posted 32-byte PCIe writes) straight into it, so a call is
// two launches and ONE stream synchronisation -- no copies, no host round trip between solve
// and sampling.
constexpr size_t kStageHeadBytes = 65536;
constexpr size_t kStageRowBytes = 4u << 20;

static int ensure_stage(ltp_planner* p) {
  if (!p->h_stage) LTP_CUDA(cudaHostAlloc(&p->h_stage, kStageHeadBytes + kStageRowBytes, cudaHostAllocMapped));
  return LTP_OK;
}

static void launch_row_latency_sampler(ltp_planner* p, int64_t n, const double* q_0, const double* v_0,
                                       const double* a_0, const ltp_solution* ds, int32_t horizon, int64_t stride,
                                       double* rows, size_t field, uint8_t* row_ok, cudaStream_t st) {
  ProfScope ps(p, LTP_PROFILE_SAMPLE_ROWS, st);
  ltp_sample_row_latency_kernel<<<(unsigned)(n * p->params.dof), 32, 0, st>>>(
      p->params, n, q_0, v_0, a_0, to_dev(ds), horizon, stride, rows, rows + field, rows + 2 * field, rows + 3 * field,
      row_ok);
  p->launches++;
}

// solve + sample n <= 32 problems through the staging block; rows are left there:
// field f of row r (= problem * dof + joint) starts at (*rows) + (f * n * dof + r) * dstride
static int plan_small_run(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                          const double* a_0, int32_t horizon, int64_t dstride, bool want_rows, const double** rows,
                          const int32_t** lens, const uint8_t** row_ok) {
  const int dof = p->params.dof;
  const size_t dn = (size_t)dof * (size_t)n;
  int rc = ensure_stage(p);
  if (rc != LTP_OK) return rc;
  unsigned char* hs = (unsigned char*)p->h_stage;
  double* h_in = (double*)hs;                         // [4][dof][n]
  int32_t* h_len = (int32_t*)(hs + 4 * dn * 8);       // [n]     written by the solve kernel
  uint8_t* h_reached = (uint8_t*)(h_len + n);         // [n]     written by the solve kernel
  uint8_t* h_row_ok = h_reached + n;                  // [n*dof] written by the sampler
  double* h_rows = (double*)(hs + kStageHeadBytes);   // [4][n*dof][dstride]
  const double* user_in[4] = {q_goal, q_0, v_0, a_0};
  for (int i = 0; i < 4; ++i) std::memcpy(h_in + i * dn, user_in[i], dn * 8);
  ltp_solution ds;
  const size_t sol_bytes = carve_solution(nullptr, dof, n, &ds);
  rc = ensure_scratch(p, sol_bytes);
  if (rc != LTP_OK) return rc;
  carve_solution((unsigned char*)p->d_scratch, dof, n, &ds);
  ds.t_opt = nullptr; ds.opt_case = nullptr; ds.ts_case = nullptr; ds.final_case = nullptr;
  ds.v_drive = nullptr;
  ds.traj_len = h_len;
  ds.reached = h_reached;
  cudaStream_t st = p->stream;
  rc = ltp_solve_batch(p, n, h_in, h_in + dn, h_in + 2 * dn, h_in + 3 * dn, &ds, st);
  if (rc != LTP_OK) return rc;
  if (want_rows) {
    launch_row_latency_sampler(p, n, h_in + dn, h_in + 2 * dn, h_in + 3 * dn, &ds, horizon, dstride, h_rows,
                               dn * (size_t)dstride, h_row_ok, st);
    LTP_CUDA(cudaGetLastError());
  }
  LTP_CUDA(cudaStreamSynchronize(st));
  *rows = h_rows;
  *lens = h_len;
  *row_ok = h_row_ok;
  return LTP_OK;
}

static int plan_host_small(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                           const double* a_0, int32_t horizon, int64_t capacity, double* q, double* v, double* a,
                           double* j, int32_t* traj_len, uint8_t* success, int64_t* needed) {
  const int dof = p->params.dof;
  const size_t dn = (size_t)dof * (size_t)n;
  const int64_t dstride = (capacity + 3) / 4 * 4;
  const bool rows_ok = q && v && a && j;
  const double* h_rows;
  const int32_t* h_len;
  const uint8_t* h_row_ok;
  int rc = plan_small_run(p, n, q_goal, q_0, v_0, a_0, horizon, dstride, rows_ok, &h_rows, &h_len, &h_row_ok);
  if (rc != LTP_OK) return rc;
  int64_t need = horizon;
  for (int64_t i = 0; i < n; ++i) {
    traj_len[i] = h_len[i];
    if (horizon == 0) need = h_len[i] > need ? h_len[i] : need;
  }
  if (needed) *needed = need;
  if (need > capacity || !rows_ok) {
    for (int64_t i = 0; i < n; ++i) success[i] = 0;
    return need > capacity ? LTP_ERR_CAPACITY : LTP_ERR_ARG;
  }
  const size_t field = dn * (size_t)dstride;
  double* user_rows[4] = {q, v, a, j};
  for (int f = 0; f < 4; ++f)
    for (size_t r = 0; r < dn; ++r)
      std::memcpy(user_rows[f] + r * (size_t)capacity, h_rows + f * field + r * (size_t)dstride, (size_t)need * 8);
  for (int64_t i = 0; i < n; ++i) {
    uint8_t ok = 1;
    for (int k = 0; k < dof; ++k) ok &= h_row_ok[i * dof + k];
    success[i] = ok;
  }
  return LTP_OK;
}

int ltp_plan_one_view(ltp_planner* p, const double* q_goal, const double* q_0, const double* v_0, const double* a_0,
                      const double** rows4, int64_t* row_stride, int32_t* length, uint8_t* success) {
  if (!p || !q_goal || !q_0 || !v_0 || !a_0 || !rows4 || !row_stride || !length || !success || p->params.dof < 1)
    return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  const int dof = p->params.dof;
  int64_t dstride = (int64_t)(kStageRowBytes / (4 * (size_t)dof * 8)) & ~(int64_t)3;
  if (dstride > 16384) dstride = 16384;
  const double* h_rows;
  const int32_t* h_len;
  const uint8_t* h_row_ok;
  int rc = plan_small_run(p, 1, q_goal, q_0, v_0, a_0, 0, dstride, true, &h_rows, &h_len, &h_row_ok);
  if (rc != LTP_OK) return rc;
  *length = h_len[0];
  *row_stride = dstride;
  *success = 0;
  for (int f = 0; f < 4; ++f) rows4[f] = h_rows + (size_t)f * (size_t)dof * (size_t)dstride;
  if (h_len[0] > dstride) return LTP_ERR_CAPACITY;
  uint8_t ok = 1;
  for (int k = 0; k < dof; ++k) ok &= h_row_ok[k];
  *success = ok;
  return LTP_OK;
}

int ltp_plan_host(ltp_planner* p, int64_t n, const double* q_goal, const double* q_0, const double* v_0,
                  const double* a_0, int32_t horizon, int64_t capacity, double* q, double* v, double* a,
                  double* j, int32_t* traj_len, uint8_t* success, int64_t* needed) {
  if (!p || n < 0 || !q_goal || !q_0 || !v_0 || !a_0 || !traj_len || !success || p->params.dof < 1 ||
      horizon < 0 || capacity < 0)
    return LTP_ERR_ARG;
  if (needed) *needed = 0;
  if (n == 0) return LTP_OK;
  DeviceGuard g(p->device);
  const int dof = p->params.dof;
  const size_t dn = (size_t)dof * (size_t)n;
  if (n <= kTile && capacity >= 4 && horizon <= capacity && 4 * dn * 8 + 6 * dn <= kStageHeadBytes &&
      4 * dn * (size_t)((capacity + 3) / 4 * 4) * 8 <= kStageRowBytes)
    return plan_host_small(p, n, q_goal, q_0, v_0, a_0, horizon, capacity, q, v, a, j, traj_len, success, needed);
  ltp_solution ds;
  const size_t sol_bytes = carve_solution(nullptr, dof, n, &ds);
  const size_t in_bytes = up(dn * 8, 256);
  const size_t succ_bytes = up((size_t)n, 256);
  // device rows use a stride rounded up to 4 samples so that the vector-store path is taken
  const int64_t dstride = (capacity + 3) / 4 * 4;
  const size_t row_bytes = up(dn * (size_t)dstride * 8, 256);
  int rc = ensure_scratch(p, 4 * in_bytes + sol_bytes + succ_bytes + 4 * row_bytes);
  if (rc != LTP_OK) return rc;
  unsigned char* base = (unsigned char*)p->d_scratch;
  double* d_in[4];
  for (int i = 0; i < 4; ++i) d_in[i] = (double*)(base + i * in_bytes);
  size_t off = 4 * in_bytes;
  off += carve_solution(base + off, dof, n, &ds);
  ds.t_opt = nullptr; ds.opt_case = nullptr; ds.ts_case = nullptr; ds.final_case = nullptr;
  uint8_t* d_succ = base + off;
  off += succ_bytes;
  double* d_rows[4];
  for (int i = 0; i < 4; ++i) d_rows[i] = (double*)(base + off + i * row_bytes);
  cudaStream_t st = p->stream;
  const double* h_in[4] = {q_goal, q_0, v_0, a_0};
  for (int i = 0; i < 4; ++i) LTP_CUDA(cudaMemcpyAsync(d_in[i], h_in[i], dn * 8, cudaMemcpyHostToDevice, st));
  rc = ltp_solve_batch(p, n, d_in[0], d_in[1], d_in[2], d_in[3], &ds, st);
  if (rc != LTP_OK) return rc;
  LTP_CUDA(cudaMemcpyAsync(traj_len, ds.traj_len, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
  LTP_CUDA(cudaStreamSynchronize(st));
  int64_t need = horizon;
  if (horizon == 0)
    for (int64_t i = 0; i < n; ++i) need = traj_len[i] > need ? traj_len[i] : need;
  if (needed) *needed = need;
  if (need > capacity || !q || !v || !a || !j) {
    for (int64_t i = 0; i < n; ++i) success[i] = 0;
    return need > capacity ? LTP_ERR_CAPACITY : LTP_ERR_ARG;
  }
  rc = ltp_sample_batch(p, n, d_in[1], d_in[2], d_in[3], &ds, horizon, LTP_LAYOUT_ROWS, dstride, d_rows[0], d_rows[1],
                        d_rows[2], d_rows[3], d_succ, st);
  if (rc != LTP_OK) return rc;
  double* h_rows[4] = {q, v, a, j};
  for (int i = 0; i < 4; ++i)
    LTP_CUDA(cudaMemcpy2DAsync(h_rows[i], (size_t)capacity * 8, d_rows[i], (size_t)dstride * 8,
                               (size_t)need * 8, dn, cudaMemcpyDeviceToHost, st));
  LTP_CUDA(cudaMemcpyAsync(success, d_succ, (size_t)n, cudaMemcpyDeviceToHost, st));
  LTP_CUDA(cudaStreamSynchronize(st));
  return LTP_OK;
}

int ltp_opt_braking_host(ltp_planner* p, int joint, double v_0, double a_0, double* q_stop,
                         double* t_rel3, double* dir) {
  if (!p || joint < 0 || joint >= p->params.dof || !q_stop || !t_rel3 || !dir) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  int rc = ensure_stage(p);
  if (rc != LTP_OK) return rc;
  double* h = (double*)p->h_stage;  // [0]=v0 [1]=a0 [2]=q [3]=dir [4..6]=t_rel
  h[0] = v_0; h[1] = a_0;
  cudaStream_t st = p->stream;
  ltp_opt_braking_kernel<<<1, 32, 0, st>>>(p->params, joint, 1, h, h + 1, h + 2, h + 4, h + 3);
  p->launches++;
  LTP_CUDA(cudaGetLastError());
  LTP_CUDA(cudaStreamSynchronize(st));
  *q_stop = h[2];
  *dir = h[3];
  t_rel3[0] = h[4]; t_rel3[1] = h[5]; t_rel3[2] = h[6];
  return LTP_OK;
}

int ltp_opt_switch_times_host(ltp_planner* p, int joint, double q_goal, double q_0, double v_0,
                              double a_0, double v_drive, double* t7, double* dir, uint8_t* mod,
                              uint8_t* kase, uint8_t* ok) {
  if (!p || joint < 0 || joint >= p->params.dof || !t7 || !dir || !mod || !ok) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  int rc = ensure_stage(p);
  if (rc != LTP_OK) return rc;
  double* h = (double*)p->h_stage;  // in: 0..4; out: t 5..11, dir 12; bytes at 16*8
  uint8_t* hb = (uint8_t*)(h + 16);
  h[0] = q_goal; h[1] = q_0; h[2] = v_0; h[3] = a_0; h[4] = v_drive;
  for (int k = 5; k < 13; ++k) h[k] = 0.0;  // the kernel leaves t untouched on the cc:340-344 failure
  cudaStream_t st = p->stream;
  ltp_opt_switch_times_kernel<<<1, 32, 0, st>>>(p->params, joint, 1, h, h + 1, h + 2, h + 3, h + 4, h + 5,
                                                h + 12, hb, hb + 1, hb + 2);
  p->launches++;
  LTP_CUDA(cudaGetLastError());
  LTP_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 7; ++k) t7[k] = h[5 + k];
  *dir = h[12];
  *mod = hb[0];
  if (kase) *kase = hb[1];
  *ok = hb[2];
  return LTP_OK;
}

int ltp_time_scaling_host(ltp_planner* p, int joint, double q_goal, double q_0, double v_0, double a_0,
                          double dir, double t_required, double* t7, double* v_drive, uint8_t* mod,
                          uint8_t* ts_case, uint8_t* ok) {
  if (!p || joint < 0 || joint >= p->params.dof || !t7 || !v_drive || !mod || !ok) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  int rc = ensure_stage(p);
  if (rc != LTP_OK) return rc;
  double* h = (double*)p->h_stage;  // in 0..5; out t 6..12, v_drive 13
  uint8_t* hb = (uint8_t*)(h + 16);
  h[0] = q_goal; h[1] = q_0; h[2] = v_0; h[3] = a_0; h[4] = dir; h[5] = t_required;
  for (int k = 6; k < 14; ++k) h[k] = 0.0;
  cudaStream_t st = p->stream;
  ltp_time_scaling_kernel<<<1, 32, 0, st>>>(p->params, joint, 1, h, h + 1, h + 2, h + 3, h + 4, h + 5, h + 6,
                                            h + 13, hb, hb + 1, hb + 2, hb + 3);
  p->launches++;
  LTP_CUDA(cudaGetLastError());
  LTP_CUDA(cudaStreamSynchronize(st));
  for (int k = 0; k < 7; ++k) t7[k] = h[6 + k];
  *v_drive = h[13];
  *mod = hb[0];
  if (ts_case) *ts_case = hb[1];
  *ok = hb[3];
  return LTP_OK;
}

int ltp_get_trajectory_host(ltp_planner* p, const double* t7, const double* dir, const uint8_t* mod,
                            const double* q_0, const double* v_0, const double* a_0,
                            const double* v_drive, int64_t capacity, double* q, double* v, double* a,
                            double* j, int32_t* length, int64_t* needed) {
  if (!p || !t7 || !dir || !mod || !q_0 || !v_0 || !a_0 || !v_drive || !length || capacity < 0 ||
      p->params.dof < 1)
    return LTP_ERR_ARG;
  const int dof = p->params.dof;
  // cc:716-719 on the host (same IEEE operations as the device would perform)
  int len = 0;
  for (int i = 0; i < dof; ++i) {
    const int li = ltp::samples_for(t7[7 * i + 6], p->params.ts);
    len = li > len ? li : len;
  }
  *length = len;
  if (needed) *needed = len;
  if (len > capacity) return LTP_ERR_CAPACITY;
  if (len == 0) return LTP_OK;
  if (!q || !v || !a || !j) return LTP_ERR_ARG;
  DeviceGuard g(p->device);
  const int64_t dstride = ((int64_t)len + 3) / 4 * 4;
  ltp_solution ds;
  if (4 * (size_t)dof * (size_t)dstride * 8 <= kStageRowBytes) {
    // everything through the mapped staging block: no copies, one launch, one synchronisation
    int rc = ensure_stage(p);
    if (rc != LTP_OK) return rc;
    double* h = (double*)p->h_stage;
    std::memset(&ds, 0, sizeof ds);
    ds.t_scaled = h;                       // [dof][1] records of 8 doubles (the block is page-aligned)
    ds.dir = h + 8 * dof;
    double* h_in = h + 9 * dof;            // q_0, v_0, a_0
    ds.traj_len = (int32_t*)(h + 12 * dof);
    ds.mod = (uint8_t*)(ds.traj_len + 2);
    ds.reached = ds.mod + dof;
    uint8_t* row_ok = ds.reached + 1;
    double* h_rows = (double*)((unsigned char*)p->h_stage + kStageHeadBytes);
    for (int i = 0; i < dof; ++i) {
      for (int k = 0; k < 7; ++k) ds.t_scaled[8 * i + k] = t7[7 * i + k];
      ds.t_scaled[8 * i + 7] = v_drive[i];
      ds.dir[i] = dir[i];
      h_in[i] = q_0[i]; h_in[dof + i] = v_0[i]; h_in[2 * dof + i] = a_0[i];
      ds.mod[i] = mod[i];
    }
    ds.traj_len[0] = len;
    ds.reached[0] = 1;
    cudaStream_t st = p->stream;
    const size_t field = (size_t)dof * (size_t)dstride;
    launch_row_latency_sampler(p, 1, h_in, h_in + dof, h_in + 2 * dof, &ds, 0, dstride, h_rows, field, row_ok, st);
    LTP_CUDA(cudaGetLastError());
    LTP_CUDA(cudaStreamSynchronize(st));
    double* user_rows[4] = {q, v, a, j};
    for (int f = 0; f < 4; ++f)
      for (int r = 0; r < dof; ++r)
        std::memcpy(user_rows[f] + (size_t)r * (size_t)capacity, h_rows + f * field + (size_t)r * (size_t)dstride,
                    (size_t)len * 8);
    return LTP_OK;
  }
  const size_t sol_bytes = carve_solution(nullptr, dof, 1, &ds);
  const size_t in_bytes = up((size_t)dof * 8, 256);
  const size_t row_bytes = up((size_t)dof * (size_t)dstride * 8, 256);
  int rc = ensure_scratch(p, 3 * in_bytes + sol_bytes + 256 + 4 * row_bytes);
  if (rc != LTP_OK) return rc;
  unsigned char* base = (unsigned char*)p->d_scratch;
  double* d_in[3];
  for (int i = 0; i < 3; ++i) d_in[i] = (double*)(base + i * in_bytes);
  size_t off = 3 * in_bytes;
  off += carve_solution(base + off, dof, 1, &ds);
  uint8_t* d_succ = base + off;
  off += 256;
  double* d_rows[4];
  for (int i = 0; i < 4; ++i) d_rows[i] = (double*)(base + off + i * row_bytes);
  cudaStream_t st = p->stream;
  // [dof][7] + v_drive -> [dof][1] records
  double tt[8 * LTP_MAX_DOF];
  for (int i = 0; i < dof; ++i) {
    for (int k = 0; k < 7; ++k) tt[8 * i + k] = t7[7 * i + k];
    tt[8 * i + 7] = v_drive[i];
  }
  const uint8_t one = 1;
  const int32_t len32 = len;
  LTP_CUDA(cudaMemcpyAsync(ds.t_scaled, tt, sizeof(double) * 8 * dof, cudaMemcpyHostToDevice, st));
  LTP_CUDA(cudaMemcpyAsync(ds.dir, dir, sizeof(double) * dof, cudaMemcpyHostToDevice, st));
  LTP_CUDA(cudaMemcpyAsync(ds.mod, mod, dof, cudaMemcpyHostToDevice, st));
  LTP_CUDA(cudaMemcpyAsync(ds.traj_len, &len32, 4, cudaMemcpyHostToDevice, st));
  LTP_CUDA(cudaMemcpyAsync(ds.reached, &one, 1, cudaMemcpyHostToDevice, st));
  const double* h_in[3] = {q_0, v_0, a_0};
  for (int i = 0; i < 3; ++i) LTP_CUDA(cudaMemcpyAsync(d_in[i], h_in[i], sizeof(double) * dof, cudaMemcpyHostToDevice, st));
  LTP_CUDA(cudaStreamSynchronize(st));  // the small host arrays above live on this stack frame
  rc = ltp_sample_batch(p, 1, d_in[0], d_in[1], d_in[2], &ds, 0, LTP_LAYOUT_ROWS, dstride, d_rows[0], d_rows[1], d_rows[2],
                        d_rows[3], d_succ, st);
  if (rc != LTP_OK) return rc;
  double* h_rows[4] = {q, v, a, j};
  for (int i = 0; i < 4; ++i)
    LTP_CUDA(cudaMemcpy2DAsync(h_rows[i], (size_t)capacity * 8, d_rows[i], (size_t)dstride * 8,
                               (size_t)len * 8, dof, cudaMemcpyDeviceToHost, st));
  LTP_CUDA(cudaStreamSynchronize(st));
  return LTP_OK;
}

}  // extern "C"
