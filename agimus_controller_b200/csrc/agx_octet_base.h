// agx_octet_base.h — lane/octet programming model shared by every kernel of the solve path.
//
// Work mapping ("octet"): 8 consecutive lanes of a warp own one 7-DoF entity — lane j owns joint j
// (and, in the Riccati sweep, columns j and j+7 of every 14-wide matrix); lane 7 idles on benign
// data.  A warp therefore carries 4 entities.  Lanes exchange data through a per-octet shared-memory
// board, separated by octet-masked __syncwarp().
//
// The kernels are ordinary CUDA.  For machines without a GPU the same source also compiles with g++
// against tests/emul/cpu_simt.h (test infrastructure: CUDA threads become fibers, __syncwarp and
// __shfl_sync become cooperative barriers), which is how the kernels are checked against the oracle
// in the CPU test-suite.  Nothing under agimus_controller_b200/ ever loads the emulation.
#ifndef AGX_OCTET_BASE_H_
#define AGX_OCTET_BASE_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(AGX_EMULATE)
#define AGX_GPU 1
#define AGX_DEV __device__ __forceinline__
#define AGX_RSQRT(x) rsqrt(x)
#define AGX_SINCOS(x, s, c) sincos((x), (s), (c))
#define AGX_SMEM(name) extern __shared__ __align__(16) double name[]
#define AGX_PREFETCH(p) asm volatile("prefetch.global.L1 [%0];" ::"l"(p))
// 16-byte asynchronous copy global -> shared (LDGSTS), no registers involved
#define AGX_CP_ASYNC16(dst_smem, src_gmem)                                                       \
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), \
               "l"(src_gmem))
#define AGX_CP_ASYNC8(dst_smem, src_gmem)                                                        \
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), \
               "l"(src_gmem))
#define AGX_CP_ASYNC_COMMIT() asm volatile("cp.async.commit_group;")
#define AGX_CP_ASYNC_WAIT_ALL() asm volatile("cp.async.wait_group 0;")
// FP64 tensor-core tile product D(8x8) = A(8x4) B(4x8) + C (DMMA): lane T holds a = A[T/4][T%4],
// b = B[T%4][T/4], c/d = C[T/4][2(T%4) + {0,1}]
#define AGX_DMMA(d0, d1, a, b, c0, c1)                                                              \
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"      \
               : "=d"(d0), "=d"(d1)                                                                 \
               : "d"(a), "d"(b), "d"(c0), "d"(c1))
#else
#define AGX_GPU 0
#define AGX_DEV inline
#define AGX_RSQRT(x) (1.0 / sqrt(x))
#define AGX_SINCOS(x, s, c) \
  do {                      \
    *(s) = sin(x);          \
    *(c) = cos(x);          \
  } while (0)
#define AGX_SMEM(name) double* name = reinterpret_cast<double*>(simt::g_smem)
#define AGX_PREFETCH(p) ((void)(p))
#define AGX_CP_ASYNC16(dst_smem, src_gmem) memcpy((dst_smem), (src_gmem), 16)
#define AGX_CP_ASYNC8(dst_smem, src_gmem) memcpy((dst_smem), (src_gmem), 8)
#define AGX_CP_ASYNC_COMMIT() ((void)0)
#define AGX_CP_ASYNC_WAIT_ALL() ((void)0)
#define AGX_DMMA(d0, d1, a, b, c0, c1) agx_emul_dmma((d0), (d1), (a), (b), (c0), (c1))
#endif

// device wall clock in nanoseconds (%globaltimer); the emulator build reads the host's steady clock
#if AGX_GPU
__device__ __forceinline__ long long agx_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}
#else
#include <chrono>
inline long long agx_now_ns() {
  return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#endif

namespace agx {

constexpr int NJ = 7;            // joints handled by one octet
constexpr int NX = 2 * NJ;       // state dimension
constexpr int REF_SIZE = 6 * NJ + 20;  // [xref 14][wx 14][uref 7][wu 7][Rref 9][pref 3][wpose 6][wcol 2]

// ---- device model table (doubles): joint fields are stored field-major / joint-minor so that
// lane j reads model[f * 8 + j]; tail holds gravity and the task frame.
constexpr int MF_RP = 0;     // 9 fields: placement rotation (row-major)
constexpr int MF_PP = 9;     // 3: placement translation
constexpr int MF_MASS = 12;  // 1
constexpr int MF_COM = 13;   // 3
constexpr int MF_INERTIA = 16;  // 6: xx xy xz yy yz zz about the COM
constexpr int MF_ARM = 22;   // 1
constexpr int MF_NFIELDS = 24;
constexpr int MT_GRAV = MF_NFIELDS * 8;  // 3
constexpr int MT_FR = MT_GRAV + 3;       // 9 frame rotation
constexpr int MT_FP = MT_FR + 9;         // 3 frame translation
// collision geometry (A10): capsule c = [a0 3][a1 3][radius][parent joint or -1] in the parent frame, then
// [n_pairs][alpha][pair0 a][pair0 b][pair1 a][pair1 b]
constexpr int MAX_CAPS = 4;
constexpr int MAX_PAIRS = 2;
constexpr int MT_CAP = MT_FP + 3 + 1;           // 208
constexpr int MT_COL = MT_CAP + 8 * MAX_CAPS;   // 240
constexpr int MODEL_SIZE = MT_COL + 8;          // 248 doubles
constexpr int N_COST_TERMS = 13;  // agx_cost_terms row: [state, control, goal, r6 (6), collision cost (2), distance (2)]

// ---- compact node records written by calc_diff, read by the Riccati sweep (instead of the 658 dense
// doubles of Fx, Fu, Lx, Lu, Lxx, Lxu, Luu per node).
// Dynamics record (octet kernel): field-major / lane-minor, rec[k * 8 + lane] — an octet's store of one
// field is one 64-byte line.
constexpr int REC_FIELDS = 23;
constexpr int REC_SIZE = 8 * REC_FIELDS;  // 184 doubles (1472 B)
constexpr int RK_AQ = 0;     // 7: dt * da/dq [:, j]
constexpr int RK_AV = 7;     // 7: dt * da/dv [:, j]
constexpr int RK_MI = 14;    // 7: dt * Minv [:, j]
constexpr int RK_QN = 21;    // xnext (q part)
constexpr int RK_VN = 22;    // xnext (v part)
// Cost record (thread-per-node kernel): 64 consecutive doubles per node (512 B), all scaled by
// s = dt for running nodes and 1 for the terminal node.
constexpr int CREC_SIZE = 64;
constexpr int CK_LQQ = 0;    // 28: Lqq, symmetric, packed lower triangle column by column (lidx)
constexpr int CK_LVV = 28;   // 7: diagonal of Lvv
constexpr int CK_LUU = 35;   // 7: diagonal of Luu
constexpr int CK_LQ = 42;    // 7
constexpr int CK_LV = 49;    // 7
constexpr int CK_LU = 56;    // 7
constexpr int CK_COST = 63;  // node cost

// ---- per-problem solver state (SoA arrays, one entry per problem; device memory)
struct SolverState {
  double* xreg;       // current regularisation (xreg == ureg)
  double* cost;       // cost of the current candidate
  double* dg;         // expected-improvement terms (updateExpectedImprovement)
  double* dq;
  double* stop;       // |d1 + d2/2|
  int32_t* is_feasible;
  int32_t* was_feasible;
  int32_t* recalc;    // the dynamics part of the node records must be recomputed (a step was accepted)
  int32_t* done;      // problem finished (converged or failed): kernels skip it
  int32_t* status;
  int32_t* iters;
  int32_t* cur;       // which of the two (xs, us) buffers holds the current candidate
  double* dv;         // gap term of the expected improvement for the alpha = 1 trial
  int32_t* recalc_cost;  // the cost part of the node records must be recomputed for the candidate
  int32_t* pending;   // the alpha = 1 trial was rejected: the line search continues with smaller steps
  int32_t* roll_ok;   // the alpha = 1 rollout met no NaN / failed factorisation
  long long* t0;      // [1] device time stamp (ns) of the start of the solve (max_solve_time)
};

// ---- 3-vector helpers (per-lane, register resident)
AGX_DEV void cross3(const double* a, const double* b, double* o) {
  const double x = a[1] * b[2] - a[2] * b[1];
  const double y = a[2] * b[0] - a[0] * b[2];
  const double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
AGX_DEV double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
AGX_DEV double dot6(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}
AGX_DEV void mv3(const double* R, const double* x, double* o) {
  const double a = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
  const double b = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
  const double c = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
  o[0] = a; o[1] = b; o[2] = c;
}
AGX_DEV void mtv3(const double* R, const double* x, double* o) {
  const double a = R[0] * x[0] + R[3] * x[1] + R[6] * x[2];
  const double b = R[1] * x[0] + R[4] * x[1] + R[7] * x[2];
  const double c = R[2] * x[0] + R[5] * x[1] + R[8] * x[2];
  o[0] = a; o[1] = b; o[2] = c;
}
// symmetric 3x3 stored as xx xy xz yy yz zz
AGX_DEV void symv3(const double* S, const double* x, double* o) {
  const double a = S[0] * x[0] + S[1] * x[1] + S[2] * x[2];
  const double b = S[1] * x[0] + S[3] * x[1] + S[4] * x[2];
  const double c = S[2] * x[0] + S[4] * x[1] + S[5] * x[2];
  o[0] = a; o[1] = b; o[2] = c;
}
// spatial cross products, vectors are [lin(3); ang(3)]
AGX_DEV void crm6(const double* a, const double* b, double* o) {  // motion x motion
  double t1[3], t2[3], t3[3];
  cross3(a + 3, b, t1);
  cross3(a, b + 3, t2);
  cross3(a + 3, b + 3, t3);
  o[0] = t1[0] + t2[0]; o[1] = t1[1] + t2[1]; o[2] = t1[2] + t2[2];
  o[3] = t3[0]; o[4] = t3[1]; o[5] = t3[2];
}
AGX_DEV void crf6(const double* a, const double* f, double* o) {  // motion x* force
  double t1[3], t2[3], t3[3];
  cross3(a + 3, f, t1);
  cross3(a + 3, f + 3, t2);
  cross3(a, f, t3);
  o[0] = t1[0]; o[1] = t1[1]; o[2] = t1[2];
  o[3] = t2[0] + t3[0]; o[4] = t2[1] + t3[1]; o[5] = t2[2] + t3[2];
}
// rigid-body inertia about the world origin: Y = (m, mc = m*c, Ibar = Ic - m [c]x[c]x), 10 numbers
// layout: Y[0] = m, Y[1..3] = mc, Y[4..9] = Ibar (xx xy xz yy yz zz)
AGX_DEV void inertia_apply(const double* Y, const double* mo, double* f) {
  double wxmc[3], Iw[3], mcxv[3];
  cross3(mo + 3, Y + 1, wxmc);
  symv3(Y + 4, mo + 3, Iw);
  cross3(Y + 1, mo, mcxv);
  f[0] = Y[0] * mo[0] + wxmc[0];
  f[1] = Y[0] * mo[1] + wxmc[1];
  f[2] = Y[0] * mo[2] + wxmc[2];
  f[3] = Iw[0] + mcxv[0];
  f[4] = Iw[1] + mcxv[1];
  f[5] = Iw[2] + mcxv[2];
}

}  // namespace agx
#endif  // AGX_OCTET_BASE_H_
