// agx_kernels.cuh — the CUDA kernels of the batched solve path for the 7-joint chain (fp64, sm_100a).
//
//   calc_diff_kernel          one octet per (problem, node): problem.calc + calcDiff -> compact dynamics record
//                             (and the cost record where it is stale)
//   node_cost_kernel          one THREAD per (problem, node): cost terms and their Gauss-Newton derivatives -> cost record
//   backward_mma_kernel       one WARP per problem on the FP64 tensor cores (agx_riccati_mma.cuh): gaps, Riccati sweep
//                             t = T-1..0 with regularisation retries, gains K/k, expected-improvement terms
//                             (SolverFDDP::backwardPass, computeGains, updateExpectedImprovement);
//   backward_kernel           the same sweep with DFMA, one octet per problem (cross-check, AGX_BW=octet)
//   rollout_try_kernel        one octet per problem: nonlinear rollout of the trial step with gap contraction
//   rollout_try2_kernel       the same with two warps per group of four problems (latency mode)
//   accept_linesearch_kernel  acceptance test, deferred / in-line line search over alpha = 2^-n, regularisation update,
//                             stop criterion (SolverFDDP::forwardPass, tryStep, expectedImprovement, solve loop body)
//   calc / rollout / integrate / rnea / shift / cost_terms / expand / gather_refs / mask_refs / extract_gradients /
//   init / finalize kernels   the remaining entry points of include/agx.h
// The SQP mode is in agx_sqp.cuh, the kernels for general kinematic trees in agx_tree.cuh.
//
// Reference entry points replaced: solver.solve at
// agimus_controller/agimus_controller/ocp_base_croco.py:172 and the Crocoddyl objects built at
// :36-64; algorithm statements in SURVEY.md Appendix B.4/B.5.
//
// No host synchronisation happens inside a solve: all per-problem decisions (step acceptance,
// regularisation, termination, the max_solve_time deadline) are taken on the device and kept in SolverState.
#ifndef AGX_KERNELS_CUH_
#define AGX_KERNELS_CUH_

#include "agx_octet_base.h"
#include "agx_dynamics.inl"
#include "agx_node.inl"

#ifndef AGX_CD_THREADS
#define AGX_CD_THREADS 64
#endif
// AGX_CD_SYNC: CTA-wide barriers at the phase boundaries of calc_diff keep the warps of a CTA at the same place in the
// (long, fully unrolled) instruction stream, so that they share instruction fetches
#ifdef AGX_CD_SYNC
#define AGX_CD_PHASE() __syncthreads()
#else
#define AGX_CD_PHASE() ((void)0)
#endif

namespace agx {

struct Problem {
  const double* model;  // [n_models][MODEL_SIZE]
  const double* refs;   // [B][T+1][REF_SIZE]
  const double* dts;    // [T]
  int n_models;         // 1 (shared) or B (one table per problem)
  int B, T;
};

struct FddpOpts {
  double reg_min, reg_max, reg_incfactor, reg_decfactor;
  double th_grad, th_stepdec, th_stepinc, th_acceptstep, th_acceptnegstep, th_stop;
  double reg_init;
  int fixed_iters, n_alphas;
  int max_iter;  // iteration budget of every problem (a problem whose search was deferred finishes a round later)
  int defer;     // rounds a problem may fall behind by deferring rejected step lengths into the next round's forward pass (0: in-line search only; see accept_linesearch_kernel)
  long long max_solve_ns;  // max_solve_time in ns (0: none), measured on the device clock from SolverState::t0
  int accept_rule;         // agx_fddp_opts::accept_rule
};

// workspace of a solve (device pointers owned by the handle)
struct Work {
  double* xs;    // [2][B][T+1][NX]   candidate / trial, selected per problem by SolverState::cur
  double* us;    // [2][B][T][NJ]
  double* rec;   // [B][T+1][REC_SIZE]   dynamics records
  double* crec;  // [B][T+1][CREC_SIZE]  cost records
  double* fs;    // [B][T+1][NX]      gaps
  double* gv;    // [B][T+1][NX]      Vxx_t fs_t
  double* K;     // [B][T][NJ][NX]
  double* k;     // [B][T][NJ]
  const double* x0;  // [B][NX]
};

// doubles of shared memory per octet in the node kernels; the odd stride puts the two octets of a
// half-warp on different banks (64-bit accesses are served per half-warp)
constexpr int OCT_BOARD = BRD_B + BRD_C + 1;

#define AGX_OCTET_SETUP()                                   \
  const int j = (int)(threadIdx.x & 7u);                    \
  const unsigned omask = 0xFFu << (threadIdx.x & 24u);      \
  const int oct_in_cta = (int)(threadIdx.x >> 3);           \
  const int octs_per_cta = (int)(blockDim.x >> 3);          \
  const long long ent = (long long)blockIdx.x * octs_per_cta + oct_in_cta;

AGX_DEV const double* model_of(const Problem& P, int b) {
  return P.model + (P.n_models > 1 ? (size_t)b * MODEL_SIZE : 0);
}

AGX_DEV void lane_load_state(LaneDyn& d, int j, const double* x, const double* u) {
  const bool live = j < NJ;
  d.q = live ? x[j] : 0.0;
  d.qd = live ? x[NJ + j] : 0.0;
  d.u = (live && u) ? u[j] : 0.0;
}

// ---------------------------------------------------------------------------------------------
// Dynamics part of calc + calcDiff of one running node: xnext, dt da/dq, dt da/dv, dt Minv columns
// into the compact record.  A failed factorisation leaves NaNs (they make the sweep fail).
AGX_DEV void node_dyn_diff(LaneDyn& d, int j, unsigned omask, const double* __restrict__ model, double dt, double* sb,
                           double* sc, double* __restrict__ rec) {
  node_kinematics(d, j, omask, model);
  AGX_CD_PHASE();
  double L[28], rinv[NJ];
  const bool ok = node_forward_dynamics<true>(d, j, omask, model, sb, sc, L, rinv);
  AGX_CD_PHASE();
  rec[RK_QN * 8 + j] = ok ? d.q + (d.qd * dt + d.qdd * (dt * dt)) : nan("");
  rec[RK_VN * 8 + j] = ok ? d.qd + d.qdd * dt : nan("");
  node_rnea_derivatives(d, j, omask, sb);
  AGX_CD_PHASE();
  AGX_OSYNC();
  factor_reload(sc, L, rinv);
  double col[NJ];
  solve_column(L, rinv, d.tq, -dt, col);
#pragma unroll
  for (int i = 0; i < NJ; ++i) rec[(RK_AQ + i) * 8 + j] = col[i];
  solve_column(L, rinv, d.tv, -dt, col);
#pragma unroll
  for (int i = 0; i < NJ; ++i) rec[(RK_AV + i) * 8 + j] = col[i];
  double ej[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) ej[i] = (i == j) ? 1.0 : 0.0;
  solve_column(L, rinv, ej, dt, col);
#pragma unroll
  for (int i = 0; i < NJ; ++i) rec[(RK_MI + i) * 8 + j] = col[i];
}

// calc of one node: cost (scaled) and this lane's entries of xnext.  Returns false on failure.
template <bool COL = false>
AGX_DEV bool node_calc(LaneDyn& d, int j, unsigned omask, const double* __restrict__ model,
                       const double* __restrict__ ref, double dt, bool terminal, double* sb, double* sc, double* cost,
                       double* qn, double* vn) {
  node_kinematics(d, j, omask, model);
  const double l = node_costs<false, COL>(d, j, omask, model, ref, terminal, sb, nullptr, nullptr, nullptr, nullptr);
  if (terminal) {
    *cost = l;
    *qn = d.q;
    *vn = d.qd;
    return true;
  }
  double L[28], rinv[NJ];
  const bool ok = node_forward_dynamics<false>(d, j, omask, model, sb, sc, L, rinv);
  *cost = dt * l;
  *qn = d.q + (d.qd * dt + d.qdd * (dt * dt));
  *vn = d.qd + d.qdd * dt;
  return ok;
}

// ---------------------------------------------------------------------------------------------
// xs/us addressing: `cur` (may be null) selects one of two stacked buffers per problem
AGX_DEV size_t buf_of(const int32_t* cur, int b, bool other) {
  if (!cur) return 0;
  const int c = cur[b] & 1;
  return (size_t)(other ? (c ^ 1) : c);
}

// problem.calcDiff: one octet per (problem, node).  Always the dynamics record; the cost record only where
// `recalc_cost` says so (null = everywhere) — inside a solve the cost records normally come from
// node_cost_kernel, which evaluated them for the accepted trial already.
#ifndef AGX_CD_MINB
#define AGX_CD_MINB 4
#endif
// COL: the models carry collision pairs (A10); the plain instantiation is the one every benchmark runs.
template <bool COL>
__global__ void __launch_bounds__(AGX_CD_THREADS, AGX_CD_MINB)
calc_diff_kernel(Problem P, const double* __restrict__ xs, const double* __restrict__ us,
                                 const int32_t* __restrict__ cur, const int32_t* __restrict__ recalc,
                                 const int32_t* __restrict__ recalc_cost, int cost_everywhere,
                                 const int32_t* __restrict__ done, double* __restrict__ rec, double* __restrict__ crec) {
  AGX_SMEM(smem);
  const int j = (int)(threadIdx.x & 7u);
  const int oct_in_cta = (int)(threadIdx.x >> 3);
  const long long ent = (long long)blockIdx.x * (int)(blockDim.x >> 3) + oct_in_cta;
  // Whole-warp collectives: the four octets of a warp run the same instruction stream (octets that have nothing
  // to do EXIT, they never branch around a collective), so shuffles and barriers use the constant full mask and
  // compile to plain SHFL / WARPSYNC without the divergence-safe MATCH / VOTE prologue of a run-time mask.
  const unsigned omask = 0xffffffffu;
  const int T1 = P.T + 1;
  if (ent >= (long long)P.B * T1) return;
  const int b = (int)(ent / T1), t = (int)(ent % T1);
  // the per-problem flags are independent loads: issue them together, then branch (a chain of dependent
  // global loads at the start of every octet's work is pure exposed latency at two warps per scheduler)
  const int f_done = done ? done[b] : 0;
  const int f_dyn = recalc ? recalc[b] : 1;
  const int f_cost = recalc_cost ? recalc_cost[b] : 0;
  // cost_everywhere: 0 = cost records where they are stale, 1 = everywhere, 2 = everywhere and no dynamics records,
  // 3 = as 2 on the OTHER trajectory buffer (the trial of the line search): the latency mode's cost evaluation
  const int f_cur = cur ? ((cur[b] ^ (cost_everywhere == 3 ? 1 : 0)) & 1) : 0;
  if (f_done) return;
  const bool do_dyn = f_dyn != 0 && cost_everywhere < 2;
  const bool do_cost = cost_everywhere || f_cost != 0;
  if (!do_dyn && !do_cost) return;
  double* sb = smem + oct_in_cta * OCT_BOARD;
  double* sc = sb + BRD_B;
  const size_t buf = (size_t)f_cur;
  const double* x = xs + ((buf * P.B + b) * T1 + t) * NX;
  const bool terminal = t == P.T;
  const bool live = j < NJ;
  const double* model = model_of(P, b);
  double* R = rec + (size_t)ent * REC_SIZE;
  LaneDyn d;
  lane_load_state(d, j, x, terminal ? nullptr : us + ((buf * P.B + b) * P.T + t) * NJ);
  if (__any_sync(omask, do_cost)) {
    // octet version of the cost record (same numbers as thread_node_cost up to rounding); the decision is
    // warp-uniform, octets that did not ask for it compute and discard
    LaneDyn dk = d;
    node_kinematics(dk, j, omask, model);
    const double* ref = P.refs + (size_t)ent * REF_SIZE;
    double* C = crec + (size_t)ent * CREC_SIZE;
    const double s = terminal ? 1.0 : P.dts[t];
    double lq, lv, lu, Lqq[NJ];
    const double l = node_costs<true, COL>(dk, j, omask, model, ref, terminal, sb, &lq, &lv, &lu, Lqq);
    if (do_cost) {
      if (live) {
#pragma unroll
        for (int i = 0; i < NJ; ++i)
          if (i >= j) C[CK_LQQ + lidx(i, j)] = s * Lqq[i];
        C[CK_LVV + j] = s * ref[NX + NJ + j];
        C[CK_LUU + j] = terminal ? 0.0 : s * ref[2 * NX + NJ + j];
        C[CK_LQ + j] = s * lq;
        C[CK_LV + j] = s * lv;
        C[CK_LU + j] = s * lu;
      } else {
        C[CK_COST] = s * l;
      }
    }
  }
  if (!do_dyn) return;
  if (terminal) {
    // terminal model (dt = 0): xnext = x, no dynamics derivatives
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      R[(RK_AQ + i) * 8 + j] = 0.0;
      R[(RK_AV + i) * 8 + j] = 0.0;
      R[(RK_MI + i) * 8 + j] = 0.0;
    }
    R[RK_QN * 8 + j] = live ? x[j] : 0.0;
    R[RK_VN * 8 + j] = live ? x[NJ + j] : 0.0;
    return;
  }
  node_dyn_diff(d, j, omask, model, P.dts[t], sb, sc, R);
}

// problem.calc / calcDiff, cost part: ONE THREAD per (problem, node).  `other` selects the trial buffer
// (the candidate being evaluated by the line search); `gate` (may be null) restricts the work to problems
// whose flag is non-zero.  DERIV: the cost part of the node record is (re)written.
constexpr int COST_STAGE = 32 * (CREC_SIZE + 1) + 32;  // doubles of shared memory per warp of node_cost_kernel
template <bool DERIV, bool COL>
__global__ void node_cost_kernel(Problem P, const double* __restrict__ xs, const double* __restrict__ us,
                                 const int32_t* __restrict__ cur, int other, const int32_t* __restrict__ done,
                                 const int32_t* __restrict__ gate, double* __restrict__ crec,
                                 double* __restrict__ out_cost) {
  AGX_SMEM(smem);
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = (int)(threadIdx.x & 31u), warp = (int)(threadIdx.x >> 5);
  const int T1 = P.T + 1;
  const long long N = (long long)P.B * T1;
  const int b = (int)((n < N ? n : N - 1) / T1), t = (int)((n < N ? n : N - 1) % T1);
  // independent flag loads first (see calc_diff_kernel)
  const int f_done = done ? done[b] : 0;
  const int f_gate = gate ? gate[b] : 1;
  const int f_cur = cur ? (cur[b] & 1) : 0;
  const bool active = n < N && !f_done && f_gate != 0;
  // the 64 outputs of a node are staged in shared memory (row stride 65: conflict-free) and written out by
  // the whole warp as one contiguous 16 KB block: a thread-per-node store would touch 32 sectors per instruction
  double* stage = smem + warp * COST_STAGE;
  if (active) {
    const size_t buf = (size_t)(cur ? (other ? (f_cur ^ 1) : f_cur) : 0);
    const bool terminal = t == P.T;
    const double* x = xs + ((buf * P.B + b) * T1 + t) * NX;
    const double* u = terminal ? nullptr : us + ((buf * P.B + b) * P.T + t) * NJ;
    {
      // the reference record (496 B) is read late, in the gradient loop, where its first-touch latency was 20 % of
      // this kernel's stall samples: ask for its lines now
      const double* rp = P.refs + (size_t)n * REF_SIZE;
      AGX_PREFETCH(rp); AGX_PREFETCH(rp + 16); AGX_PREFETCH(rp + 32); AGX_PREFETCH(rp + 48); AGX_PREFETCH(rp + REF_SIZE - 1);
      if (u) { AGX_PREFETCH(u); AGX_PREFETCH(u + NJ - 1); }
    }
    const double c = thread_node_cost<DERIV, COL>(model_of(P, b), P.refs + (size_t)n * REF_SIZE, x, u, terminal,
                                                  terminal ? 1.0 : P.dts[t],
                                                  DERIV ? stage + lane * (CREC_SIZE + 1) : nullptr);
    if (out_cost) out_cost[n] = c;
  }
  if (DERIV) {
    double* flags = stage + 32 * (CREC_SIZE + 1);
    flags[lane] = active ? 1.0 : 0.0;
    __syncwarp();
    const long long n0 = n - lane;
#pragma unroll 4
    for (int it = 0; it < CREC_SIZE; ++it) {
      const int idx = it * 32 + lane, node = idx >> 6, k = idx & (CREC_SIZE - 1);
      if (flags[node] != 0.0) crec[(size_t)(n0 + node) * CREC_SIZE + k] = stage[node * (CREC_SIZE + 1) + k];
    }
  }
}

template <bool COL>
__global__ void calc_kernel(Problem P, const double* __restrict__ xs, const double* __restrict__ us,
                            double* __restrict__ out_cost, double* __restrict__ out_xnext) {
  AGX_SMEM(smem);
  AGX_OCTET_SETUP();
  const int T1 = P.T + 1;
  if (ent >= (long long)P.B * T1) return;
  const int b = (int)(ent / T1), t = (int)(ent % T1);
  double* sb = smem + oct_in_cta * OCT_BOARD;
  double* sc = sb + BRD_B;
  const bool terminal = t == P.T;
  LaneDyn d;
  lane_load_state(d, j, xs + (size_t)ent * NX, terminal ? nullptr : us + ((size_t)b * P.T + t) * NJ);
  double c, qn, vn;
  const bool ok = node_calc<COL>(d, j, omask, model_of(P, b), P.refs + (size_t)ent * REF_SIZE,
                                 terminal ? 0.0 : P.dts[t], terminal, sb, sc, &c, &qn, &vn);
  if (!ok) c = nan("");
  if (out_cost && j == 0) out_cost[ent] = c;
  if (out_xnext && j < NJ) {
    out_xnext[(size_t)ent * NX + j] = qn;
    out_xnext[(size_t)ent * NX + NJ + j] = vn;
  }
}

// dense view of the records (problem.calcDiff data: Fx Fu Lx Lu Lxx Lxu Luu); one thread per (node, row)
__global__ void expand_kernel(Problem P, const double* __restrict__ rec, const double* __restrict__ crec,
                              double* out_cost, double* out_xnext,
                              double* Fx, double* Fu, double* Lx, double* Lu, double* Lxx, double* Lxu, double* Luu) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = P.T + 1;
  const long long n = gid / NX;
  const int r = (int)(gid % NX);
  if (n >= (long long)P.B * T1) return;
  const int t = (int)(n % T1);
  const bool terminal = t == P.T;
  const double h = terminal ? 0.0 : P.dts[t];
  const double* R = rec + (size_t)n * REC_SIZE;
  const double* C = crec + (size_t)n * CREC_SIZE;
  const int i = r % NJ;        // joint row inside the q / v block
  const bool top = r < NJ;     // q rows
  if (out_cost && r == 0) out_cost[n] = C[CK_COST];
  if (out_xnext) out_xnext[n * NX + r] = R[(top ? RK_QN : RK_VN) * 8 + i];
  if (Lx) Lx[n * NX + r] = C[(top ? CK_LQ : CK_LV) + i];
  if (Lu && top) Lu[n * NJ + i] = C[CK_LU + i];
#pragma unroll
  for (int c = 0; c < NJ; ++c) {
    const double aq = R[(RK_AQ + i) * 8 + c], av = R[(RK_AV + i) * 8 + c], mi = R[(RK_MI + i) * 8 + c];
    const double s = top ? h : 1.0;
    if (Fx) {
      Fx[(n * NX + r) * NX + c] = s * aq + ((top && c == i) ? 1.0 : 0.0);
      Fx[(n * NX + r) * NX + NJ + c] =
          terminal ? ((!top && c == i) ? 1.0 : 0.0) : s * (av + ((c == i) ? 1.0 : 0.0));
    }
    if (Fu) Fu[(n * NX + r) * NJ + c] = s * mi;
    if (Lxx) {
      Lxx[(n * NX + r) * NX + c] = top ? C[CK_LQQ + (i >= c ? lidx(i, c) : lidx(c, i))] : 0.0;
      Lxx[(n * NX + r) * NX + NJ + c] = (!top && c == i) ? C[CK_LVV + i] : 0.0;
    }
    if (Lxu) Lxu[(n * NX + r) * NJ + c] = 0.0;
    if (Luu && top) Luu[(n * NJ + i) * NJ + c] = (c == i) ? C[CK_LUU + i] : 0.0;
  }
}

// ---------------------------------------------------------------------------------------------
// Riccati sweep.  Lane j owns columns j and j+7 of every 14-wide matrix.  With
//   G = [dt aq, I + dt av] (7x14), S = [dt I; I] (14x7), N = dt Minv:  Fx = [I 0; 0 0] + S G, Fu = S N
// the products collapse to 7-deep ones:  Z = S^T V', Vs = Z S, W = Vs G + [Z_q 0],
//   Qxx = Lxx + [V'_qq 0; 0 0] + G^T W + [Z_q^T G; 0],  Qux = N^T W,  Quu = Luu + N^T Vs N,
//   Qx = Lx + [v'_q; 0] + G^T S^T v',  Qu = Lu + N^T S^T v'.
constexpr int FW_BOARD = OCT_BOARD + 1 + 16 + 2 * 98;  // forward kernels: node boards + dx[14] + two staged gain blocks (even: 16-byte aligned octet boards)
// shared-memory board of one octet (doubles).  Regions that are never live together share storage.
constexpr int BW_VS = 0;      // [7][8]   Vs ........ later the Quu columns handed to the factorisation
constexpr int BW_L = 0;
constexpr int BW_G = 56;      // [7][16]  G ......... later Qux
constexpr int BW_QUX = 56;
constexpr int BW_ZQ = 168;    // [7][8]
constexpr int BW_N = 224;     // [7][8]
constexpr int BW_SV = 280;    // [8]
constexpr int BW_QU = 288;    // [8]
constexpr int BW_FS = 296;    // [16]
constexpr int BW_V = 312;     // [14][16] Qxx, then the unsymmetrised Vxx
constexpr int BW_ST = 536;    // staged records of the NEXT node: dynamics [184], cost [64], gap row [16]
constexpr int ST_REC = BW_ST, ST_CREC = BW_ST + REC_SIZE, ST_FS = BW_ST + REC_SIZE + CREC_SIZE;
constexpr int BW_SIZE = 536 + REC_SIZE + CREC_SIZE + 16 + 8;  // +8 doubles: octets of a half-warp 16 banks apart

// asynchronous copy of node t's records (and gap row) into the stage buffer: 131 chunks of 16 B over 8 lanes
AGX_DEV void stage_node(double* sm, const double* __restrict__ rec, const double* __restrict__ crec,
                        const double* __restrict__ fs, int j) {
  constexpr int NREC = REC_SIZE / 2, NCREC = CREC_SIZE / 2, NFS = NX / 2;
  for (int c = j; c < NREC + NCREC + NFS; c += 8) {
    if (c < NREC) AGX_CP_ASYNC16(sm + ST_REC + 2 * c, rec + 2 * c);
    else if (c < NREC + NCREC) AGX_CP_ASYNC16(sm + ST_CREC + 2 * (c - NREC), crec + 2 * (c - NREC));
    else AGX_CP_ASYNC16(sm + ST_FS + 2 * (c - NREC - NCREC), fs + 2 * (c - NREC - NCREC));
  }
  AGX_CP_ASYNC_COMMIT();
}

__global__ void backward_kernel(Problem P, Work W, SolverState S, FddpOpts O) {
  AGX_SMEM(smem);
  AGX_OCTET_SETUP();
  const int b = (int)ent;
  if (b >= P.B) return;
  if (S.done[b] || S.pending[b]) return;  // pending: the candidate did not change, its sweep is still valid
  double* sm = smem + oct_in_cta * BW_SIZE;
  const int T = P.T, T1 = T + 1;
  const bool live = j < NJ;
  const int jj = live ? j : 0;
  const size_t buf = buf_of(S.cur, b, false);
  const double* xs = W.xs + (buf * P.B + b) * (size_t)T1 * NX;
  const double* rec0 = W.rec + (size_t)b * T1 * REC_SIZE;
  const double* crec0 = W.crec + (size_t)b * T1 * CREC_SIZE;
  double* fsb = W.fs + (size_t)b * T1 * NX;
  double* gvb = W.gv + (size_t)b * T1 * NX;
  double* Kb = W.K + (size_t)b * T * NJ * NX;
  double* kb = W.k + (size_t)b * T * NJ;
  const bool feasible = S.is_feasible[b] != 0;
  double xreg = S.xreg[b];

  // total cost of the candidate and the gaps (SolverAbstract::computeDynamicFeasibility)
  double cost = 0.0;
  {
    double part = 0.0;
    for (int t = j; t <= T; t += 8) part += crec0[(size_t)t * CREC_SIZE + CK_COST];
    cost = octet_sum(part, omask);
  }
  if (!feasible && live) {
    fsb[j] = W.x0[(size_t)b * NX + j] - xs[j];
    fsb[NJ + j] = W.x0[(size_t)b * NX + NJ + j] - xs[NJ + j];
    for (int t = 0; t < T; ++t) {
      const double* R = rec0 + (size_t)t * REC_SIZE;
      fsb[(t + 1) * NX + j] = R[RK_QN * 8 + j] - xs[(t + 1) * NX + j];
      fsb[(t + 1) * NX + NJ + j] = R[RK_VN * 8 + j] - xs[(t + 1) * NX + NJ + j];
    }
  }
  AGX_OSYNC();

  bool failed = !(cost == cost) ? true : false;  // a NaN node cost marks a failed calcDiff
  double dg = 0.0, dq = 0.0;
  for (;;) {
    bool ok = !failed;
    double V0[NX], V1[NX], vx0, vx1;
    double dgp = 0.0, dqp = 0.0;  // per-lane partial sums
    if (ok) {
      // the first running node's records start flowing into shared memory while the terminal node is handled
      stage_node(sm, rec0 + (size_t)(T - 1) * REC_SIZE, crec0 + (size_t)(T - 1) * CREC_SIZE, fsb + (size_t)(T - 1) * NX, j);
      // ---- terminal node
      const double* C = crec0 + (size_t)T * CREC_SIZE;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        V0[i] = live ? C[CK_LQQ + (i >= jj ? lidx(i, jj) : lidx(jj, i))] : 0.0;
        V0[NJ + i] = 0.0;
        V1[i] = 0.0;
        V1[NJ + i] = 0.0;
      }
      const double lvv = live ? C[CK_LVV + jj] : 0.0;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        if (i == j) { V0[i] += xreg; V1[NJ + i] = lvv + xreg; }
      }
      vx0 = live ? C[CK_LQ + jj] : 0.0;
      vx1 = live ? C[CK_LV + jj] : 0.0;
      if (!feasible) {
        if (live) { sm[BW_FS + j] = fsb[T * NX + j]; sm[BW_FS + NJ + j] = fsb[T * NX + NJ + j]; }
        AGX_OSYNC();
        double g0 = 0.0, g1 = 0.0;
#pragma unroll
        for (int r = 0; r < NX; ++r) { g0 += V0[r] * sm[BW_FS + r]; g1 += V1[r] * sm[BW_FS + r]; }
        vx0 += g0; vx1 += g1;
        if (live) {
          gvb[T * NX + j] = g0; gvb[T * NX + NJ + j] = g1;
          dgp -= vx0 * sm[BW_FS + j] + vx1 * sm[BW_FS + NJ + j];
          dqp += g0 * sm[BW_FS + j] + g1 * sm[BW_FS + NJ + j];
        }
        AGX_OSYNC();
      }
    }
    // ---- running nodes
    for (int t = T - 1; ok && t >= 0; --t) {
      const double h = P.dts[t];
      // node t's records were staged during the previous node: wait, copy this lane's slice to registers,
      // then let the next node's records flow in behind the arithmetic
      AGX_CP_ASYNC_WAIT_ALL();
      AGX_OSYNC();
      double G0[NJ], G1[NJ], Nc[NJ], Lqq[NJ];
      {
        const double* R = sm + ST_REC;
        const double* C = sm + ST_CREC;
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
          G0[i] = live ? R[(RK_AQ + i) * 8 + jj] : 0.0;
          G1[i] = (live ? R[(RK_AV + i) * 8 + jj] : 0.0) + ((i == j) ? 1.0 : 0.0);
          Nc[i] = live ? R[(RK_MI + i) * 8 + jj] : 0.0;
          Lqq[i] = live ? C[CK_LQQ + (i >= jj ? lidx(i, jj) : lidx(jj, i))] : 0.0;
        }
      }
      const double lvv = live ? sm[ST_CREC + CK_LVV + jj] : 0.0, luu = live ? sm[ST_CREC + CK_LUU + jj] : 0.0;
      const double lq = live ? sm[ST_CREC + CK_LQ + jj] : 0.0, lv = live ? sm[ST_CREC + CK_LV + jj] : 0.0;
      const double lu = live ? sm[ST_CREC + CK_LU + jj] : 0.0;
      const double fs0 = live ? sm[ST_FS + jj] : 0.0, fs1 = live ? sm[ST_FS + NJ + jj] : 0.0;
      AGX_OSYNC();
      if (t > 0)
        stage_node(sm, rec0 + (size_t)(t - 1) * REC_SIZE, crec0 + (size_t)(t - 1) * CREC_SIZE, fsb + (size_t)(t - 1) * NX, j);
      double Z0[NJ], Z1[NJ];
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        Z0[i] = h * V0[i] + V0[NJ + i];
        Z1[i] = h * V1[i] + V1[NJ + i];
      }
      const double sv = h * vx0 + vx1;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        sm[BW_VS + i * 8 + j] = h * Z0[i] + Z1[i];
        sm[BW_G + i * 16 + j] = G0[i];
        sm[BW_G + i * 16 + 8 + j] = G1[i];
        sm[BW_ZQ + i * 8 + j] = Z0[i];
        sm[BW_N + i * 8 + j] = Nc[i];
      }
      sm[BW_SV + j] = sv;
      if (!feasible && live) { sm[BW_FS + j] = fs0; sm[BW_FS + NJ + j] = fs1; }
      AGX_OSYNC();
      // W = Vs G + [Zq 0] ; VN = Vs N
      double W0[NJ], W1[NJ], VN[NJ];
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        double a0 = Z0[i], a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int m = 0; m < NJ; ++m) {
          const double vs = sm[BW_VS + i * 8 + m];
          a0 += vs * G0[m];
          a1 += vs * G1[m];
          a2 += vs * Nc[m];
        }
        W0[i] = a0; W1[i] = a1; VN[i] = a2;
      }
      // Qxx columns = Lxx + [V'qq 0; 0 0] + G^T W + [Zq^T G; 0]; parked on the board (their registers are needed
      // by the factorisation and the gains)
#pragma unroll
      for (int r = 0; r < NJ; ++r) {
        double a0 = Lqq[r] + V0[r], a1 = 0.0;
        double b0 = 0.0, b1 = (r == j) ? lvv : 0.0;
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
          const double gq = sm[BW_G + i * 16 + r], gv = sm[BW_G + i * 16 + 8 + r], zq = sm[BW_ZQ + i * 8 + r];
          a0 += gq * W0[i] + zq * G0[i];
          a1 += gq * W1[i] + zq * G1[i];
          b0 += gv * W0[i];
          b1 += gv * W1[i];
        }
        sm[BW_V + r * 16 + j] = a0;
        sm[BW_V + r * 16 + 8 + j] = a1;
        sm[BW_V + (NJ + r) * 16 + j] = b0;
        sm[BW_V + (NJ + r) * 16 + 8 + j] = b1;
      }
      // Qux columns, Quu column, Qx, Qu
      double U0[NJ], U1[NJ], Quu[NJ];
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int m = 0; m < NJ; ++m) {
          const double n = sm[BW_N + m * 8 + i];
          a0 += n * W0[m];
          a1 += n * W1[m];
          a2 += n * VN[m];
        }
        U0[i] = a0; U1[i] = a1;
        Quu[i] = a2 + ((i == j) ? (luu + xreg) : 0.0);
      }
      double qx0 = lq + vx0, qx1 = lv, qu = lu;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const double s = sm[BW_SV + i];
        qx0 += G0[i] * s;
        qx1 += G1[i] * s;
        qu += Nc[i] * s;
      }
      if (!live) {
#pragma unroll
        for (int i = 0; i < NJ; ++i) Quu[i] = (i == 0) ? 1.0 : 0.0;
      }
      AGX_OSYNC();  // all reads of VS / G / ZQ / N done: their storage is reused for Quu and Qux
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        sm[BW_QUX + i * 16 + j] = U0[i];
        sm[BW_QUX + i * 16 + 8 + j] = U1[i];
        sm[BW_L + i * 8 + j] = Quu[i];
      }
      sm[BW_QU + j] = qu;
      AGX_OSYNC();
      // computeGains: every lane factors Quu in registers (no barrier inside the factorisation)
      double K0[NJ], K1[NJ], kk[NJ];
      {
        double L[28], rinv[NJ];
        ok = chol7_registers(sm + BW_L, L, rinv);
        if (!ok) break;
#pragma unroll
        for (int i = 0; i < NJ; ++i) { K0[i] = U0[i]; K1[i] = U1[i]; kk[i] = sm[BW_QU + i]; }
        chol_solve7(L, rinv, K0);
        chol_solve7(L, rinv, K1);
        chol_solve7(L, rinv, kk);
      }
      // value function: Vx = Qx - K^T Qu ; Vxx = Qxx - Qxu K (updated in place on the board)
      double nvx0 = qx0, nvx1 = qx1;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const double qui = sm[BW_QU + i];
        nvx0 -= K0[i] * qui;
        nvx1 -= K1[i] * qui;
      }
      double Q0[NX], Q1[NX];
#pragma unroll
      for (int r = 0; r < NJ; ++r) {
        double a0 = sm[BW_V + r * 16 + j], a1 = sm[BW_V + r * 16 + 8 + j];
        double b0 = sm[BW_V + (NJ + r) * 16 + j], b1 = sm[BW_V + (NJ + r) * 16 + 8 + j];
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
          const double xq = sm[BW_QUX + i * 16 + r], xv = sm[BW_QUX + i * 16 + 8 + r];
          a0 -= xq * K0[i];
          a1 -= xq * K1[i];
          b0 -= xv * K0[i];
          b1 -= xv * K1[i];
        }
        sm[BW_V + r * 16 + j] = a0;
        sm[BW_V + r * 16 + 8 + j] = a1;
        sm[BW_V + (NJ + r) * 16 + j] = b0;
        sm[BW_V + (NJ + r) * 16 + 8 + j] = b1;
        Q0[r] = a0; Q1[r] = a1; Q0[NJ + r] = b0; Q1[NJ + r] = b1;
      }
      AGX_OSYNC();
      // symmetrise (row c of the unsymmetrised matrix is read back from the board)
#pragma unroll
      for (int r = 0; r < NJ; ++r) {
        V0[r] = 0.5 * (Q0[r] + sm[BW_V + jj * 16 + r]);
        V0[NJ + r] = 0.5 * (Q0[NJ + r] + sm[BW_V + jj * 16 + 8 + r]);
        V1[r] = 0.5 * (Q1[r] + sm[BW_V + (NJ + jj) * 16 + r]);
        V1[NJ + r] = 0.5 * (Q1[NJ + r] + sm[BW_V + (NJ + jj) * 16 + 8 + r]);
      }
#pragma unroll
      for (int i = 0; i < NJ; ++i)
        if (i == j) { V0[i] += xreg; V1[NJ + i] += xreg; }
      if (!live) {
#pragma unroll
        for (int r = 0; r < NX; ++r) { V0[r] = 0.0; V1[r] = 0.0; }
      }
      vx0 = nvx0; vx1 = nvx1;
      // expected improvement pieces and gap terms
      if (live) {
        double quuk = 0.0;
#pragma unroll
        for (int i = 0; i < NJ; ++i) quuk += sm[BW_L + i * 8 + j] * kk[i];
        double kj = 0.0;
#pragma unroll
        for (int i = 0; i < NJ; ++i)
          if (i == j) kj = kk[i];
        dgp += qu * kj;
        dqp -= kj * quuk;
        kb[t * NJ + j] = kj;
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
          Kb[(t * NJ + i) * NX + j] = K0[i];
          Kb[(t * NJ + i) * NX + NJ + j] = K1[i];
        }
      }
      if (!feasible) {
        double g0 = 0.0, g1 = 0.0;
#pragma unroll
        for (int r = 0; r < NX; ++r) { g0 += V0[r] * sm[BW_FS + r]; g1 += V1[r] * sm[BW_FS + r]; }
        vx0 += g0; vx1 += g1;
        if (live) {
          gvb[t * NX + j] = g0; gvb[t * NX + NJ + j] = g1;
          dgp -= vx0 * sm[BW_FS + j] + vx1 * sm[BW_FS + NJ + j];
          dqp += g0 * sm[BW_FS + j] + g1 * sm[BW_FS + NJ + j];
        }
      }
      AGX_OSYNC();
    }
    AGX_CP_ASYNC_WAIT_ALL();  // nothing may still be in flight when the sweep is abandoned or restarted
    AGX_OSYNC();
    if (ok) {
      // non-finite value function = failed sweep (SolverDDP::backwardPass raises on NaN)
      double chk = vx0 + vx1;
#pragma unroll
      for (int r = 0; r < NX; ++r) chk += V0[r] + V1[r];
      chk = octet_sum(live ? chk : 0.0, omask);
      if (!(chk - chk == 0.0)) ok = false;
    }
    if (ok) {
      dg = octet_sum(dgp, omask);
      dq = octet_sum(dqp, omask);
      break;
    }
    // increaseRegularization and retry without recalc
    failed = false;
    xreg *= O.reg_incfactor;
    if (xreg > O.reg_max) xreg = O.reg_max;
    if (xreg == O.reg_max) {
      if (j == 0) { S.status[b] = 2; S.done[b] = 1; }
      break;
    }
  }
  if (j == 0) {
    S.xreg[b] = xreg;
    S.cost[b] = cost;
    S.dg[b] = dg;
    S.dq[b] = dq;
  }
}

// ---------------------------------------------------------------------------------------------
}  // namespace agx
#include "agx_riccati_mma.cuh"
#include "agx_sqp.cuh"
namespace agx {

// Forward pass of one FDDP iteration, split so that the common case costs little:
//   rollout_try_kernel  (octet per problem)  nonlinear rollout with alpha = 1, dynamics only
//   node_cost_kernel    (thread per node)    costs (+ their derivatives) of the trial trajectory
//   accept_linesearch_kernel (octet per problem) dV / dVexp acceptance test, buffer flip, reg / stop logic;
//                                            only if the alpha = 1 trial is rejected: alpha = 1/2, 1/4, ...
//                                            with the costs evaluated in line
// (SolverFDDP::forwardPass / tryStep / expectedImprovement and the tail of the solve loop).

// end-of-iteration bookkeeping shared by accept_kernel and linesearch_kernel (one thread per problem)
AGX_DEV void finish_iteration(const SolverState& S, const FddpOpts& O, int b, bool accepted, double steplength,
                              bool feasible, double cost_try, int obuf, bool fast_path) {
  bool was_feasible = S.was_feasible[b] != 0;
  if (accepted) {
    was_feasible = feasible;
    S.was_feasible[b] = feasible ? 1 : 0;
    S.is_feasible[b] = (feasible || steplength == 1.0) ? 1 : 0;
    S.cost[b] = cost_try;
    S.cur[b] = (int32_t)obuf;
    S.recalc[b] = 1;
    S.recalc_cost[b] = fast_path ? 0 : 1;  // the fast path already wrote the cost records of the new candidate
  } else {
    S.recalc[b] = 0;
    S.recalc_cost[b] = 1;  // the trial's cost records overwrote the candidate's
  }
  double xreg = S.xreg[b];
  int status = 1, done = 0;
  if (steplength > O.th_stepdec) {
    xreg /= O.reg_decfactor;
    if (xreg < O.reg_min) xreg = O.reg_min;
  }
  if (steplength <= O.th_stepinc) {
    xreg *= O.reg_incfactor;
    if (xreg > O.reg_max) xreg = O.reg_max;
    if (xreg == O.reg_max) { status = 2; done = 1; }
  }
  S.xreg[b] = xreg;
  S.iters[b] += 1;
  if (!done && !O.fixed_iters && was_feasible && S.stop[b] < O.th_stop) { status = 0; done = 1; }
  if (!done && S.iters[b] >= O.max_iter) done = 1;  // budget used (status stays MAXITER)
  // max_solve_time (ocp_base_croco.py:70-71): checked at the end of an iteration, as the reference's solver does
  if (!done && O.max_solve_ns > 0 && agx_now_ns() - *S.t0 > O.max_solve_ns) { status = 5; done = 1; }
  if (done) { S.status[b] = status; S.done[b] = 1; }
}

// dV / dVexp test of SolverFDDP::solve
// (accept_rule: 0 = Crocoddyl >= 2.0 — |d1| and no negative-expectation step from a feasible candidate; 1 = 1.x)
AGX_DEV bool accept_step(const FddpOpts& O, double dV, double d1, double dVexp, bool feasible) {
  const bool legacy = O.accept_rule != 0;
  if (dVexp >= 0.0) return (legacy ? d1 : fabs(d1)) < O.th_grad || dV > O.th_acceptstep * dVexp;
  return (legacy || !feasible) && dV > O.th_acceptnegstep * dVexp;
}

__global__ void rollout_try_kernel(Problem P, Work W, SolverState S) {
  AGX_SMEM(smem);
  const int j = (int)(threadIdx.x & 7u);
  const int oct_in_cta = (int)(threadIdx.x >> 3);
  const int b = (int)blockIdx.x * (int)(blockDim.x >> 3) + oct_in_cta;
  // whole-warp collectives (see calc_diff_kernel): every octet of a warp walks the same T nodes, finished
  // problems exit before the first collective
  const unsigned omask = 0xffffffffu;
  if (b >= P.B) return;
  if (S.done[b]) return;
  double* sb = smem + oct_in_cta * FW_BOARD;
  double* sc = sb + BRD_B;
  double* sdx = sc + BRD_C;  // [14]
  double* sK = sdx + 16;     // [2][7][14] gain blocks of the current / next node
  const int T = P.T, T1 = T + 1;
  const bool live = j < NJ;
  const int jj = live ? j : 0;
  const double* model = model_of(P, b);
  const size_t buf = buf_of(S.cur, b, false), obuf = buf ^ 1;
  const double* xs = W.xs + (buf * P.B + b) * (size_t)T1 * NX;
  const double* us = W.us + (buf * P.B + b) * (size_t)T * NJ;
  double* xt = W.xs + (obuf * P.B + b) * (size_t)T1 * NX;
  double* ut = W.us + (obuf * P.B + b) * (size_t)T * NJ;
  const double* gvb = W.gv + (size_t)b * T1 * NX;
  const double* Kb = W.K + (size_t)b * T * NJ * NX;
  const double* kb = W.k + (size_t)b * T * NJ;
  const bool feasible = S.is_feasible[b] != 0;
  // a problem whose alpha = 1 trial was rejected in the previous round tries alpha = 1/2 here (deferred line search)
  const double alpha = ldexp(1.0, -S.pending[b]);  // pending = n: the search stands at step length 2^-n
  const double* fsb = W.fs + (size_t)b * T1 * NX;
  // per-node inputs are fetched one node ahead (scalars into registers, gain rows into L1)
  struct NodeIn { double us, kff, dt, xsq, xsv, gq, gv, fq, fv; };
  auto fetch = [&](int t, NodeIn& n) {
    const bool run = t < T;
    n.us = (live && run) ? us[t * NJ + jj] : 0.0;
    n.kff = (live && run) ? kb[t * NJ + jj] : 0.0;
    n.dt = run ? P.dts[t] : 0.0;
    n.xsq = live ? xs[t * NX + jj] : 0.0;
    n.xsv = live ? xs[t * NX + NJ + jj] : 0.0;
    const bool gaps = live && !feasible;
    n.gq = gaps ? gvb[t * NX + jj] : 0.0;
    n.gv = gaps ? gvb[t * NX + NJ + jj] : 0.0;
    const bool contract = gaps && alpha != 1.0;  // xs_try = xhat + (alpha - 1) fs
    n.fq = contract ? fsb[t * NX + jj] : 0.0;
    n.fv = contract ? fsb[t * NX + NJ + jj] : 0.0;
    // the node's 7x14 gain block (784 B) is copied global -> shared asynchronously: 49 chunks of 16 B over 8 lanes
    if (run) {
      double* dst = sK + (t & 1) * 98;
      const double* srck = Kb + (size_t)t * NJ * NX;
      for (int c = j; c < 49; c += 8) AGX_CP_ASYNC16(dst + 2 * c, srck + 2 * c);
    }
    AGX_CP_ASYNC_COMMIT();
  };
  double xq = live ? W.x0[(size_t)b * NX + jj] : 0.0, xv = live ? W.x0[(size_t)b * NX + NJ + jj] : 0.0;
  double dvp = 0.0;
  bool ok = true;
  NodeIn cur;
  fetch(0, cur);
  for (int t = 0; t <= T; ++t) {
    NodeIn nxt;
    AGX_CP_ASYNC_WAIT_ALL();  // node t's gain block has landed (it was issued one node ago)
    if (t < T) fetch(t + 1, nxt);
    if (alpha != 1.0) {
      xq += cur.fq * (alpha - 1.0);
      xv += cur.fv * (alpha - 1.0);
    }
    const double dxq = live ? xq - cur.xsq : 0.0, dxv = live ? xv - cur.xsv : 0.0;
    if (live) {
      xt[t * NX + j] = xq;
      xt[t * NX + NJ + j] = xv;
    }
    dvp += cur.gq * dxq + cur.gv * dxv;
    if (t == T) break;
    if (live) { sdx[j] = dxq; sdx[NJ + j] = dxv; }
    AGX_OSYNC();
    double s = 0.0;
#pragma unroll
    for (int m = 0; m < NX; ++m) s += sK[(t & 1) * 98 + jj * NX + m] * sdx[m];
    LaneDyn d;
    d.q = xq; d.qd = xv;
    d.u = live ? cur.us - cur.kff * alpha - s : 0.0;
    if (live) ut[t * NJ + j] = d.u;
    node_kinematics(d, j, omask, model);
    double L[28], rinv[NJ];
    const bool okn = node_forward_dynamics<false>(d, j, omask, model, sb, sc, L, rinv);
    ok = ok && okn;
    xq = d.q + (d.qd * cur.dt + d.qdd * (cur.dt * cur.dt));
    xv = d.qd + d.qdd * cur.dt;
    cur = nxt;
    AGX_OSYNC();
  }
  const double dv = feasible ? 0.0 : octet_sum(dvp, omask);
  if (j == 0) {
    S.dv[b] = dv;
    S.roll_ok[b] = ok ? 1 : 0;
  }
}

// The same forward pass with TWO warps on every group of four problems.  A node's forward dynamics is one long
// dependent chain; its two halves — the mass matrix with its factorisation, and the bias forces — only share the
// kinematics.  Warp 0 (role M) runs kinematics, composite inertias, mass matrix, Cholesky; warp 1 (role B) runs the
// feedback control, kinematics, velocities, bias forces.  They meet twice per node through shared memory: B hands over
// the control and the bias torques, M solves, integrates and hands back the next state.  Same arithmetic per quantity
// as rollout_try_kernel.  CTA = 64 threads = 4 problems; shared memory per problem: FW_BOARD doubles (boards, dx, gain
// blocks) + 48 (exchange: next state [2][8], control [8], bias [8], spare).
constexpr int FW2_BOARD = FW_BOARD + 48;
__global__ void __launch_bounds__(64) rollout_try2_kernel(Problem P, Work W, SolverState S) {
  AGX_SMEM(smem);
  const int j = (int)(threadIdx.x & 7u);
  const int oct = (int)((threadIdx.x >> 3) & 3u);
  const int role = (int)(threadIdx.x >> 5);  // 0: mass matrix and solve, 1: control and bias forces
  const int b = (int)blockIdx.x * 4 + oct;
  const unsigned omask = 0xffffffffu;        // whole-warp collectives: the four octets of a warp run one stream
  const bool valid = b < P.B && !S.done[b < P.B ? b : 0];
  {
    // nothing to do for the whole CTA: leave (decided by every thread from the same four flags: no barrier needed)
    bool any = false;
    for (int k = 0; k < 4; ++k) {
      const int bk = (int)blockIdx.x * 4 + k;
      any = any || (bk < P.B && !S.done[bk]);
    }
    if (!any) return;
  }
  const int bb = valid ? b : 0;               // idle octets walk along on problem 0's data and write nothing
  double* sb = smem + oct * FW2_BOARD;
  double* sc = sb + BRD_B;
  double* sdx = sc + BRD_C;   // [14]
  double* sK = sdx + 16;      // [2][7][14]
  double* sx = sK + 2 * 98;   // exchange: [0..15] next state (q 0..7, v 8..15), [16..23] control, [24..31] bias torque
  const int T = P.T, T1 = T + 1;
  const bool live = j < NJ;
  const int jj = live ? j : 0;
  const double* model = model_of(P, bb);
  const size_t buf = buf_of(S.cur, bb, false), obuf = buf ^ 1;
  const double* xs = W.xs + (buf * P.B + bb) * (size_t)T1 * NX;
  const double* us = W.us + (buf * P.B + bb) * (size_t)T * NJ;
  double* xt = W.xs + (obuf * P.B + bb) * (size_t)T1 * NX;
  double* ut = W.us + (obuf * P.B + bb) * (size_t)T * NJ;
  const double* gvb = W.gv + (size_t)bb * T1 * NX;
  const double* fsb = W.fs + (size_t)bb * T1 * NX;
  const double* Kb = W.K + (size_t)bb * T * NJ * NX;
  const double* kb = W.k + (size_t)bb * T * NJ;
  const bool feasible = S.is_feasible[bb] != 0;
  const double alpha = ldexp(1.0, -S.pending[bb]);
  const bool contract = !feasible && alpha != 1.0;
  double xq = live ? W.x0[(size_t)bb * NX + jj] : 0.0, xv = live ? W.x0[(size_t)bb * NX + NJ + jj] : 0.0;
  bool ok = true;
  if (role == 1) {
    // ---------------------------------------------------------------- role B: control, kinematics, bias forces
    struct NodeIn { double us, kff, xsq, xsv, gq, gv, fq, fv; };
    auto fetch = [&](int t, NodeIn& n) {
      const bool run = t < T;
      n.us = (live && run) ? us[t * NJ + jj] : 0.0;
      n.kff = (live && run) ? kb[t * NJ + jj] : 0.0;
      n.xsq = live ? xs[t * NX + jj] : 0.0;
      n.xsv = live ? xs[t * NX + NJ + jj] : 0.0;
      const bool gaps = live && !feasible;
      n.gq = gaps ? gvb[t * NX + jj] : 0.0;
      n.gv = gaps ? gvb[t * NX + NJ + jj] : 0.0;
      n.fq = (live && contract) ? fsb[t * NX + jj] : 0.0;
      n.fv = (live && contract) ? fsb[t * NX + NJ + jj] : 0.0;
      if (run) {
        double* dst = sK + (t & 1) * 98;
        const double* srck = Kb + (size_t)t * NJ * NX;
        for (int c = j; c < 49; c += 8) AGX_CP_ASYNC16(dst + 2 * c, srck + 2 * c);
      }
      AGX_CP_ASYNC_COMMIT();
    };
    double dvp = 0.0;
    NodeIn cur;
    fetch(0, cur);
    const double zero6[6] = {0, 0, 0, 0, 0, 0};
    const double agrav[6] = {-model[MT_GRAV + 0], -model[MT_GRAV + 1], -model[MT_GRAV + 2], 0, 0, 0};
    for (int t = 0; t <= T; ++t) {
      NodeIn nxt;
      AGX_CP_ASYNC_WAIT_ALL();
      if (t < T) fetch(t + 1, nxt);
      if (contract) { xq += cur.fq * (alpha - 1.0); xv += cur.fv * (alpha - 1.0); }
      const double dxq = live ? xq - cur.xsq : 0.0, dxv = live ? xv - cur.xsv : 0.0;
      if (live && valid) { xt[t * NX + j] = xq; xt[t * NX + NJ + j] = xv; }
      dvp += cur.gq * dxq + cur.gv * dxv;
      if (t == T) break;
      if (live) { sdx[j] = dxq; sdx[NJ + j] = dxv; }
      __syncwarp();
      double s = 0.0;
#pragma unroll
      for (int m = 0; m < NX; ++m) s += sK[(t & 1) * 98 + jj * NX + m] * sdx[m];
      LaneDyn d;
      d.q = xq; d.qd = xv;
      d.u = live ? cur.us - cur.kff * alpha - s : 0.0;
      if (live && valid) ut[t * NJ + j] = d.u;
      node_kinematics(d, j, omask, model);
      scan_prefix_excl<6>(d.s, d.vp, zero6, j, omask);
      body_terms(d, j, model, nullptr, false);
      scan_prefix_excl<6>(d.g, d.a0p, agrav, j, omask);
      body_force(d);
      scan_suffix_incl<6>(d.Z + 22, j, omask);
      sx[16 + j] = d.u;
      sx[24 + j] = dot6(d.J, d.Z + 22);
      __syncthreads();   // (1) control and bias torques are on the board
      __syncthreads();   // (2) role M has written the next state
      xq = live ? sx[j] : 0.0;
      xv = live ? sx[8 + j] : 0.0;
      cur = nxt;
    }
    const double dv = feasible ? 0.0 : octet_sum(dvp, omask);
    if (j == 0 && valid) S.dv[b] = dv;
  } else {
    // ---------------------------------------------------------------- role M: mass matrix, factorisation, solve
    for (int t = 0; t < T; ++t) {
      if (contract && live) { xq += fsb[t * NX + jj] * (alpha - 1.0); xv += fsb[t * NX + NJ + jj] * (alpha - 1.0); }
      const double dt = P.dts[t];
      LaneDyn d;
      d.q = xq; d.qd = xv; d.u = 0.0;
      node_kinematics(d, j, omask, model);
      body_inertia(d, j, model);
      scan_suffix_incl<10>(d.Z, j, omask);
      inertia_apply(d.Z, d.J, d.dFda);
      double* o = sb + j * 18;
#pragma unroll
      for (int k = 0; k < 6; ++k) { o[k] = d.J[k]; o[6 + k] = d.dFda[k]; }
      __syncwarp();
      mass_column(d, j, model, sb);
#pragma unroll
      for (int i = 0; i < NJ; ++i) sc[i * 8 + j] = d.Mc[i];
      __syncwarp();
      double L[28], rinv[NJ];
      const bool okn = chol7_registers(sc, L, rinv);
      ok = ok && okn;
      __syncthreads();   // (1) wait for the control and the bias torques
      double rhs[NJ];
#pragma unroll
      for (int i = 0; i < NJ; ++i) rhs[i] = sx[16 + i] - sx[24 + i];
      chol_solve7(L, rinv, rhs);
      double qdd = 0.0;
#pragma unroll
      for (int i = 0; i < NJ; ++i)
        if (i == j) qdd = rhs[i];
      xq = d.q + (d.qd * dt + qdd * (dt * dt));
      xv = d.qd + qdd * dt;
      sx[j] = xq;
      sx[8 + j] = xv;
      __syncthreads();   // (2) the next state is on the board
    }
    if (j == 0 && valid) S.roll_ok[b] = ok ? 1 : 0;
  }
}

// Acceptance test of the alpha = 1 trial (its costs come from node_cost_kernel) and, only if it is rejected,
// the remaining step lengths alpha = 2^-ia, ia >= 1, with the costs evaluated in line.
//
// Deferred search (O.defer): a rejected alpha = 1 trial does not start the in-line search — a whole-horizon rollout
// that every other problem of the batch would wait for.  The problem is marked `pending` instead: it sits out the
// next round's calc_diff and sweep (its candidate did not change) and its alpha = 1/2 trial goes through the next
// round's rollout_try / node_cost launches together with everybody else's alpha = 1 trial.  The sequence of operations
// of the problem is SolverFDDP's; only the round in which its alpha = 1/2 trial runs moves.  A problem defers once
// (round - iterations < 1), so one extra round after the budget lets every problem finish its max_iter iterations.
// Generalised: `pending` holds the index n of the step length 2^-n the search stands at, and a problem may defer
// O.defer times over a solve (it is then O.defer rounds behind); the host queues max_iter + O.defer rounds.
template <bool COL>
__global__ void accept_linesearch_kernel(Problem P, Work W, SolverState S, FddpOpts O, int round_host,
                                         const int32_t* __restrict__ round_dev) {
  // the round index comes from the host loop, or from the device counter of the tick graph (agx_api.cu: TickGraph)
  const int round = round_dev ? *round_dev : round_host;
  AGX_SMEM(smem);
  AGX_OCTET_SETUP();
  const int b = (int)ent;
  if (b >= P.B) return;
  if (S.done[b]) return;
  const int pend = S.pending[b];
  const double a1 = ldexp(1.0, -pend);  // the step length the fast path evaluated this round
  {
    const double* crec0 = W.crec + (size_t)b * (P.T + 1) * CREC_SIZE;
    double part = 0.0;
    for (int t = j; t <= P.T; t += 8) part += crec0[(size_t)t * CREC_SIZE + CK_COST];
    const double cost1 = octet_sum(part, omask);
    const bool finite = S.roll_ok[b] != 0 && (cost1 - cost1 == 0.0);
    bool acc1 = false;
    double stop1 = S.stop[b];
    if (finite) {
      const double dv = S.dv[b];
      const double d1 = S.dg[b] + dv, d2 = S.dq[b] - 2.0 * dv;
      stop1 = fabs(d1 + 0.5 * d2);
      acc1 = accept_step(O, S.cost[b] - cost1, d1, a1 * (d1 + 0.5 * a1 * d2), S.is_feasible[b] != 0);
    }
    const bool last_alpha = O.n_alphas <= pend + 1;
    // a problem may run up to O.defer rounds behind the batch (one round per deferred step length)
    const bool defer = !acc1 && !last_alpha && O.defer > 0 && round - S.iters[b] < O.defer;
    AGX_OSYNC();  // every lane has read the state before lane 0 updates it
    if (acc1 || last_alpha) {
      if (j == 0) {
        S.stop[b] = stop1;
        S.pending[b] = 0;
        finish_iteration(S, O, b, acc1, a1, S.is_feasible[b] != 0, cost1, (S.cur[b] & 1) ^ 1, true);
      }
      return;
    }
    if (defer) {
      if (j == 0) {
        S.stop[b] = stop1;
        S.pending[b] = pend + 1;
        S.recalc[b] = 0;       // same candidate: its dynamics records stay valid ...
        S.recalc_cost[b] = 0;  // ... and its cost records are not needed before the search ends
      }
      return;
    }
    if (j == 0) { S.stop[b] = stop1; S.pending[b] = 0; }
    AGX_OSYNC();
  }
  double* sb = smem + oct_in_cta * FW_BOARD;
  double* sc = sb + BRD_B;
  double* sdx = sc + BRD_C;  // [14]
  const int T = P.T, T1 = T + 1;
  const bool live = j < NJ;
  const int jj = live ? j : 0;
  const double* model = model_of(P, b);
  const size_t buf = buf_of(S.cur, b, false), obuf = buf ^ 1;
  const double* xs = W.xs + (buf * P.B + b) * (size_t)T1 * NX;
  const double* us = W.us + (buf * P.B + b) * (size_t)T * NJ;
  double* xt = W.xs + (obuf * P.B + b) * (size_t)T1 * NX;
  double* ut = W.us + (obuf * P.B + b) * (size_t)T * NJ;
  const double* fsb = W.fs + (size_t)b * T1 * NX;
  const double* gvb = W.gv + (size_t)b * T1 * NX;
  const double* Kb = W.K + (size_t)b * T * NJ * NX;
  const double* kb = W.k + (size_t)b * T * NJ;
  const double* refs = P.refs + (size_t)b * T1 * REF_SIZE;
  const bool feasible = S.is_feasible[b] != 0;
  const double cost = S.cost[b], dg = S.dg[b], dq = S.dq[b];
  const double x0q = live ? W.x0[(size_t)b * NX + jj] : 0.0, x0v = live ? W.x0[(size_t)b * NX + NJ + jj] : 0.0;

  double steplength = 1.0, cost_try = 0.0, stop = S.stop[b];
  bool accepted = false;
  for (int ia = pend + 1; ia < O.n_alphas; ++ia) {
    steplength = ldexp(1.0, -ia);
    const bool contract = !feasible;
    double xq = x0q, xv = x0v;
    double ctry = 0.0, dvp = 0.0;
    bool ok = true;
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        // the next node's operands (state, gap, Vxx fs, gain row, control, reference record) are asked for now: this
        // loop is one dependent chain per node and it has no other way to hide their latency
        const int tn = t + 1;
        AGX_PREFETCH(xs + tn * NX + jj); AGX_PREFETCH(xs + tn * NX + NJ + jj);
        if (!feasible) { AGX_PREFETCH(fsb + tn * NX + jj); AGX_PREFETCH(gvb + tn * NX + jj); }
        AGX_PREFETCH(refs + (size_t)tn * REF_SIZE + 8 * j);
        if (tn < T) {
          AGX_PREFETCH(Kb + ((size_t)tn * NJ + jj) * NX); AGX_PREFETCH(Kb + ((size_t)tn * NJ + jj) * NX + NX - 1);
          AGX_PREFETCH(us + tn * NJ + jj); AGX_PREFETCH(kb + tn * NJ + jj);
        }
      }
      double tq = xq, tv = xv;
      if (contract && live) {
        tq += fsb[t * NX + j] * (steplength - 1.0);
        tv += fsb[t * NX + NJ + j] * (steplength - 1.0);
      }
      const double dxq = live ? tq - xs[t * NX + jj] : 0.0, dxv = live ? tv - xs[t * NX + NJ + jj] : 0.0;
      if (live) {
        xt[t * NX + j] = tq;
        xt[t * NX + NJ + j] = tv;
        if (!feasible) dvp += gvb[t * NX + j] * dxq + gvb[t * NX + NJ + j] * dxv;
      }
      LaneDyn d;
      d.q = tq; d.qd = tv; d.u = 0.0;
      const bool terminal = t == T;
      if (!terminal) {
        if (live) { sdx[j] = dxq; sdx[NJ + j] = dxv; }
        AGX_OSYNC();
        if (live) {
          double s = 0.0;
#pragma unroll
          for (int m = 0; m < NX; ++m) s += Kb[(t * NJ + j) * NX + m] * sdx[m];
          d.u = us[t * NJ + j] - kb[t * NJ + j] * steplength - s;
          ut[t * NJ + j] = d.u;
        }
      }
      double c, qn, vn;
      const bool okn = node_calc<COL>(d, j, omask, model, refs + (size_t)t * REF_SIZE, terminal ? 0.0 : P.dts[t],
                                      terminal, sb, sc, &c, &qn, &vn);
      ok = ok && okn;
      ctry += c;
      xq = qn; xv = vn;
      if (!(ctry - ctry == 0.0)) { ok = false; break; }  // NaN / inf: reject this step length
    }
    if (!ok) continue;
    cost_try = ctry;
    const double dv = feasible ? 0.0 : octet_sum(dvp, omask);
    const double dV = cost - cost_try;
    const double d1 = dg + dv, d2 = dq - 2.0 * dv;
    stop = fabs(d1 + 0.5 * d2);
    const double dVexp = steplength * (d1 + 0.5 * steplength * d2);
    accepted = accept_step(O, dV, d1, dVexp, feasible);
    if (accepted) break;
  }
  if (j == 0) {
    S.stop[b] = stop;
    finish_iteration(S, O, b, accepted, steplength, feasible, cost_try, (int)obuf, false);
  }
}

// problem.rollout(us): one octet per problem
__global__ void rollout_kernel(Problem P, const double* __restrict__ x0, const double* __restrict__ us,
                               double* __restrict__ out_xs) {
  AGX_SMEM(smem);
  AGX_OCTET_SETUP();
  const int b = (int)ent;
  if (b >= P.B) return;
  double* sb = smem + oct_in_cta * OCT_BOARD;
  double* sc = sb + BRD_B;
  const int T = P.T, T1 = T + 1;
  const bool live = j < NJ;
  const int jj = live ? j : 0;
  double xq = live ? x0[(size_t)b * NX + jj] : 0.0, xv = live ? x0[(size_t)b * NX + NJ + jj] : 0.0;
  double* xo = out_xs + (size_t)b * T1 * NX;
  if (live) { xo[j] = xq; xo[NJ + j] = xv; }
  for (int t = 0; t < T; ++t) {
    LaneDyn d;
    d.q = xq; d.qd = xv; d.u = live ? us[((size_t)b * T + t) * NJ + jj] : 0.0;
    double c, qn, vn;
    const bool ok = node_calc(d, j, omask, model_of(P, b), P.refs + ((size_t)b * T1 + t) * REF_SIZE, P.dts[t], false,
                              sb, sc, &c, &qn, &vn);
    xq = ok ? qn : nan("");
    xv = ok ? vn : nan("");
    if (live) { xo[(t + 1) * NX + j] = xq; xo[(t + 1) * NX + NJ + j] = xv; }
  }
}

// Reference window: refs[b][t][:] = stream[b or 0][start_b + horizon_idx[t]][:] — the device-side replacement of the
// per-tick horizon extraction (TrajectoryBuffer.horizon, trajectory.py:199-222) and of the per-node setter loop /
// circularAppend of ocp_croco_generic.py:855-892: the whole reference stream stays on the device.
__global__ void gather_refs_kernel(int B, int T1, const double* __restrict__ stream, int n_streams, int n_points,
                                   const int32_t* __restrict__ start, int start0, const int32_t* __restrict__ hidx,
                                   double* __restrict__ refs) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = gid / REF_SIZE;
  const int k = (int)(gid % REF_SIZE);
  if (n >= (long long)B * T1) return;
  const int b = (int)(n / T1), t = (int)(n % T1);
  int p = (start ? start[b] : start0) + hidx[t];
  if (p >= n_points) p = n_points - 1;  // buffer under-run: repeat the last point (agimus_controller.py:492-503)
  if (p < 0) p = 0;
  refs[gid] = stream[((size_t)(n_streams > 1 ? b : 0) * n_points + p) * REF_SIZE + k];
}

// per-cost evaluation: one thread per node -> [state_reg, control_reg, goal_tracking, r6(6), collision cost (2),
// collision distance (2)] (N_COST_TERMS doubles)
template <bool COL>
__global__ void cost_terms_kernel(Problem P, const double* __restrict__ xs, const double* __restrict__ us,
                                  double* __restrict__ out_terms) {
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = P.T + 1;
  if (n >= (long long)P.B * T1) return;
  const int b = (int)(n / T1), t = (int)(n % T1);
  const bool terminal = t == P.T;
  double terms[N_COST_TERMS];
  thread_node_cost<false, COL>(model_of(P, b), P.refs + (size_t)n * REF_SIZE, xs + (size_t)n * NX,
                          terminal ? nullptr : us + ((size_t)b * P.T + t) * NJ, terminal, 1.0, nullptr, terms);
#pragma unroll
  for (int k = 0; k < N_COST_TERMS; ++k) out_terms[(size_t)n * N_COST_TERMS + k] = terms[k];
}

// Warm start by shifting the previous solution by the first time step
// (WarmStartShiftPreviousSolution.shift, warm_start_shift_previous_solution.py:85-104): nodes whose step equals
// dt0 take the next node's state / control, coarser nodes are re-integrated over dt0 with their own control.
__global__ void shift_kernel(Problem P, const double* __restrict__ xs, const double* __restrict__ us,
                             double* __restrict__ out_xs, double* __restrict__ out_us) {
  AGX_SMEM(smem);
  AGX_OCTET_SETUP();
  const int T = P.T, T1 = T + 1;
  if (ent >= (long long)P.B * T1) return;
  const int b = (int)(ent / T1), i = (int)(ent % T1);
  double* sb = smem + oct_in_cta * OCT_BOARD;
  double* sc = sb + BRD_B;
  const bool live = j < NJ;
  const double* xb = xs + (size_t)b * T1 * NX;
  const double* ub = us + (size_t)b * T * NJ;
  double* xo = out_xs + ((size_t)b * T1 + i) * NX;
  if (i == T) {
    if (live) { xo[j] = xb[T * NX + j]; xo[NJ + j] = xb[T * NX + NJ + j]; }
    return;
  }
  double* uo = out_us + ((size_t)b * T + i) * NJ;
  const double dt0 = P.dts[0];
  if (P.dts[i] == dt0) {
    if (live) {
      xo[j] = xb[(i + 1) * NX + j];
      xo[NJ + j] = xb[(i + 1) * NX + NJ + j];
      uo[j] = ub[(i < T - 1 ? i + 1 : i) * NJ + j];
    }
    return;
  }
  LaneDyn d;
  lane_load_state(d, j, xb + (size_t)i * NX, ub + (size_t)i * NJ);
  node_kinematics(d, j, omask, model_of(P, b));
  double L[28], rinv[NJ];
  const bool ok = node_forward_dynamics<false>(d, j, omask, model_of(P, b), sb, sc, L, rinv);
  if (live) {
    xo[j] = ok ? d.q + (d.qd * dt0 + d.qdd * (dt0 * dt0)) : nan("");
    xo[NJ + j] = ok ? d.qd + d.qdd * dt0 : nan("");
    uo[j] = d.u;
  }
}

// IntegratedActionModelEuler.calc -> xnext for n independent (x, u) pairs (costs skipped)
__global__ void integrate_kernel(const double* __restrict__ models, int per_row, const double* __restrict__ x,
                                 const double* __restrict__ u, double dt, int n, double* __restrict__ out) {
  AGX_SMEM(smem);
  AGX_OCTET_SETUP();
  if (ent >= n) return;
  const double* model = models + (per_row ? (size_t)ent * MODEL_SIZE : 0);
  double* sb = smem + oct_in_cta * OCT_BOARD;
  double* sc = sb + BRD_B;
  LaneDyn d;
  lane_load_state(d, j, x + (size_t)ent * NX, u + (size_t)ent * NJ);
  node_kinematics(d, j, omask, model);
  double L[28], rinv[NJ];
  const bool ok = node_forward_dynamics<false>(d, j, omask, model, sb, sc, L, rinv);
  if (j < NJ) {
    out[(size_t)ent * NX + j] = ok ? d.q + (d.qd * dt + d.qdd * (dt * dt)) : nan("");
    out[(size_t)ent * NX + NJ + j] = ok ? d.qd + d.qdd * dt : nan("");
  }
}

// pin.rnea(q, v, a) for n independent triples: tau = nle(q, v) + M(q) a (no armature)
__global__ void rnea_kernel(const double* __restrict__ models, int per_row, const double* __restrict__ q,
                            const double* __restrict__ v, const double* __restrict__ a, int n,
                            double* __restrict__ out_tau) {
  AGX_SMEM(smem);
  AGX_OCTET_SETUP();
  if (ent >= n) return;
  const double* model = models + (per_row ? (size_t)ent * MODEL_SIZE : 0);
  double* sb = smem + oct_in_cta * OCT_BOARD;
  double* sc = sb + BRD_B;
  const bool live = j < NJ;
  const int jj = live ? j : 0;
  LaneDyn d;
  d.q = live ? q[(size_t)ent * NJ + jj] : 0.0;
  d.qd = live ? v[(size_t)ent * NJ + jj] : 0.0;
  d.u = 0.0;
  d.qdd = live ? a[(size_t)ent * NJ + jj] : 0.0;
  node_kinematics(d, j, omask, model);
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  const double agrav[6] = {-model[MT_GRAV + 0], -model[MT_GRAV + 1], -model[MT_GRAV + 2], 0, 0, 0};
  scan_prefix_excl<6>(d.s, d.vp, zero6, j, omask);
  body_terms(d, j, model, nullptr, false);
  // bias acceleration including the joint accelerations: g = c qd + J qdd
#pragma unroll
  for (int k = 0; k < 6; ++k) d.g[k] += d.J[k] * d.qdd;
  scan_prefix_excl<6>(d.g, d.a0p, agrav, j, omask);
  body_force(d);
  scan_suffix_incl<6>(d.Z + 22, j, omask);
  (void)sb; (void)sc;
  if (live) out_tau[(size_t)ent * NJ + j] = dot6(d.J, d.Z + 22);
}

// update_geometry_placement: new end points / radius of one collision capsule in every model table
__global__ void set_capsule_kernel(double* __restrict__ model, int n_models, int capsule, double a0x, double a0y,
                                   double a0z, double a1x, double a1y, double a1z, double radius) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n_models) return;
  double* c = model + (size_t)i * MODEL_SIZE + MT_CAP + 8 * capsule;
  c[0] = a0x; c[1] = a0y; c[2] = a0z; c[3] = a1x; c[4] = a1y; c[5] = a1z; c[6] = radius;
}

// number of problems still running (early-exit check of long iteration budgets)
__global__ void count_live_kernel(int B, const int32_t* __restrict__ done, int32_t* __restrict__ out) {
  const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (b < B && !done[b]) atomicAdd(out, 1);
}

// ---------------------------------------------------------------------------------------------
__global__ void init_kernel(Problem P, Work W, SolverState S, FddpOpts O, const double* __restrict__ xs_ws,
                            const double* __restrict__ us_ws) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = P.T + 1;
  const long long nxs = (long long)P.B * T1 * NX, nus = (long long)P.B * P.T * NJ;
  if (gid < nxs) W.xs[gid] = xs_ws[gid];
  if (gid < nus) W.us[gid] = us_ws[gid];
  if (gid == 0) *S.t0 = agx_now_ns();
  if (gid < P.B) {
    const int b = (int)gid;
    S.xreg[b] = (O.reg_init == O.reg_init) ? O.reg_init : O.reg_min;
    S.cost[b] = 0.0; S.dg[b] = 0.0; S.dq[b] = 0.0; S.stop[b] = 0.0;
    S.is_feasible[b] = 0; S.was_feasible[b] = 0; S.recalc[b] = 1; S.done[b] = 0;
    S.status[b] = 1; S.iters[b] = 0; S.cur[b] = 0;
    S.dv[b] = 0.0; S.recalc_cost[b] = 1; S.pending[b] = 0; S.roll_ok[b] = 0;
  }
}

__global__ void finalize_kernel(Problem P, Work W, SolverState S, double* out_xs, double* out_us, double* out_K,
                                double* out_k, double* out_cost, int32_t* out_iters, int32_t* out_status,
                                double* out_stop) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = P.T + 1;
  const long long per_xs = (long long)T1 * NX, per_us = (long long)P.T * NJ, per_K = (long long)P.T * NJ * NX;
  if (gid < P.B * per_xs) {
    const int b = (int)(gid / per_xs);
    out_xs[gid] = W.xs[(size_t)(S.cur[b] & 1) * P.B * per_xs + gid];
  }
  if (gid < P.B * per_us) {
    const int b = (int)(gid / per_us);
    out_us[gid] = W.us[(size_t)(S.cur[b] & 1) * P.B * per_us + gid];
    if (out_k) out_k[gid] = W.k[gid];
  }
  if (out_K && out_K != W.K && gid < P.B * per_K) out_K[gid] = W.K[gid];
  if (gid < P.B) {
    out_cost[gid] = S.cost[gid];
    out_iters[gid] = S.iters[gid];
    out_status[gid] = S.status[gid];
    if (out_stop) out_stop[gid] = S.stop[gid];
  }
}

// ---- per-cost derivatives (agx_cost_derivatives): the references with every weight but one cost's zeroed, and the
// gradient part of the cost records unscaled into [n][AGX_N_COSTS][nx] / [n][AGX_N_COSTS][nv]
// ---- the MPC tick as ONE graph launch (agx_api.cu: TickGraph) ------------------------------------------------------
// The graph's kernel arguments are fixed when it is built, the caller's buffers change from tick to tick: the first and
// the last kernel of the graph read the caller's pointers from a small table in device memory that agx_solve refreshes
// with one copy before each launch.
struct IoTable {
  const double* x0; const double* xs_ws; const double* us_ws;
  double* out_xs; double* out_us; double* out_K; double* out_k; double* out_cost;
  int32_t* out_iters; int32_t* out_status; double* out_stop;
};

// init_kernel + the x0 copy; the cost records are written by the node_cost launch that follows in the graph, so the
// first calc_diff of the loop finds recalc_cost = 0
__global__ void init_io_kernel(Problem P, int nx, int nv, Work W, SolverState S, FddpOpts O,
                               const IoTable* __restrict__ io, double* __restrict__ x0_dst,
                               int32_t* __restrict__ round_ctr) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = P.T + 1;
  const long long nxs = (long long)P.B * T1 * nx, nus = (long long)P.B * P.T * nv;
  if (gid < nxs) W.xs[gid] = io->xs_ws[gid];
  if (gid < nus) W.us[gid] = io->us_ws[gid];
  if (gid < (long long)P.B * nx) x0_dst[gid] = io->x0[gid];
  if (gid == 0) { *S.t0 = agx_now_ns(); *round_ctr = 0; }
  if (gid < P.B) {
    const int b = (int)gid;
    S.xreg[b] = (O.reg_init == O.reg_init) ? O.reg_init : O.reg_min;
    S.cost[b] = 0.0; S.dg[b] = 0.0; S.dq[b] = 0.0; S.stop[b] = 0.0;
    S.is_feasible[b] = 0; S.was_feasible[b] = 0; S.recalc[b] = 1; S.done[b] = 0;
    S.status[b] = 1; S.iters[b] = 0; S.cur[b] = 0;
    S.dv[b] = 0.0; S.recalc_cost[b] = 0; S.pending[b] = 0; S.roll_ok[b] = 0;
  }
}

__global__ void finalize_io_kernel(Problem P, int nx, int nv, Work W, SolverState S, const IoTable* __restrict__ io) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = P.T + 1;
  const long long per_xs = (long long)T1 * nx, per_us = (long long)P.T * nv, per_K = (long long)P.T * nv * nx;
  if (gid < P.B * per_xs) {
    const int b = (int)(gid / per_xs);
    io->out_xs[gid] = W.xs[(size_t)(S.cur[b] & 1) * P.B * per_xs + gid];
  }
  if (gid < P.B * per_us) {
    const int b = (int)(gid / per_us);
    io->out_us[gid] = W.us[(size_t)(S.cur[b] & 1) * P.B * per_us + gid];
    if (io->out_k) io->out_k[gid] = W.k[gid];
  }
  if (io->out_K && gid < P.B * per_K) io->out_K[gid] = W.K[gid];
  if (gid < P.B) {
    io->out_cost[gid] = S.cost[gid];
    io->out_iters[gid] = S.iters[gid];
    io->out_status[gid] = S.status[gid];
    if (io->out_stop) io->out_stop[gid] = S.stop[gid];
  }
}

#if AGX_GPU
// last kernel of the loop body: one more round while some problem is unfinished and the budget allows it
__global__ void loop_condition_kernel(int B, const int32_t* __restrict__ done, int32_t* __restrict__ round_ctr,
                                      int rounds, cudaGraphConditionalHandle handle) {
  int mine = 0;
  for (int b = (int)threadIdx.x; b < B; b += (int)blockDim.x) mine |= !done[b];
  const int live = __syncthreads_or(mine);
  if (threadIdx.x == 0) {
    const int r = *round_ctr + 1;
    *round_ctr = r;
    cudaGraphSetConditional(handle, (live && r < rounds) ? 1u : 0u);
  }
}
#endif

__global__ void mask_refs_kernel(long long n_nodes, int nv, int ref_size, int slot, const double* __restrict__ refs,
                                 double* __restrict__ out) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_nodes * ref_size) return;
  const int k = (int)(gid % ref_size), nx = 2 * nv, o = 2 * nx + 2 * nv;
  double v = refs[gid];
  const bool is_wx = k >= nx && k < 2 * nx, is_wu = k >= 2 * nx + nv && k < o, is_wp = k >= o + 12 && k < o + 18;
  const bool is_c0 = k == o + 18, is_c1 = k == o + 19;
  if ((is_wx && slot != 0) || (is_wu && slot != 1) || (is_wp && slot != 2) || (is_c0 && slot != 3) || (is_c1 && slot != 4))
    v = 0.0;
  out[gid] = v;
}
__global__ void extract_gradients_kernel(long long n_nodes, int T1, int nv, int crec_size, int off_lq, int off_lv,
                                         int off_lu, const double* __restrict__ dts, int slot, int n_slots,
                                         const double* __restrict__ crec, double* __restrict__ out_Lx,
                                         double* __restrict__ out_Lu) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n_nodes * nv) return;
  const long long n = gid / nv;
  const int i = (int)(gid % nv), t = (int)(n % T1);
  const double inv = t == T1 - 1 ? 1.0 : 1.0 / dts[t];
  const double* C = crec + (size_t)n * crec_size;
  if (out_Lx) {
    out_Lx[((size_t)n * n_slots + slot) * 2 * nv + i] = C[off_lq + i] * inv;
    out_Lx[((size_t)n * n_slots + slot) * 2 * nv + nv + i] = C[off_lv + i] * inv;
  }
  if (out_Lu) out_Lu[((size_t)n * n_slots + slot) * nv + i] = t == T1 - 1 ? 0.0 : C[off_lu + i] * inv;
}

// ---- the same three kernels with run-time sizes (general-tree path, agx_tree.cuh)
__global__ void gather_refs_kernel_n(int B, int T1, int ref_size, const double* __restrict__ stream, int n_streams,
                                     int n_points, const int32_t* __restrict__ start, int start0,
                                     const int32_t* __restrict__ hidx, double* __restrict__ refs) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = gid / ref_size;
  const int k = (int)(gid % ref_size);
  if (n >= (long long)B * T1) return;
  const int b = (int)(n / T1), t = (int)(n % T1);
  int p = (start ? start[b] : start0) + hidx[t];
  if (p >= n_points) p = n_points - 1;
  if (p < 0) p = 0;
  refs[gid] = stream[((size_t)(n_streams > 1 ? b : 0) * n_points + p) * ref_size + k];
}

__global__ void init_kernel_n(Problem P, int nx, int nv, Work W, SolverState S, FddpOpts O,
                              const double* __restrict__ xs_ws, const double* __restrict__ us_ws) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = P.T + 1;
  const long long nxs = (long long)P.B * T1 * nx, nus = (long long)P.B * P.T * nv;
  if (gid < nxs) W.xs[gid] = xs_ws[gid];
  if (gid < nus) W.us[gid] = us_ws[gid];
  if (gid == 0) *S.t0 = agx_now_ns();
  if (gid < P.B) {
    const int b = (int)gid;
    S.xreg[b] = (O.reg_init == O.reg_init) ? O.reg_init : O.reg_min;
    S.cost[b] = 0.0; S.dg[b] = 0.0; S.dq[b] = 0.0; S.stop[b] = 0.0;
    S.is_feasible[b] = 0; S.was_feasible[b] = 0; S.recalc[b] = 1; S.done[b] = 0;
    S.status[b] = 1; S.iters[b] = 0; S.cur[b] = 0;
    S.dv[b] = 0.0; S.recalc_cost[b] = 1; S.pending[b] = 0; S.roll_ok[b] = 0;
  }
}

__global__ void finalize_kernel_n(Problem P, int nx, int nv, Work W, SolverState S, double* out_xs, double* out_us,
                                  double* out_K, double* out_k, double* out_cost, int32_t* out_iters,
                                  int32_t* out_status, double* out_stop) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = P.T + 1;
  const long long per_xs = (long long)T1 * nx, per_us = (long long)P.T * nv, per_K = (long long)P.T * nv * nx;
  if (gid < P.B * per_xs) {
    const int b = (int)(gid / per_xs);
    out_xs[gid] = W.xs[(size_t)(S.cur[b] & 1) * P.B * per_xs + gid];
  }
  if (gid < P.B * per_us) {
    const int b = (int)(gid / per_us);
    out_us[gid] = W.us[(size_t)(S.cur[b] & 1) * P.B * per_us + gid];
    if (out_k) out_k[gid] = W.k[gid];
  }
  if (out_K && out_K != W.K && gid < P.B * per_K) out_K[gid] = W.K[gid];
  if (gid < P.B) {
    out_cost[gid] = S.cost[gid];
    out_iters[gid] = S.iters[gid];
    out_status[gid] = S.status[gid];
    if (out_stop) out_stop[gid] = S.stop[gid];
  }
}

}  // namespace agx
#include "agx_tree.cuh"
#endif  // AGX_KERNELS_CUH_
