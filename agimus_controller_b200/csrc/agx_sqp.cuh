// agx_sqp.cuh — the SQP mode: mim_solvers.SolverCSQP as the reference configures it (ocp_base_croco.py:64-75, :172)
// with no constraint active.  It shares calc_diff / node_cost / the Riccati sweep with the FDDP path; what is new is
// the LINEAR rollout of the QP solution, the QP multipliers and the KKT norm (sqp_direction_kernel), and the merit
// line search on xs + a dx, us + a du (sqp_try_kernel + sqp_accept_kernel), all decided per problem on the device.
//
// Workspace reuse: dx lives in Work::gv, du overwrites Work::k, the per-node (cost, gap) pairs of a trial step go to
// slots 0 / 1 of Work::fs (the gaps were consumed by the direction kernel by then).  SolverState reuse: dg = merit of
// the candidate, stop = KKT norm, pending = line search running, roll_ok = index n of the step length 2^-n.
#ifndef AGX_SQP_CUH_
#define AGX_SQP_CUH_

namespace agx {

struct SqpOpts {
  double sigma, reg, mu, tol;
  int n_alphas;
  // SolverDDP's regularisation schedule, inherited by the mim_solvers solvers: floor `reg`, x reg_factor after a failed
  // factorisation or a step length <= th_stepinc (which includes a failed line search), / reg_factor after a step
  // length > th_stepdec; reaching reg_max ends the problem
  double reg_max, reg_factor, th_stepdec, th_stepinc;
  long long max_solve_ns;  // as FddpOpts::max_solve_ns
};

AGX_DEV double octet_max(double x, unsigned omask) {
  x = fmax(x, __shfl_xor_sync(omask, x, 1, 8));
  x = fmax(x, __shfl_xor_sync(omask, x, 2, 8));
  x = fmax(x, __shfl_xor_sync(omask, x, 4, 8));
  return x;
}

// One octet per problem.  Forward: dx_0 = fs_0, du_t = -k_t - K_t dx_t, dx_{t+1} = Fx dx_t + Fu du_t + fs_{t+1} with
// Fx = [I + dt G_q, dt (I + G_v); G_q, I + G_v], Fu = [dt N; N] read from the dynamics record (G_q = dt da/dq,
// G_v = dt da/dv, N = dt Minv).  Backward: l_T = Lx_T + Lxx_T dx_T, l_t = Lx_t + Lxx_t dx_t + Fx^T l_{t+1};
// KKT = max(|Lx + Fx^T l' - l|, |Lu + Fu^T l'|, |fs|) (SolverCSQP::checkKKTConditions).
// `pend` counts the problems that enter step length n of the line search (pend[0] is written here, pend[n + 1] by
// sqp_accept_kernel): a try / accept launch that finds its counter at zero returns at once.
__global__ void sqp_direction_kernel(Problem P, Work W, SolverState S, SqpOpts Q, int32_t* __restrict__ pend) {
  AGX_OCTET_SETUP();
  const int b = (int)ent;
  if (b >= P.B) return;
  if (S.done[b]) return;
  const int T = P.T, T1 = T + 1;
  const bool live = j < NJ;
  const int jj = live ? j : 0;
  const double* fsb = W.fs + (size_t)b * T1 * NX;
  double* dxb = W.gv + (size_t)b * T1 * NX;
  const double* Kb = W.K + (size_t)b * T * NJ * NX;
  double* kb = W.k + (size_t)b * T * NJ;
  const double* rec0 = W.rec + (size_t)b * T1 * REC_SIZE;
  const double* crec0 = W.crec + (size_t)b * T1 * CREC_SIZE;

  double dq = live ? fsb[jj] : 0.0, dv = live ? fsb[NJ + jj] : 0.0;
  double gl1 = fabs(dq) + fabs(dv), ginf = fmax(fabs(dq), fabs(dv));
  // the sweep is one dependent chain per node; its operands do not depend on it, so node t + 1's are fetched into
  // registers before node t's chain starts (row jj of K, G_q, G_v, N; k; the gap)
  struct FwdIn { double Kq[NJ], Kv[NJ], aq[NJ], av[NJ], mi[NJ], k, fq, fv, dt; };
  auto fetch_fwd = [&](int t, FwdIn& o) {
    const double* Kr = Kb + ((size_t)t * NJ + jj) * NX;
    const double* R = rec0 + (size_t)t * REC_SIZE;
#pragma unroll
    for (int m = 0; m < NJ; ++m) {
      o.Kq[m] = Kr[m]; o.Kv[m] = Kr[NJ + m];
      o.aq[m] = R[(RK_AQ + jj) * 8 + m]; o.av[m] = R[(RK_AV + jj) * 8 + m]; o.mi[m] = R[(RK_MI + jj) * 8 + m];
    }
    o.k = kb[t * NJ + jj];
    o.fq = live ? fsb[(t + 1) * NX + jj] : 0.0;
    o.fv = live ? fsb[(t + 1) * NX + NJ + jj] : 0.0;
    o.dt = P.dts[t];
  };
  FwdIn in;
  if (T > 0) fetch_fwd(0, in);
  for (int t = 0; t < T; ++t) {
    const FwdIn c = in;
    if (t + 1 < T) fetch_fwd(t + 1, in);
    if (live) { dxb[t * NX + j] = dq; dxb[t * NX + NJ + j] = dv; }
    double dqm[NJ], dvm[NJ], dum[NJ];
#pragma unroll
    for (int m = 0; m < NJ; ++m) { dqm[m] = __shfl_sync(omask, dq, m, 8); dvm[m] = __shfl_sync(omask, dv, m, 8); }
    double s = -c.k;
#pragma unroll
    for (int m = 0; m < NJ; ++m) s -= c.Kq[m] * dqm[m] + c.Kv[m] * dvm[m];
    const double du = live ? s : 0.0;
    if (live) kb[t * NJ + j] = du;
#pragma unroll
    for (int m = 0; m < NJ; ++m) dum[m] = __shfl_sync(omask, du, m, 8);
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < NJ; ++m) acc += c.aq[m] * dqm[m] + c.av[m] * dvm[m] + c.mi[m] * dum[m];
    const double fq = c.fq, fv = c.fv;
    const double dt = c.dt;
    const double dvn = dv + acc + fv;
    const double dqn = dq + dt * (dv + acc) + fq;
    gl1 += fabs(fq) + fabs(fv);
    ginf = fmax(ginf, fmax(fabs(fq), fabs(fv)));
    dq = live ? dqn : 0.0;
    dv = live ? dvn : 0.0;
  }
  if (live) { dxb[T * NX + j] = dq; dxb[T * NX + NJ + j] = dv; }

  // multipliers and stationarity
  double lq, lv, kkt;
  {
    const double* C = crec0 + (size_t)T * CREC_SIZE;
    double hq = 0.0;
#pragma unroll
    for (int m = 0; m < NJ; ++m)
      hq += C[CK_LQQ + (jj >= m ? lidx_(jj, m) : lidx_(m, jj))] * __shfl_sync(omask, dq, m, 8);
    const double hv = C[CK_LVV + jj] * dv;
    lq = C[CK_LQ + jj] + hq;
    lv = C[CK_LV + jj] + hv;
    kkt = live ? fmax(fabs(hq), fabs(hv)) : 0.0;
  }
  // adjoint sweep, same prefetching (column jj of G_q, G_v, N; row jj of Lqq; the node's dx)
  struct BwdIn { double aq[NJ], av[NJ], mi[NJ], lqq[NJ], lvv, lq, lv, lu, xq, xv, dt; };
  auto fetch_bwd = [&](int t, BwdIn& o) {
    const double* R = rec0 + (size_t)t * REC_SIZE;
    const double* C = crec0 + (size_t)t * CREC_SIZE;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      o.aq[i] = R[(RK_AQ + i) * 8 + jj]; o.av[i] = R[(RK_AV + i) * 8 + jj]; o.mi[i] = R[(RK_MI + i) * 8 + jj];
      o.lqq[i] = C[CK_LQQ + (jj >= i ? lidx_(jj, i) : lidx_(i, jj))];
    }
    o.lvv = C[CK_LVV + jj]; o.lq = C[CK_LQ + jj]; o.lv = C[CK_LV + jj]; o.lu = C[CK_LU + jj];
    o.xq = live ? dxb[t * NX + jj] : 0.0;
    o.xv = live ? dxb[t * NX + NJ + jj] : 0.0;
    o.dt = P.dts[t];
  };
  BwdIn bin;
  if (T > 0) fetch_bwd(T - 1, bin);
  for (int t = T - 1; t >= 0; --t) {
    const BwdIn c = bin;
    if (t > 0) fetch_bwd(t - 1, bin);
    const double dt = c.dt;
    const double w = live ? dt * lq + lv : 0.0;
    double su = c.lu, aq = 0.0, av = 0.0;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const double wi = __shfl_sync(omask, w, i, 8);
      su += c.mi[i] * wi;
      aq += c.aq[i] * wi;
      av += c.av[i] * wi;
    }
    const double xq = c.xq, xv = c.xv;
    double hq = 0.0;
#pragma unroll
    for (int m = 0; m < NJ; ++m) hq += c.lqq[m] * __shfl_sync(omask, xq, m, 8);
    const double hv = c.lvv * xv;
    const double nlq = c.lq + hq + lq + aq;
    const double nlv = c.lv + hv + w + av;
    if (live) kkt = fmax(kkt, fmax(fabs(su), fmax(fabs(hq), fabs(hv))));
    lq = nlq;
    lv = nlv;
  }
  // NaNs must survive the max: fmax drops them
  const double bad = octet_sum((live && !(lq - lq == 0.0 && lv - lv == 0.0)) ? 1.0 : 0.0, omask);
  kkt = fmax(octet_max(kkt, omask), octet_max(ginf, omask));
  gl1 = octet_sum(live ? gl1 : 0.0, omask);
  if (j == 0) {
    if (bad != 0.0) {
      S.stop[b] = nan("");
      S.status[b] = 3;
      S.done[b] = 1;
    } else {
      S.stop[b] = kkt;
      if (kkt <= Q.tol) {
        S.status[b] = 0;
        S.done[b] = 1;
      } else {
        S.dg[b] = S.cost[b] + Q.mu * gl1;
        S.pending[b] = 1;
        S.roll_ok[b] = 0;
        atomicAdd(pend, 1);
      }
    }
  }
}

// SolverCSQP::tryStep for the step length 2^-n the problem is at: one octet per (problem, node) evaluates the node
// cost and the gap to the next trial state of xs + a dx, us + a du, which it writes into the trial buffer.
// `pend` = counter of the problems entering this step length; the tick graph passes the counter array and the
// device-side index of the step length (`idx`), the stream path the counter itself (`idx` null)
template <bool COL>
__global__ void sqp_try_kernel(Problem P, Work W, SolverState S, const int32_t* __restrict__ pend,
                               const int32_t* __restrict__ idx) {
  if (idx) pend += *idx;
  if (*pend == 0) return;
  AGX_SMEM(smem);
  AGX_OCTET_SETUP();
  const int T = P.T, T1 = T + 1;
  double* sb = smem + oct_in_cta * OCT_BOARD;
  double* sc = sb + BRD_B;
  // grid-stride over the (problem, node) entries: after the first step length most problems have accepted, and a
  // launch that finds nothing to do must cost next to nothing
  const long long total = (long long)P.B * T1, stride = (long long)gridDim.x * octs_per_cta;
  for (long long e = ent; e < total; e += stride) {
  const int b = (int)(e / T1), t = (int)(e % T1);
  if (S.done[b] || !S.pending[b]) continue;
  const double a = ldexp(1.0, -S.roll_ok[b]);
  const size_t cur = (size_t)(S.cur[b] & 1), oth = cur ^ 1;
  const bool live = j < NJ, terminal = t == T;
  const int jj = live ? j : 0;
  const double* xs = W.xs + (cur * P.B + b) * (size_t)T1 * NX;
  const double* us = W.us + (cur * P.B + b) * (size_t)T * NJ;
  double* xt = W.xs + (oth * P.B + b) * (size_t)T1 * NX;
  double* ut = W.us + (oth * P.B + b) * (size_t)T * NJ;
  const double* dx = W.gv + (size_t)b * T1 * NX;
  const double* du = W.k + (size_t)b * T * NJ;
  LaneDyn d;
  d.q = live ? xs[t * NX + jj] + a * dx[t * NX + jj] : 0.0;
  d.qd = live ? xs[t * NX + NJ + jj] + a * dx[t * NX + NJ + jj] : 0.0;
  d.u = (live && !terminal) ? us[t * NJ + jj] + a * du[t * NJ + jj] : 0.0;
  const double q0 = d.q, v0 = d.qd;
  if (live) {
    xt[t * NX + j] = d.q;
    xt[t * NX + NJ + j] = d.qd;
    if (!terminal) ut[t * NJ + j] = d.u;
  }
  double c, qn, vn;
  const bool ok = node_calc<COL>(d, j, omask, model_of(P, b), P.refs + (size_t)e * REF_SIZE, terminal ? 0.0 : P.dts[t],
                                 terminal, sb, sc, &c, &qn, &vn);
  double g = 0.0;
  if (live && !terminal) {
    const double nq = xs[(t + 1) * NX + jj] + a * dx[(t + 1) * NX + jj];
    const double nvv = xs[(t + 1) * NX + NJ + jj] + a * dx[(t + 1) * NX + NJ + jj];
    g = fabs(qn - nq) + fabs(vn - nvv);
  }
  if (live && t == 0) g += fabs(W.x0[(size_t)b * NX + jj] - q0) + fabs(W.x0[(size_t)b * NX + NJ + jj] - v0);
  g = octet_sum(g, omask);
  if (j == 0) {
    double* out = W.fs + ((size_t)b * T1 + t) * NX;
    out[0] = ok ? c : nan("");
    out[1] = g;
  }
  AGX_OSYNC();  // the boards are reused by the next entry
  }
}

// merit_try < merit: take the step; otherwise the next step length (SolverCSQP::solve, merit line search)
__global__ void sqp_accept_kernel(Problem P, Work W, SolverState S, SqpOpts Q, int32_t* __restrict__ pend,
                                  const int32_t* __restrict__ idx) {
  if (idx) pend += *idx;
  if (*pend == 0) return;
  const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (b >= P.B) return;
  if (S.done[b] || !S.pending[b]) return;
  const int T1 = P.T + 1;
  const double* r = W.fs + (size_t)b * T1 * NX;
  double c = 0.0, g = 0.0;
  for (int t = 0; t < T1; ++t) { c += r[t * NX]; g += r[t * NX + 1]; }
  const double mt = c + Q.mu * g;
  const int n_now = S.roll_ok[b];
  bool finished = false;  // the line search of this iteration is over (step taken, or every step length refused)
  if (mt < S.dg[b]) {
    S.cur[b] ^= 1;
    finished = true;
  } else if (n_now + 1 >= Q.n_alphas) {
    finished = true;
  } else {
    S.roll_ok[b] = n_now + 1;
    atomicAdd(pend + 1, 1);
  }
  if (finished) {
    S.pending[b] = 0;
    S.iters[b] += 1;  // the solver counts every pass of its loop, whether the step was taken or every length refused
    const double steplength = ldexp(1.0, -n_now);
    double reg = S.xreg[b];
    if (steplength > Q.th_stepdec) reg = fmax(reg / Q.reg_factor, Q.reg);
    if (steplength <= Q.th_stepinc) {
      reg = fmin(reg * Q.reg_factor, Q.reg_max);
      if (reg == Q.reg_max) { S.status[b] = 2; S.done[b] = 1; }
    }
    S.xreg[b] = reg;
    if (!S.done[b] && Q.max_solve_ns > 0 && agx_now_ns() - *S.t0 > Q.max_solve_ns) { S.status[b] = 5; S.done[b] = 1; }
  }
}

// before the last sweep: every problem takes part again, with the solver's proximal sigma on top of the regularisation
__global__ void sqp_final_prepare_kernel(int B, SolverState S, SqpOpts Q) {
  const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (b >= B) return;
  S.xreg[b] = Q.sigma + S.xreg[b];
  S.is_feasible[b] = 0;
  if (S.status[b] != 3) S.done[b] = 0;
}

#if AGX_GPU
// ---- tick graph of agx_solve_sqp (agx_api.cu): loop control on the device ------------------------------------------
// before the line-search loop of an iteration: first step length, and enter the loop only if somebody searches
__global__ void sqp_arm_linesearch_kernel(int32_t* __restrict__ n_ctr, const int32_t* __restrict__ pend,
                                          cudaGraphConditionalHandle inner) {
  *n_ctr = 0;
  cudaGraphSetConditional(inner, pend[0] != 0 ? 1u : 0u);
}
// last kernel of the line-search loop: next step length while some problem entered it
__global__ void sqp_linesearch_condition_kernel(int32_t* __restrict__ n_ctr, const int32_t* __restrict__ pend,
                                                int n_alphas, cudaGraphConditionalHandle inner) {
  const int n = *n_ctr + 1;
  *n_ctr = n;
  cudaGraphSetConditional(inner, (n < n_alphas && pend[n] != 0) ? 1u : 0u);
}
#endif

}  // namespace agx
#endif  // AGX_SQP_CUH_
