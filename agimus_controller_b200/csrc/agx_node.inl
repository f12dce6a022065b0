// agx_node.inl — one horizon node on one octet: IntegratedActionModelEuler::calc / calcDiff.
//
// Replaces, for the solve path, what the reference builds at
// agimus_controller/agimus_controller/ocp/ocp_croco_generic.py:687-745 (DifferentialActionModel-
// FreeFwdDynamics + CostModelSum{state, control, frame placement} + IntegratedActionModelEuler) and
// evaluates through problem.calc / problem.calcDiff (ocp_base_croco.py:172,
// agimus_controller_ros/agimus_controller_ros/mpc_debugger_node.py:300-301).
//
// Built from the lane phases of agx_dynamics.inl; `sb` (BRD_B doubles) and `sc` (BRD_C doubles) are this
// octet's shared-memory boards; the chain recursions themselves run as register scans (warp shuffles).

namespace agx {

#define AGX_OSYNC() __syncwarp(omask)

AGX_DEV double octet_sum(double x, unsigned omask) {
  x += __shfl_xor_sync(omask, x, 1, 8);
  x += __shfl_xor_sync(omask, x, 2, 8);
  x += __shfl_xor_sync(omask, x, 4, 8);
  return x;
}

// forward kinematics: world placements (prefix product over the chain), joint axes, s = J qd
AGX_DEV void node_kinematics(LaneDyn& d, int j, unsigned omask, const double* __restrict__ model) {
  kin_local(d, j, model);
  scan_se3_prefix(d, j, omask);
  kin_axis(d, j);
}

// Forward dynamics a = (M + armature)^-1 (u - nle) in the world-frame formulation.  On exit:
// d.qdd = this joint's acceleration, L/rinv = Cholesky factor of M + armature (every lane holds the
// whole factor; also left on board `sc` as L[i][k] at sc[i*8+k], 1/L[k][k] at sc[k*8+7]), board `sb` =
// [J dFda BS b u] per lane.  Returns false when the factorisation fails (octet-uniform).
template <bool DERIV>
AGX_DEV bool node_forward_dynamics(LaneDyn& d, int j, unsigned omask, const double* __restrict__ model, double* sb,
                                   double* sc, double* L, double* rinv) {
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  const double agrav[6] = {-model[MT_GRAV + 0], -model[MT_GRAV + 1], -model[MT_GRAV + 2], 0, 0, 0};
  scan_prefix_excl<6>(d.s, d.vp, zero6, j, omask);
  body_terms(d, j, model, nullptr, DERIV);
  scan_prefix_excl<6>(d.g, d.a0p, agrav, j, omask);
  body_force(d);
  if (DERIV) {
    scan_suffix_incl<28>(d.Z, j, omask);
  } else {
    scan_suffix_incl<10>(d.Z, j, omask);       // composite inertia
    scan_suffix_incl<6>(d.Z + 22, j, omask);   // composite bias force
  }
  column_terms<DERIV>(d, j, sb);
  sb[j * 18 + 16] = d.u;
  AGX_OSYNC();
  mass_column(d, j, model, sb);
#pragma unroll
  for (int i = 0; i < NJ; ++i) sc[i * 8 + j] = d.Mc[i];
  AGX_OSYNC();
  const bool ok = chol7_registers(sc, L, rinv);
  double rhs[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) rhs[i] = sb[i * 18 + 16] - sb[i * 18 + 15];
  chol_solve7(L, rinv, rhs);
  d.qdd = 0.0;
#pragma unroll
  for (int i = 0; i < NJ; ++i)
    if (i == j) d.qdd = rhs[i];
  if (DERIV) {
    // park the factor on the board so that its registers are free during the derivative phase
    AGX_OSYNC();
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
      if (k == j) {
#pragma unroll
        for (int i = 0; i < NJ; ++i)
          if (i >= k) sc[i * 8 + k] = L[lidx(i, k)];
        sc[k * 8 + 7] = rinv[k];
      }
    }
  }
  return ok;
}
// reload the parked factor (after an octet barrier)
AGX_DEV void factor_reload(const double* sc, double* L, double* rinv) {
#pragma unroll
  for (int k = 0; k < NJ; ++k) {
#pragma unroll
    for (int i = 0; i < NJ; ++i)
      if (i >= k) L[lidx(i, k)] = sc[i * 8 + k];
    rinv[k] = sc[k * 8 + 7];
  }
}

// dtau/dq, dtau/dv columns (computeRNEADerivatives at the forward-dynamics acceleration)
AGX_DEV void node_rnea_derivatives(LaneDyn& d, int j, unsigned omask, const double* sb) {
  // acceleration added by qdd: da_j = sum_{l<=j} J_l qdd_l ; force added: suffix sum of Y_l da_l
  double jq[6], dap[6], da[6], dfc[6];
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 6; ++k) jq[k] = d.J[k] * d.qdd;
  scan_prefix_excl<6>(jq, dap, zero6, j, omask);
#pragma unroll
  for (int k = 0; k < 6; ++k) da[k] = dap[k] + jq[k];
  // sum_{l >= j} Y_l da_l = Yc_j da_j + sum_{m > j} (Yc_m J_m) qdd_m: the own inertia Y is not needed any more
  double fq[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) fq[k] = d.dFda[k] * d.qdd;
  scan_suffix_incl<6>(fq, j, omask);
  inertia_apply(d.Z, da, dfc);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double up = __shfl_down_sync(omask, fq[k], 1, 8);
    dfc[k] += (j + 1 < 8) ? up : 0.0;
  }
  double dFdq[6], dFdv[6];
  deriv_columns(d, j, dap, dfc, dFdq, dFdv);
  deriv_fill(d, j, dFdq, dFdv, sb);
}

// Weighted-quadratic costs of one node.  pose residual r6 = log6(Mref^-1 oMf); when DERIV the
// Gauss-Newton terms Lq_j (this lane's entry) and Lqq[:, j] are produced as well.
// Returns the (unscaled) node cost, identical on every lane.
template <bool DERIV, bool COL = false>
AGX_DEV double node_costs(const LaneDyn& d, int j, unsigned omask, const double* __restrict__ model,
                          const double* __restrict__ ref, bool terminal, double* srq /*[8][6] scratch board*/, double* Lq,
                          double* Lv, double* Lu, double* Lqq /*7*/) {
  const bool live = j < NJ;
  const int jj = live ? j : 0;
  // state / control regularisation (ResidualModelState, ResidualModelControl; A7, A8)
  const double rq = d.q - ref[jj], rv = d.qd - ref[NJ + jj];
  const double wq = live ? ref[NX + jj] : 0.0, wv = live ? ref[NX + NJ + jj] : 0.0;
  const double ru = d.u - ref[2 * NX + jj];
  const double wu = (live && !terminal) ? ref[2 * NX + NJ + jj] : 0.0;
  double part = 0.5 * wq * rq * rq + 0.5 * wv * rv * rv + 0.5 * wu * ru * ru;
  // frame placement (A9): joint frame_parent's world placement is broadcast to the octet
  const int fpar = (int)model[MT_FP + 3];
  double R6[9], p6[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) R6[k] = __shfl_sync(omask, d.R[k], fpar, 8);
#pragma unroll
  for (int k = 0; k < 3; ++k) p6[k] = __shfl_sync(omask, d.p[k], fpar, 8);
  const double* Rref = ref + 2 * NX + 2 * NJ;
  const double* pref = Rref + 9;
  const double* wp = pref + 3;
  double Rf[9], pf[3], r6[6], Jl[18];
  frame_residual(R6, p6, model, Rref, pref, Rf, pf, r6, DERIV ? Jl : nullptr);
  const bool tworld = model[MT_COL + 6] != 0.0;  // ResidualModelFrameTranslation: linear part p_f - pref in the world
  if (tworld) {
#pragma unroll
    for (int k = 0; k < 3; ++k) r6[k] = pf[k] - pref[k];
  }
  double cpose = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) cpose += 0.5 * wp[k] * r6[k] * r6[k];
  double cost = octet_sum(part, omask) + cpose;
  // collision pairs (A10): lane j holds its own entry of each gradient row; the closest points are computed
  // redundantly on every lane from capsule end points broadcast by the parent joint's lane
  double crq[MAX_PAIRS] = {0, 0}, g1[MAX_PAIRS] = {0, 0}, g2[MAX_PAIRS] = {0, 0};
  if (COL) {
    const int npairs = (int)model[MT_COL];
    const double alpha = model[MT_COL + 1];
    const double* wc = wp + 6;
#pragma unroll
    for (int k = 0; k < MAX_PAIRS; ++k) {
      // the broadcasts run unconditionally: octets that share a warp may carry different models
      double e[2][6];
      int jpar[2];
      double rad = 0.0;
#pragma unroll
      for (int side = 0; side < 2; ++side) {
        const int ic = (int)model[MT_COL + 2 + 2 * k + side];
        const double* a = model + MT_CAP + 8 * ic;
        const int par = (int)a[7];
        jpar[side] = par;
        rad += a[6];
        double w[6];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          w[r] = d.p[r] + (d.R[3 * r] * a[0] + d.R[3 * r + 1] * a[1] + d.R[3 * r + 2] * a[2]);
          w[3 + r] = d.p[r] + (d.R[3 * r] * a[3] + d.R[3 * r + 1] * a[4] + d.R[3 * r + 2] * a[5]);
        }
#pragma unroll
        for (int m = 0; m < 6; ++m) {
          const double v = __shfl_sync(omask, w[m], par < 0 ? 0 : par, 8);
          e[side][m] = par < 0 ? a[m] : v;
        }
      }
      if (k < npairs) {
        double ca[3], cb[3], nn[3];
        const double len = segment_pair(e[0], e[0] + 3, e[1], e[1] + 3, ca, cb, nn);
        const double r = len - rad;
        double a, ar, arr;
        quadexp(r, alpha, a, ar, arr);
        cost += wc[k] * a;
        if (DERIV) {
          g1[k] = wc[k] * ar;
          g2[k] = wc[k] * arr;
          double wa[3], wb[3];
          cross3(d.J + 3, ca, wa);
          cross3(d.J + 3, cb, wb);
          double da = 0.0, db = 0.0;
#pragma unroll
          for (int m = 0; m < 3; ++m) { da += nn[m] * (wa[m] + d.J[m]); db += nn[m] * (wb[m] + d.J[m]); }
          crq[k] = ((j <= jpar[0]) ? da : 0.0) - ((j <= jpar[1]) ? db : 0.0);
        }
      }
    }
  }
  if (DERIV) {
    // column j of the LOCAL frame Jacobian: oMf^-1 acting on the world axis J_j (zero for joints
    // that do not move the frame), then Rq[:, j] = Jlog6 * that
    double t[3], pw[3], cl[3], ca[3];
    cross3(pf, d.J + 3, pw);
    const double moves = (j <= fpar) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) t[k] = moves * (d.J[k] - pw[k]);
    mtv3(Rf, t, cl);
    double ja[3] = {moves * d.J[3], moves * d.J[4], moves * d.J[5]};
    mtv3(Rf, ja, ca);
    double rqc[6];
    const double* A = Jl;
    const double* Bm = Jl + 9;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      rqc[i] = A[3 * i] * cl[0] + A[3 * i + 1] * cl[1] + A[3 * i + 2] * cl[2] + Bm[3 * i] * ca[0] +
               Bm[3 * i + 1] * ca[1] + Bm[3 * i + 2] * ca[2];
      rqc[3 + i] = A[3 * i] * ca[0] + A[3 * i + 1] * ca[1] + A[3 * i + 2] * ca[2];
    }
    if (tworld) {  // d(p_f)/dq_j = world velocity of the frame origin under joint j
#pragma unroll
      for (int k = 0; k < 3; ++k) rqc[k] = t[k];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) srq[j * 6 + k] = rqc[k];
    AGX_OSYNC();
    double wr[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) wr[k] = wp[k] * rqc[k];
    double lq = wq * rq;
#pragma unroll
    for (int k = 0; k < 6; ++k) lq += rqc[k] * (wp[k] * r6[k]);
    if (COL) {
#pragma unroll
      for (int k = 0; k < MAX_PAIRS; ++k) lq += g1[k] * crq[k];
    }
    *Lq = lq;
    *Lv = wv * rv;
    *Lu = wu * ru;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      double h = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) h += srq[i * 6 + k] * wr[k];
      Lqq[i] = h + ((i == j) ? wq : 0.0);
    }
    if (COL) {
#pragma unroll
      for (int k = 0; k < MAX_PAIRS; ++k) {
        const double gk = g2[k] * crq[k];
#pragma unroll
        for (int i = 0; i < NJ; ++i) Lqq[i] += __shfl_sync(omask, crq[k], i, 8) * gk;
      }
    }
  }
  AGX_OSYNC();
  return cost;
}

// ---------------------------------------------------------------------------------------------
// One node's cost terms on ONE THREAD (32 nodes per warp): sequential forward kinematics, frame
// placement residual r = log6(Mref^-1 oMf) with Pinocchio's branches, Rq = Jlog6 * fJf, weighted-quad
// Gauss-Newton terms.  The log maps are scalar code: run once per octet they waste 7/8 of the lanes,
// run one node per thread they do not.  With DERIV the cost record of the node is written
// (scaled by s = dt, or 1 for the terminal node).  Returns the scaled node cost.
// With COL the capsule pairs of the model table add w a(r) per pair (a = QuadExp of the signed distance r) and its
// Gauss-Newton terms w a' Rq, w a'' Rq Rq^T.
template <bool DERIV, bool COL = false>
AGX_DEV double thread_node_cost(const double* __restrict__ model, const double* __restrict__ ref,
                                const double* __restrict__ x, const double* __restrict__ u, bool terminal, double s,
                                double* __restrict__ rec, double* __restrict__ terms = nullptr) {
  const int fpar = (int)model[MT_FP + 3];
  // forward kinematics down the chain; world joint axes J_i = [p_i x z_i; z_i]
  double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, p[3] = {0, 0, 0};
  double Rf0[9], pf0[3];
  double Jw[NJ][6];
  double cw[COL ? MAX_CAPS : 1][6];  // capsule end points in the world
  if (COL) {
#pragma unroll
    for (int c = 0; c < MAX_CAPS; ++c)
#pragma unroll
      for (int k = 0; k < 6; ++k) cw[c][k] = model[MT_CAP + 8 * c + k];  // world-fixed capsules stay as they are
  }
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    double sq, cq;
    AGX_SINCOS(x[i], &sq, &cq);
    double Rl[9], pl[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double a = model[(MF_RP + 3 * r) * 8 + i], b = model[(MF_RP + 3 * r + 1) * 8 + i];
      Rl[3 * r] = cq * a + sq * b;
      Rl[3 * r + 1] = cq * b - sq * a;
      Rl[3 * r + 2] = model[(MF_RP + 3 * r + 2) * 8 + i];
      pl[r] = model[(MF_PP + r) * 8 + i];
    }
    double Rn[9], pn[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) Rn[3 * r + c] = R[3 * r] * Rl[c] + R[3 * r + 1] * Rl[3 + c] + R[3 * r + 2] * Rl[6 + c];
      pn[r] = p[r] + (R[3 * r] * pl[0] + R[3 * r + 1] * pl[1] + R[3 * r + 2] * pl[2]);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = Rn[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) p[k] = pn[k];
    if (DERIV) {
      const double z[3] = {R[2], R[5], R[8]};
      double pz[3];
      cross3(p, z, pz);
      const double moves = (COL || i <= fpar) ? 1.0 : 0.0;  // COL keeps every axis and masks at the point of use
#pragma unroll
      for (int k = 0; k < 3; ++k) { Jw[i][k] = moves * pz[k]; Jw[i][3 + k] = moves * z[k]; }
    }
    if (COL) {
#pragma unroll
      for (int c = 0; c < MAX_CAPS; ++c)
        if ((int)model[MT_CAP + 8 * c + 7] == i) {
          const double* a = model + MT_CAP + 8 * c;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            cw[c][r] = p[r] + (R[3 * r] * a[0] + R[3 * r + 1] * a[1] + R[3 * r + 2] * a[2]);
            cw[c][3 + r] = p[r] + (R[3 * r] * a[3] + R[3 * r + 1] * a[4] + R[3 * r + 2] * a[5]);
          }
        }
    }
    if (i == fpar) {
#pragma unroll
      for (int k = 0; k < 9; ++k) Rf0[k] = R[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) pf0[k] = p[k];
    }
  }
  const double* Rref = ref + 2 * NX + 2 * NJ;
  const double* pref = Rref + 9;
  const double* wp = pref + 3;
  double Rf[9], pf[3], r6[6], Jl[18];
  frame_residual(Rf0, pf0, model, Rref, pref, Rf, pf, r6, DERIV ? Jl : nullptr);
  const bool tworld = model[MT_COL + 6] != 0.0;  // ResidualModelFrameTranslation: linear part p_f - pref in the world
  if (tworld) {
#pragma unroll
    for (int k = 0; k < 3; ++k) r6[k] = pf[k] - pref[k];
  }
  double cost = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) cost += 0.5 * wp[k] * r6[k] * r6[k];
  // collision pairs: gradient rows crq, gains g1 = w a', g2 = w a''
  double crq[COL ? MAX_PAIRS : 1][NJ], g1[MAX_PAIRS] = {0, 0}, g2[MAX_PAIRS] = {0, 0}, ccost[MAX_PAIRS] = {0, 0},
                                       cdist[MAX_PAIRS] = {0, 0};
  if (COL) {
    const int npairs = (int)model[MT_COL];
    const double alpha = model[MT_COL + 1];
    const double* wc = wp + 6;
#pragma unroll
    for (int k = 0; k < MAX_PAIRS; ++k) {
#pragma unroll
      for (int i = 0; i < NJ; ++i) crq[k][i] = 0.0;
      if (k < npairs) {
        const int ia = (int)model[MT_COL + 2 + 2 * k], ib = (int)model[MT_COL + 3 + 2 * k];
        double ea[6], eb[6];
#pragma unroll
        for (int c = 0; c < MAX_CAPS; ++c) {  // static indexing keeps cw in registers
          if (c == ia) {
#pragma unroll
            for (int m = 0; m < 6; ++m) ea[m] = cw[c][m];
          }
          if (c == ib) {
#pragma unroll
            for (int m = 0; m < 6; ++m) eb[m] = cw[c][m];
          }
        }
        double ca[3], cb[3], nn[3];
        const double len = segment_pair(ea, ea + 3, eb, eb + 3, ca, cb, nn);
        const double r = len - model[MT_CAP + 8 * ia + 6] - model[MT_CAP + 8 * ib + 6];
        double a, ar, arr;
        quadexp(r, alpha, a, ar, arr);
        ccost[k] = wc[k] * a;
        cdist[k] = r;
        cost += ccost[k];
        if (DERIV) {
          g1[k] = wc[k] * ar;
          g2[k] = wc[k] * arr;
          const int ja = (int)model[MT_CAP + 8 * ia + 7], jb = (int)model[MT_CAP + 8 * ib + 7];
          // d r / d q_i = n . (z_i x (ca - p_i)) [i <= ja] - n . (z_i x (cb - p_i)) [i <= jb]
#pragma unroll
          for (int i = 0; i < NJ; ++i) {
            double wa[3], wb[3];
            cross3(Jw[i] + 3, ca, wa);
            cross3(Jw[i] + 3, cb, wb);
            double da = 0.0, db = 0.0;
#pragma unroll
            for (int m = 0; m < 3; ++m) { da += nn[m] * (wa[m] + Jw[i][m]); db += nn[m] * (wb[m] + Jw[i][m]); }
            crq[k][i] = ((i <= ja) ? da : 0.0) - ((i <= jb) ? db : 0.0);
          }
        }
      }
    }
  }
  if (terms) {
    // per-cost view (mpc_debugger_node.py:294-323): [state_reg, control_reg, goal_tracking] values and the
    // frame-placement residual; unscaled (differential) costs
    double cs = 0.0, cu = 0.0;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      const double rq = x[i] - ref[i], rv = x[NJ + i] - ref[NJ + i];
      cs += 0.5 * ref[NX + i] * rq * rq + 0.5 * ref[NX + NJ + i] * rv * rv;
      if (!terminal) {
        const double ru = u[i] - ref[2 * NX + i];
        cu += 0.5 * ref[2 * NX + NJ + i] * ru * ru;
      }
    }
    terms[0] = cs; terms[1] = cu; terms[2] = cost - (ccost[0] + ccost[1]);
#pragma unroll
    for (int k = 0; k < 6; ++k) terms[3 + k] = r6[k];
#pragma unroll
    for (int k = 0; k < MAX_PAIRS; ++k) { terms[9 + k] = ccost[k]; terms[9 + MAX_PAIRS + k] = cdist[k]; }
  }
  double wr6[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) wr6[k] = wp[k] * r6[k];
  // Rq columns (in place over Jw) and the per-joint gradient / diagonal terms
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double rq = x[i] - ref[i], rv = x[NJ + i] - ref[NJ + i];
    const double wq = ref[NX + i], wv = ref[NX + NJ + i];
    const double ru = terminal ? 0.0 : u[i] - ref[2 * NX + i];
    const double wu = terminal ? 0.0 : ref[2 * NX + NJ + i];
    cost += 0.5 * wq * rq * rq + 0.5 * wv * rv * rv + 0.5 * wu * ru * ru;
    if (DERIV) {
      double t[3], pw[3], cl[3], ca[3];
      if (COL && i > fpar) {
#pragma unroll
        for (int k = 0; k < 6; ++k) Jw[i][k] = 0.0;
      }
      cross3(pf, Jw[i] + 3, pw);
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] = Jw[i][k] - pw[k];
      mtv3(Rf, t, cl);
      mtv3(Rf, Jw[i] + 3, ca);
      const double* A = Jl;
      const double* Bm = Jl + 9;
      double lq = wq * rq;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        Jw[i][k] = A[3 * k] * cl[0] + A[3 * k + 1] * cl[1] + A[3 * k + 2] * cl[2] + Bm[3 * k] * ca[0] +
                   Bm[3 * k + 1] * ca[1] + Bm[3 * k + 2] * ca[2];
        Jw[i][3 + k] = A[3 * k] * ca[0] + A[3 * k + 1] * ca[1] + A[3 * k + 2] * ca[2];
      }
      if (tworld) {
#pragma unroll
        for (int k = 0; k < 3; ++k) Jw[i][k] = t[k];
      }
#pragma unroll
      for (int k = 0; k < 6; ++k) lq += Jw[i][k] * wr6[k];
      if (COL) {
#pragma unroll
        for (int k = 0; k < MAX_PAIRS; ++k) lq += g1[k] * crq[k][i];
      }
      rec[CK_LVV + i] = s * wv;
      rec[CK_LUU + i] = s * wu;
      rec[CK_LQ + i] = s * lq;
      rec[CK_LV + i] = s * (wv * rv);
      rec[CK_LU + i] = s * (wu * ru);
    }
  }
  if (DERIV) {
    // Lqq = Rq^T diag(w) Rq + diag(wq): symmetric, packed lower triangle
#pragma unroll
    for (int k = 0; k < NJ; ++k) {
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        if (i >= k) {
          double h = (i == k) ? ref[NX + i] : 0.0;
#pragma unroll
          for (int m = 0; m < 6; ++m) h += Jw[i][m] * (wp[m] * Jw[k][m]);
          if (COL) {
#pragma unroll
            for (int m = 0; m < MAX_PAIRS; ++m) h += g2[m] * crq[m][i] * crq[m][k];
          }
          rec[CK_LQQ + lidx(i, k)] = s * h;
        }
      }
    }
    rec[CK_COST] = s * cost;
  }
  return s * cost;
}

// solve (M + armature) X = rhs for this lane's column with the register-resident factor
AGX_DEV void solve_column(const double* L, const double* rinv, const double* rhs, double scale, double* out) {
  double t[NJ];
#pragma unroll
  for (int i = 0; i < NJ; ++i) t[i] = rhs[i];
  chol_solve7(L, rinv, t);
#pragma unroll
  for (int i = 0; i < NJ; ++i) out[i] = scale * t[i];
}

}  // namespace agx
