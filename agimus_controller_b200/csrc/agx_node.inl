// agx_node.inl — one horizon node on one octet: IntegratedActionModelEuler::calc / calcDiff.
//
// Replaces, for the solve path, what the reference builds at
// agimus_controller/agimus_controller/ocp/ocp_croco_generic.py:687-745 (DifferentialActionModel-
// FreeFwdDynamics + CostModelSum{state, control, frame placement} + IntegratedActionModelEuler) and
// evaluates through problem.calc / problem.calcDiff (ocp_base_croco.py:172,
// agimus_controller_ros/agimus_controller_ros/mpc_debugger_node.py:300-301).
//
// Built from the lane phases of agx_dynamics.inl; `sa` (BRD_A doubles) and `sb` (BRD_B doubles) are
// this octet's shared-memory boards.

namespace agx {

#define AGX_OSYNC() __syncwarp(omask)

AGX_DEV double octet_sum(double x, unsigned omask) {
  x += __shfl_xor_sync(omask, x, 1, 8);
  x += __shfl_xor_sync(omask, x, 2, 8);
  x += __shfl_xor_sync(omask, x, 4, 8);
  return x;
}

// forward kinematics: world placements (prefix product over the chain), joint axes, s = J qd
AGX_DEV void node_kinematics(LaneDyn& d, int j, unsigned omask, const double* __restrict__ model, double* sa) {
  kin_local(d, j, model);
#pragma unroll
  for (int dist = 1; dist < 8; dist <<= 1) {
    se3_store(d, j, sa);
    AGX_OSYNC();
    se3_combine(d, j, dist, sa);
    AGX_OSYNC();
  }
  kin_axis(d, j);
}

// Forward dynamics a = (M + armature)^-1 (u - nle) in the world-frame formulation.  On exit:
// d.qdd = this joint's acceleration, qdd_all = all seven, L/rinv = Cholesky factor of M + armature
// (every lane holds the whole factor), board `sb` = [J dFda BS b u] per lane.  Returns false when the
// factorisation fails (octet-uniform).
template <bool DERIV>
AGX_DEV bool node_forward_dynamics(LaneDyn& d, int j, unsigned omask, const double* __restrict__ model, double* sa,
                                   double* sb, double* L, double* rinv, double* qdd_all) {
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  const double agrav[6] = {-model[MT_GRAV + 0], -model[MT_GRAV + 1], -model[MT_GRAV + 2], 0, 0, 0};
  vec6_store(d.s, j, sa);
  AGX_OSYNC();
  vec6_prefix_excl(d.vp, j, zero6, sa);
  AGX_OSYNC();
  body_terms(d, j, model, nullptr, DERIV);
  vec6_store(d.g, j, sa);
  AGX_OSYNC();
  vec6_prefix_excl(d.a0p, j, agrav, sa);
  AGX_OSYNC();
  body_force(d);
  comp_store(d, j, sa);
  AGX_OSYNC();
  comp_suffix(d, j, sa);
  AGX_OSYNC();
  column_terms(d, j, sb);
  sb[j * 18 + 16] = d.u;
  AGX_OSYNC();
  mass_column(d, j, model, sb);
#pragma unroll
  for (int k = 0; k < NJ; ++k) {
    chol_pivot(d.Mc, j, k, sa);
    AGX_OSYNC();
    chol_update(d.Mc, j, k, sa);
  }
  const bool ok = chol_load(sa, L, rinv);
#pragma unroll
  for (int i = 0; i < NJ; ++i) qdd_all[i] = sb[i * 18 + 16] - sb[i * 18 + 15];
  chol_solve7(L, rinv, qdd_all);
  d.qdd = 0.0;
#pragma unroll
  for (int i = 0; i < NJ; ++i)
    if (i == j) d.qdd = qdd_all[i];
  AGX_OSYNC();  // board A (the factor) may be overwritten from here on
  return ok;
}

// dtau/dq, dtau/dv columns (computeRNEADerivatives at the forward-dynamics acceleration)
AGX_DEV void node_rnea_derivatives(LaneDyn& d, int j, unsigned omask, double* sa, const double* sb) {
  // acceleration added by qdd: da_j = sum_{l<=j} J_l qdd_l ; force added: suffix sum of Y_l da_l
  double jq[6], dap[6], da[6], yda[6], dfc[6];
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 6; ++k) jq[k] = d.J[k] * d.qdd;
  vec6_store(jq, j, sa);
  AGX_OSYNC();
  vec6_prefix_excl(dap, j, zero6, sa);
#pragma unroll
  for (int k = 0; k < 6; ++k) da[k] = dap[k] + jq[k];
  inertia_apply(d.Y, da, yda);
#pragma unroll
  for (int k = 0; k < 6; ++k) dfc[k] = yda[k];
  vec6_store(yda, j, sa + 48);
  AGX_OSYNC();
  vec6_suffix_incl(dfc, j, sa + 48);
  double dFdq[6], dFdv[6];
  deriv_columns(d, j, dap, dfc, dFdq, dFdv);
  deriv_fill(d, j, dFdq, dFdv, sb);
  AGX_OSYNC();
}

// Weighted-quadratic costs of one node.  pose residual r6 = log6(Mref^-1 oMf); when DERIV the
// Gauss-Newton terms Lq_j (this lane's entry) and Lqq[:, j] are produced as well.
// Returns the (unscaled) node cost, identical on every lane.
template <bool DERIV>
AGX_DEV double node_costs(const LaneDyn& d, int j, unsigned omask, const double* __restrict__ model,
                          const double* __restrict__ ref, bool terminal, double* sa, double* Lq, double* Lv, double* Lu,
                          double* Lqq /*7*/) {
  const bool live = j < NJ;
  const int jj = live ? j : 0;
  // state / control regularisation (ResidualModelState, ResidualModelControl; A7, A8)
  const double rq = d.q - ref[jj], rv = d.qd - ref[NJ + jj];
  const double wq = live ? ref[NX + jj] : 0.0, wv = live ? ref[NX + NJ + jj] : 0.0;
  const double ru = d.u - ref[2 * NX + jj];
  const double wu = (live && !terminal) ? ref[2 * NX + NJ + jj] : 0.0;
  double part = 0.5 * wq * rq * rq + 0.5 * wv * rv * rv + 0.5 * wu * ru * ru;
  // frame placement (A9): joint frame_parent's world placement is broadcast through the board
  const int fpar = (int)model[MT_FP + 3];
  if (j == fpar) {
#pragma unroll
    for (int k = 0; k < 9; ++k) sa[k] = d.R[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) sa[9 + k] = d.p[k];
  }
  AGX_OSYNC();
  double R6[9], p6[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) R6[k] = sa[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) p6[k] = sa[9 + k];
  const double* Rref = ref + 2 * NX + 2 * NJ;
  const double* pref = Rref + 9;
  const double* wp = pref + 3;
  double Rf[9], pf[3], r6[6], Jl[18];
  frame_residual(R6, p6, model, Rref, pref, Rf, pf, r6, DERIV ? Jl : nullptr);
  double cpose = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) cpose += 0.5 * wp[k] * r6[k] * r6[k];
  const double cost = octet_sum(part, omask) + cpose;
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
    double* srq = sa + 16;  // [8][6]
#pragma unroll
    for (int k = 0; k < 6; ++k) srq[j * 6 + k] = rqc[k];
    AGX_OSYNC();
    double wr[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) wr[k] = wp[k] * rqc[k];
    double lq = wq * rq;
#pragma unroll
    for (int k = 0; k < 6; ++k) lq += rqc[k] * (wp[k] * r6[k]);
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
  }
  AGX_OSYNC();
  return cost;
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
