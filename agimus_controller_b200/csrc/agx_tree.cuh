// agx_tree.cuh — the solve path for GENERAL kinematic trees: any parent table, revolute / prismatic joints about any
// axis, nv <= 16 (the 9-DoF Panda with its two finger joints branching off the hand is BASELINE config 4).
//
// Reference: the reference locks the finger joints only optionally
// (agimus_controller/agimus_controller/factory/robot_model.py:231-259); with `moving_joint_names` naming them the
// reduced model has nq = 9 and DifferentialActionModelFreeFwdDynamics / IntegratedActionModelEuler
// (ocp/ocp_croco_generic.py:687-745) run on the branching tree with two prismatic joints.
//
// Mapping: a GROUP of 16 consecutive lanes owns one entity (a (problem, node) pair or a problem); lane j owns joint j
// and column j of M, dtau/dq, dtau/dv; a warp carries two entities.  Same world-frame formulation as the 7-joint
// chain kernels (agx_dynamics.inl), with the chain scans replaced by tree scans:
//   * "sum / product over the ancestors" (placements, velocities, accelerations) = POINTER JUMPING over the parent
//     table: in round r every lane combines with the partial result of its 2^r-th ancestor (log2(depth) rounds of
//     width-16 shuffles, any tree shape);
//   * "sum over the subtree" (composite inertia / momentum / B block / force) = every lane adds the values of the
//     lanes named in its descendant mask;
//   * entry (i, j) of M, dtau/dq, dtau/dv is non-zero only when i is an ancestor of j or j an ancestor of i
//     (ancestor / descendant masks from the table).
// Joint transforms: revolute = Rodrigues rotation about the joint axis, prismatic = translation along it; world motion
// axes J = [p x Ra; Ra] / [Ra; 0].
//
// The kernels here are templated on NV and instantiated by agx_api.cu for the sizes listed in AGX_TREE_DISPATCH; the
// 7-joint serial revolute-z chain keeps its own tuned kernels (agx_kernels.cuh) unless AGX_TREE=1 asks for these
// (which is how the two paths are cross-checked).
#ifndef AGX_TREE_CUH_
#define AGX_TREE_CUH_

namespace agx {
namespace tree {

constexpr int GW = 16;  // lanes per entity

// ---- device model table (doubles): joint fields are field-major / lane-minor, tm[f * 16 + j]
constexpr int TF_RP = 0;        // 9 placement rotation (row-major)
constexpr int TF_PP = 9;        // 3 placement translation
constexpr int TF_MASS = 12;
constexpr int TF_COM = 13;      // 3
constexpr int TF_INERTIA = 16;  // 6
constexpr int TF_ARM = 22;
constexpr int TF_AXIS = 23;     // 3 joint axis (joint frame)
constexpr int TF_JTYPE = 26;
constexpr int TF_PARENT = 27;   // parent joint or -1
constexpr int TF_SUB = 28;      // bit k set: k is j or a descendant of j
constexpr int TF_ANC = 29;      // bit i set: i is j or an ancestor of j
constexpr int TF_NFIELDS = 30;
constexpr int TT_GRAV = TF_NFIELDS * GW;  // 3
constexpr int TT_FR = TT_GRAV + 3;        // 9 task-frame rotation
constexpr int TT_FP = TT_FR + 9;          // 3 task-frame translation, then the frame's parent joint
constexpr int TT_CAP = TT_FP + 4;         // capsules: [a0 3][a1 3][radius][parent joint or -1] each
constexpr int TT_COL = TT_CAP + 8 * MAX_CAPS;  // [n_pairs][alpha][pair0 a][pair0 b][pair1 a][pair1 b]
constexpr int TMODEL_SIZE = TT_COL + 8;

// ---- per-NV layouts of the records and boards
template <int NV>
struct TL {
  static constexpr int NX = 2 * NV;
  static constexpr int REF = 6 * NV + 20;
  // dynamics record, field-major / lane-minor (stride 16): dt da/dq, dt da/dv, dt Minv rows, xnext
  static constexpr int RK_AQ = 0, RK_AV = NV, RK_MI = 2 * NV, RK_QN = 3 * NV, RK_VN = 3 * NV + 1;
  static constexpr int REC = (3 * NV + 2) * GW;
  // cost record
  static constexpr int NTRI = NV * (NV + 1) / 2;
  static constexpr int CK_LQQ = 0, CK_LVV = NTRI, CK_LUU = NTRI + NV, CK_LQ = NTRI + 2 * NV, CK_LV = NTRI + 3 * NV,
                       CK_LU = NTRI + 4 * NV, CK_COST = NTRI + 5 * NV;
  static constexpr int CREC = (CK_COST + 2) & ~1;
  static constexpr int ROUNDS = NV > 8 ? 4 : (NV > 4 ? 3 : (NV > 2 ? 2 : 1));  // pointer-jumping rounds: 2^ROUNDS >= NV
  // per-group shared-memory boards
  static constexpr int SB = 0;                  // [16][18]  J dFda BS b u per lane
  static constexpr int SC = GW * 18;            // [NV][17]  mass matrix, then its factor (slot 16 of row k = 1/L[k][k])
  static constexpr int SQ = SC + NV * (GW + 1); // [16][6]   pose-residual Jacobian columns
  static constexpr int SX = SQ + GW * 6;        // [NX]      dx of the forward pass
  static constexpr int BOARD = (SX + NX + 1) & ~1;
};
constexpr int CS = GW + 1;  // row stride of the mass-matrix board

AGX_DEV constexpr int tidx(int nv, int i, int k) { return k * nv - (k * (k - 1)) / 2 + (i - k); }  // packed lower, i >= k

#define AGX_TREE_SETUP()                                        \
  const int j = (int)(threadIdx.x & 15u);                       \
  const unsigned gm = 0xFFFFu << (threadIdx.x & 16u);           \
  const int grp_in_cta = (int)(threadIdx.x >> 4);               \
  const int grps_per_cta = (int)(blockDim.x >> 4);              \
  const long long ent = (long long)blockIdx.x * grps_per_cta + grp_in_cta;
#define AGX_GSYNC() __syncwarp(gm)

AGX_DEV double gsum(double x, unsigned gm) {
  x += __shfl_xor_sync(gm, x, 1, GW);
  x += __shfl_xor_sync(gm, x, 2, GW);
  x += __shfl_xor_sync(gm, x, 4, GW);
  x += __shfl_xor_sync(gm, x, 8, GW);
  return x;
}
AGX_DEV int shfl_int(unsigned gm, int v, int src) { return (int)__shfl_sync(gm, v, src, GW); }

template <int NV>
struct TLane {
  double q, qd, u, qdd, b;
  double R[9], p[3], J[6], s[6], vp[6], v[6], c[6], g[6], a0p[6], Y[10], Z[28], dFda[6], BS[3];
  double Mc[NV], tq[NV], tv[NV];
  int par;
  unsigned sub, anc;
};

// ---------------------------------------------------------------- tree scans
// x_j <- x_j + sum over the strict ancestors of j
template <int N, int ROUNDS>
AGX_DEV void scan_anc_incl(double* x, int par, int j, unsigned gm) {
  int a = par;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const int src = a >= 0 ? a : j;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double t = __shfl_sync(gm, x[k], src, GW);
      if (a >= 0) x[k] += t;
    }
    const int an = shfl_int(gm, a, src);
    a = a >= 0 ? an : -1;
  }
}
// out = seed + sum over the strict ancestors of j
template <int N, int ROUNDS>
AGX_DEV void scan_anc_excl(const double* x, double* out, const double* seed, int par, int j, unsigned gm) {
  double acc[N];
#pragma unroll
  for (int k = 0; k < N; ++k) acc[k] = x[k];
  scan_anc_incl<N, ROUNDS>(acc, par, j, gm);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double t = __shfl_sync(gm, acc[k], par >= 0 ? par : j, GW);
    out[k] = seed[k] + (par >= 0 ? t : 0.0);
  }
}
// x_j <- sum over j and its descendants
template <int N, int NV, bool ROLLED = false>
AGX_DEV void subtree_sum(double* x, unsigned sub, unsigned gm) {
  double acc[N];
#pragma unroll
  for (int m = 0; m < N; ++m) acc[m] = 0.0;
  // a real loop over the source lanes (the source lane of a shuffle may be a run-time value): the unrolled form was
  // NV x N shuffle / add pairs of straight-line code, and the node kernels are bound by instruction fetch
  // (ROLLED for the node kernels, which are bound by instruction fetch; unrolled for the sequential kernels, which
  // are bound by the latency of this very chain)
  if (ROLLED) {
#pragma unroll 1
    for (int k = 0; k < NV; ++k) {
      const bool in = ((sub >> k) & 1u) != 0u;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const double t = __shfl_sync(gm, x[m], k, GW);
        if (in) acc[m] += t;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const bool in = ((sub >> k) & 1u) != 0u;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const double t = __shfl_sync(gm, x[m], k, GW);
        if (in) acc[m] += t;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < N; ++m) x[m] = acc[m];
}

// ---------------------------------------------------------------- kinematics
template <int NV>
AGX_DEV void lane_load(TLane<NV>& d, int j, const double* x, const double* u) {
  const bool live = j < NV;
  d.q = live ? x[j] : 0.0;
  d.qd = live ? x[NV + j] : 0.0;
  d.u = (live && u) ? u[j] : 0.0;
  d.qdd = 0.0;
}

// world placements (pointer-jumping prefix product), world joint axes, s = J qd
template <int NV>
AGX_DEV void kinematics(TLane<NV>& d, int j, unsigned gm, const double* __restrict__ tm) {
  const bool live = j < NV;
  const int jj = live ? j : 0;
  d.par = live ? (int)tm[TF_PARENT * GW + jj] : -1;
  d.sub = live ? (unsigned)tm[TF_SUB * GW + jj] : 0u;
  d.anc = live ? (unsigned)tm[TF_ANC * GW + jj] : 0u;
  const bool revolute = (int)tm[TF_JTYPE * GW + jj] == 0;
  const double ax[3] = {tm[(TF_AXIS + 0) * GW + jj], tm[(TF_AXIS + 1) * GW + jj], tm[(TF_AXIS + 2) * GW + jj]};
  if (live) {
    double Rp[9], pp[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) Rp[k] = tm[(TF_RP + k) * GW + jj];
#pragma unroll
    for (int k = 0; k < 3; ++k) pp[k] = tm[(TF_PP + k) * GW + jj];
    if (revolute) {
      double sn, cs;
      AGX_SINCOS(d.q, &sn, &cs);
      const double vc = 1.0 - cs;
      double Rj[9];
      Rj[0] = cs + ax[0] * ax[0] * vc;         Rj[1] = ax[0] * ax[1] * vc - ax[2] * sn; Rj[2] = ax[0] * ax[2] * vc + ax[1] * sn;
      Rj[3] = ax[1] * ax[0] * vc + ax[2] * sn; Rj[4] = cs + ax[1] * ax[1] * vc;         Rj[5] = ax[1] * ax[2] * vc - ax[0] * sn;
      Rj[6] = ax[2] * ax[0] * vc - ax[1] * sn; Rj[7] = ax[2] * ax[1] * vc + ax[0] * sn; Rj[8] = cs + ax[2] * ax[2] * vc;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) d.R[3 * r + c] = Rp[3 * r] * Rj[c] + Rp[3 * r + 1] * Rj[3 + c] + Rp[3 * r + 2] * Rj[6 + c];
#pragma unroll
      for (int k = 0; k < 3; ++k) d.p[k] = pp[k];
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) d.R[k] = Rp[k];
#pragma unroll
      for (int r = 0; r < 3; ++r) d.p[r] = pp[r] + (Rp[3 * r] * ax[0] + Rp[3 * r + 1] * ax[1] + Rp[3 * r + 2] * ax[2]) * d.q;
    }
  } else {
    d.R[0] = 1; d.R[1] = 0; d.R[2] = 0; d.R[3] = 0; d.R[4] = 1; d.R[5] = 0; d.R[6] = 0; d.R[7] = 0; d.R[8] = 1;
    d.p[0] = d.p[1] = d.p[2] = 0;
  }
  // prefix product over the ancestors
  int a = d.par;
#pragma unroll
  for (int r = 0; r < TL<NV>::ROUNDS; ++r) {
    const int src = a >= 0 ? a : j;
    double o[12];
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = __shfl_sync(gm, d.R[k], src, GW);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[9 + k] = __shfl_sync(gm, d.p[k], src, GW);
    const int an = shfl_int(gm, a, src);
    if (a >= 0) {
      double Rn[9], pn[3];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          Rn[3 * rr + c] = o[3 * rr] * d.R[c] + o[3 * rr + 1] * d.R[3 + c] + o[3 * rr + 2] * d.R[6 + c];
        pn[rr] = o[9 + rr] + (o[3 * rr] * d.p[0] + o[3 * rr + 1] * d.p[1] + o[3 * rr + 2] * d.p[2]);
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) d.R[k] = Rn[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) d.p[k] = pn[k];
    }
    a = a >= 0 ? an : -1;
  }
  // world motion axis
  double z[3];
  mv3(d.R, ax, z);
  if (!live) {
#pragma unroll
    for (int k = 0; k < 6; ++k) d.J[k] = 0.0;
  } else if (revolute) {
    double pz[3];
    cross3(d.p, z, pz);
#pragma unroll
    for (int k = 0; k < 3; ++k) { d.J[k] = pz[k]; d.J[3 + k] = z[k]; }
  } else {
#pragma unroll
    for (int k = 0; k < 3; ++k) { d.J[k] = z[k]; d.J[3 + k] = 0.0; }
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) d.s[k] = d.J[k] * d.qd;
}

// ---------------------------------------------------------------- NV x NV Cholesky in registers (every lane, redundantly)
template <int NV>
AGX_DEV bool chol_registers(const double* sm_M, int stride, double* A, double* rinv) {
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i >= k) A[tidx(NV, i, k)] = sm_M[i * stride + k];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double dkk = A[tidx(NV, k, k)];
#pragma unroll
    for (int m = 0; m < NV; ++m)
      if (m < k) dkk -= A[tidx(NV, k, m)] * A[tidx(NV, k, m)];
    ok = ok && (dkk > 0.0);
    const double r = AGX_RSQRT(dkk);
    A[tidx(NV, k, k)] = dkk * r;
    rinv[k] = r;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (i > k) {
        double t = A[tidx(NV, i, k)];
#pragma unroll
        for (int m = 0; m < NV; ++m)
          if (m < k) t -= A[tidx(NV, i, m)] * A[tidx(NV, k, m)];
        A[tidx(NV, i, k)] = t * r;
      }
    }
  }
  return ok;
}
template <int NV>
AGX_DEV void chol_solve(const double* L, const double* rinv, double* r) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    double s = r[i];
#pragma unroll
    for (int m = 0; m < i; ++m) s -= L[tidx(NV, i, m)] * r[m];
    r[i] = s * rinv[i];
  }
#pragma unroll
  for (int i = NV - 1; i >= 0; --i) {
    double s = r[i];
#pragma unroll
    for (int m = i + 1; m < NV; ++m) s -= L[tidx(NV, m, i)] * r[m];
    r[i] = s * rinv[i];
  }
}

// ---------------------------------------------------------------- forward dynamics a = (M + armature)^-1 (u - nle)
// On exit d.qdd, the factor in registers (L, rinv) and, with DERIV, parked on board sc.
template <bool DERIV, int NV>
AGX_DEV bool forward_dynamics(TLane<NV>& d, int j, unsigned gm, const double* __restrict__ tm, double* sb, double* sc,
                              double* L, double* rinv) {
  constexpr int RD = TL<NV>::ROUNDS;
  const bool live = j < NV;
  const int jj = live ? j : 0;
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  const double agrav[6] = {-tm[TT_GRAV + 0], -tm[TT_GRAV + 1], -tm[TT_GRAV + 2], 0, 0, 0};
  scan_anc_excl<6, RD>(d.s, d.vp, zero6, d.par, j, gm);
  body_motion(d);
  {
    double mass = 0, com[3] = {0, 0, 0}, I6[6] = {0, 0, 0, 0, 0, 0};
    if (live) {
      mass = tm[TF_MASS * GW + jj];
#pragma unroll
      for (int k = 0; k < 3; ++k) com[k] = tm[(TF_COM + k) * GW + jj];
#pragma unroll
      for (int k = 0; k < 6; ++k) I6[k] = tm[(TF_INERTIA + k) * GW + jj];
    }
    body_inertia_from(d, mass, com, I6);
  }
  body_momentum(d, DERIV);
  scan_anc_excl<6, RD>(d.g, d.a0p, agrav, d.par, j, gm);
  body_force(d);
  if (DERIV) {
    // in three pieces: the accumulators of one 28-wide sum would double the live composites
    subtree_sum<10, NV, true>(d.Z, d.sub, gm);
    subtree_sum<12, NV, true>(d.Z + 10, d.sub, gm);
    subtree_sum<6, NV, true>(d.Z + 22, d.sub, gm);
  } else {
    subtree_sum<10, NV>(d.Z, d.sub, gm);
    subtree_sum<6, NV>(d.Z + 22, d.sub, gm);
  }
  column_terms<DERIV>(d, j, sb);
  sb[j * 18 + 16] = d.u;
  AGX_GSYNC();
  // column j of M + armature: entry (i, j) = J_i . dFda_j for i an ancestor (or j itself), dFda_i . J_j for i a descendant
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double* o = sb + i * 18;
    const double up = dot6(o, d.dFda);
    const double lo = dot6(o + 6, d.J);
    d.Mc[i] = ((d.anc >> i) & 1u) ? up : (((d.sub >> i) & 1u) ? lo : 0.0);
  }
  {
    const double arm = live ? tm[TF_ARM * GW + jj] : 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i == j) d.Mc[i] += arm;
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sc[i * CS + j] = d.Mc[i];
  }
  AGX_GSYNC();
  const bool ok = chol_registers<NV>(sc, CS, L, rinv);
  double rhs[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) rhs[i] = sb[i * 18 + 16] - sb[i * 18 + 15];
  chol_solve<NV>(L, rinv, rhs);
  d.qdd = 0.0;
#pragma unroll
  for (int i = 0; i < NV; ++i)
    if (i == j) d.qdd = rhs[i];
  if (DERIV) {
    // park the factor on the board so that its registers are free during the derivative phase (column k by lane k)
    AGX_GSYNC();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      if (k == j) {
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (i >= k) sc[i * CS + k] = L[tidx(NV, i, k)];
        sc[k * CS + GW] = rinv[k];
      }
    }
  }
  return ok;
}
// reload the parked factor (after a group barrier)
template <int NV>
AGX_DEV void factor_reload(const double* sc, double* L, double* rinv) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (i >= k) L[tidx(NV, i, k)] = sc[i * CS + k];
    rinv[k] = sc[k * CS + GW];
  }
}

// dtau/dq, dtau/dv columns (computeRNEADerivatives at the forward-dynamics acceleration)
template <int NV>
AGX_DEV void rnea_derivatives(TLane<NV>& d, int j, unsigned gm, const double* sb) {
  constexpr int RD = TL<NV>::ROUNDS;
  double jq[6], dap[6], da[6], dfc[6];
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < 6; ++k) jq[k] = d.J[k] * d.qdd;
  scan_anc_excl<6, RD>(jq, dap, zero6, d.par, j, gm);
#pragma unroll
  for (int k = 0; k < 6; ++k) da[k] = dap[k] + jq[k];
  inertia_apply(d.Y, da, dfc);
  subtree_sum<6, NV, true>(dfc, d.sub, gm);
  double dFdq[6], dFdv[6];
  deriv_columns(d, j, dap, dfc, dFdq, dFdv);  // leaves A_j in d.g
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double* o = sb + i * 18;  // J_i, dFda_i, BS_i
    const double uq = dot6(o, dFdq);
    const double uv = dot6(o, dFdv);
    const double lq = dot6(o + 6, d.g) + dot3(o + 12, d.c + 3);
    const double lv = 2.0 * dot6(o + 6, d.c) + dot3(o + 12, d.J + 3);
    const bool up = ((d.anc >> i) & 1u) != 0u, lo = ((d.sub >> i) & 1u) != 0u;
    d.tq[i] = up ? uq : (lo ? lq : 0.0);
    d.tv[i] = up ? uv : (lo ? lv : 0.0);
  }
}

// dynamics part of calc + calcDiff of a running node -> dynamics record
template <int NV>
AGX_DEV void node_dyn_diff(TLane<NV>& d, int j, unsigned gm, const double* __restrict__ tm, double dt, double* sb,
                           double* sc, double* __restrict__ rec) {
  using Lt = TL<NV>;
  const bool live = j < NV;
  double L[Lt::NTRI], rinv[NV];
  const bool ok = forward_dynamics<true, NV>(d, j, gm, tm, sb, sc, L, rinv);
  if (live) {
    rec[Lt::RK_QN * GW + j] = ok ? d.q + (d.qd * dt + d.qdd * (dt * dt)) : nan("");
    rec[Lt::RK_VN * GW + j] = ok ? d.qd + d.qdd * dt : nan("");
  }
  rnea_derivatives<NV>(d, j, gm, sb);
  AGX_GSYNC();
  factor_reload<NV>(sc, L, rinv);
  double col[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) col[i] = d.tq[i];
  chol_solve<NV>(L, rinv, col);
  if (live) {
#pragma unroll
    for (int i = 0; i < NV; ++i) rec[(Lt::RK_AQ + i) * GW + j] = -dt * col[i];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) col[i] = d.tv[i];
  chol_solve<NV>(L, rinv, col);
  if (live) {
#pragma unroll
    for (int i = 0; i < NV; ++i) rec[(Lt::RK_AV + i) * GW + j] = -dt * col[i];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) col[i] = (i == j) ? 1.0 : 0.0;
  chol_solve<NV>(L, rinv, col);
  if (live) {
#pragma unroll
    for (int i = 0; i < NV; ++i) rec[(Lt::RK_MI + i) * GW + j] = dt * col[i];
  }
  AGX_GSYNC();
}

// ---------------------------------------------------------------- costs of one node (after kinematics)
// Returns the unscaled node cost (identical on every lane).  DERIV: Gauss-Newton Lq_j, Lv_j, Lu_j and column j of
// Lqq.  terms (may be null): the per-cost view of agx_cost_terms, written by lane 0.
template <bool DERIV, int NV>
AGX_DEV double node_costs(const TLane<NV>& d, int j, unsigned gm, const double* __restrict__ tm,
                          const double* __restrict__ ref, bool terminal, double* srq, double* Lq, double* Lv, double* Lu,
                          double* Lqq, double* terms = nullptr) {
  constexpr int NX = 2 * NV;
  const bool live = j < NV;
  const int jj = live ? j : 0;
  const double rq = d.q - ref[jj], rv = d.qd - ref[NV + jj];
  const double wq = live ? ref[NX + jj] : 0.0, wv = live ? ref[NX + NV + jj] : 0.0;
  const double ru = d.u - ref[2 * NX + jj];
  const double wu = (live && !terminal) ? ref[2 * NX + NV + jj] : 0.0;
  const double cs = 0.5 * wq * rq * rq + 0.5 * wv * rv * rv, cu = 0.5 * wu * ru * ru;
  // frame placement: the world placement of the frame's parent joint is broadcast to the group
  const int fpar = (int)tm[TT_FP + 3];
  const unsigned fanc = (unsigned)tm[TF_ANC * GW + fpar];
  double R6[9], p6[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) R6[k] = __shfl_sync(gm, d.R[k], fpar, GW);
#pragma unroll
  for (int k = 0; k < 3; ++k) p6[k] = __shfl_sync(gm, d.p[k], fpar, GW);
  const double* Rref = ref + 2 * NX + 2 * NV;
  const double* pref = Rref + 9;
  const double* wp = pref + 3;
  double Rf[9], pf[3], r6[6], Jl[18];
  frame_residual_at(R6, p6, tm + TT_FR, tm + TT_FP, Rref, pref, Rf, pf, r6, DERIV ? Jl : nullptr);
  const bool tworld = tm[TT_COL + 6] != 0.0;  // ResidualModelFrameTranslation: linear part p_f - pref in the world
  if (tworld) {
#pragma unroll
    for (int k = 0; k < 3; ++k) r6[k] = pf[k] - pref[k];
  }
  double cpose = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) cpose += 0.5 * wp[k] * r6[k] * r6[k];
  const double cst = gsum(cs, gm), cct = gsum(cu, gm);
  double cost = (cst + cct) + cpose;
  // collision pairs
  double crq[MAX_PAIRS] = {0, 0}, g1[MAX_PAIRS] = {0, 0}, g2[MAX_PAIRS] = {0, 0}, ccost[MAX_PAIRS] = {0, 0},
         cdist[MAX_PAIRS] = {0, 0};
  const int npairs = (int)tm[TT_COL];
  if (npairs > 0) {
    const double alpha = tm[TT_COL + 1];
    const double* wc = wp + 6;
#pragma unroll
    for (int k = 0; k < MAX_PAIRS; ++k) {
      if (k < npairs) {  // group-uniform: every lane of the group reads the same model table
        double e[2][6];
        int jpar[2];
        double rad = 0.0;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          const int ic = (int)tm[TT_COL + 2 + 2 * k + side];
          const double* a = tm + TT_CAP + 8 * ic;
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
            const double v = __shfl_sync(gm, w[m], par < 0 ? 0 : par, GW);
            e[side][m] = par < 0 ? a[m] : v;
          }
        }
        double ca[3], cb[3], nn[3];
        const double len = segment_pair(e[0], e[0] + 3, e[1], e[1] + 3, ca, cb, nn);
        const double r = len - rad;
        double a, ar, arr;
        quadexp(r, alpha, a, ar, arr);
        ccost[k] = wc[k] * a;
        cdist[k] = r;
        cost += ccost[k];
        if (DERIV) {
          g1[k] = wc[k] * ar;
          g2[k] = wc[k] * arr;
          double wa[3], wb[3];
          cross3(d.J + 3, ca, wa);
          cross3(d.J + 3, cb, wb);
          double da = 0.0, db = 0.0;
#pragma unroll
          for (int m = 0; m < 3; ++m) { da += nn[m] * (wa[m] + d.J[m]); db += nn[m] * (wb[m] + d.J[m]); }
          const bool ma = jpar[0] >= 0 && (((unsigned)tm[TF_ANC * GW + (jpar[0] < 0 ? 0 : jpar[0])] >> j) & 1u);
          const bool mb = jpar[1] >= 0 && (((unsigned)tm[TF_ANC * GW + (jpar[1] < 0 ? 0 : jpar[1])] >> j) & 1u);
          crq[k] = (ma ? da : 0.0) - (mb ? db : 0.0);
        }
      }
    }
  }
  if (terms && j == 0) {
    terms[0] = cst; terms[1] = cct; terms[2] = cpose;
#pragma unroll
    for (int k = 0; k < 6; ++k) terms[3 + k] = r6[k];
#pragma unroll
    for (int k = 0; k < MAX_PAIRS; ++k) { terms[9 + k] = ccost[k]; terms[9 + MAX_PAIRS + k] = cdist[k]; }
  }
  if (DERIV) {
    // column j of the LOCAL frame Jacobian (oMf^-1 acting on the world axis J_j; zero for joints that do not move the
    // frame), then Rq[:, j] = Jlog6 * that
    double t[3], pw[3], cl[3], ca[3];
    cross3(pf, d.J + 3, pw);
    const double moves = ((fanc >> j) & 1u) ? 1.0 : 0.0;
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
    AGX_GSYNC();
    double wr[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) wr[k] = wp[k] * rqc[k];
    double lq = wq * rq;
#pragma unroll
    for (int k = 0; k < 6; ++k) lq += rqc[k] * (wp[k] * r6[k]);
#pragma unroll
    for (int k = 0; k < MAX_PAIRS; ++k) lq += g1[k] * crq[k];
    *Lq = lq;
    *Lv = wv * rv;
    *Lu = wu * ru;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double h = 0.0;
#pragma unroll
      for (int k = 0; k < 6; ++k) h += srq[i * 6 + k] * wr[k];
      Lqq[i] = h + ((i == j) ? wq : 0.0);
    }
    if (npairs > 0) {
#pragma unroll
      for (int k = 0; k < MAX_PAIRS; ++k) {
        const double gk = g2[k] * crq[k];
#pragma unroll
        for (int i = 0; i < NV; ++i) Lqq[i] += __shfl_sync(gm, crq[k], i, GW) * gk;
      }
    }
  }
  AGX_GSYNC();
  return cost;
}

// calc of one node: cost (scaled) and this lane's entries of xnext.  Returns false on failure.
template <int NV>
AGX_DEV bool node_calc(TLane<NV>& d, int j, unsigned gm, const double* __restrict__ tm, const double* __restrict__ ref,
                       double dt, bool terminal, double* brd, double* cost, double* qn, double* vn) {
  using Lt = TL<NV>;
  kinematics<NV>(d, j, gm, tm);
  const double l = node_costs<false, NV>(d, j, gm, tm, ref, terminal, brd + Lt::SQ, nullptr, nullptr, nullptr, nullptr);
  if (terminal) {
    *cost = l;
    *qn = d.q;
    *vn = d.qd;
    return true;
  }
  double L[Lt::NTRI], rinv[NV];
  const bool ok = forward_dynamics<false, NV>(d, j, gm, tm, brd + Lt::SB, brd + Lt::SC, L, rinv);
  *cost = dt * l;
  *qn = d.q + (d.qd * dt + d.qdd * (dt * dt));
  *vn = d.qd + d.qdd * dt;
  AGX_GSYNC();
  return ok;
}

AGX_DEV const double* tmodel_of(const Problem& P, int b) {
  return P.model + (P.n_models > 1 ? (size_t)b * TMODEL_SIZE : 0);
}

// ================================================================ kernels
// problem.calc + calcDiff: one group per (problem, node) -> dynamics record + cost record
// PART: 0 = cost record and dynamics record, 1 = cost record only, 2 = dynamics record only.  The solve launches the
// two halves as separate kernels: one kernel with both is 8 000 straight-line instructions per warp (130 KB of SASS),
// more than the instruction cache holds, and ran with `no_instruction` as its first stall reason.
template <int NV, int PART = 0>
__global__ void __launch_bounds__(64) tree_calc_diff_kernel(Problem P, const double* __restrict__ xs,
                                                           const double* __restrict__ us, const int32_t* __restrict__ cur,
                                                           const int32_t* __restrict__ recalc,
                                                           const int32_t* __restrict__ done, double* __restrict__ rec,
                                                           double* __restrict__ crec) {
  using Lt = TL<NV>;
  constexpr int NX = 2 * NV;
  AGX_SMEM(smem);
  AGX_TREE_SETUP();
  const int T1 = P.T + 1;
  if (ent >= (long long)P.B * T1) return;
  const int b = (int)(ent / T1), t = (int)(ent % T1);
  if (done && done[b]) return;
  if (recalc && !recalc[b]) return;
  double* brd = smem + grp_in_cta * Lt::BOARD;
  const size_t buf = (size_t)(cur ? (cur[b] & 1) : 0);
  const double* x = xs + ((buf * P.B + b) * T1 + t) * NX;
  const bool terminal = t == P.T, live = j < NV;
  const double* tm = tmodel_of(P, b);
  double* R = rec + (size_t)ent * Lt::REC;
  double* C = crec + (size_t)ent * Lt::CREC;
  const double* ref = P.refs + (size_t)ent * Lt::REF;
  if (PART != 2) {
    // the reference record is read late (weights, targets): ask for its lines now, one line per lane
    if (j * 16 < Lt::REF) AGX_PREFETCH(ref + j * 16);
  }
  TLane<NV> d;
  lane_load<NV>(d, j, x, terminal ? nullptr : us + ((buf * P.B + b) * P.T + t) * NV);
  kinematics<NV>(d, j, gm, tm);
  if (PART != 2) {
    const double s = terminal ? 1.0 : P.dts[t];
    double lq, lv, lu, Lqq[NV];
    const double l = node_costs<true, NV>(d, j, gm, tm, ref, terminal, brd + Lt::SQ, &lq, &lv, &lu, Lqq);
    if (live) {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (i >= j) C[Lt::CK_LQQ + tidx(NV, i, j)] = s * Lqq[i];
      C[Lt::CK_LVV + j] = s * ref[NX + NV + j];
      C[Lt::CK_LUU + j] = terminal ? 0.0 : s * ref[2 * NX + NV + j];
      C[Lt::CK_LQ + j] = s * lq;
      C[Lt::CK_LV + j] = s * lv;
      C[Lt::CK_LU + j] = s * lu;
    }
    if (j == 0) C[Lt::CK_COST] = s * l;
  }
  if (PART == 1) return;
  if (terminal) {
    if (live) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        R[(Lt::RK_AQ + i) * GW + j] = 0.0;
        R[(Lt::RK_AV + i) * GW + j] = 0.0;
        R[(Lt::RK_MI + i) * GW + j] = 0.0;
      }
      R[Lt::RK_QN * GW + j] = x[j];
      R[Lt::RK_VN * GW + j] = x[NV + j];
    }
    return;
  }
  node_dyn_diff<NV>(d, j, gm, tm, P.dts[t], brd + Lt::SB, brd + Lt::SC, R);
}

template <int NV>
__global__ void tree_calc_kernel(Problem P, const double* __restrict__ xs, const double* __restrict__ us,
                                 double* __restrict__ out_cost, double* __restrict__ out_xnext) {
  using Lt = TL<NV>;
  constexpr int NX = 2 * NV;
  AGX_SMEM(smem);
  AGX_TREE_SETUP();
  const int T1 = P.T + 1;
  if (ent >= (long long)P.B * T1) return;
  const int b = (int)(ent / T1), t = (int)(ent % T1);
  double* brd = smem + grp_in_cta * Lt::BOARD;
  const bool terminal = t == P.T;
  TLane<NV> d;
  lane_load<NV>(d, j, xs + (size_t)ent * NX, terminal ? nullptr : us + ((size_t)b * P.T + t) * NV);
  double c, qn, vn;
  const bool ok = node_calc<NV>(d, j, gm, tmodel_of(P, b), P.refs + (size_t)ent * Lt::REF, terminal ? 0.0 : P.dts[t],
                                terminal, brd, &c, &qn, &vn);
  if (!ok) c = nan("");
  if (out_cost && j == 0) out_cost[ent] = c;
  if (out_xnext && j < NV) {
    out_xnext[(size_t)ent * NX + j] = qn;
    out_xnext[(size_t)ent * NX + NV + j] = vn;
  }
}

// per-cost view: [state_reg, control_reg, goal_tracking, r6 (6), collision cost (2), distance (2)] per node
template <int NV>
__global__ void tree_cost_terms_kernel(Problem P, const double* __restrict__ xs, const double* __restrict__ us,
                                       double* __restrict__ out_terms) {
  using Lt = TL<NV>;
  constexpr int NX = 2 * NV;
  AGX_SMEM(smem);
  AGX_TREE_SETUP();
  const int T1 = P.T + 1;
  if (ent >= (long long)P.B * T1) return;
  const int b = (int)(ent / T1), t = (int)(ent % T1);
  double* brd = smem + grp_in_cta * Lt::BOARD;
  const bool terminal = t == P.T;
  const double* tm = tmodel_of(P, b);
  TLane<NV> d;
  lane_load<NV>(d, j, xs + (size_t)ent * NX, terminal ? nullptr : us + ((size_t)b * P.T + t) * NV);
  kinematics<NV>(d, j, gm, tm);
  node_costs<false, NV>(d, j, gm, tm, P.refs + (size_t)ent * Lt::REF, terminal, brd + Lt::SQ, nullptr, nullptr, nullptr,
                        nullptr, out_terms + (size_t)ent * N_COST_TERMS);
}

// dense view of the records (problem.calcDiff data); one thread per (node, row)
template <int NV>
__global__ void tree_expand_kernel(Problem P, const double* __restrict__ rec, const double* __restrict__ crec,
                                   double* out_cost, double* out_xnext, double* Fx, double* Fu, double* Lx, double* Lu,
                                   double* Lxx, double* Lxu, double* Luu) {
  using Lt = TL<NV>;
  constexpr int NX = 2 * NV;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int T1 = P.T + 1;
  const long long n = gid / NX;
  const int r = (int)(gid % NX);
  if (n >= (long long)P.B * T1) return;
  const int t = (int)(n % T1);
  const bool terminal = t == P.T;
  const double h = terminal ? 0.0 : P.dts[t];
  const double* R = rec + (size_t)n * Lt::REC;
  const double* C = crec + (size_t)n * Lt::CREC;
  const int i = r % NV;
  const bool top = r < NV;
  if (out_cost && r == 0) out_cost[n] = C[Lt::CK_COST];
  if (out_xnext) out_xnext[n * NX + r] = R[(top ? Lt::RK_QN : Lt::RK_VN) * GW + i];
  if (Lx) Lx[n * NX + r] = C[(top ? Lt::CK_LQ : Lt::CK_LV) + i];
  if (Lu && top) Lu[n * NV + i] = C[Lt::CK_LU + i];
  for (int c = 0; c < NV; ++c) {
    const double aq = R[(Lt::RK_AQ + i) * GW + c], av = R[(Lt::RK_AV + i) * GW + c], mi = R[(Lt::RK_MI + i) * GW + c];
    const double s = top ? h : 1.0;
    if (Fx) {
      Fx[(n * NX + r) * NX + c] = s * aq + ((top && c == i) ? 1.0 : 0.0);
      Fx[(n * NX + r) * NX + NV + c] = terminal ? ((!top && c == i) ? 1.0 : 0.0) : s * (av + ((c == i) ? 1.0 : 0.0));
    }
    if (Fu) Fu[(n * NX + r) * NV + c] = s * mi;
    if (Lxx) {
      Lxx[(n * NX + r) * NX + c] = top ? C[Lt::CK_LQQ + (i >= c ? tidx(NV, i, c) : tidx(NV, c, i))] : 0.0;
      Lxx[(n * NX + r) * NX + NV + c] = (!top && c == i) ? C[Lt::CK_LVV + i] : 0.0;
    }
    if (Lxu) Lxu[(n * NX + r) * NV + c] = 0.0;
    if (Luu && top) Luu[(n * NV + i) * NV + c] = (c == i) ? C[Lt::CK_LUU + i] : 0.0;
  }
}

// ---------------------------------------------------------------- Riccati sweep: one WARP per problem on the FP64
// tensor cores (mma.sync m8n8k4, SASS DMMA), any nv.  Same algebra as the chain kernels (agx_kernels.cuh):
// Fx = [I 0; 0 0] + S G, Fu = S N with G = [dt aq, I + dt av] (NV x 2NV), S = [dt I; I], N = dt Minv:
//   Z = S^T V', Vs = Z S, W = Vs G + [Zq 0], Qxx = Lxx + [V'qq 0; 0 0] + G^T W + [Zq^T G; 0], Qux = N^T W,
//   Quu = Luu + N^T Vs N, Qx = Lx + [v'q; 0] + G^T S^T v', Qu = Lu + N^T S^T v', Vxx = Qxx - Qux^T K.
// The state keeps its natural order [q; v]; every matrix lives on a zero-padded shared-memory board (NV -> KP = 4-
// multiple in the contraction dimension, RP = 8-multiple in rows; 2NV -> NP = 8-multiple) and every product is a loop
// of 8 x 8 x 4 tile products whose operands are read from the boards: lane T (g = T / 4, q = T % 4) holds
// a = A[g][4 kc + q], b = B[4 kc + q][g] and the accumulator pair C[g][2q], C[g][2q + 1].  nv = 9: 132 tile products
// per node.  The Quu Cholesky runs redundantly in registers; lanes 0..2NV each solve one column of [Qux | Qu].
template <int NV>
struct BWL {
  static constexpr int N = 2 * NV;
  static constexpr int KP = (NV + 3) & ~3;     // contraction length (multiple of 4)
  static constexpr int RP = (NV + 7) & ~7;     // rows of the NV-row boards (multiple of 8)
  static constexpr int NP = (N + 7) & ~7;      // padded state dimension
  static constexpr int KC = KP / 4, RT = RP / 8, NT = NP / 8;
  static constexpr int LDN = NP + 4, LDR = RP + 4;   // row strides (doubles)
  static constexpr int V = 0;                  // [NP][LDN]  V' (value Hessian of the next node)
  static constexpr int G = V + NP * LDN;       // [KP][LDN]
  static constexpr int NN = G + KP * LDN;      // [KP][LDR]  N
  static constexpr int Z = NN + KP * LDR;      // [RP][LDN]  Z, later [Qux]
  static constexpr int VS = Z + RP * LDN;      // [RP][LDR]  Vs, later Quu
  static constexpr int W = VS + RP * LDR;      // [RP][LDN]  W, later K
  static constexpr int VN = W + RP * LDN;      // [RP][LDR]
  static constexpr int VX = VN + RP * LDR;     // [NP]
  static constexpr int QX = VX + NP;           // [NP]
  static constexpr int FS = QX + NP;           // [NP]
  static constexpr int SV = FS + NP;           // [RP]
  static constexpr int QU = SV + RP;           // [RP]
  static constexpr int KF = QU + RP;           // [RP]  feed-forward k
  static constexpr int ZERO_END = (KF + RP + 1) & ~1;   // boards up to here are zero-padded work space
  static constexpr int CB = ZERO_END;          // [CREC]  cost record of the current node
  static constexpr int ST = CB + TL<NV>::CREC; // staged records of the NEXT node: dynamics, cost, gap row
  static constexpr int ST_REC = ST, ST_CREC = ST + TL<NV>::REC, ST_FS = ST_CREC + TL<NV>::CREC;
  static constexpr int SIZE = (ST_FS + N + 1) & ~1;
};

// asynchronous copy of node t's records (and gap row) into the stage buffer: 16-byte chunks over the 32 lanes
template <int NV>
AGX_DEV void tree_stage_node(double* sm, const double* __restrict__ rec, const double* __restrict__ crec,
                             const double* __restrict__ fs, bool gaps, int lane) {
  using B_ = BWL<NV>;
  constexpr int NREC = TL<NV>::REC / 2, NCREC = TL<NV>::CREC / 2, NFS = NV;  // 2 NV doubles = NV chunks
  for (int c = lane; c < NREC + NCREC + NFS; c += 32) {
    if (c < NREC) AGX_CP_ASYNC16(sm + B_::ST_REC + 2 * c, rec + 2 * c);
    else if (c < NREC + NCREC) AGX_CP_ASYNC16(sm + B_::ST_CREC + 2 * (c - NREC), crec + 2 * (c - NREC));
    else if (gaps) AGX_CP_ASYNC16(sm + B_::ST_FS + 2 * (c - NREC - NCREC), fs + 2 * (c - NREC - NCREC));
  }
  AGX_CP_ASYNC_COMMIT();
}

template <int NV>
__global__ void __launch_bounds__(32) tree_backward_kernel(Problem P, Work W, SolverState S, FddpOpts O) {
  using Lt = TL<NV>;
  using B_ = BWL<NV>;
  constexpr int N = 2 * NV, KC = B_::KC, RT = B_::RT, NT = B_::NT, LDN = B_::LDN, LDR = B_::LDR, NP = B_::NP, RP = B_::RP;
  AGX_SMEM(sm);
  const int lane = (int)(threadIdx.x & 31u);
  const int g = lane >> 2, q = lane & 3;
  const int b = (int)blockIdx.x;
  if (b >= P.B) return;
  if (S.done[b] || S.pending[b]) return;
  const int T = P.T, T1 = T + 1;
  const size_t buf = buf_of(S.cur, b, false);
  const double* xs = W.xs + (buf * P.B + b) * (size_t)T1 * N;
  const double* rec0 = W.rec + (size_t)b * T1 * Lt::REC;
  const double* crec0 = W.crec + (size_t)b * T1 * Lt::CREC;
  double* fsb = W.fs + (size_t)b * T1 * N;
  double* gvb = W.gv + (size_t)b * T1 * N;
  double* Kb = W.K + (size_t)b * T * NV * N;
  double* kb = W.k + (size_t)b * T * NV;
  const bool feasible = S.is_feasible[b] != 0;
  double xreg = S.xreg[b];

  double cost;
  {
    double part = 0.0;
    for (int t = lane; t <= T; t += 32) part += crec0[(size_t)t * Lt::CREC + Lt::CK_COST];
    cost = warp_sum(part);
  }
  if (!feasible && lane < N) {
    const int c = lane, jj = c < NV ? c : c - NV;
    const int rk = c < NV ? Lt::RK_QN : Lt::RK_VN;
    fsb[c] = W.x0[(size_t)b * N + c] - xs[c];
    for (int t = 0; t < T; ++t) fsb[(t + 1) * N + c] = rec0[(size_t)t * Lt::REC + rk * GW + jj] - xs[(t + 1) * N + c];
  }
  // the pads of every board stay zero for the whole sweep: only valid entries are ever rewritten
  for (int idx = lane; idx < B_::ZERO_END; idx += 32) sm[idx] = 0.0;
  __syncwarp();

  // tile products: acc += A(rows r0.., k) B(k, cols c0..); *_t: the A operand is stored transposed ([k][row])
  auto mma_nn = [&](double* acc, const double* A, int lda, int r0, const double* Bm, int ldb, int c0) {
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      const double a = A[(r0 + g) * lda + 4 * kc + q], bb = Bm[(4 * kc + q) * ldb + c0 + g];
      AGX_DMMA(acc[0], acc[1], a, bb, acc[0], acc[1]);
    }
  };
  auto mma_tn = [&](double* acc, const double* A, int lda, int r0, const double* Bm, int ldb, int c0, double sgn, int row_lim) {
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      const double a = (r0 + g < row_lim) ? sgn * A[(4 * kc + q) * lda + r0 + g] : 0.0, bb = Bm[(4 * kc + q) * ldb + c0 + g];
      AGX_DMMA(acc[0], acc[1], a, bb, acc[0], acc[1]);
    }
  };
  auto store_tile = [&](double* C, int ldc, int r0, int c0, const double* acc) {
    C[(r0 + g) * ldc + c0 + 2 * q] = acc[0];
    C[(r0 + g) * ldc + c0 + 2 * q + 1] = acc[1];
  };

  bool failed = !(cost == cost);
  double dg = 0.0, dq = 0.0;
  for (;;) {
    bool ok = !failed;
    double dgp = 0.0, dqp = 0.0;
    if (ok) {
      // the first running node's records start flowing into shared memory while the terminal node is handled
      tree_stage_node<NV>(sm, rec0 + (size_t)(T - 1) * Lt::REC, crec0 + (size_t)(T - 1) * Lt::CREC,
                          fsb + (size_t)(T - 1) * N, !feasible, lane);
      // ---- terminal node: Vxx = Lxx (+ xreg), Vx = Lx (+ Vxx fs)
      const double* C = crec0 + (size_t)T * Lt::CREC;
      for (int idx = lane; idx < N * N; idx += 32) {
        const int r = idx / N, c = idx % N;
        double v = 0.0;
        if (r < NV && c < NV) v = C[Lt::CK_LQQ + (r >= c ? tidx(NV, r, c) : tidx(NV, c, r))];
        else if (r == c) v = C[Lt::CK_LVV + r - NV];
        if (r == c) v += xreg;
        sm[B_::V + r * LDN + c] = v;
      }
      if (lane < N) {
        sm[B_::VX + lane] = lane < NV ? C[Lt::CK_LQ + lane] : C[Lt::CK_LV + lane - NV];
        sm[B_::FS + lane] = feasible ? 0.0 : fsb[T * N + lane];
      }
      __syncwarp();
      if (!feasible) {
        if (lane < N) {
          double gg = 0.0;
          const double f = sm[B_::FS + lane];
          for (int c = 0; c < N; ++c) gg += sm[B_::V + lane * LDN + c] * sm[B_::FS + c];
          const double vx = sm[B_::VX + lane] + gg;
          sm[B_::VX + lane] = vx;
          gvb[T * N + lane] = gg;
          dgp -= vx * f;
          dqp += gg * f;
        }
        __syncwarp();
      }
    }
    for (int t = T - 1; ok && t >= 0; --t) {
      const double h = P.dts[t];
      // node t's records were staged during the previous node; they are unpacked onto the boards, then the next
      // node's records start flowing in behind the arithmetic
      AGX_CP_ASYNC_WAIT_ALL();
      __syncwarp();
      const double* R = sm + B_::ST_REC;
      const double* C = sm + B_::CB;
      for (int idx = lane; idx < Lt::CREC; idx += 32) sm[B_::CB + idx] = sm[B_::ST_CREC + idx];
      // node operands: G = [dt aq, I + dt av], N = dt Minv; Z = S^T V', sv = S^T v'
      for (int idx = lane; idx < NV * N; idx += 32) {
        const int i = idx / N, c = idx % N;
        sm[B_::G + i * LDN + c] = c < NV ? R[(Lt::RK_AQ + i) * GW + c] : R[(Lt::RK_AV + i) * GW + (c - NV)] + ((c - NV == i) ? 1.0 : 0.0);
        sm[B_::Z + i * LDN + c] = h * sm[B_::V + i * LDN + c] + sm[B_::V + (NV + i) * LDN + c];
      }
      for (int idx = lane; idx < NV * NV; idx += 32) {
        const int i = idx / NV, c = idx % NV;
        sm[B_::NN + i * LDR + c] = R[(Lt::RK_MI + i) * GW + c];
      }
      if (lane < N) sm[B_::FS + lane] = feasible ? 0.0 : sm[B_::ST_FS + lane];
      if (lane < NV) sm[B_::SV + lane] = h * sm[B_::VX + lane] + sm[B_::VX + NV + lane];
      __syncwarp();
      if (t > 0)
        tree_stage_node<NV>(sm, rec0 + (size_t)(t - 1) * Lt::REC, crec0 + (size_t)(t - 1) * Lt::CREC,
                            fsb + (size_t)(t - 1) * N, !feasible, lane);
      for (int idx = lane; idx < NV * NV; idx += 32) {
        const int i = idx / NV, m = idx % NV;
        sm[B_::VS + i * LDR + m] = h * sm[B_::Z + i * LDN + m] + sm[B_::Z + i * LDN + NV + m];
      }
      __syncwarp();
      // W = Vs G + [Zq 0], VN = Vs N
#pragma unroll
      for (int rt = 0; rt < RT; ++rt) {
#pragma unroll
        for (int ct = 0; ct < NT; ++ct) {
          double acc[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 8 * ct + 2 * q + e;
            acc[e] = c < NV ? sm[B_::Z + (8 * rt + g) * LDN + c] : 0.0;
          }
          mma_nn(acc, sm + B_::VS, LDR, 8 * rt, sm + B_::G, LDN, 8 * ct);
          store_tile(sm + B_::W, LDN, 8 * rt, 8 * ct, acc);
        }
#pragma unroll
        for (int ct = 0; ct < RT; ++ct) {
          double acc[2] = {0.0, 0.0};
          mma_nn(acc, sm + B_::VS, LDR, 8 * rt, sm + B_::NN, LDR, 8 * ct);
          store_tile(sm + B_::VN, LDR, 8 * rt, 8 * ct, acc);
        }
      }
      __syncwarp();
      // Qxx = Lxx + [V'qq 0; 0 0] + G^T W + [Zq^T G; 0] (kept in registers), Qx
      double Qt[NT][NT][2];
#pragma unroll
      for (int rt = 0; rt < NT; ++rt)
#pragma unroll
        for (int ct = 0; ct < NT; ++ct) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = 8 * rt + g, c = 8 * ct + 2 * q + e;
            double a = 0.0;
            if (r < NV && c < NV) a = C[Lt::CK_LQQ + (r >= c ? tidx(NV, r, c) : tidx(NV, c, r))] + sm[B_::V + r * LDN + c];
            else if (r == c && r < N) a = C[Lt::CK_LVV + r - NV];
            Qt[rt][ct][e] = a;
          }
          mma_tn(Qt[rt][ct], sm + B_::G, LDN, 8 * rt, sm + B_::W, LDN, 8 * ct, 1.0, N);
          if (8 * rt < NV) mma_tn(Qt[rt][ct], sm + B_::Z, LDN, 8 * rt, sm + B_::G, LDN, 8 * ct, 1.0, NV);
        }
      if (lane < N) {
        const int r = lane;
        double a = r < NV ? C[Lt::CK_LQ + r] + sm[B_::VX + r] : C[Lt::CK_LV + r - NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) a += sm[B_::G + i * LDN + r] * sm[B_::SV + i];
        sm[B_::QX + r] = a;
      }
      if (lane < NV) {
        double a = C[Lt::CK_LU + lane];
#pragma unroll
        for (int m = 0; m < NV; ++m) a += sm[B_::NN + m * LDR + lane] * sm[B_::SV + m];
        sm[B_::QU + lane] = a;
      }
      __syncwarp();  // every lane is done reading Z and Vs: Qux and Quu take their boards
      // Qux = N^T W -> Z board, Quu = Luu + N^T VN (+ ureg) -> Vs board
#pragma unroll
      for (int rt = 0; rt < RT; ++rt) {
#pragma unroll
        for (int ct = 0; ct < NT; ++ct) {
          double acc[2] = {0.0, 0.0};
          mma_tn(acc, sm + B_::NN, LDR, 8 * rt, sm + B_::W, LDN, 8 * ct, 1.0, NV);
          store_tile(sm + B_::Z, LDN, 8 * rt, 8 * ct, acc);
        }
#pragma unroll
        for (int ct = 0; ct < RT; ++ct) {
          double acc[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int i = 8 * rt + g, c = 8 * ct + 2 * q + e;
            acc[e] = (i == c && i < NV) ? C[Lt::CK_LUU + i] + xreg : 0.0;
          }
          mma_tn(acc, sm + B_::NN, LDR, 8 * rt, sm + B_::VN, LDR, 8 * ct, 1.0, NV);
          store_tile(sm + B_::VS, LDR, 8 * rt, 8 * ct, acc);
        }
      }
      __syncwarp();
      // computeGains: every lane factors Quu in registers; lanes 0..N solve one column of [Qux | Qu] each
      {
        double L[Lt::NTRI], rinv[NV];
        ok = chol_registers<NV>(sm + B_::VS, LDR, L, rinv);
        if (!ok) break;
        if (lane <= N) {
          const int c = lane;
          double col[NV];
#pragma unroll
          for (int i = 0; i < NV; ++i) col[i] = c < N ? sm[B_::Z + i * LDN + c] : sm[B_::QU + i];
          chol_solve<NV>(L, rinv, col);
          if (c < N) {
#pragma unroll
            for (int i = 0; i < NV; ++i) { sm[B_::W + i * LDN + c] = col[i]; Kb[(t * NV + i) * N + c] = col[i]; }
#pragma unroll
            for (int i = NV; i < RP; ++i) sm[B_::W + i * LDN + c] = 0.0;
          } else {
            double qk = 0.0;
#pragma unroll
            for (int i = 0; i < NV; ++i) { sm[B_::KF + i] = col[i]; kb[t * NV + i] = col[i]; qk += sm[B_::QU + i] * col[i]; }
            dgp += qk;   // Qu . k
            dqp -= qk;   // k . Quu k = k . Qu
          }
        } else if (lane < NP + 1) {
          // pad columns of the K board (it held W): back to zero
          const int c = lane - 1;
#pragma unroll
          for (int i = 0; i < RP; ++i) sm[B_::W + i * LDN + c] = 0.0;
        }
      }
      __syncwarp();
      // Vxx = Qxx - Qux^T K (unsymmetrised, in registers), Vx = Qx - K^T Qu
#pragma unroll
      for (int rt = 0; rt < NT; ++rt)
#pragma unroll
        for (int ct = 0; ct < NT; ++ct) mma_tn(Qt[rt][ct], sm + B_::Z, LDN, 8 * rt, sm + B_::W, LDN, 8 * ct, -1.0, N);
      double nvx = 0.0;
      if (lane < N) {
        nvx = sm[B_::QX + lane];
#pragma unroll
        for (int i = 0; i < NV; ++i) nvx -= sm[B_::W + i * LDN + lane] * sm[B_::QU + i];
      }
      // symmetrise through the V board: unsymmetrised tiles out, (r, c) and (c, r) back in
#pragma unroll
      for (int rt = 0; rt < NT; ++rt)
#pragma unroll
        for (int ct = 0; ct < NT; ++ct) store_tile(sm + B_::V, LDN, 8 * rt, 8 * ct, Qt[rt][ct]);
      __syncwarp();
#pragma unroll
      for (int rt = 0; rt < NT; ++rt)
#pragma unroll
        for (int ct = 0; ct < NT; ++ct)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int r = 8 * rt + g, c = 8 * ct + 2 * q + e;
            const double tv = sm[B_::V + c * LDN + r];
            Qt[rt][ct][e] = (r < N && c < N) ? 0.5 * (Qt[rt][ct][e] + tv) + ((r == c) ? xreg : 0.0) : 0.0;
          }
      __syncwarp();
#pragma unroll
      for (int rt = 0; rt < NT; ++rt)
#pragma unroll
        for (int ct = 0; ct < NT; ++ct) store_tile(sm + B_::V, LDN, 8 * rt, 8 * ct, Qt[rt][ct]);
      if (lane < N) sm[B_::VX + lane] = nvx;
      __syncwarp();
      if (!feasible) {
        if (lane < N) {
          double gg = 0.0;
          for (int c = 0; c < N; ++c) gg += sm[B_::V + lane * LDN + c] * sm[B_::FS + c];
          const double f = sm[B_::FS + lane];
          const double vx = sm[B_::VX + lane] + gg;
          sm[B_::VX + lane] = vx;
          gvb[t * N + lane] = gg;
          dgp -= vx * f;
          dqp += gg * f;
        }
      }
      __syncwarp();
    }
    if (ok) {
      double chk = lane < N ? sm[B_::VX + lane] : 0.0;
      for (int idx = lane; idx < N * N; idx += 32) chk += sm[B_::V + (idx / N) * LDN + (idx % N)];
      chk = warp_sum(chk);
      if (!(chk - chk == 0.0)) ok = false;
    }
    __syncwarp();
    if (ok) {
      dg = warp_sum(dgp);
      dq = warp_sum(dqp);
      break;
    }
    failed = false;
    xreg *= O.reg_incfactor;
    if (xreg > O.reg_max) xreg = O.reg_max;
    if (xreg == O.reg_max) {
      if (lane == 0) { S.status[b] = 2; S.done[b] = 1; }
      break;
    }
    // a restarted sweep starts from clean boards (a failed one may have left non-finite numbers in the pads)
    AGX_CP_ASYNC_WAIT_ALL();
    __syncwarp();
    for (int idx = lane; idx < B_::ZERO_END; idx += 32) sm[idx] = 0.0;
    __syncwarp();
  }
  if (lane == 0) {
    S.xreg[b] = xreg;
    S.cost[b] = cost;
    S.dg[b] = dg;
    S.dq[b] = dq;
  }
}

// ---------------------------------------------------------------- forward pass: one group per problem runs the whole
// line search (SolverFDDP::forwardPass / tryStep / expectedImprovement and the tail of the solve loop) with the
// node costs evaluated in line.
template <int NV>
__global__ void tree_forward_kernel(Problem P, Work W, SolverState S, FddpOpts O) {
  using Lt = TL<NV>;
  constexpr int NX = 2 * NV;
  AGX_SMEM(smem);
  AGX_TREE_SETUP();
  const int b = (int)ent;
  if (b >= P.B) return;
  if (S.done[b]) return;
  double* brd = smem + grp_in_cta * Lt::BOARD;
  double* sdx = brd + Lt::SX;
  const int T = P.T, T1 = T + 1;
  const bool live = j < NV;
  const int jj = live ? j : 0;
  const double* tm = tmodel_of(P, b);
  const size_t buf = buf_of(S.cur, b, false), obuf = buf ^ 1;
  const double* xs = W.xs + (buf * P.B + b) * (size_t)T1 * NX;
  const double* us = W.us + (buf * P.B + b) * (size_t)T * NV;
  double* xt = W.xs + (obuf * P.B + b) * (size_t)T1 * NX;
  double* ut = W.us + (obuf * P.B + b) * (size_t)T * NV;
  const double* fsb = W.fs + (size_t)b * T1 * NX;
  const double* gvb = W.gv + (size_t)b * T1 * NX;
  const double* Kb = W.K + (size_t)b * T * NV * NX;
  const double* kb = W.k + (size_t)b * T * NV;
  const double* refs = P.refs + (size_t)b * T1 * Lt::REF;
  const bool feasible = S.is_feasible[b] != 0;
  const double cost = S.cost[b], dg = S.dg[b], dq = S.dq[b];
  const double x0q = live ? W.x0[(size_t)b * NX + jj] : 0.0, x0v = live ? W.x0[(size_t)b * NX + NV + jj] : 0.0;
  double steplength = 1.0, cost_try = 0.0, stop = S.stop[b];
  bool accepted = false;
  AGX_GSYNC();  // every lane has read the state before lane 0 updates it
  for (int ia = 0; ia < O.n_alphas; ++ia) {
    steplength = ldexp(1.0, -ia);
    const bool contract = !feasible && steplength != 1.0;
    double xq = x0q, xv = x0v;
    double ctry = 0.0, dvp = 0.0;
    bool ok = true;
    for (int t = 0; t <= T; ++t) {
      if (t < T) {
        // the next node's operands are asked for now: this loop is one dependent chain per node
        const int tn = t + 1;
        AGX_PREFETCH(xs + tn * NX + jj); AGX_PREFETCH(xs + tn * NX + NV + jj);
        if (!feasible) { AGX_PREFETCH(fsb + tn * NX + jj); AGX_PREFETCH(gvb + tn * NX + jj); AGX_PREFETCH(gvb + tn * NX + NV + jj); }
        AGX_PREFETCH(refs + (size_t)tn * Lt::REF + 8 * j);
        if (tn < T) {
          AGX_PREFETCH(Kb + ((size_t)tn * NV + jj) * NX); AGX_PREFETCH(Kb + ((size_t)tn * NV + jj) * NX + NX - 1);
          AGX_PREFETCH(us + tn * NV + jj); AGX_PREFETCH(kb + tn * NV + jj);
        }
      }
      double tq = xq, tv = xv;
      if (contract && live) {
        tq += fsb[t * NX + j] * (steplength - 1.0);
        tv += fsb[t * NX + NV + j] * (steplength - 1.0);
      }
      const double dxq = live ? tq - xs[t * NX + jj] : 0.0, dxv = live ? tv - xs[t * NX + NV + jj] : 0.0;
      if (live) {
        xt[t * NX + j] = tq;
        xt[t * NX + NV + j] = tv;
        if (!feasible) dvp += gvb[t * NX + j] * dxq + gvb[t * NX + NV + j] * dxv;
      }
      TLane<NV> d;
      d.q = tq; d.qd = tv; d.u = 0.0; d.qdd = 0.0;
      const bool terminal = t == T;
      if (!terminal) {
        if (live) { sdx[j] = dxq; sdx[NV + j] = dxv; }
        AGX_GSYNC();
        if (live) {
          double s = 0.0;
          for (int m = 0; m < NX; ++m) s += Kb[(t * NV + j) * NX + m] * sdx[m];
          d.u = us[t * NV + j] - kb[t * NV + j] * steplength - s;
          ut[t * NV + j] = d.u;
        }
      }
      double c, qn, vn;
      const bool okn = node_calc<NV>(d, j, gm, tm, refs + (size_t)t * Lt::REF, terminal ? 0.0 : P.dts[t], terminal, brd,
                                     &c, &qn, &vn);
      ok = ok && okn;
      ctry += c;
      xq = qn; xv = vn;
      if (!(ctry - ctry == 0.0)) { ok = false; break; }  // NaN / inf: reject this step length (group-uniform)
    }
    if (!ok) continue;
    cost_try = ctry;
    const double dv = feasible ? 0.0 : gsum(dvp, gm);
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

// problem.rollout(us)
template <int NV>
__global__ void tree_rollout_kernel(Problem P, const double* __restrict__ x0, const double* __restrict__ us,
                                    double* __restrict__ out_xs) {
  using Lt = TL<NV>;
  constexpr int NX = 2 * NV;
  AGX_SMEM(smem);
  AGX_TREE_SETUP();
  const int b = (int)ent;
  if (b >= P.B) return;
  double* brd = smem + grp_in_cta * Lt::BOARD;
  const int T = P.T, T1 = T + 1;
  const bool live = j < NV;
  const int jj = live ? j : 0;
  const double* tm = tmodel_of(P, b);
  double xq = live ? x0[(size_t)b * NX + jj] : 0.0, xv = live ? x0[(size_t)b * NX + NV + jj] : 0.0;
  double* xo = out_xs + (size_t)b * T1 * NX;
  if (live) { xo[j] = xq; xo[NV + j] = xv; }
  for (int t = 0; t < T; ++t) {
    TLane<NV> d;
    d.q = xq; d.qd = xv; d.u = live ? us[((size_t)b * T + t) * NV + jj] : 0.0; d.qdd = 0.0;
    kinematics<NV>(d, j, gm, tm);
    double L[Lt::NTRI], rinv[NV];
    const bool ok = forward_dynamics<false, NV>(d, j, gm, tm, brd + Lt::SB, brd + Lt::SC, L, rinv);
    const double dt = P.dts[t];
    xq = ok ? d.q + (d.qd * dt + d.qdd * (dt * dt)) : nan("");
    xv = ok ? d.qd + d.qdd * dt : nan("");
    if (live) { xo[(t + 1) * NX + j] = xq; xo[(t + 1) * NX + NV + j] = xv; }
    AGX_GSYNC();
  }
}

// IntegratedActionModelEuler.calc -> xnext for n independent (x, u) pairs; model m_b = models[b] when per_row
template <int NV>
__global__ void tree_integrate_kernel(const double* __restrict__ models, int per_row, const double* __restrict__ x,
                                      const double* __restrict__ u, double dt, int n, double* __restrict__ out) {
  using Lt = TL<NV>;
  constexpr int NX = 2 * NV;
  AGX_SMEM(smem);
  AGX_TREE_SETUP();
  if (ent >= n) return;
  double* brd = smem + grp_in_cta * Lt::BOARD;
  const double* tm = models + (per_row ? (size_t)ent * TMODEL_SIZE : 0);
  TLane<NV> d;
  lane_load<NV>(d, j, x + (size_t)ent * NX, u + (size_t)ent * NV);
  kinematics<NV>(d, j, gm, tm);
  double L[Lt::NTRI], rinv[NV];
  const bool ok = forward_dynamics<false, NV>(d, j, gm, tm, brd + Lt::SB, brd + Lt::SC, L, rinv);
  if (j < NV) {
    out[(size_t)ent * NX + j] = ok ? d.q + (d.qd * dt + d.qdd * (dt * dt)) : nan("");
    out[(size_t)ent * NX + NV + j] = ok ? d.qd + d.qdd * dt : nan("");
  }
}

// pin.rnea(q, v, a): tau = nle(q, v) + M(q) a (no armature)
template <int NV>
__global__ void tree_rnea_kernel(const double* __restrict__ models, int per_row, const double* __restrict__ q,
                                 const double* __restrict__ v, const double* __restrict__ a, int n,
                                 double* __restrict__ out_tau) {
  constexpr int RD = TL<NV>::ROUNDS;
  AGX_TREE_SETUP();
  if (ent >= n) return;
  const bool live = j < NV;
  const int jj = live ? j : 0;
  const double* tm = models + (per_row ? (size_t)ent * TMODEL_SIZE : 0);
  TLane<NV> d;
  d.q = live ? q[(size_t)ent * NV + jj] : 0.0;
  d.qd = live ? v[(size_t)ent * NV + jj] : 0.0;
  d.u = 0.0;
  d.qdd = live ? a[(size_t)ent * NV + jj] : 0.0;
  kinematics<NV>(d, j, gm, tm);
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  const double agrav[6] = {-tm[TT_GRAV + 0], -tm[TT_GRAV + 1], -tm[TT_GRAV + 2], 0, 0, 0};
  scan_anc_excl<6, RD>(d.s, d.vp, zero6, d.par, j, gm);
  body_motion(d);
  {
    double mass = 0, com[3] = {0, 0, 0}, I6[6] = {0, 0, 0, 0, 0, 0};
    if (live) {
      mass = tm[TF_MASS * GW + jj];
#pragma unroll
      for (int k = 0; k < 3; ++k) com[k] = tm[(TF_COM + k) * GW + jj];
#pragma unroll
      for (int k = 0; k < 6; ++k) I6[k] = tm[(TF_INERTIA + k) * GW + jj];
    }
    body_inertia_from(d, mass, com, I6);
  }
  body_momentum(d, false);
#pragma unroll
  for (int k = 0; k < 6; ++k) d.g[k] += d.J[k] * d.qdd;
  scan_anc_excl<6, RD>(d.g, d.a0p, agrav, d.par, j, gm);
  body_force(d);
  subtree_sum<6, NV>(d.Z + 22, d.sub, gm);
  if (live) out_tau[(size_t)ent * NV + j] = dot6(d.J, d.Z + 22);
}

// warm start by shifting the previous solution by the first time step (warm_start_shift_previous_solution.py:85-104)
template <int NV>
__global__ void tree_shift_kernel(Problem P, const double* __restrict__ xs, const double* __restrict__ us,
                                  double* __restrict__ out_xs, double* __restrict__ out_us) {
  using Lt = TL<NV>;
  constexpr int NX = 2 * NV;
  AGX_SMEM(smem);
  AGX_TREE_SETUP();
  const int T = P.T, T1 = T + 1;
  if (ent >= (long long)P.B * T1) return;
  const int b = (int)(ent / T1), i = (int)(ent % T1);
  double* brd = smem + grp_in_cta * Lt::BOARD;
  const bool live = j < NV;
  const double* xb = xs + (size_t)b * T1 * NX;
  const double* ub = us + (size_t)b * T * NV;
  double* xo = out_xs + ((size_t)b * T1 + i) * NX;
  if (i == T) {
    if (live) { xo[j] = xb[T * NX + j]; xo[NV + j] = xb[T * NX + NV + j]; }
    return;
  }
  double* uo = out_us + ((size_t)b * T + i) * NV;
  const double dt0 = P.dts[0];
  if (P.dts[i] == dt0) {
    if (live) {
      xo[j] = xb[(i + 1) * NX + j];
      xo[NV + j] = xb[(i + 1) * NX + NV + j];
      uo[j] = ub[(i < T - 1 ? i + 1 : i) * NV + j];
    }
    return;
  }
  const double* tm = tmodel_of(P, b);
  TLane<NV> d;
  lane_load<NV>(d, j, xb + (size_t)i * NX, ub + (size_t)i * NV);
  kinematics<NV>(d, j, gm, tm);
  double L[Lt::NTRI], rinv[NV];
  const bool ok = forward_dynamics<false, NV>(d, j, gm, tm, brd + Lt::SB, brd + Lt::SC, L, rinv);
  if (live) {
    xo[j] = ok ? d.q + (d.qd * dt0 + d.qdd * (dt0 * dt0)) : nan("");
    xo[NV + j] = ok ? d.qd + d.qdd * dt0 : nan("");
    uo[j] = d.u;
  }
}

// ---------------------------------------------------------------- SQP mode on general trees (agx_sqp.cuh restated for
// 16-lane groups): the linear rollout of the QP solution, the QP multipliers and the KKT norm.  Lane j owns row j.
AGX_DEV double gmax(double x, unsigned gm) {
  x = fmax(x, __shfl_xor_sync(gm, x, 1, GW));
  x = fmax(x, __shfl_xor_sync(gm, x, 2, GW));
  x = fmax(x, __shfl_xor_sync(gm, x, 4, GW));
  x = fmax(x, __shfl_xor_sync(gm, x, 8, GW));
  return x;
}

template <int NV>
__global__ void tree_sqp_direction_kernel(Problem P, Work W, SolverState S, SqpOpts Q, int32_t* __restrict__ pend) {
  using Lt = TL<NV>;
  constexpr int NX = 2 * NV;
  AGX_TREE_SETUP();
  const int b = (int)ent;
  if (b >= P.B) return;
  if (S.done[b]) return;
  const int T = P.T, T1 = T + 1;
  const bool live = j < NV;
  const int jj = live ? j : 0;
  const double* fsb = W.fs + (size_t)b * T1 * NX;
  double* dxb = W.gv + (size_t)b * T1 * NX;
  const double* Kb = W.K + (size_t)b * T * NV * NX;
  double* kb = W.k + (size_t)b * T * NV;
  const double* rec0 = W.rec + (size_t)b * T1 * Lt::REC;
  const double* crec0 = W.crec + (size_t)b * T1 * Lt::CREC;

  double dq = live ? fsb[jj] : 0.0, dv = live ? fsb[NV + jj] : 0.0;
  double gl1 = fabs(dq) + fabs(dv), ginf = fmax(fabs(dq), fabs(dv));
  for (int t = 0; t < T; ++t) {
    const double* Kr = Kb + ((size_t)t * NV + jj) * NX;
    const double* R = rec0 + (size_t)t * Lt::REC;
    if (live) { dxb[t * NX + j] = dq; dxb[t * NX + NV + j] = dv; }
    double dqm[NV], dvm[NV];
#pragma unroll
    for (int m = 0; m < NV; ++m) { dqm[m] = __shfl_sync(gm, dq, m, GW); dvm[m] = __shfl_sync(gm, dv, m, GW); }
    double s = -kb[t * NV + jj];
#pragma unroll
    for (int m = 0; m < NV; ++m) s -= Kr[m] * dqm[m] + Kr[NV + m] * dvm[m];
    const double du = live ? s : 0.0;
    if (live) kb[t * NV + j] = du;
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < NV; ++m) {
      const double dum = __shfl_sync(gm, du, m, GW);
      acc += R[(Lt::RK_AQ + jj) * GW + m] * dqm[m] + R[(Lt::RK_AV + jj) * GW + m] * dvm[m] + R[(Lt::RK_MI + jj) * GW + m] * dum;
    }
    const double fq = live ? fsb[(t + 1) * NX + jj] : 0.0, fv = live ? fsb[(t + 1) * NX + NV + jj] : 0.0;
    const double dt = P.dts[t];
    const double dvn = dv + acc + fv;
    const double dqn = dq + dt * (dv + acc) + fq;
    gl1 += fabs(fq) + fabs(fv);
    ginf = fmax(ginf, fmax(fabs(fq), fabs(fv)));
    dq = live ? dqn : 0.0;
    dv = live ? dvn : 0.0;
  }
  if (live) { dxb[T * NX + j] = dq; dxb[T * NX + NV + j] = dv; }
  // multipliers and stationarity
  double lq, lv, kkt;
  {
    const double* C = crec0 + (size_t)T * Lt::CREC;
    double hq = 0.0;
#pragma unroll
    for (int m = 0; m < NV; ++m)
      hq += C[Lt::CK_LQQ + (jj >= m ? tidx(NV, jj, m) : tidx(NV, m, jj))] * __shfl_sync(gm, dq, m, GW);
    const double hv = C[Lt::CK_LVV + jj] * dv;
    lq = C[Lt::CK_LQ + jj] + hq;
    lv = C[Lt::CK_LV + jj] + hv;
    kkt = live ? fmax(fabs(hq), fabs(hv)) : 0.0;
  }
  for (int t = T - 1; t >= 0; --t) {
    const double* R = rec0 + (size_t)t * Lt::REC;
    const double* C = crec0 + (size_t)t * Lt::CREC;
    const double dt = P.dts[t];
    const double w = live ? dt * lq + lv : 0.0;
    double su = C[Lt::CK_LU + jj], aq = 0.0, av = 0.0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const double wi = __shfl_sync(gm, w, i, GW);
      su += R[(Lt::RK_MI + i) * GW + jj] * wi;
      aq += R[(Lt::RK_AQ + i) * GW + jj] * wi;
      av += R[(Lt::RK_AV + i) * GW + jj] * wi;
    }
    const double xq = live ? dxb[t * NX + jj] : 0.0, xv = live ? dxb[t * NX + NV + jj] : 0.0;
    double hq = 0.0;
#pragma unroll
    for (int m = 0; m < NV; ++m)
      hq += C[Lt::CK_LQQ + (jj >= m ? tidx(NV, jj, m) : tidx(NV, m, jj))] * __shfl_sync(gm, xq, m, GW);
    const double hv = C[Lt::CK_LVV + jj] * xv;
    const double nlq = C[Lt::CK_LQ + jj] + hq + lq + aq;
    const double nlv = C[Lt::CK_LV + jj] + hv + w + av;
    if (live) kkt = fmax(kkt, fmax(fabs(su), fmax(fabs(hq), fabs(hv))));
    lq = nlq;
    lv = nlv;
  }
  const double bad = gsum((live && !(lq - lq == 0.0 && lv - lv == 0.0)) ? 1.0 : 0.0, gm);
  kkt = fmax(gmax(kkt, gm), gmax(ginf, gm));
  gl1 = gsum(live ? gl1 : 0.0, gm);
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

// SolverCSQP::tryStep for the step length 2^-n the problem is at: one group per (problem, node)
template <int NV>
__global__ void tree_sqp_try_kernel(Problem P, Work W, SolverState S, const int32_t* __restrict__ pend) {
  if (*pend == 0) return;
  using Lt = TL<NV>;
  constexpr int NX = 2 * NV;
  AGX_SMEM(smem);
  AGX_TREE_SETUP();
  const int T = P.T, T1 = T + 1;
  double* brd = smem + grp_in_cta * Lt::BOARD;
  const long long total = (long long)P.B * T1, stride = (long long)gridDim.x * grps_per_cta;
  for (long long e = ent; e < total; e += stride) {
    const int b = (int)(e / T1), t = (int)(e % T1);
    if (S.done[b] || !S.pending[b]) continue;
    const double a = ldexp(1.0, -S.roll_ok[b]);
    const size_t cur = (size_t)(S.cur[b] & 1), oth = cur ^ 1;
    const bool live = j < NV, terminal = t == T;
    const int jj = live ? j : 0;
    const double* xs = W.xs + (cur * P.B + b) * (size_t)T1 * NX;
    const double* us = W.us + (cur * P.B + b) * (size_t)T * NV;
    double* xt = W.xs + (oth * P.B + b) * (size_t)T1 * NX;
    double* ut = W.us + (oth * P.B + b) * (size_t)T * NV;
    const double* dx = W.gv + (size_t)b * T1 * NX;
    const double* du = W.k + (size_t)b * T * NV;
    TLane<NV> d;
    d.q = live ? xs[t * NX + jj] + a * dx[t * NX + jj] : 0.0;
    d.qd = live ? xs[t * NX + NV + jj] + a * dx[t * NX + NV + jj] : 0.0;
    d.u = (live && !terminal) ? us[t * NV + jj] + a * du[t * NV + jj] : 0.0;
    d.qdd = 0.0;
    const double q0 = d.q, v0 = d.qd;
    if (live) {
      xt[t * NX + j] = d.q;
      xt[t * NX + NV + j] = d.qd;
      if (!terminal) ut[t * NV + j] = d.u;
    }
    double c, qn, vn;
    const bool ok = node_calc<NV>(d, j, gm, tmodel_of(P, b), P.refs + (size_t)e * Lt::REF, terminal ? 0.0 : P.dts[t],
                                  terminal, brd, &c, &qn, &vn);
    double g = 0.0;
    if (live && !terminal) {
      const double nq = xs[(t + 1) * NX + jj] + a * dx[(t + 1) * NX + jj];
      const double nvv = xs[(t + 1) * NX + NV + jj] + a * dx[(t + 1) * NX + NV + jj];
      g = fabs(qn - nq) + fabs(vn - nvv);
    }
    if (live && t == 0) g += fabs(W.x0[(size_t)b * NX + jj] - q0) + fabs(W.x0[(size_t)b * NX + NV + jj] - v0);
    g = gsum(g, gm);
    if (j == 0) {
      double* out = W.fs + ((size_t)b * T1 + t) * NX;
      out[0] = ok ? c : nan("");
      out[1] = g;
    }
    AGX_GSYNC();
  }
}

// merit_try < merit: take the step; otherwise the next step length (run-time state dimension)
__global__ void sqp_accept_kernel_n(Problem P, int nx, Work W, SolverState S, SqpOpts Q, int32_t* __restrict__ pend) {
  if (*pend == 0) return;
  const int b = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (b >= P.B) return;
  if (S.done[b] || !S.pending[b]) return;
  const int T1 = P.T + 1;
  const double* r = W.fs + (size_t)b * T1 * nx;
  double c = 0.0, g = 0.0;
  for (int t = 0; t < T1; ++t) { c += r[t * nx]; g += r[t * nx + 1]; }
  const double mt = c + Q.mu * g;
  const int n_now = S.roll_ok[b];
  bool finished = false;
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
    S.iters[b] += 1;
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

__global__ void tree_set_capsule_kernel(double* __restrict__ model, int n_models, int capsule, double a0x, double a0y,
                                        double a0z, double a1x, double a1y, double a1z, double radius) {
  const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (i >= n_models) return;
  double* c = model + (size_t)i * TMODEL_SIZE + TT_CAP + 8 * capsule;
  c[0] = a0x; c[1] = a0y; c[2] = a0z; c[3] = a1x; c[4] = a1y; c[5] = a1z; c[6] = radius;
}

}  // namespace tree
}  // namespace agx
#endif  // AGX_TREE_CUH_
