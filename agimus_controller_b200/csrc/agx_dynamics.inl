// agx_dynamics.inl — world-frame rigid-body dynamics of one 7-DoF chain on one octet.
//
// Replaces, for the solve path, the Pinocchio routines Crocoddyl calls inside
// DifferentialActionModelFreeFwdDynamics::calc/calcDiff (built by the reference at
// agimus_controller/agimus_controller/ocp/ocp_croco_generic.py:687-711, armature :802):
// computeAllTerms (+ armature on diag M, Cholesky, Minv), computeRNEADerivatives,
// forwardKinematics / updateFramePlacements / getFrameJacobian(LOCAL), log6 / Jlog6.
//
// Formulation (not a port: Pinocchio recurses joint by joint in body frames).  Everything is
// expressed in the WORLD frame, which turns the tree recursions into scans over the 7 lanes:
//   * placements  oM_j  = inclusive prefix PRODUCT of the local transforms (3 Hillis-Steele steps);
//   * velocities  v_j   = prefix SUM of J_l qd_l; accelerations likewise;
//   * composite inertia Yc_j, composite momentum hc_j, composite "B" matrix and composite force
//     fc_j = suffix SUMS over the lanes (28 numbers per body);
//   * dtau/dq, dtau/dv, M: lane j owns COLUMN j; entry (i,j) needs one dot product with a 6-vector
//     owned by lane i, read as a broadcast from the octet's shared-memory board.
// The composite B matrix (d/dv of the bias force, Pinocchio's doYcrb) is kept in its reduced form
//   B = [[0, -2[hf]x], [0, Sym - [hn]x]]   (composite momentum (hf,hn) + a symmetric 3x3),
// see DESIGN.md "calc_diff".
//
// Every phase is a plain inline function of ONE lane's state.  The scans run in registers (width-8 warp
// shuffles); the column phases read the other lanes' 6-vectors from a small shared-memory board.

namespace agx {

// Per-lane state that lives across phases.
struct LaneDyn {
  double q, qd, u;   // this joint's position, velocity, torque
  double R[9];       // world rotation of joint frame j
  double p[3];       // world position
  double J[6];       // world joint axis (motion vector) [p x z; z]
  double s[6];       // J * qd
  double vp[6];      // parent velocity
  double v[6];       // body velocity
  double c[6];       // vp x J                      (dV/dq column)
  double g[6];       // scratch for the scans
  double a0p[6];     // parent acceleration with qdd = 0 (gravity included)
  double Y[10];      // own body inertia, world frame about the origin (m, mc, Ibar)
  double Z[28];      // composites: [0..9] Yc, [10..15] hc, [16..21] Sym_c, [22..27] fc
  double dFda[6];    // Yc J
  double BS[3];      // angular part of Bc^T J (linear part is 0)
  double Mc[7];      // column j of M + armature (lower part destroyed by the Cholesky)
  double Minv[7];    // column j of (M + armature)^-1
  double b;          // nle_j = bias torque
  double qdd;        // forward-dynamics acceleration of this joint
  double tq[7];      // dtau/dq column j
  double tv[7];      // dtau/dv column j
};

// board sizes (doubles) used by the dynamics phases
constexpr int BRD_B = 144;  // region B: [8][18] = J(6) dFda(6) BS(3) b(1) u(1), persistent during derivatives
constexpr int BRD_C = 64;   // region C: [7][8] mass matrix, then its Cholesky factor (slot 7 of row k = 1/L[k][k])

// ---------------------------------------------------------------- kinematics
AGX_DEV void kin_local(LaneDyn& d, int j, const double* __restrict__ model) {
  if (j < NJ) {
    double s, c;
    AGX_SINCOS(d.q, &s, &c);
    const double r0 = model[(MF_RP + 0) * 8 + j], r1 = model[(MF_RP + 1) * 8 + j], r2 = model[(MF_RP + 2) * 8 + j];
    const double r3 = model[(MF_RP + 3) * 8 + j], r4 = model[(MF_RP + 4) * 8 + j], r5 = model[(MF_RP + 5) * 8 + j];
    const double r6 = model[(MF_RP + 6) * 8 + j], r7 = model[(MF_RP + 7) * 8 + j], r8 = model[(MF_RP + 8) * 8 + j];
    // placement rotation times Rz(q)
    d.R[0] = c * r0 + s * r1; d.R[1] = c * r1 - s * r0; d.R[2] = r2;
    d.R[3] = c * r3 + s * r4; d.R[4] = c * r4 - s * r3; d.R[5] = r5;
    d.R[6] = c * r6 + s * r7; d.R[7] = c * r7 - s * r6; d.R[8] = r8;
    d.p[0] = model[(MF_PP + 0) * 8 + j];
    d.p[1] = model[(MF_PP + 1) * 8 + j];
    d.p[2] = model[(MF_PP + 2) * 8 + j];
  } else {
    d.R[0] = 1; d.R[1] = 0; d.R[2] = 0; d.R[3] = 0; d.R[4] = 1; d.R[5] = 0; d.R[6] = 0; d.R[7] = 0; d.R[8] = 1;
    d.p[0] = d.p[1] = d.p[2] = 0;
  }
}
AGX_DEV void kin_axis(LaneDyn& d, int j) {
  const double z[3] = {d.R[2], d.R[5], d.R[8]};
  double pz[3];
  cross3(d.p, z, pz);
  const double live = (j < NJ) ? 1.0 : 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    d.J[k] = live * pz[k];
    d.J[3 + k] = live * z[k];
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) d.s[k] = d.J[k] * d.qd;
}

// ---------------------------------------------------------------- register scans over the octet (warp shuffles)
// The chain recursions are scans over the 8 lanes; Hillis-Steele with __shfl_up/down of width 8 keeps
// them in registers (no shared-memory round trip, no barrier).  Lane 7 carries neutral elements.
template <int N>
AGX_DEV void scan_prefix_incl(double* x, int j, unsigned omask) {
#pragma unroll
  for (int dist = 1; dist < 8; dist <<= 1) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double t = __shfl_up_sync(omask, x[k], dist, 8);
      if (j >= dist) x[k] += t;
    }
  }
}
// out = seed + sum_{l < j} x_l
template <int N>
AGX_DEV void scan_prefix_excl(const double* x, double* out, const double* seed, int j, unsigned omask) {
  double acc[N];
#pragma unroll
  for (int k = 0; k < N; ++k) acc[k] = x[k];
  scan_prefix_incl<N>(acc, j, omask);
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const double t = __shfl_up_sync(omask, acc[k], 1, 8);
    out[k] = seed[k] + ((j >= 1) ? t : 0.0);
  }
}
// x_j <- sum_{l >= j} x_l
template <int N>
AGX_DEV void scan_suffix_incl(double* x, int j, unsigned omask) {
#pragma unroll
  for (int dist = 1; dist < 8; dist <<= 1) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const double t = __shfl_down_sync(omask, x[k], dist, 8);
      if (j + dist < 8) x[k] += t;
    }
  }
}
// world placements: inclusive prefix PRODUCT of the local transforms
AGX_DEV void scan_se3_prefix(LaneDyn& d, int j, unsigned omask) {
#pragma unroll
  for (int dist = 1; dist < 8; dist <<= 1) {
    double o[12];
#pragma unroll
    for (int k = 0; k < 9; ++k) o[k] = __shfl_up_sync(omask, d.R[k], dist, 8);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[9 + k] = __shfl_up_sync(omask, d.p[k], dist, 8);
    if (j >= dist) {
      double Rn[9], pn[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
          Rn[3 * r + c] = o[3 * r] * d.R[c] + o[3 * r + 1] * d.R[3 + c] + o[3 * r + 2] * d.R[6 + c];
        pn[r] = o[9 + r] + (o[3 * r] * d.p[0] + o[3 * r + 1] * d.p[1] + o[3 * r + 2] * d.p[2]);
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) d.R[k] = Rn[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) d.p[k] = pn[k];
    }
  }
}

// ---------------------------------------------------------------- body quantities (after the velocity scan)
// world-frame inertia of this lane's body about the origin: Y = (m, m c, Ibar) and the start of the composite scan
template <class LD>
AGX_DEV void body_inertia_from(LD& d, double mass, const double* com, const double* I6) {
  double cw[3];
  mv3(d.R, com, cw);
#pragma unroll
  for (int k = 0; k < 3; ++k) cw[k] += d.p[k];
  // Iw = R Ic R^T  (symmetric)
  double RI[9];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    RI[3 * r + 0] = d.R[3 * r] * I6[0] + d.R[3 * r + 1] * I6[1] + d.R[3 * r + 2] * I6[2];
    RI[3 * r + 1] = d.R[3 * r] * I6[1] + d.R[3 * r + 1] * I6[3] + d.R[3 * r + 2] * I6[4];
    RI[3 * r + 2] = d.R[3 * r] * I6[2] + d.R[3 * r + 1] * I6[4] + d.R[3 * r + 2] * I6[5];
  }
  double Iw[6];
  Iw[0] = RI[0] * d.R[0] + RI[1] * d.R[1] + RI[2] * d.R[2];
  Iw[1] = RI[0] * d.R[3] + RI[1] * d.R[4] + RI[2] * d.R[5];
  Iw[2] = RI[0] * d.R[6] + RI[1] * d.R[7] + RI[2] * d.R[8];
  Iw[3] = RI[3] * d.R[3] + RI[4] * d.R[4] + RI[5] * d.R[5];
  Iw[4] = RI[3] * d.R[6] + RI[4] * d.R[7] + RI[5] * d.R[8];
  Iw[5] = RI[6] * d.R[6] + RI[7] * d.R[7] + RI[8] * d.R[8];
  const double cc = dot3(cw, cw);
  d.Y[0] = mass;
  d.Y[1] = mass * cw[0]; d.Y[2] = mass * cw[1]; d.Y[3] = mass * cw[2];
  d.Y[4] = Iw[0] + mass * (cc - cw[0] * cw[0]);
  d.Y[5] = Iw[1] - mass * cw[0] * cw[1];
  d.Y[6] = Iw[2] - mass * cw[0] * cw[2];
  d.Y[7] = Iw[3] + mass * (cc - cw[1] * cw[1]);
  d.Y[8] = Iw[4] - mass * cw[1] * cw[2];
  d.Y[9] = Iw[5] + mass * (cc - cw[2] * cw[2]);
#pragma unroll
  for (int k = 0; k < 10; ++k) d.Z[k] = d.Y[k];
}
AGX_DEV void body_inertia(LaneDyn& d, int j, const double* __restrict__ model) {
  double mass = 0, com[3] = {0, 0, 0}, I6[6] = {0, 0, 0, 0, 0, 0};
  if (j < NJ) {
    mass = model[MF_MASS * 8 + j];
#pragma unroll
    for (int k = 0; k < 3; ++k) com[k] = model[(MF_COM + k) * 8 + j];
#pragma unroll
    for (int k = 0; k < 6; ++k) I6[k] = model[(MF_INERTIA + k) * 8 + j];
  }
  body_inertia_from(d, mass, com, I6);
}

// with_B: also the Sym block of the B matrix (derivatives only)
// first half: velocity of this body, dV/dq column, bias acceleration term g = c * qd
template <class LD>
AGX_DEV void body_motion(LD& d) {
#pragma unroll
  for (int k = 0; k < 6; ++k) d.v[k] = d.vp[k] + d.s[k];
  crm6(d.vp, d.J, d.c);
#pragma unroll
  for (int k = 0; k < 6; ++k) d.g[k] = d.c[k] * d.qd;
}
// second half (after the body inertia d.Y is known): momentum h = Y v and, with_B, the Sym block
template <class LD>
AGX_DEV void body_momentum(LD& d, bool with_B) {
  inertia_apply(d.Y, d.v, d.Z + 10);
  if (with_B) {
    // Sym = P + P^T - (mc vl^T + vl mc^T) + 2 (vl . mc) I,  P = [w]x Ibar
    const double* w = d.v + 3;
    const double* vl = d.v;
    const double* mc = d.Y + 1;
    const double* I = d.Y + 4;
    const double c0[3] = {I[0], I[1], I[2]}, c1[3] = {I[1], I[3], I[4]}, c2[3] = {I[2], I[4], I[5]};
    double P0[3], P1[3], P2[3];  // columns of P
    cross3(w, c0, P0);
    cross3(w, c1, P1);
    cross3(w, c2, P2);
    const double vm2 = 2.0 * dot3(vl, mc);
    d.Z[16] = 2.0 * P0[0] - 2.0 * mc[0] * vl[0] + vm2;                 // xx
    d.Z[17] = P1[0] + P0[1] - (mc[0] * vl[1] + vl[0] * mc[1]);         // xy
    d.Z[18] = P2[0] + P0[2] - (mc[0] * vl[2] + vl[0] * mc[2]);         // xz
    d.Z[19] = 2.0 * P1[1] - 2.0 * mc[1] * vl[1] + vm2;                 // yy
    d.Z[20] = P2[1] + P1[2] - (mc[1] * vl[2] + vl[1] * mc[2]);         // yz
    d.Z[21] = 2.0 * P2[2] - 2.0 * mc[2] * vl[2] + vm2;                 // zz
  } else {
#pragma unroll
    for (int k = 16; k < 22; ++k) d.Z[k] = 0;
  }
}
AGX_DEV void body_terms(LaneDyn& d, int j, const double* __restrict__ model, const double* grav_acc, bool with_B) {
  (void)grav_acc;
  body_motion(d);
  body_inertia(d, j, model);
  body_momentum(d, with_B);
}
// bias force with qdd = 0: f0 = Y a0 + v x* h   (after the acceleration scan filled a0p)
template <class LD>
AGX_DEV void body_force(LD& d) {
  double a0[6], Ya[6], vh[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) a0[k] = d.a0p[k] + d.g[k];
  inertia_apply(d.Y, a0, Ya);
  crf6(d.v, d.Z + 10, vh);
#pragma unroll
  for (int k = 0; k < 6; ++k) d.Z[22 + k] = Ya[k] + vh[k];
}
// column quantities: nle, dFda, BS; stored on board B as [J(6) dFda(6) BS(3) b(1)] stride 18
template <bool DERIV, class LD>
AGX_DEV void column_terms(LD& d, int j, double* sbb) {
  d.b = dot6(d.J, d.Z + 22);
  inertia_apply(d.Z, d.J, d.dFda);
  double* o = sbb + j * 18;
#pragma unroll
  for (int k = 0; k < 6; ++k) { o[k] = d.J[k]; o[6 + k] = d.dFda[k]; }
  if (DERIV) {
    // BS = angular part of Bc^T J = 2 hf x J_lin + (Sym + [hn]x) J_ang
    const double* hf = d.Z + 10;
    const double* hn = d.Z + 13;
    double t1[3], t2[3], t3[3];
    cross3(hf, d.J, t1);
    symv3(d.Z + 16, d.J + 3, t2);
    cross3(hn, d.J + 3, t3);
#pragma unroll
    for (int k = 0; k < 3; ++k) { d.BS[k] = 2.0 * t1[k] + t2[k] + t3[k]; o[12 + k] = d.BS[k]; }
  }
  o[15] = d.b;
}
// column j of M + armature:  M[i][j] = J_min . dFda_max
AGX_DEV void mass_column(LaneDyn& d, int j, const double* __restrict__ model, const double* sbb) {
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double* o = sbb + i * 18;
    const double up = dot6(o, d.dFda);       // i <= j : J_i . dFda_j
    const double lo = dot6(o + 6, d.J);      // i >  j : dFda_i . J_j
    d.Mc[i] = (i <= j) ? up : lo;
  }
  const double arm = model[MF_ARM * 8 + j];  // lane 7's slot holds 0
#pragma unroll
  for (int i = 0; i < NJ; ++i)
    if (i == j) d.Mc[i] += arm;  // static indices only: a dynamic d.Mc[j] would push the lane state to local memory
}

// ---------------------------------------------------------------- 7x7 Cholesky
// Redundant in-register factorisation: every lane loads the lower triangle of the 7x7 matrix stored
// on the board as M[i * 8 + k] and factors it (no barrier inside).  Same arithmetic order as the
// oracle's column loop.
AGX_DEV constexpr int lidx_(int i, int k) { return k * NJ - (k * (k - 1)) / 2 + (i - k); }
AGX_DEV bool chol7_registers(const double* sm_M, double* A /*28*/, double* rinv /*7*/) {
#pragma unroll
  for (int k = 0; k < NJ; ++k)
#pragma unroll
    for (int i = 0; i < NJ; ++i)
      if (i >= k) A[lidx_(i, k)] = sm_M[i * 8 + k];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < NJ; ++k) {
    double dkk = A[lidx_(k, k)];
#pragma unroll
    for (int m = 0; m < NJ; ++m)
      if (m < k) dkk -= A[lidx_(k, m)] * A[lidx_(k, m)];
    ok = ok && (dkk > 0.0);
    const double r = AGX_RSQRT(dkk);
    A[lidx_(k, k)] = dkk * r;
    rinv[k] = r;
#pragma unroll
    for (int i = 0; i < NJ; ++i) {
      if (i > k) {
        double t = A[lidx_(i, k)];
#pragma unroll
        for (int m = 0; m < NJ; ++m)
          if (m < k) t -= A[lidx_(i, m)] * A[lidx_(k, m)];
        A[lidx_(i, k)] = t * r;
      }
    }
  }
  return ok;
}
// packed index of L[i][k], i >= k, column-major packing as filled by chol_load
AGX_DEV constexpr int lidx(int i, int k) { return k * NJ - (k * (k - 1)) / 2 + (i - k); }
// solve (L L^T) x = r in place, r has 7 entries
AGX_DEV void chol_solve7(const double* L, const double* rinv, double* r) {
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    double s = r[i];
#pragma unroll
    for (int m = 0; m < i; ++m) s -= L[lidx(i, m)] * r[m];
    r[i] = s * rinv[i];
  }
#pragma unroll
  for (int i = NJ - 1; i >= 0; --i) {
    double s = r[i];
#pragma unroll
    for (int m = i + 1; m < NJ; ++m) s -= L[lidx(m, i)] * r[m];
    r[i] = s * rinv[i];
  }
}

// ---------------------------------------------------------------- second pass (with qdd) and derivative columns
// after the qdd prefix scan: d.g holds (parent) delta acceleration from qdd
template <class LD>
AGX_DEV void deriv_columns(LD& d, int j, const double* dap /*prefix of J qdd*/, const double* dfc /*suffix of Y da*/,
                           double* dFdq, double* dFdv) {
  // parent acceleration incl. qdd, dA/dq column
  double ap[6], A[6], t1[6], t2[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) ap[k] = d.a0p[k] + dap[k];
  crm6(ap, d.J, t1);
  crm6(d.vp, d.c, t2);
#pragma unroll
  for (int k = 0; k < 6; ++k) A[k] = t1[k] + t2[k];
  // composite force with qdd
  double fc[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) fc[k] = d.Z[22 + k] + dfc[k];
  const double* hf = d.Z + 10;
  const double* hn = d.Z + 13;
  // dFdv = Yc (2c) + Bc J
  double c2[6], y1[6], bl[3], ba1[3], ba2[3];
#pragma unroll
  for (int k = 0; k < 6; ++k) c2[k] = 2.0 * d.c[k];
  inertia_apply(d.Z, c2, y1);
  cross3(hf, d.J + 3, bl);
  symv3(d.Z + 16, d.J + 3, ba1);
  cross3(hn, d.J + 3, ba2);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dFdv[k] = y1[k] - 2.0 * bl[k];
    dFdv[3 + k] = y1[3 + k] + ba1[k] - ba2[k];
  }
  // dFdq = Yc A + Bc c + J x* fc
  double y2[6], jf[6];
  inertia_apply(d.Z, A, y2);
  cross3(hf, d.c + 3, bl);
  symv3(d.Z + 16, d.c + 3, ba1);
  cross3(hn, d.c + 3, ba2);
  crf6(d.J, fc, jf);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dFdq[k] = y2[k] - 2.0 * bl[k] + jf[k];
    dFdq[3 + k] = y2[3 + k] + ba1[k] - ba2[k] + jf[3 + k];
  }
  // keep A in g for the lower-triangle entries
#pragma unroll
  for (int k = 0; k < 6; ++k) d.g[k] = A[k];
}
// columns j of dtau/dq and dtau/dv (d.g = A_j)
AGX_DEV void deriv_fill(LaneDyn& d, int j, const double* dFdq, const double* dFdv, const double* sbb) {
#pragma unroll
  for (int i = 0; i < NJ; ++i) {
    const double* o = sbb + i * 18;  // J_i, dFda_i, BS_i
    const double uq = dot6(o, dFdq);
    const double uv = dot6(o, dFdv);
    const double lq = dot6(o + 6, d.g) + dot3(o + 12, d.c + 3);
    const double lv = 2.0 * dot6(o + 6, d.c) + dot3(o + 12, d.J + 3);
    d.tq[i] = (i <= j) ? uq : lq;
    d.tv[i] = (i <= j) ? uv : lv;
  }
}

// ---------------------------------------------------------------- SE3 logarithms (Pinocchio's formulas and branches)
#define AGX_TAYLOR_PREC3 1.220703125e-4  /* eps^(1/4) */
#define AGX_PI 3.14159265358979323846
// log3 with its by-products: theta, cos(theta) = (tr R - 1)/2 (clamped) and sin(theta) = sqrt((1-c)(1+c))
// (theta in [0, pi], so sin >= 0): same functions of R as sin/cos(acos(.)) without the extra sincos.
AGX_DEV void log3(const double* R, double& theta, double& ct, double& st, double* w) {
  const double tr = R[0] + R[4] + R[8];
  if (tr >= 3.0) { theta = 0.0; ct = 1.0; }
  else if (tr <= -1.0) { theta = AGX_PI; ct = -1.0; }
  else { ct = (tr - 1.0) / 2.0; theta = acos(ct); }
  st = sqrt((1.0 - ct) * (1.0 + ct));
  if (theta >= AGX_PI - 1e-2) {
    const double cphi = -(tr - 1.0) / 2.0;
    const double beta = theta * theta / (1.0 + cphi);
    const double t0 = (R[0] + cphi) * beta, t1 = (R[4] + cphi) * beta, t2 = (R[8] + cphi) * beta;
    w[0] = (R[7] > R[5] ? 1.0 : -1.0) * (t0 > 0.0 ? sqrt(t0) : 0.0);
    w[1] = (R[2] > R[6] ? 1.0 : -1.0) * (t1 > 0.0 ? sqrt(t1) : 0.0);
    w[2] = (R[3] > R[1] ? 1.0 : -1.0) * (t2 > 0.0 ? sqrt(t2) : 0.0);
  } else {
    const double t = ((theta > AGX_TAYLOR_PREC3) ? theta / st : 1.0) / 2.0;
    w[0] = t * (R[7] - R[5]);
    w[1] = t * (R[2] - R[6]);
    w[2] = t * (R[3] - R[1]);
  }
}
// r = log6(R, p) ([lin; ang]); if Jl != nullptr also the blocks of Jlog6 = [[A, B],[0, A]]: Jl[0..8] = A, Jl[9..17] = B.
// Pinocchio's coefficients share sub-expressions: alpha(log6) = diag(Jlog3) = theta sin / (2 (1 - cos)) and
// beta(log6) = alpha(Jlog3) = beta(Jlog6) = 1/theta^2 - sin / (2 theta (1 - cos)); they are formed once.
AGX_DEV void log6_and_jac(const double* R, const double* p, double* r, double* Jl) {
  double t, ct, st, w[3];
  log3(R, t, ct, st, w);
  const double t2 = t * t;
  double alpha, beta, bdot;
  if (t < AGX_TAYLOR_PREC3) {
    alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0;
    beta = 1.0 / 12.0 + t2 / 720.0;
    bdot = 1.0 / 360.0;
  } else {
    const double tinv = 1.0 / t, t2inv = tinv * tinv;
    const double inv_2_2ct = 0.5 / (1.0 - ct);
    alpha = t * st * inv_2_2ct;
    beta = t2inv - st * tinv * inv_2_2ct;
    bdot = -2.0 * t2inv * t2inv + (1.0 + st * tinv) * t2inv * inv_2_2ct;
  }
  double wxp[3];
  cross3(w, p, wxp);
  const double wp = dot3(w, p);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    r[k] = alpha * p[k] - 0.5 * wxp[k] + (beta * wp) * w[k];
    r[3 + k] = w[k];
  }
  if (!Jl) return;
  // Jlog3 = beta w w^T + diag I + 1/2 [w]x   (diag: Taylor branch 1 - theta^2/12, else alpha)
  const double diag = (t < AGX_TAYLOR_PREC3) ? 0.5 * (2.0 - t2 / 6.0) : alpha;
  double* A = Jl;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) A[3 * i + c] = beta * w[i] * w[c];
  A[0] += diag; A[4] += diag; A[8] += diag;
  A[1] -= 0.5 * w[2]; A[2] += 0.5 * w[1];
  A[3] += 0.5 * w[2]; A[5] -= 0.5 * w[0];
  A[6] -= 0.5 * w[1]; A[7] += 0.5 * w[0];
  double v3[3], C[9];
#pragma unroll
  for (int k = 0; k < 3; ++k) v3[k] = (bdot * wp) * w[k] - (t2 * bdot + 2.0 * beta) * p[k];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) C[3 * i + c] = v3[i] * w[c] + beta * w[i] * p[c];
  C[0] += wp * beta; C[4] += wp * beta; C[8] += wp * beta;
  C[1] -= 0.5 * p[2]; C[2] += 0.5 * p[1];
  C[3] += 0.5 * p[2]; C[5] -= 0.5 * p[0];
  C[6] -= 0.5 * p[1]; C[7] += 0.5 * p[0];
  double* B = Jl + 9;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) B[3 * i + c] = C[3 * i] * A[c] + C[3 * i + 1] * A[3 + c] + C[3 * i + 2] * A[6 + c];
}

// frame placement residual: given joint-6 world placement (R6, p6) -> r (6) and, if Jl, Jlog6 blocks and oMf
AGX_DEV void frame_residual_at(const double* R6, const double* p6, const double* __restrict__ FR,
                               const double* __restrict__ FP, const double* __restrict__ Rref,
                               const double* __restrict__ pref, double* Rf, double* pf, double* r, double* Jl) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) Rf[3 * i + c] = R6[3 * i] * FR[c] + R6[3 * i + 1] * FR[3 + c] + R6[3 * i + 2] * FR[6 + c];
    pf[i] = p6[i] + (R6[3 * i] * FP[0] + R6[3 * i + 1] * FP[1] + R6[3 * i + 2] * FP[2]);
  }
  double Rr[9], pr[3], dp[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int c = 0; c < 3; ++c) Rr[3 * i + c] = Rref[i] * Rf[c] + Rref[3 + i] * Rf[3 + c] + Rref[6 + i] * Rf[6 + c];
#pragma unroll
  for (int k = 0; k < 3; ++k) dp[k] = pf[k] - pref[k];
  mtv3(Rref, dp, pr);
  log6_and_jac(Rr, pr, r, Jl);
}
AGX_DEV void frame_residual(const double* R6, const double* p6, const double* __restrict__ model,
                            const double* __restrict__ Rref, const double* __restrict__ pref, double* Rf, double* pf,
                            double* r, double* Jl) {
  frame_residual_at(R6, p6, model + MT_FR, model + MT_FP, Rref, pref, Rf, pf, r, Jl);
}

// ---------------------------------------------------------------------------------------------
// Collision residual (A10): colmpc::ResidualDistanceCollision on a capsule pair + ActivationModelQuadExp
// (ocp/ocp_croco_generic.py:119-147, :499-535; ocp_traj_tracking_collision_avoidance.yaml:36-46).
AGX_DEV double clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }

// Closest points of the segments [a0,a1] and [b0,b1] (world frame): ca, cb, the unit direction n = (ca-cb)/|ca-cb|;
// returns |ca - cb|.  Degenerate (point-like) segments and the parallel case take the first end point.
AGX_DEV double segment_pair(const double* a0, const double* a1, const double* b0, const double* b1, double* ca,
                            double* cb, double* n) {
  const double eps = 1e-12;
  double d1[3], d2[3], r[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { d1[k] = a1[k] - a0[k]; d2[k] = b1[k] - b0[k]; r[k] = a0[k] - b0[k]; }
  const double a = dot3(d1, d1), e = dot3(d2, d2), f = dot3(d2, r);
  double s = 0.0, t = 0.0;
  if (a <= eps && e <= eps) {
    s = t = 0.0;
  } else if (a <= eps) {
    t = clamp01(f / e);
  } else {
    const double c = dot3(d1, r);
    if (e <= eps) {
      s = clamp01(-c / a);
    } else {
      const double b = dot3(d1, d2), denom = a * e - b * b;
      s = (denom > eps * a * e) ? clamp01((b * f - c * e) / denom) : 0.0;
      t = (b * s + f) / e;
      if (t < 0.0) { t = 0.0; s = clamp01(-c / a); }
      else if (t > 1.0) { t = 1.0; s = clamp01((b - c) / a); }
    }
  }
  double dd[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { ca[k] = a0[k] + s * d1[k]; cb[k] = b0[k] + t * d2[k]; dd[k] = ca[k] - cb[k]; }
  const double len = sqrt(dot3(dd, dd));
  const double inv = len > 1e-14 ? 1.0 / len : 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) n[k] = dd[k] * inv;
  return len;
}

// a = exp(-r^2/alpha) with its first two derivatives in r
AGX_DEV void quadexp(double r, double alpha, double& a, double& ar, double& arr) {
  const double ia = 1.0 / alpha;
  a = exp(-r * r * ia);
  ar = -2.0 * r * ia * a;
  arr = (4.0 * r * r * ia * ia - 2.0 * ia) * a;
}

}  // namespace agx
