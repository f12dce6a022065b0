// agx_oracle.cpp — CPU restatement of the OCP solve path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load this library; the product (agimus_controller_b200/) never does.
//
// What it restates (reference paths relative to /root/reference; the arithmetic itself lives in
// third-party C++ that is NOT in the reference tree, so each block restates the published
// algorithm of the dependency named there — versions are pinned only transitively through
// flake.lock: gepetto/nix@6d2dbabb81b7, nixpkgs@0182a3613243):
//   * Pinocchio   rnea / crba / computeRNEADerivatives / forwardKinematics / getFrameJacobian(LOCAL)
//                 / log3 / Jlog3 / log6 / Jlog6          (called inside Crocoddyl; direct call
//                 agimus_controller/agimus_controller/warm_start_reference.py:77-87)
//   * Crocoddyl   DifferentialActionModelFreeFwdDynamics (+armature), IntegratedActionModelEuler,
//                 CostModelSum / CostModelResidual / ActivationModelWeightedQuad,
//                 ResidualModel{State,Control,FramePlacement}, ShootingProblem::{calc,calcDiff,rollout},
//                 SolverFDDP                              (built at
//                 agimus_controller/agimus_controller/ocp/ocp_croco_generic.py:687-745, :798-812 and
//                 agimus_controller/agimus_controller/ocp_base_croco.py:36-80; solve at :142-182)
//   * mim_solvers SolverCSQP (the solver the reference really instantiates, ocp_base_croco.py:64-75) in the form it
//                 takes with no active constraint: Gauss-Newton SQP with a regularised Riccati QP solve, KKT stop,
//                 L1 merit line search, gains from a last sweep with the proximal sigma (struct Sqp).
//   * colmpc      ResidualDistanceCollision on capsule pairs + ActivationModelQuadExp (PARITY UNPINNED: no source,
//                 no golden vector).
//
// Parity pinning: the reference's golden file agimus_controller/tests/resources/simple_ocp_croco_results.pkl pins
// the Panda table, forward dynamics and the cost stack (KAT-1/2), the Riccati gains to 1e-11 (KAT-3) and — through
// the SQP mode replaying the reference's own test — the whole solve at the reference test's 6 decimals (KAT-9,
// tests/test_oracle_golden.py); test_sin_wave_cartesian_space.py:190-218 pins the kinematics (KAT-4).
// What stays "parity unpinned": the FDDP-specific decisions (no reference test stores an FDDP result) and the
// collision residuals.
//
// Build: see oracle/Makefile (g++ -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off).
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/agx.h"

namespace {

constexpr int MAXV = AGX_MAX_NV;
constexpr int MAXX = 2 * MAXV;

// ----------------------------------------------------------------------------- small algebra
inline void cross3(const double* a, const double* b, double* o) {
  double x = a[1] * b[2] - a[2] * b[1];
  double y = a[2] * b[0] - a[0] * b[2];
  double z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
inline void mv3(const double* R, const double* x, double* o) {  // o = R x
  double a = R[0] * x[0] + R[1] * x[1] + R[2] * x[2];
  double b = R[3] * x[0] + R[4] * x[1] + R[5] * x[2];
  double c = R[6] * x[0] + R[7] * x[1] + R[8] * x[2];
  o[0] = a; o[1] = b; o[2] = c;
}
inline void mtv3(const double* R, const double* x, double* o) {  // o = R^T x
  double a = R[0] * x[0] + R[3] * x[1] + R[6] * x[2];
  double b = R[1] * x[0] + R[4] * x[1] + R[7] * x[2];
  double c = R[2] * x[0] + R[5] * x[1] + R[8] * x[2];
  o[0] = a; o[1] = b; o[2] = c;
}
inline void mm3(const double* A, const double* B, double* C) {  // C = A B
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  std::memcpy(C, t, sizeof t);
}
inline void mtm3(const double* A, const double* B, double* C) {  // C = A^T B
  double t[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) t[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
  std::memcpy(C, t, sizeof t);
}
inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double dot6(const double* a, const double* b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

struct SE3 {
  double R[9];
  double p[3];
};
inline void se3_mul(const SE3& A, const SE3& B, SE3& C) {
  SE3 t;
  mm3(A.R, B.R, t.R);
  mv3(A.R, B.p, t.p);
  for (int k = 0; k < 3; ++k) t.p[k] += A.p[k];
  C = t;
}
// spatial vectors: [0..2] linear, [3..5] angular (Pinocchio ordering)
inline void mot_act(const SE3& M, const double* m, double* o) {  // child -> parent
  double w[3], v[3], pw[3];
  mv3(M.R, m + 3, w);
  mv3(M.R, m, v);
  cross3(M.p, w, pw);
  for (int k = 0; k < 3; ++k) { o[k] = v[k] + pw[k]; o[3 + k] = w[k]; }
}
inline void mot_actinv(const SE3& M, const double* m, double* o) {  // parent -> child
  double pw[3], t[3];
  cross3(M.p, m + 3, pw);
  for (int k = 0; k < 3; ++k) t[k] = m[k] - pw[k];
  double a[3], b[3];
  mtv3(M.R, t, a);
  mtv3(M.R, m + 3, b);
  for (int k = 0; k < 3; ++k) { o[k] = a[k]; o[3 + k] = b[k]; }
}
inline void frc_act(const SE3& M, const double* f, double* o) {  // child -> parent
  double fl[3], n[3], pf[3];
  mv3(M.R, f, fl);
  mv3(M.R, f + 3, n);
  cross3(M.p, fl, pf);
  for (int k = 0; k < 3; ++k) { o[k] = fl[k]; o[3 + k] = n[k] + pf[k]; }
}
inline void crm(const double* a, const double* b, double* o) {  // motion x motion
  double t1[3], t2[3], t3[3];
  cross3(a + 3, b, t1);
  cross3(a, b + 3, t2);
  cross3(a + 3, b + 3, t3);
  for (int k = 0; k < 3; ++k) { o[k] = t1[k] + t2[k]; o[3 + k] = t3[k]; }
}
inline void crf(const double* a, const double* f, double* o) {  // motion x* force
  double t1[3], t2[3], t3[3];
  cross3(a + 3, f, t1);
  cross3(a + 3, f + 3, t2);
  cross3(a, f, t3);
  for (int k = 0; k < 3; ++k) { o[k] = t1[k]; o[3 + k] = t2[k] + t3[k]; }
}
inline void skew(const double* a, double* S) {  // S = [a]x  (3x3)
  S[0] = 0; S[1] = -a[2]; S[2] = a[1];
  S[3] = a[2]; S[4] = 0; S[5] = -a[0];
  S[6] = -a[1]; S[7] = a[0]; S[8] = 0;
}
inline void mv6(const double* A, const double* x, double* o) {
  double t[6];
  for (int i = 0; i < 6; ++i) t[i] = dot6(A + 6 * i, x);
  std::memcpy(o, t, sizeof t);
}
inline void mtv6(const double* A, const double* x, double* o) {
  double t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) t[j] += A[6 * i + j] * x[i];
  std::memcpy(o, t, sizeof t);
}
inline void rodrigues(const double* n, double q, double* R) {
  double s = std::sin(q), c = std::cos(q), v = 1.0 - c;
  R[0] = c + n[0] * n[0] * v;        R[1] = n[0] * n[1] * v - n[2] * s; R[2] = n[0] * n[2] * v + n[1] * s;
  R[3] = n[1] * n[0] * v + n[2] * s; R[4] = c + n[1] * n[1] * v;        R[5] = n[1] * n[2] * v - n[0] * s;
  R[6] = n[2] * n[0] * v - n[1] * s; R[7] = n[2] * n[1] * v + n[0] * s; R[8] = c + n[2] * n[2] * v;
}
inline void inertia3(const double* I6, double* I) {
  I[0] = I6[0]; I[1] = I6[1]; I[2] = I6[2];
  I[3] = I6[1]; I[4] = I6[3]; I[5] = I6[4];
  I[6] = I6[2]; I[7] = I6[4]; I[8] = I6[5];
}
// body inertia (m, c, Ic) applied to a motion, all in the same frame
inline void inertia_apply(double m, const double* c, const double* Ic, const double* mo, double* f) {
  double wc[3], fl[3], n[3], cf[3];
  cross3(mo + 3, c, wc);
  for (int k = 0; k < 3; ++k) fl[k] = m * (mo[k] + wc[k]);
  mv3(Ic, mo + 3, n);
  cross3(c, fl, cf);
  for (int k = 0; k < 3; ++k) { f[k] = fl[k]; f[3 + k] = n[k] + cf[k]; }
}

// joint transform liMi = placement * joint(q), and motion subspace S (body frame)
inline void joint_calc(const agx_model& m, int i, double q, SE3& liMi, double* S) {
  SE3 P, J;
  std::memcpy(P.R, m.placement_R[i], sizeof P.R);
  std::memcpy(P.p, m.placement_p[i], sizeof P.p);
  if (m.jtype[i] == AGX_JOINT_REVOLUTE) {
    rodrigues(m.axis[i], q, J.R);
    J.p[0] = J.p[1] = J.p[2] = 0;
    S[0] = S[1] = S[2] = 0;
    S[3] = m.axis[i][0]; S[4] = m.axis[i][1]; S[5] = m.axis[i][2];
  } else {
    J.R[0] = 1; J.R[1] = 0; J.R[2] = 0; J.R[3] = 0; J.R[4] = 1; J.R[5] = 0; J.R[6] = 0; J.R[7] = 0; J.R[8] = 1;
    for (int k = 0; k < 3; ++k) J.p[k] = m.axis[i][k] * q;
    S[0] = m.axis[i][0]; S[1] = m.axis[i][1]; S[2] = m.axis[i][2];
    S[3] = S[4] = S[5] = 0;
  }
  se3_mul(P, J, liMi);
}

// ----------------------------------------------------------------------------- Pinocchio::rnea
// Recursive Newton-Euler in body frames (Featherstone), gravity as base acceleration -g.
void rnea(const agx_model& m, const double* q, const double* v, const double* a, double* tau) {
  const int nv = m.nv;
  SE3 liMi[MAXV];
  double S[MAXV][6], vi[MAXV][6], ai[MAXV][6], fi[MAXV][6];
  for (int i = 0; i < nv; ++i) {
    joint_calc(m, i, q[i], liMi[i], S[i]);
    double vp[6] = {0, 0, 0, 0, 0, 0}, ap[6] = {-m.gravity[0], -m.gravity[1], -m.gravity[2], 0, 0, 0};
    if (m.parent[i] >= 0) {
      std::memcpy(vp, vi[m.parent[i]], sizeof vp);
      std::memcpy(ap, ai[m.parent[i]], sizeof ap);
    }
    double vx[6], ax[6], vj[6], cx[6];
    mot_actinv(liMi[i], vp, vx);
    mot_actinv(liMi[i], ap, ax);
    for (int k = 0; k < 6; ++k) { vj[k] = S[i][k] * v[i]; vi[i][k] = vx[k] + vj[k]; }
    crm(vi[i], vj, cx);
    for (int k = 0; k < 6; ++k) ai[i][k] = ax[k] + S[i][k] * a[i] + cx[k];
    double Ic[9], h[6], Ia[6], vh[6];
    inertia3(m.inertia[i], Ic);
    inertia_apply(m.mass[i], m.com[i], Ic, vi[i], h);
    inertia_apply(m.mass[i], m.com[i], Ic, ai[i], Ia);
    crf(vi[i], h, vh);
    for (int k = 0; k < 6; ++k) fi[i][k] = Ia[k] + vh[k];
  }
  for (int i = nv - 1; i >= 0; --i) {
    tau[i] = dot6(S[i], fi[i]);
    if (m.parent[i] >= 0) {
      double fp[6];
      frc_act(liMi[i], fi[i], fp);
      for (int k = 0; k < 6; ++k) fi[m.parent[i]][k] += fp[k];
    }
  }
}

// ----------------------------------------------------------------------------- forward kinematics
struct Kin {
  SE3 oMi[MAXV];
  double J[MAXV][6];  // world-frame joint motion axes (columns of the world Jacobian)
};
void forward_kinematics(const agx_model& m, const double* q, Kin& k) {
  for (int i = 0; i < m.nv; ++i) {
    SE3 li;
    double S[6];
    joint_calc(m, i, q[i], li, S);
    if (m.parent[i] >= 0) se3_mul(k.oMi[m.parent[i]], li, k.oMi[i]);
    else k.oMi[i] = li;
    mot_act(k.oMi[i], S, k.J[i]);
  }
}
void frame_placement(const agx_model& m, const Kin& k, SE3& oMf) {
  SE3 F;
  std::memcpy(F.R, m.frame_R, sizeof F.R);
  std::memcpy(F.p, m.frame_p, sizeof F.p);
  se3_mul(k.oMi[m.frame_parent], F, oMf);
}
// getFrameJacobian(LOCAL): 6 x nv, row-major
void frame_jacobian_local(const agx_model& m, const Kin& k, const SE3& oMf, double* fJ) {
  const int nv = m.nv;
  for (int i = 0; i < 6 * nv; ++i) fJ[i] = 0;
  for (int j = m.frame_parent; j >= 0; j = m.parent[j]) {
    double c[6];
    mot_actinv(oMf, k.J[j], c);
    for (int r = 0; r < 6; ++r) fJ[r * nv + j] = c[r];
  }
}

// ----------------------------------------------------------------------------- Pinocchio::crba
// Composite rigid body algorithm in body frames, dense 6x6 composite inertias.
inline void inertia_dense(double mass, const double* c, const double* Ic, double* Y) {
  double C[9], CC[9];
  skew(c, C);
  mm3(C, C, CC);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Y[6 * i + j] = (i == j) ? mass : 0.0;
      Y[6 * i + 3 + j] = -mass * C[3 * i + j];
      Y[6 * (3 + i) + j] = mass * C[3 * i + j];
      Y[6 * (3 + i) + 3 + j] = Ic[3 * i + j] - mass * CC[3 * i + j];
    }
}
inline void force_xform_dense(const SE3& M, double* X) {  // 6x6, child -> parent for forces
  double P[9], PR[9];
  skew(M.p, P);
  mm3(P, M.R, PR);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      X[6 * i + j] = M.R[3 * i + j];
      X[6 * i + 3 + j] = 0;
      X[6 * (3 + i) + j] = PR[3 * i + j];
      X[6 * (3 + i) + 3 + j] = M.R[3 * i + j];
    }
}
void crba(const agx_model& m, const double* q, double* M /* nv x nv */) {
  const int nv = m.nv;
  SE3 liMi[MAXV];
  double S[MAXV][6];
  static thread_local double Yc[MAXV][36];
  for (int i = 0; i < nv; ++i) {
    joint_calc(m, i, q[i], liMi[i], S[i]);
    double Ic[9];
    inertia3(m.inertia[i], Ic);
    inertia_dense(m.mass[i], m.com[i], Ic, Yc[i]);
  }
  for (int i = 0; i < nv * nv; ++i) M[i] = 0;
  for (int i = nv - 1; i >= 0; --i) {
    double F[6];
    mv6(Yc[i], S[i], F);
    M[i * nv + i] = dot6(S[i], F);
    if (m.parent[i] >= 0) {
      double X[36], XY[36];
      force_xform_dense(liMi[i], X);
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          double s = 0;
          for (int k = 0; k < 6; ++k) s += X[6 * r + k] * Yc[i][6 * k + c];
          XY[6 * r + c] = s;
        }
      double* Yp = Yc[m.parent[i]];
      for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) {
          double s = 0;
          for (int k = 0; k < 6; ++k) s += XY[6 * r + k] * X[6 * c + k];
          Yp[6 * r + c] += s;
        }
    }
    int j = i;
    while (m.parent[j] >= 0) {
      double Fp[6];
      frc_act(liMi[j], F, Fp);
      std::memcpy(F, Fp, sizeof F);
      j = m.parent[j];
      M[i * nv + j] = M[j * nv + i] = dot6(S[j], F);
    }
  }
}

// ----------------------------------------------------------------------------- computeRNEADerivatives
// World-frame O(n^2) algorithm (Carpentier & Mansard, RSS 2018) as Pinocchio implements it:
// tau, dtau/dq, dtau/dv and M (= dtau/da) in one sweep.
void rnea_derivatives(const agx_model& m, const double* q, const double* v, const double* a, double* tau,
                      double* dq, double* dv, double* M) {
  const int nv = m.nv;
  Kin kin;
  forward_kinematics(m, q, kin);
  static thread_local double Yc[MAXV][36], Bc[MAXV][36];
  double ov[MAXV][6], oa[MAXV][6], cc[MAXV][6], AA[MAXV][6], fc[MAXV][6];
  double dFda[MAXV][6], dFdv[MAXV][6], dFdq[MAXV][6], BS[MAXV][6];
  const double a0[6] = {-m.gravity[0], -m.gravity[1], -m.gravity[2], 0, 0, 0};
  const double zero6[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < nv; ++i) {
    const int p = m.parent[i];
    const double* vp = p >= 0 ? ov[p] : zero6;
    const double* ap = p >= 0 ? oa[p] : a0;
    const double* J = kin.J[i];
    crm(vp, J, cc[i]);  // dVdq column
    double t1[6], t2[6];
    crm(ap, J, t1);
    crm(vp, cc[i], t2);
    for (int k = 0; k < 6; ++k) {
      AA[i][k] = t1[k] + t2[k];  // dAdq column
      ov[i][k] = vp[k] + J[k] * v[i];
      oa[i][k] = ap[k] + J[k] * a[i] + cc[i][k] * v[i];
    }
    // world-frame body inertia
    double cw[3], Ic[9], RI[9], Iw[9];
    mv3(kin.oMi[i].R, m.com[i], cw);
    for (int k = 0; k < 3; ++k) cw[k] += kin.oMi[i].p[k];
    inertia3(m.inertia[i], Ic);
    mm3(kin.oMi[i].R, Ic, RI);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        Iw[3 * r + c] = RI[3 * r] * kin.oMi[i].R[3 * c] + RI[3 * r + 1] * kin.oMi[i].R[3 * c + 1] +
                        RI[3 * r + 2] * kin.oMi[i].R[3 * c + 2];
    double* Y = Yc[i];
    inertia_dense(m.mass[i], cw, Iw, Y);
    double h[6], Ya[6], vh[6];
    mv6(Y, ov[i], h);
    mv6(Y, oa[i], Ya);
    crf(ov[i], h, vh);
    for (int k = 0; k < 6; ++k) fc[i][k] = Ya[k] + vh[k];
    // B = crf(v) Y - Y crm(v) + Hx(h)
    double W[9], V[9], Hf[9], Hn[9];
    skew(ov[i] + 3, W);
    skew(ov[i], V);
    skew(h, Hf);
    skew(h + 3, Hn);
    double CF[36], CM[36];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        CF[6 * r + c] = W[3 * r + c]; CF[6 * r + 3 + c] = 0;
        CF[6 * (3 + r) + c] = V[3 * r + c]; CF[6 * (3 + r) + 3 + c] = W[3 * r + c];
        CM[6 * r + c] = W[3 * r + c]; CM[6 * r + 3 + c] = V[3 * r + c];
        CM[6 * (3 + r) + c] = 0; CM[6 * (3 + r) + 3 + c] = W[3 * r + c];
      }
    double* B = Bc[i];
    for (int r = 0; r < 6; ++r)
      for (int c = 0; c < 6; ++c) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += CF[6 * r + k] * Y[6 * k + c] - Y[6 * r + k] * CM[6 * k + c];
        B[6 * r + c] = s;
      }
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        B[6 * r + 3 + c] -= Hf[3 * r + c];
        B[6 * (3 + r) + c] -= Hf[3 * r + c];
        B[6 * (3 + r) + 3 + c] -= Hn[3 * r + c];
      }
  }
  for (int i = nv - 1; i >= 0; --i) {
    const double* J = kin.J[i];
    tau[i] = dot6(J, fc[i]);
    double c2[6], t1[6], t2[6], t3[6];
    mv6(Yc[i], J, dFda[i]);
    for (int k = 0; k < 6; ++k) c2[k] = 2.0 * cc[i][k];
    mv6(Yc[i], c2, t1);
    mv6(Bc[i], J, t2);
    for (int k = 0; k < 6; ++k) dFdv[i][k] = t1[k] + t2[k];
    mv6(Yc[i], AA[i], t1);
    mv6(Bc[i], cc[i], t2);
    crf(J, fc[i], t3);
    for (int k = 0; k < 6; ++k) dFdq[i][k] = t1[k] + t2[k] + t3[k];
    mtv6(Bc[i], J, BS[i]);
    const int p = m.parent[i];
    if (p >= 0) {
      for (int k = 0; k < 36; ++k) { Yc[p][k] += Yc[i][k]; Bc[p][k] += Bc[i][k]; }
      for (int k = 0; k < 6; ++k) fc[p][k] += fc[i][k];
    }
  }
  for (int k = 0; k < nv * nv; ++k) dq[k] = dv[k] = M[k] = 0;
  for (int j = 0; j < nv; ++j) {
    for (int i = j; i >= 0; i = m.parent[i]) {  // i ancestor-or-self of j
      const double* Ji = kin.J[i];
      dq[i * nv + j] = dot6(Ji, dFdq[j]);
      dv[i * nv + j] = dot6(Ji, dFdv[j]);
      M[i * nv + j] = dot6(Ji, dFda[j]);
      if (i != j) {
        double c2[6];
        for (int k = 0; k < 6; ++k) c2[k] = 2.0 * cc[i][k];
        dq[j * nv + i] = dot6(dFda[j], AA[i]) + dot6(BS[j], cc[i]);
        dv[j * nv + i] = dot6(dFda[j], c2) + dot6(BS[j], Ji);
        M[j * nv + i] = M[i * nv + j];
      }
    }
  }
}

// ----------------------------------------------------------------------------- dense helpers
bool cholesky(int n, const double* A, double* L) {  // lower, row-major; false if not SPD / NaN
  for (int k = 0; k < n * n; ++k) L[k] = 0;
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= L[j * n + k] * L[j * n + k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    d = std::sqrt(d);
    L[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= L[i * n + k] * L[j * n + k];
      L[i * n + j] = s / d;
    }
  }
  return true;
}
void chol_solve(int n, const double* L, double* b) {  // in place
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k];
    b[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k];
    b[i] = s / L[i * n + i];
  }
}

// ----------------------------------------------------------------------------- Pinocchio log maps
const double TAYLOR_PREC3 = std::pow(std::numeric_limits<double>::epsilon(), 0.25);

void log3(const double* R, double& theta, double* w) {
  const double PI_ = 3.14159265358979323846;
  const double tr = R[0] + R[4] + R[8];
  if (tr >= 3.0) theta = 0.0;
  else if (tr <= -1.0) theta = PI_;
  else theta = std::acos((tr - 1.0) / 2.0);
  if (theta >= PI_ - 1e-2) {
    const double cphi = -(tr - 1.0) / 2.0;
    const double beta = theta * theta / (1.0 + cphi);
    const double t0 = (R[0] + cphi) * beta, t1 = (R[4] + cphi) * beta, t2 = (R[8] + cphi) * beta;
    w[0] = (R[7] > R[5] ? 1.0 : -1.0) * (t0 > 0.0 ? std::sqrt(t0) : 0.0);
    w[1] = (R[2] > R[6] ? 1.0 : -1.0) * (t1 > 0.0 ? std::sqrt(t1) : 0.0);
    w[2] = (R[3] > R[1] ? 1.0 : -1.0) * (t2 > 0.0 ? std::sqrt(t2) : 0.0);
  } else {
    const double t = ((theta > TAYLOR_PREC3) ? theta / std::sin(theta) : 1.0) / 2.0;
    w[0] = t * (R[7] - R[5]);
    w[1] = t * (R[2] - R[6]);
    w[2] = t * (R[3] - R[1]);
  }
}
void Jlog3(double theta, const double* w, double* J) {
  double alpha, diag;
  if (theta < TAYLOR_PREC3) {
    alpha = 1.0 / 12.0 + theta * theta / 720.0;
    diag = 0.5 * (2.0 - theta * theta / 6.0);
  } else {
    const double ct = std::cos(theta), st = std::sin(theta);
    const double st_1mct = st / (1.0 - ct);
    alpha = 1.0 / (theta * theta) - st_1mct / (2.0 * theta);
    diag = 0.5 * (theta * st_1mct);
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) J[3 * i + j] = alpha * w[i] * w[j];
  J[0] += diag; J[4] += diag; J[8] += diag;
  // addSkew(0.5 w)
  J[1] -= 0.5 * w[2]; J[2] += 0.5 * w[1];
  J[3] += 0.5 * w[2]; J[5] -= 0.5 * w[0];
  J[6] -= 0.5 * w[1]; J[7] += 0.5 * w[0];
}
void log6(const SE3& M, double* out /* [lin; ang] */) {
  double t, w[3];
  log3(M.R, t, w);
  const double t2 = t * t;
  double alpha, beta;
  if (t < TAYLOR_PREC3) {
    alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0;
    beta = 1.0 / 12.0 + t2 / 720.0;
  } else {
    const double st = std::sin(t), ct = std::cos(t);
    alpha = t * st / (2.0 * (1.0 - ct));
    beta = 1.0 / t2 - st / (2.0 * t * (1.0 - ct));
  }
  double wxp[3];
  cross3(w, M.p, wxp);
  const double wp = dot3(w, M.p);
  for (int k = 0; k < 3; ++k) {
    out[k] = alpha * M.p[k] - 0.5 * wxp[k] + (beta * wp) * w[k];
    out[3 + k] = w[k];
  }
}
void Jlog6(const SE3& M, double* J /* 6x6 row-major */) {
  double t, w[3];
  log3(M.R, t, w);
  const double t2 = t * t;
  double A[9];
  Jlog3(t, w, A);
  double beta, bdot;
  if (t < TAYLOR_PREC3) {
    beta = 1.0 / 12.0 + t2 / 720.0;
    bdot = 1.0 / 360.0;
  } else {
    const double tinv = 1.0 / t, t2inv = tinv * tinv;
    const double st = std::sin(t), ct = std::cos(t);
    const double inv_2_2ct = 1.0 / (2.0 * (1.0 - ct));
    beta = t2inv - st * tinv * inv_2_2ct;
    bdot = -2.0 * t2inv * t2inv + (1.0 + st * tinv) * t2inv * inv_2_2ct;
  }
  const double* p = M.p;
  const double wTp = dot3(w, p);
  double v3[3];
  for (int k = 0; k < 3; ++k) v3[k] = (bdot * wTp) * w[k] - (t2 * bdot + 2.0 * beta) * p[k];
  double C[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[3 * i + j] = v3[i] * w[j] + beta * w[i] * p[j];
  C[0] += wTp * beta; C[4] += wTp * beta; C[8] += wTp * beta;
  C[1] -= 0.5 * p[2]; C[2] += 0.5 * p[1];
  C[3] += 0.5 * p[2]; C[5] -= 0.5 * p[0];
  C[6] -= 0.5 * p[1]; C[7] += 0.5 * p[0];
  double Bm[9];
  mm3(C, A, Bm);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      J[6 * i + j] = A[3 * i + j];
      J[6 * i + 3 + j] = Bm[3 * i + j];
      J[6 * (3 + i) + j] = 0;
      J[6 * (3 + i) + 3 + j] = A[3 * i + j];
    }
}

// ----------------------------------------------------------------------------- node (action) model
struct NodeRef {  // view into one reference record
  const double *xref, *wx, *uref, *wu, *Rref, *pref, *wpose, *wcol;
};
inline NodeRef make_ref(const double* r, int nv) {
  const int nx = 2 * nv;
  NodeRef o;
  o.xref = r; o.wx = r + nx; o.uref = r + 2 * nx; o.wu = r + 2 * nx + nv;
  o.Rref = r + 2 * nx + 2 * nv; o.pref = o.Rref + 9; o.wpose = o.pref + 3; o.wcol = o.wpose + 6;
  return o;
}

struct NodeData {
  double xnext[MAXX], cost;
  double Fx[MAXX * MAXX], Fu[MAXX * MAXV];
  double Lx[MAXX], Lu[MAXV], Lxx[MAXX * MAXX], Lxu[MAXX * MAXV], Luu[MAXV * MAXV];
  // intermediate (kept for the per-cost API / debugging)
  double a[MAXV], Minv[MAXV * MAXV], rpose[6], rcol[AGX_MAX_COLLISION_PAIRS];
};

// frame-placement residual r = log6(Mref^-1 oMf) and its Jacobian Rq = Jlog6 * fJf (6 x nv)
void pose_residual(const agx_model& m, const Kin& kin, const NodeRef& ref, double* r, double* Rq) {
  SE3 oMf, rMf;
  frame_placement(m, kin, oMf);
  mtm3(ref.Rref, oMf.R, rMf.R);
  double d[3] = {oMf.p[0] - ref.pref[0], oMf.p[1] - ref.pref[1], oMf.p[2] - ref.pref[2]};
  mtv3(ref.Rref, d, rMf.p);
  log6(rMf, r);
  // ResidualModelFrameTranslation (ocp_croco_generic.py:252-303): r = p_f - pref in the world, Rq = oRf fJf[:3]
  const bool tworld = m.pose_mode == AGX_POSE_TRANSLATION_WORLD;
  if (tworld) for (int k = 0; k < 3; ++k) r[k] = d[k];
  if (Rq) {
    const int nv = m.nv;
    double Jl[36], fJ[6 * MAXV];
    Jlog6(rMf, Jl);
    frame_jacobian_local(m, kin, oMf, fJ);
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < nv; ++j) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += Jl[6 * i + k] * fJ[k * nv + j];
        Rq[i * nv + j] = s;
      }
    if (tworld)
      for (int j = 0; j < nv; ++j) {
        const double l[3] = {fJ[0 * nv + j], fJ[1 * nv + j], fJ[2 * nv + j]};
        double w[3];
        mv3(oMf.R, l, w);
        for (int i = 0; i < 3; ++i) Rq[i * nv + j] = w[i];
      }
  }
}

// ----------------------------------------------------------------------------- collision residual (A10)
// colmpc::ResidualDistanceCollision on a capsule pair, followed by colmpc::ActivationModelQuadExp
// (ocp/ocp_croco_generic.py:119-147, :499-535; ocp/ocp_traj_tracking_collision_avoidance.yaml:36-46).  colmpc is a
// third-party dependency that is not in the reference tree, so this restates its published formulas:
//   r(q)   = dist(shape_a, shape_b)           (coal signed distance; capsule pair = segment distance - r_a - r_b)
//   dr/dq  = n^T (J_a(c_a) - J_b(c_b))        n = (c_a - c_b)/|c_a - c_b|, J(c) the world linear Jacobian of the
//                                             point c rigidly attached to the capsule's parent joint
//   a(r)   = exp(-r^2/alpha), a' = -2 r/alpha a, a'' = (4 r^2/alpha^2 - 2/alpha) a
// PARITY UNPINNED for this row: no golden vector of the reference exercises it.
struct CapsuleHit {
  double dist;       // signed distance between the capsules
  double ca[3], cb[3], n[3];  // closest points on the two axes (world), unit direction cb -> ca
};
inline double clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }
// closest points of two segments [a0,a1], [b0,b1] (Ericson, Real-Time Collision Detection, 5.1.9)
void capsule_pair(const double* a0, const double* a1, double ra, const double* b0, const double* b1, double rb,
                  CapsuleHit& h) {
  const double eps = 1e-12;
  double d1[3], d2[3], r[3];
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
  for (int k = 0; k < 3; ++k) { h.ca[k] = a0[k] + s * d1[k]; h.cb[k] = b0[k] + t * d2[k]; dd[k] = h.ca[k] - h.cb[k]; }
  const double len = std::sqrt(dot3(dd, dd));
  const double inv = len > 1e-14 ? 1.0 / len : 0.0;
  for (int k = 0; k < 3; ++k) h.n[k] = dd[k] * inv;
  h.dist = len - ra - rb;
}
// world end points of capsule c
inline void capsule_world(const agx_model& m, const Kin& kin, int c, double* a0, double* a1) {
  const int par = m.cap_parent[c];
  if (par < 0) {
    for (int k = 0; k < 3; ++k) { a0[k] = m.cap_a0[c][k]; a1[k] = m.cap_a1[c][k]; }
    return;
  }
  double t0[3], t1[3];
  mv3(kin.oMi[par].R, m.cap_a0[c], t0);
  mv3(kin.oMi[par].R, m.cap_a1[c], t1);
  for (int k = 0; k < 3; ++k) { a0[k] = kin.oMi[par].p[k] + t0[k]; a1[k] = kin.oMi[par].p[k] + t1[k]; }
}
// distance of pair k and (optionally) its gradient Rq[nv]
double collision_residual(const agx_model& m, const Kin& kin, int k, double* Rq) {
  const int ia = m.pair_a[k], ib = m.pair_b[k];
  double a0[3], a1[3], b0[3], b1[3];
  capsule_world(m, kin, ia, a0, a1);
  capsule_world(m, kin, ib, b0, b1);
  CapsuleHit h;
  capsule_pair(a0, a1, m.cap_radius[ia], b0, b1, m.cap_radius[ib], h);
  if (Rq) {
    for (int j = 0; j < m.nv; ++j) Rq[j] = 0.0;
    // world linear velocity of a point c attached to joint `par` due to joint j: z_j x (c - p_j) = w x c + v,
    // with the world axis J_j = [v; w] of forward_kinematics
    for (int side = 0; side < 2; ++side) {
      const double* c = side == 0 ? h.ca : h.cb;
      const double sgn = side == 0 ? 1.0 : -1.0;
      for (int j = m.cap_parent[side == 0 ? ia : ib]; j >= 0; j = m.parent[j]) {
        double wxc[3];
        cross3(kin.J[j] + 3, c, wxc);
        double vel[3] = {wxc[0] + kin.J[j][0], wxc[1] + kin.J[j][1], wxc[2] + kin.J[j][2]};
        Rq[j] += sgn * dot3(h.n, vel);
      }
    }
  }
  return h.dist;
}
// QuadExp activation of a scalar residual
inline void quadexp(double r, double alpha, double& a, double& ar, double& arr) {
  a = std::exp(-r * r / alpha);
  ar = -2.0 * r / alpha * a;
  arr = (4.0 * r * r / (alpha * alpha) - 2.0 / alpha) * a;
}

// DifferentialActionModelFreeFwdDynamics::calc (armature path): a = (M + diag(arm))^-1 (u - nle)
bool forward_dynamics(const agx_model& m, const double* q, const double* v, const double* u, double* a,
                      double* Minv /* may be null */) {
  const int nv = m.nv;
  double M[MAXV * MAXV], L[MAXV * MAXV], b[MAXV], zero[MAXV] = {0};
  crba(m, q, M);
  for (int i = 0; i < nv; ++i) M[i * nv + i] += m.armature[i];
  rnea(m, q, v, zero, b);
  if (!cholesky(nv, M, L)) return false;
  if (Minv) {
    for (int j = 0; j < nv; ++j) {
      double e[MAXV] = {0};
      e[j] = 1.0;
      chol_solve(nv, L, e);
      for (int i = 0; i < nv; ++i) Minv[i * nv + j] = e[i];
    }
    for (int i = 0; i < nv; ++i) {
      double s = 0;
      for (int j = 0; j < nv; ++j) s += Minv[i * nv + j] * (u[j] - b[j]);
      a[i] = s;
    }
  } else {
    for (int i = 0; i < nv; ++i) a[i] = u[i] - b[i];
    chol_solve(nv, L, a);
  }
  return true;
}

// cost of one node; dt < 0 marks the terminal node (unscaled, no control)
double node_cost(const agx_model& m, const NodeRef& ref, const double* x, const double* u, bool terminal,
                 double* rpose_out) {
  const int nv = m.nv, nx = 2 * nv;
  double c = 0;
  for (int i = 0; i < nx; ++i) { double r = x[i] - ref.xref[i]; c += 0.5 * ref.wx[i] * r * r; }
  if (!terminal)
    for (int i = 0; i < nv; ++i) { double r = u[i] - ref.uref[i]; c += 0.5 * ref.wu[i] * r * r; }
  Kin kin;
  forward_kinematics(m, x, kin);
  double r6[6];
  pose_residual(m, kin, ref, r6, nullptr);
  for (int i = 0; i < 6; ++i) c += 0.5 * ref.wpose[i] * r6[i] * r6[i];
  if (rpose_out) std::memcpy(rpose_out, r6, sizeof r6);
  for (int k = 0; k < m.n_pairs; ++k) {
    double a, ar, arr;
    quadexp(collision_residual(m, kin, k, nullptr), m.col_alpha, a, ar, arr);
    c += ref.wcol[k] * a;
  }
  return c;
}

// IntegratedActionModelEuler::calc
bool node_calc(const agx_model& m, const double* refrec, double dt, bool terminal, const double* x, const double* u,
               double* xnext, double* cost) {
  const int nv = m.nv, nx = 2 * nv;
  NodeRef ref = make_ref(refrec, nv);
  const double l = node_cost(m, ref, x, u, terminal, nullptr);
  if (terminal) {
    for (int i = 0; i < nx; ++i) xnext[i] = x[i];
    *cost = l;
    return true;
  }
  // Crocoddyl's calc always forms Minv (cholesky::computeMinv) and a = Minv (u - nle)
  double a[MAXV], Minv[MAXV * MAXV];
  if (!forward_dynamics(m, x, x + nv, u, a, Minv)) return false;
  for (int i = 0; i < nv; ++i) {
    const double dq = x[nv + i] * dt + a[i] * (dt * dt);
    const double dvv = a[i] * dt;
    xnext[i] = x[i] + dq;
    xnext[nv + i] = x[nv + i] + dvv;
  }
  *cost = dt * l;
  return true;
}

// IntegratedActionModelEuler::calc + calcDiff
bool node_calc_diff(const agx_model& m, const double* refrec, double dt, bool terminal, const double* x,
                    const double* u, NodeData& d) {
  const int nv = m.nv, nx = 2 * nv;
  NodeRef ref = make_ref(refrec, nv);
  const double* q = x;
  const double* v = x + nv;
  for (int i = 0; i < nx * nx; ++i) d.Fx[i] = d.Lxx[i] = 0;
  for (int i = 0; i < nx * nv; ++i) d.Fu[i] = d.Lxu[i] = 0;
  for (int i = 0; i < nv * nv; ++i) d.Luu[i] = 0;
  for (int i = 0; i < nv; ++i) d.Lu[i] = 0;
  // ---- costs (differential) ----
  double l = 0;
  double Lx[MAXX], Lu[MAXV];
  for (int i = 0; i < nx; ++i) {
    const double r = x[i] - ref.xref[i];
    l += 0.5 * ref.wx[i] * r * r;
    Lx[i] = ref.wx[i] * r;
    d.Lxx[i * nx + i] = ref.wx[i];
  }
  if (!terminal)
    for (int i = 0; i < nv; ++i) {
      const double r = u[i] - ref.uref[i];
      l += 0.5 * ref.wu[i] * r * r;
      Lu[i] = ref.wu[i] * r;
      d.Luu[i * nv + i] = ref.wu[i];
    }
  Kin kin;
  forward_kinematics(m, q, kin);
  double r6[6], Rq[6 * MAXV];
  pose_residual(m, kin, ref, r6, Rq);
  std::memcpy(d.rpose, r6, sizeof r6);
  for (int k = 0; k < 6; ++k) l += 0.5 * ref.wpose[k] * r6[k] * r6[k];
  for (int i = 0; i < nv; ++i) {
    double s = 0;
    for (int k = 0; k < 6; ++k) s += Rq[k * nv + i] * (ref.wpose[k] * r6[k]);
    Lx[i] += s;
    for (int j = 0; j < nv; ++j) {
      double h = 0;
      for (int k = 0; k < 6; ++k) h += Rq[k * nv + i] * ref.wpose[k] * Rq[k * nv + j];
      d.Lxx[i * nx + j] += h;
    }
  }
  for (int k = 0; k < m.n_pairs; ++k) {
    double Cq[MAXV], a, ar, arr;
    const double r = collision_residual(m, kin, k, Cq);
    quadexp(r, m.col_alpha, a, ar, arr);
    d.rcol[k] = r;
    l += ref.wcol[k] * a;
    for (int i = 0; i < nv; ++i) {
      Lx[i] += ref.wcol[k] * ar * Cq[i];
      for (int j = 0; j < nv; ++j) d.Lxx[i * nx + j] += ref.wcol[k] * arr * Cq[i] * Cq[j];
    }
  }
  if (terminal) {
    for (int i = 0; i < nx; ++i) { d.xnext[i] = x[i]; d.Lx[i] = Lx[i]; d.Fx[i * nx + i] = 1.0; }
    d.cost = l;
    for (int i = 0; i < nv; ++i) d.a[i] = 0;
    return true;
  }
  // ---- dynamics ----
  if (!forward_dynamics(m, q, v, u, d.a, d.Minv)) return false;
  double tau[MAXV], dtq[MAXV * MAXV], dtv[MAXV * MAXV], Mm[MAXV * MAXV];
  rnea_derivatives(m, q, v, d.a, tau, dtq, dtv, Mm);
  double aq[MAXV * MAXV], av[MAXV * MAXV];
  for (int i = 0; i < nv; ++i)
    for (int j = 0; j < nv; ++j) {
      double s1 = 0, s2 = 0;
      for (int k = 0; k < nv; ++k) {
        s1 += d.Minv[i * nv + k] * dtq[k * nv + j];
        s2 += d.Minv[i * nv + k] * dtv[k * nv + j];
      }
      aq[i * nv + j] = -s1;
      av[i * nv + j] = -s2;
    }
  const double dt2 = dt * dt;
  for (int i = 0; i < nv; ++i) {
    d.xnext[i] = q[i] + (v[i] * dt + d.a[i] * dt2);
    d.xnext[nv + i] = v[i] + d.a[i] * dt;
    for (int j = 0; j < nv; ++j) {
      d.Fx[i * nx + j] = aq[i * nv + j] * dt2;
      d.Fx[i * nx + nv + j] = av[i * nv + j] * dt2;
      d.Fx[(nv + i) * nx + j] = aq[i * nv + j] * dt;
      d.Fx[(nv + i) * nx + nv + j] = av[i * nv + j] * dt;
      d.Fu[i * nv + j] = d.Minv[i * nv + j] * dt2;
      d.Fu[(nv + i) * nv + j] = d.Minv[i * nv + j] * dt;
    }
    d.Fx[i * nx + nv + i] += dt;
  }
  for (int i = 0; i < nx; ++i) d.Fx[i * nx + i] += 1.0;
  d.cost = dt * l;
  for (int i = 0; i < nx; ++i) d.Lx[i] = dt * Lx[i];
  for (int i = 0; i < nv; ++i) d.Lu[i] = dt * Lu[i];
  for (int i = 0; i < nx * nx; ++i) d.Lxx[i] *= dt;
  for (int i = 0; i < nv * nv; ++i) d.Luu[i] *= dt;
  return true;
}

// ----------------------------------------------------------------------------- SolverFDDP
struct Fddp {
  const agx_model* m;
  const double* refs;  // [T+1][rs]
  const double* dts;
  int T, nv, nx, rs;
  std::vector<NodeData> nd;
  std::vector<double> xs, us, fs, xs_try, us_try, K, k, Vxx, Vx, Qu, Quuk;
  double cost, cost_try, xreg, ureg, dg, dq_, dv_, d1, d2, dVexp, dV, stop, steplength;
  bool is_feasible, was_feasible;
  double x0[MAXX];

  void init(const agx_model* m_, const double* refs_, const double* dts_, int T_) {
    m = m_; refs = refs_; dts = dts_; T = T_;
    nv = m->nv; nx = 2 * nv; rs = agx_ref_size(nv);
    nd.resize(T + 1);
    xs.assign((T + 1) * nx, 0); xs_try = xs; fs = xs;
    us.assign(T * nv, 0); us_try = us;
    K.assign(T * nv * nx, 0); k.assign(T * nv, 0); Qu = k; Quuk = k;
    Vxx.assign((T + 1) * nx * nx, 0); Vx.assign((T + 1) * nx, 0);
  }
  // problem.calc + calcDiff at (xs, us), then the gaps (SolverAbstract::computeDynamicFeasibility)
  int node_threads = 1;  // > 1: problem.calc / calcDiff run over the nodes in parallel (ShootingProblem.nthreads,
                         // ocp_base_croco.py:62) -- the B = 1 latency baseline; the batched baseline keeps 1
  bool calc_diff() {
    double c = 0;
    if (node_threads > 1) {
      int bad = 0;
#pragma omp parallel for num_threads(node_threads) schedule(static) reduction(+ : bad)
      for (int t = 0; t <= T; ++t) {
        const bool term = (t == T);
        if (!node_calc_diff(*m, refs + t * rs, term ? 0.0 : dts[t], term, &xs[t * nx], term ? nullptr : &us[t * nv], nd[t]))
          bad += 1;
      }
      if (bad) return false;
      for (int t = 0; t <= T; ++t) c += nd[t].cost;
    } else {
      for (int t = 0; t <= T; ++t) {
        const bool term = (t == T);
        if (!node_calc_diff(*m, refs + t * rs, term ? 0.0 : dts[t], term, &xs[t * nx], term ? nullptr : &us[t * nv], nd[t]))
          return false;
        c += nd[t].cost;
      }
    }
    cost = c;
    if (!is_feasible) {
      for (int i = 0; i < nx; ++i) fs[i] = x0[i] - xs[i];
      for (int t = 0; t < T; ++t)
        for (int i = 0; i < nx; ++i) fs[(t + 1) * nx + i] = nd[t].xnext[i] - xs[(t + 1) * nx + i];
    } else if (!was_feasible) {
      std::fill(fs.begin(), fs.end(), 0.0);
    }
    return true;
  }
  // SolverDDP::backwardPass (+ FDDP gap terms); false = Cholesky failure / NaN
  bool backward_pass() {
    const int n = nx;
    double* VxxT = &Vxx[T * n * n];
    double* VxT = &Vx[T * n];
    for (int i = 0; i < n * n; ++i) VxxT[i] = nd[T].Lxx[i];
    for (int i = 0; i < n; ++i) VxT[i] = nd[T].Lx[i];
    if (!std::isnan(xreg)) for (int i = 0; i < n; ++i) VxxT[i * n + i] += xreg;
    if (!is_feasible)
      for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += VxxT[i * n + j] * fs[T * n + j];
        VxT[i] += s;
      }
    std::vector<double> FxTV(n * n), FuTV(nv * n), Qxx(n * n), Qxu(n * nv), Quu(nv * nv), L(nv * nv), Qx(n);
    for (int t = T - 1; t >= 0; --t) {
      const NodeData& d = nd[t];
      const double* Vp = &Vxx[(t + 1) * n * n];
      const double* vp = &Vx[(t + 1) * n];
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double s = 0;
          for (int l = 0; l < n; ++l) s += d.Fx[l * n + i] * Vp[l * n + j];
          FxTV[i * n + j] = s;
        }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double s = 0;
          for (int l = 0; l < n; ++l) s += FxTV[i * n + l] * d.Fx[l * n + j];
          Qxx[i * n + j] = d.Lxx[i * n + j] + s;
        }
      for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int l = 0; l < n; ++l) s += d.Fx[l * n + i] * vp[l];
        Qx[i] = d.Lx[i] + s;
      }
      for (int i = 0; i < nv; ++i)
        for (int j = 0; j < n; ++j) {
          double s = 0;
          for (int l = 0; l < n; ++l) s += d.Fu[l * nv + i] * Vp[l * n + j];
          FuTV[i * n + j] = s;
        }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < nv; ++j) {
          double s = 0;
          for (int l = 0; l < n; ++l) s += FxTV[i * n + l] * d.Fu[l * nv + j];
          Qxu[i * nv + j] = d.Lxu[i * nv + j] + s;
        }
      for (int i = 0; i < nv; ++i)
        for (int j = 0; j < nv; ++j) {
          double s = 0;
          for (int l = 0; l < n; ++l) s += FuTV[i * n + l] * d.Fu[l * nv + j];
          Quu[i * nv + j] = d.Luu[i * nv + j] + s;
        }
      double* Qut = &Qu[t * nv];
      for (int i = 0; i < nv; ++i) {
        double s = 0;
        for (int l = 0; l < n; ++l) s += d.Fu[l * nv + i] * vp[l];
        Qut[i] = d.Lu[i] + s;
      }
      if (!std::isnan(ureg)) for (int i = 0; i < nv; ++i) Quu[i * nv + i] += ureg;
      // computeGains
      if (!cholesky(nv, Quu.data(), L.data())) return false;
      double* Kt = &K[t * nv * n];
      double* kt = &k[t * nv];
      for (int j = 0; j < n; ++j) {
        double col[MAXV];
        for (int i = 0; i < nv; ++i) col[i] = Qxu[j * nv + i];
        chol_solve(nv, L.data(), col);
        for (int i = 0; i < nv; ++i) Kt[i * n + j] = col[i];
      }
      for (int i = 0; i < nv; ++i) kt[i] = Qut[i];
      chol_solve(nv, L.data(), kt);
      for (int i = 0; i < nv; ++i) {
        double s = 0;
        for (int j = 0; j < nv; ++j) s += Quu[i * nv + j] * kt[j];
        Quuk[t * nv + i] = s;
      }
      double* Vt = &Vxx[t * n * n];
      double* vt = &Vx[t * n];
      for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int l = 0; l < nv; ++l) s += Kt[l * n + i] * Qut[l];
        vt[i] = Qx[i] - s;
      }
      std::vector<double>& tmp = FxTV;  // reuse
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
          double s = 0;
          for (int l = 0; l < nv; ++l) s += Qxu[i * nv + l] * Kt[l * n + j];
          tmp[i * n + j] = Qxx[i * n + j] - s;
        }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Vt[i * n + j] = 0.5 * (tmp[i * n + j] + tmp[j * n + i]);
      if (!std::isnan(xreg)) for (int i = 0; i < n; ++i) Vt[i * n + i] += xreg;
      if (!is_feasible)
        for (int i = 0; i < n; ++i) {
          double s = 0;
          for (int j = 0; j < n; ++j) s += Vt[i * n + j] * fs[t * n + j];
          vt[i] += s;
        }
      for (int i = 0; i < n; ++i) if (!std::isfinite(vt[i])) return false;
      for (int i = 0; i < n * n; ++i) if (!std::isfinite(Vt[i])) return false;
    }
    return true;
  }
  // SolverFDDP::updateExpectedImprovement
  void update_expected_improvement() {
    const int n = nx;
    dg = 0; dq_ = 0;
    auto gap_terms = [&](int t) {
      const double* V = &Vxx[t * n * n];
      const double* f = &fs[t * n];
      double s1 = 0, s2 = 0;
      for (int i = 0; i < n; ++i) {
        s1 += Vx[t * n + i] * f[i];
        double r = 0;
        for (int j = 0; j < n; ++j) r += V[i * n + j] * f[j];
        s2 += f[i] * r;
      }
      dg -= s1;
      dq_ += s2;
    };
    if (!is_feasible) gap_terms(T);
    for (int t = 0; t < T; ++t) {
      double s1 = 0, s2 = 0;
      for (int i = 0; i < nv; ++i) { s1 += Qu[t * nv + i] * k[t * nv + i]; s2 += k[t * nv + i] * Quuk[t * nv + i]; }
      dg += s1;
      dq_ -= s2;
      if (!is_feasible) gap_terms(t);
    }
  }
  // SolverFDDP::forwardPass ; false = NaN met
  bool forward_pass(double alpha) {
    const int n = nx;
    double xnext[MAXX];
    for (int i = 0; i < n; ++i) xnext[i] = x0[i];
    cost_try = 0;
    for (int t = 0; t < T; ++t) {
      double* xt = &xs_try[t * n];
      if (is_feasible || alpha == 1.0) for (int i = 0; i < n; ++i) xt[i] = xnext[i];
      else for (int i = 0; i < n; ++i) xt[i] = xnext[i] + fs[t * n + i] * (alpha - 1.0);
      double dx[MAXX];
      for (int i = 0; i < n; ++i) dx[i] = xt[i] - xs[t * n + i];
      double* ut = &us_try[t * nv];
      for (int i = 0; i < nv; ++i) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += K[(t * nv + i) * n + j] * dx[j];
        ut[i] = us[t * nv + i] - k[t * nv + i] * alpha - s;
      }
      double c;
      if (!node_calc(*m, refs + t * rs, dts[t], false, xt, ut, xnext, &c)) return false;
      cost_try += c;
      if (!std::isfinite(cost_try)) return false;
      for (int i = 0; i < n; ++i) if (!std::isfinite(xnext[i])) return false;
    }
    double* xT = &xs_try[T * n];
    if (is_feasible || alpha == 1.0) for (int i = 0; i < n; ++i) xT[i] = xnext[i];
    else for (int i = 0; i < n; ++i) xT[i] = xnext[i] + fs[T * n + i] * (alpha - 1.0);
    double c, dummy[MAXX];
    node_calc(*m, refs + T * rs, 0.0, true, xT, nullptr, dummy, &c);
    cost_try += c;
    return std::isfinite(cost_try);
  }
  // SolverFDDP::expectedImprovement
  void expected_improvement() {
    const int n = nx;
    dv_ = 0;
    if (!is_feasible)
      for (int t = 0; t <= T; ++t) {
        const double* V = &Vxx[t * n * n];
        double s = 0;
        for (int i = 0; i < n; ++i) {
          double r = 0;
          for (int j = 0; j < n; ++j) r += V[i * n + j] * (xs[t * n + j] - xs_try[t * n + j]);
          s += fs[t * n + i] * r;
        }
        dv_ -= s;
      }
    d1 = dg + dv_;
    d2 = dq_ - 2 * dv_;
  }
  void inc_reg(const agx_fddp_opts& o) {
    xreg *= o.reg_incfactor;
    if (xreg > o.reg_max) xreg = o.reg_max;
    ureg = xreg;
  }
  void dec_reg(const agx_fddp_opts& o) {
    xreg /= o.reg_decfactor;
    if (xreg < o.reg_min) xreg = o.reg_min;
    ureg = xreg;
  }
  // SolverFDDP::solve
  int solve(const double* x0_, const double* xs0, const double* us0, int maxiter, const agx_fddp_opts& o,
            int* iters_out) {
    for (int i = 0; i < nx; ++i) x0[i] = x0_[i];
    std::memcpy(xs.data(), xs0, sizeof(double) * (T + 1) * nx);
    std::memcpy(us.data(), us0, sizeof(double) * T * nv);
    is_feasible = false;
    was_feasible = false;
    xreg = ureg = std::isnan(o.reg_init) ? o.reg_min : o.reg_init;
    bool recalc = true;
    stop = 0;
    int status = AGX_STATUS_MAXITER;
    int it = 0;
    for (it = 0; it < maxiter; ++it) {
      bool failed = false;
      while (true) {
        bool ok = true;
        if (recalc) ok = calc_diff();
        if (ok) ok = backward_pass();
        if (!ok) {
          recalc = false;
          inc_reg(o);
          if (xreg == o.reg_max) { failed = true; break; }
          continue;
        }
        break;
      }
      if (failed) { status = AGX_STATUS_REGMAX; break; }
      update_expected_improvement();
      recalc = false;
      for (int ia = 0; ia < o.n_alphas; ++ia) {
        steplength = std::ldexp(1.0, -ia);
        if (!forward_pass(steplength)) continue;
        dV = cost - cost_try;
        expected_improvement();
        dVexp = steplength * (d1 + 0.5 * steplength * d2);
        bool accept = false;
        // SolverFDDP::solve's acceptance test; the two forms are those of Crocoddyl >= 2.0 (default) and 1.x (agx.h)
        const bool legacy = o.accept_rule == AGX_ACCEPT_CROCODDYL1;
        if (dVexp >= 0) {
          if ((legacy ? d1 : std::fabs(d1)) < o.th_grad || dV > o.th_acceptstep * dVexp) accept = true;
        } else {
          if ((legacy || !is_feasible) && dV > o.th_acceptnegstep * dVexp) accept = true;
        }
        if (accept) {
          was_feasible = is_feasible;
          xs = xs_try;
          us = us_try;
          is_feasible = was_feasible || (steplength == 1.0);
          cost = cost_try;
          recalc = true;
          break;
        }
      }
      if (steplength > o.th_stepdec) dec_reg(o);
      if (steplength <= o.th_stepinc) {
        inc_reg(o);
        if (xreg == o.reg_max) { status = AGX_STATUS_REGMAX; ++it; break; }
      }
      stop = std::fabs(d1 + 0.5 * d2);
      if (!o.fixed_iters && was_feasible && stop < o.th_stop) { status = AGX_STATUS_CONVERGED; ++it; break; }
    }
    *iters_out = it;
    return status;
  }
};

inline const agx_model& model_of(const agx_model* models, int n_models, int b) {
  return models[n_models > 1 ? b : 0];
}

}  // namespace

// ============================================================================= C entry points
extern "C" {

int agx_ref_size(int nv) { return 6 * nv + 20; }

void agx_fddp_opts_default(agx_fddp_opts* o) {
  o->reg_min = 1e-9; o->reg_max = 1e9; o->reg_incfactor = 10.0; o->reg_decfactor = 10.0;
  o->th_grad = 1e-12; o->th_stepdec = 0.5; o->th_stepinc = 0.01; o->th_acceptstep = 0.1;
  o->th_acceptnegstep = 2.0; o->th_stop = 1e-9;
  o->reg_init = std::numeric_limits<double>::quiet_NaN();
  o->fixed_iters = 0; o->n_alphas = 10; o->eager_exit = 0; o->accept_rule = AGX_ACCEPT_CROCODDYL2; o->max_solve_time = 0.0;
}

void orc_rnea(const agx_model* m, const double* q, const double* v, const double* a, int n, double* tau) {
  const int nv = m->nv;
  for (int i = 0; i < n; ++i) rnea(*m, q + i * nv, v + i * nv, a + i * nv, tau + i * nv);
}
void orc_crba(const agx_model* m, const double* q, double* M) { crba(*m, q, M); }
void orc_rnea_derivatives(const agx_model* m, const double* q, const double* v, const double* a, double* tau,
                          double* dq, double* dv, double* M) {
  rnea_derivatives(*m, q, v, a, tau, dq, dv, M);
}
int orc_forward_dynamics(const agx_model* m, const double* q, const double* v, const double* u, double* a,
                         double* Minv) {
  return forward_dynamics(*m, q, v, u, a, Minv) ? 0 : -1;
}
void orc_frame_placement(const agx_model* m, const double* q, double* R, double* p) {
  Kin kin;
  forward_kinematics(*m, q, kin);
  SE3 f;
  frame_placement(*m, kin, f);
  std::memcpy(R, f.R, sizeof f.R);
  std::memcpy(p, f.p, sizeof f.p);
}
// collision pair k: signed distance, its gradient [nv] and the QuadExp activation (a, a', a'')
void orc_collision(const agx_model* m, const double* q, int k, double* dist, double* Rq, double* act) {
  Kin kin;
  forward_kinematics(*m, q, kin);
  *dist = collision_residual(*m, kin, k, Rq);
  if (act) quadexp(*dist, m->col_alpha, act[0], act[1], act[2]);
}
// frame Jacobians: LOCAL (6 x nv) and LOCAL_WORLD_ALIGNED (6 x nv)
void orc_frame_jacobian(const agx_model* m, const double* q, double* J_local, double* J_lwa) {
  Kin kin;
  forward_kinematics(*m, q, kin);
  SE3 f;
  frame_placement(*m, kin, f);
  const int nv = m->nv;
  frame_jacobian_local(*m, kin, f, J_local);
  for (int j = 0; j < nv; ++j) {
    double l[3] = {J_local[0 * nv + j], J_local[1 * nv + j], J_local[2 * nv + j]};
    double w[3] = {J_local[3 * nv + j], J_local[4 * nv + j], J_local[5 * nv + j]};
    double a[3], b[3];
    mv3(f.R, l, a);
    mv3(f.R, w, b);
    for (int r = 0; r < 3; ++r) { J_lwa[r * nv + j] = a[r]; J_lwa[(3 + r) * nv + j] = b[r]; }
  }
}
void orc_log6(const double* R, const double* p, double* out) {
  SE3 M;
  std::memcpy(M.R, R, sizeof M.R);
  std::memcpy(M.p, p, sizeof M.p);
  log6(M, out);
}
void orc_Jlog6(const double* R, const double* p, double* J) {
  SE3 M;
  std::memcpy(M.R, R, sizeof M.R);
  std::memcpy(M.p, p, sizeof M.p);
  Jlog6(M, J);
}

// problem.calc over B problems
int orc_calc(const agx_model* models, int n_models, const double* refs, const double* dts, int B, int T,
             const double* xs, const double* us, double* out_cost, double* out_xnext) {
  const int nv = models[0].nv, nx = 2 * nv, rs = agx_ref_size(nv);
  int err = 0;
#pragma omp parallel for schedule(dynamic) reduction(| : err)
  for (int bt = 0; bt < B * (T + 1); ++bt) {
    const int b = bt / (T + 1), t = bt % (T + 1);
    const bool term = t == T;
    double xn[MAXX], c;
    if (!node_calc(model_of(models, n_models, b), refs + ((size_t)b * (T + 1) + t) * rs, term ? 0.0 : dts[t], term,
                   xs + ((size_t)b * (T + 1) + t) * nx, term ? nullptr : us + ((size_t)b * T + t) * nv, xn, &c))
      err |= 1;
    if (out_cost) out_cost[bt] = c;
    if (out_xnext) std::memcpy(out_xnext + (size_t)bt * nx, xn, sizeof(double) * nx);
  }
  return err ? -1 : 0;
}

// problem.calc + calcDiff over B problems, dense outputs (same layout as agx_calc_diff)
int orc_calc_diff(const agx_model* models, int n_models, const double* refs, const double* dts, int B, int T,
                  const double* xs, const double* us, double* out_cost, double* out_xnext, double* Fx, double* Fu,
                  double* Lx, double* Lu, double* Lxx, double* Lxu, double* Luu) {
  const int nv = models[0].nv, nx = 2 * nv, rs = agx_ref_size(nv);
  int err = 0;
#pragma omp parallel for schedule(dynamic) reduction(| : err)
  for (int bt = 0; bt < B * (T + 1); ++bt) {
    const int b = bt / (T + 1), t = bt % (T + 1);
    const bool term = t == T;
    static thread_local NodeData d;
    if (!node_calc_diff(model_of(models, n_models, b), refs + ((size_t)b * (T + 1) + t) * rs, term ? 0.0 : dts[t], term,
                        xs + ((size_t)b * (T + 1) + t) * nx, term ? nullptr : us + ((size_t)b * T + t) * nv, d))
      err |= 1;
    const size_t o = bt;
    if (out_cost) out_cost[o] = d.cost;
    if (out_xnext) std::memcpy(out_xnext + o * nx, d.xnext, sizeof(double) * nx);
    if (Fx) std::memcpy(Fx + o * nx * nx, d.Fx, sizeof(double) * nx * nx);
    if (Fu) std::memcpy(Fu + o * nx * nv, d.Fu, sizeof(double) * nx * nv);
    if (Lx) std::memcpy(Lx + o * nx, d.Lx, sizeof(double) * nx);
    if (Lu) std::memcpy(Lu + o * nv, d.Lu, sizeof(double) * nv);
    if (Lxx) std::memcpy(Lxx + o * nx * nx, d.Lxx, sizeof(double) * nx * nx);
    if (Lxu) std::memcpy(Lxu + o * nx * nv, d.Lxu, sizeof(double) * nx * nv);
    if (Luu) std::memcpy(Luu + o * nv * nv, d.Luu, sizeof(double) * nv * nv);
  }
  return err ? -1 : 0;
}

// problem.rollout
int orc_rollout(const agx_model* models, int n_models, const double* refs, const double* dts, int B, int T,
                const double* x0, const double* us, double* out_xs) {
  const int nv = models[0].nv, nx = 2 * nv, rs = agx_ref_size(nv);
  int err = 0;
#pragma omp parallel for schedule(dynamic) reduction(| : err)
  for (int b = 0; b < B; ++b) {
    double* xs = out_xs + (size_t)b * (T + 1) * nx;
    std::memcpy(xs, x0 + (size_t)b * nx, sizeof(double) * nx);
    for (int t = 0; t < T; ++t) {
      double c;
      if (!node_calc(model_of(models, n_models, b), refs + ((size_t)b * (T + 1) + t) * rs, dts[t], false, xs + t * nx,
                     us + ((size_t)b * T + t) * nv, xs + (t + 1) * nx, &c))
        err |= 1;
    }
  }
  return err ? -1 : 0;
}

// IntegratedActionModelEuler.calc -> xnext for n pairs (costs ignored)
int orc_integrate(const agx_model* m, const double* x, const double* u, double dt, int n, double* out) {
  const int nv = m->nv, nx = 2 * nv;
  for (int i = 0; i < n; ++i) {
    double a[MAXV], Minv[MAXV * MAXV];
    if (!forward_dynamics(*m, x + i * nx, x + i * nx + nv, u + i * nv, a, Minv)) return -1;
    for (int j = 0; j < nv; ++j) {
      out[i * nx + j] = x[i * nx + j] + (x[i * nx + nv + j] * dt + a[j] * (dt * dt));
      out[i * nx + nv + j] = x[i * nx + nv + j] + a[j] * dt;
    }
  }
  return 0;
}

// SolverFDDP.solve over B problems; one problem per OpenMP thread (nthreads <= 0: all cores)
int orc_solve(const agx_model* models, int n_models, const double* refs, const double* dts, int B, int T,
              const double* x0, const double* xs_ws, const double* us_ws, int max_iter, const agx_fddp_opts* opts,
              double* out_xs, double* out_us, double* out_K, double* out_k, double* out_cost, int32_t* out_iters,
              int32_t* out_status, double* out_stop, int nthreads) {
  const int nv = models[0].nv, nx = 2 * nv, rs = agx_ref_size(nv);
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
  {
    Fddp s;
    bool inited = false;
#pragma omp for schedule(dynamic)
    for (int b = 0; b < B; ++b) {
      const agx_model& m = model_of(models, n_models, b);
      if (!inited || n_models > 1) { s.init(&m, nullptr, dts, T); inited = true; }
      s.m = &m;
      s.refs = refs + (size_t)b * (T + 1) * rs;
      int iters = 0;
      const int st = s.solve(x0 + (size_t)b * nx, xs_ws + (size_t)b * (T + 1) * nx, us_ws + (size_t)b * T * nv,
                             max_iter, *opts, &iters);
      std::memcpy(out_xs + (size_t)b * (T + 1) * nx, s.xs.data(), sizeof(double) * (T + 1) * nx);
      std::memcpy(out_us + (size_t)b * T * nv, s.us.data(), sizeof(double) * T * nv);
      if (out_K) std::memcpy(out_K + (size_t)b * T * nv * nx, s.K.data(), sizeof(double) * T * nv * nx);
      if (out_k) std::memcpy(out_k + (size_t)b * T * nv, s.k.data(), sizeof(double) * T * nv);
      out_cost[b] = s.cost;
      out_iters[b] = iters;
      out_status[b] = st;
      if (out_stop) out_stop[b] = s.stop;
    }
  }
  return 0;
}

// Closed-loop single-problem MPC on the CPU (the B = 1 latency baseline of bench.py; BASELINE config 1): per tick the
// horizon window of the reference stream (uniform steps: point k + t for node t; past the end the last point repeats),
// FDDP from the shifted previous solution with early exit, then the plant = the OCP's integrator.  Mirrors MPC.run
// (mpc.py:32-66) with WarmStartShiftPreviousSolution; out_ns[k] = wall time of tick k's window copy + solve
// (mpc.py:52-64 stamps the same span as duration_ocp_solve_ns).  node_threads = ShootingProblem.nthreads.
int orc_mpc_latency(const agx_model* m, const double* stream_refs, int n_points, const double* dts, int T,
                    const double* x_init, const double* xs_init, const double* us_init, int ticks, int max_iter,
                    const agx_fddp_opts* opts, int node_threads, long long* out_ns, int32_t* out_iters,
                    double* out_x_final) {
  const int nv = m->nv, nx = 2 * nv, rs = agx_ref_size(nv);
  std::vector<double> refs((size_t)(T + 1) * rs), x(x_init, x_init + nx), xs(xs_init, xs_init + (size_t)(T + 1) * nx),
      us(us_init, us_init + (size_t)T * nv), xs2(xs.size()), us2(us.size());
  Fddp s;
  s.init(m, refs.data(), dts, T);
  s.node_threads = node_threads > 1 ? node_threads : 1;
  for (int k = 0; k < ticks; ++k) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t <= T; ++t) {
      int pnt = k + t;
      if (pnt >= n_points) pnt = n_points - 1;
      std::memcpy(&refs[(size_t)t * rs], stream_refs + (size_t)pnt * rs, sizeof(double) * rs);
    }
    int iters = 0;
    s.solve(x.data(), xs.data(), us.data(), max_iter, *opts, &iters);
    const auto t1 = std::chrono::steady_clock::now();
    out_ns[k] = (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
    out_iters[k] = iters;
    // plant step with the first control, then the shifted warm start (uniform steps)
    double xn[MAXX], c;
    if (!node_calc(*m, refs.data(), dts[0], false, x.data(), s.us.data(), xn, &c)) return -1;
    for (int i = 0; i < nx; ++i) x[i] = xn[i];
    for (int t = 0; t < T; ++t) std::memcpy(&xs2[(size_t)t * nx], &s.xs[(size_t)(t + 1) * nx], sizeof(double) * nx);
    std::memcpy(&xs2[(size_t)T * nx], &s.xs[(size_t)T * nx], sizeof(double) * nx);
    for (int t = 0; t < T; ++t)
      std::memcpy(&us2[(size_t)t * nv], &s.us[(size_t)(t + 1 < T ? t + 1 : t) * nv], sizeof(double) * nv);
    xs = xs2; us = us2;
    for (int i = 0; i < nx; ++i) xs[i] = x[i];
  }
  for (int i = 0; i < nx; ++i) out_x_final[i] = x[i];
  return 0;
}

// ----------------------------------------------------------------------------- SolverCSQP, unconstrained (SQP)
// mim_solvers.SolverCSQP as the reference runs it (ocp_base_croco.py:64-75, :172) when no constraint is active.
// mim_solvers is not in the reference tree; this restates its published iteration and is pinned
// on the reference's golden file: the gains exactly (KAT-3), the iterate at the KKT stop to 6e-5 (KAT-9).
struct Sqp : Fddp {
  std::vector<double> dx, du, lam;
  double kkt = 0, merit = 0, gap_l1 = 0;

  // equality-constrained QP at the current iterate: Riccati sweep with `reg`, linear rollout, multipliers, KKT norm
  bool direction(double reg) {
    const int n = nx;
    xreg = ureg = reg;
    if (!backward_pass()) return false;
    dx.assign((T + 1) * n, 0.0); du.assign(T * nv, 0.0); lam.assign((T + 1) * n, 0.0);
    for (int i = 0; i < n; ++i) dx[i] = fs[i];
    for (int t = 0; t < T; ++t) {
      const NodeData& d = nd[t];
      for (int i = 0; i < nv; ++i) {
        double s = -k[t * nv + i];
        for (int j = 0; j < n; ++j) s -= K[(t * nv + i) * n + j] * dx[t * n + j];
        du[t * nv + i] = s;
      }
      for (int i = 0; i < n; ++i) {
        double s = fs[(t + 1) * n + i];
        for (int j = 0; j < n; ++j) s += d.Fx[i * n + j] * dx[t * n + j];
        for (int j = 0; j < nv; ++j) s += d.Fu[i * nv + j] * du[t * nv + j];
        dx[(t + 1) * n + i] = s;
      }
    }
    // multipliers of the QP (adjoint recursion) and the KKT residual of the NONLINEAR problem with them
    double v = 0;
    for (int i = 0; i < n; ++i) {
      double s = nd[T].Lx[i];
      for (int j = 0; j < n; ++j) s += nd[T].Lxx[i * n + j] * dx[T * n + j];
      lam[T * n + i] = s;
      v = std::fmax(v, std::fabs(nd[T].Lx[i] - s));
    }
    for (int t = T - 1; t >= 0; --t) {
      const NodeData& d = nd[t];
      const double* ln = &lam[(t + 1) * n];
      for (int i = 0; i < nv; ++i) {
        double s = d.Lu[i];
        for (int l = 0; l < n; ++l) s += d.Fu[l * nv + i] * ln[l];
        v = std::fmax(v, std::fabs(s));
      }
      for (int i = 0; i < n; ++i) {
        double g = d.Lx[i];
        for (int l = 0; l < n; ++l) g += d.Fx[l * n + i] * ln[l];
        double h = 0;
        for (int j = 0; j < n; ++j) h += d.Lxx[i * n + j] * dx[t * n + j];
        for (int j = 0; j < nv; ++j) h += d.Lxu[i * nv + j] * du[t * nv + j];
        lam[t * n + i] = g + h;
        v = std::fmax(v, std::fabs(h));  // Lx + Fx^T l' - l
      }
    }
    gap_l1 = 0;
    for (size_t i = 0; i < fs.size(); ++i) { v = std::fmax(v, std::fabs(fs[i])); gap_l1 += std::fabs(fs[i]); }
    kkt = v;
    return true;
  }
  // merit of (xs + a dx, us + a du); false on a failed calc
  bool try_step(double a, double mu, double* merit_try) {
    const int n = nx;
    for (int i = 0; i < (T + 1) * n; ++i) xs_try[i] = xs[i] + a * dx[i];
    for (int i = 0; i < T * nv; ++i) us_try[i] = us[i] + a * du[i];
    double c = 0, g = 0;
    for (int i = 0; i < n; ++i) g += std::fabs(x0[i] - xs_try[i]);
    for (int t = 0; t <= T; ++t) {
      const bool term = t == T;
      double xn[MAXX], ct;
      if (!node_calc(*m, refs + t * rs, term ? 0.0 : dts[t], term, &xs_try[t * n], term ? nullptr : &us_try[t * nv], xn,
                     &ct))
        return false;
      c += ct;
      if (!term)
        for (int i = 0; i < n; ++i) g += std::fabs(xn[i] - xs_try[(t + 1) * n + i]);
    }
    cost_try = c;
    *merit_try = c + mu * g;
    return true;
  }
  int solve_sqp(const double* x0_, const double* xs_ws, const double* us_ws, int max_iter, const agx_sqp_opts& o,
                int* iters_out) {
    for (int i = 0; i < nx; ++i) x0[i] = x0_[i];
    std::memcpy(xs.data(), xs_ws, sizeof(double) * (T + 1) * nx);
    std::memcpy(us.data(), us_ws, sizeof(double) * T * nv);
    is_feasible = false; was_feasible = false;
    int status = AGX_STATUS_MAXITER, iters = 0;
    stop = 0;
    bool have_diff = false;
    // the regularisation follows SolverDDP's schedule, which the mim_solvers solvers inherit: floor o.reg, x10 after a
    // failed factorisation or a step length <= th_stepinc (0.01, which includes a failed line search), /10 after a step
    // length > th_stepdec (0.5); reaching reg_max = 1e9 ends the problem
    const double reg_min = o.reg, reg_max = 1e9, reg_factor = 10.0, th_stepdec = 0.5, th_stepinc = 0.01;
    double reg = o.reg;
    for (int it = 0; it < max_iter; ++it) {
      if (!calc_diff()) { status = AGX_STATUS_NAN; break; }
      have_diff = true;
      bool dir_ok = direction(reg);
      while (!dir_ok) {
        reg = std::fmin(reg * reg_factor, reg_max);
        if (reg == reg_max) break;
        dir_ok = direction(reg);
      }
      if (!dir_ok) { status = AGX_STATUS_REGMAX; break; }
      stop = kkt;
      if (!(kkt == kkt)) { status = AGX_STATUS_NAN; break; }
      if (kkt <= o.termination_tolerance) { status = AGX_STATUS_CONVERGED; break; }
      merit = cost + o.mu * gap_l1;
      bool accepted = false;
      double steplength = 1.0;
      for (int n = 0; n < o.n_alphas; ++n) {
        double mt;
        steplength = std::ldexp(1.0, -n);
        if (!try_step(steplength, o.mu, &mt)) continue;
        if (mt < merit) { accepted = true; break; }
      }
      if (accepted) {
        xs.swap(xs_try); us.swap(us_try);
        have_diff = false;
      }
      ++iters;  // the solver's iteration counter counts every pass of its loop, step taken or not
      if (steplength > th_stepdec) reg = std::fmax(reg / reg_factor, reg_min);
      if (steplength <= th_stepinc) {
        reg = std::fmin(reg * reg_factor, reg_max);
        if (reg == reg_max) { status = AGX_STATUS_REGMAX; break; }
      }
    }
    // the gains the solver holds: its last backward pass carries sigma + reg on Quu, Qxx, Vxx_T
    if (status != AGX_STATUS_NAN) {
      if (!have_diff && !calc_diff()) status = AGX_STATUS_NAN;
      else {
        xreg = ureg = o.sigma + reg;
        if (!backward_pass() && status != AGX_STATUS_REGMAX) status = AGX_STATUS_REGMAX;
      }
    }
    *iters_out = iters;
    return status;
  }
};

// One Riccati sweep at (xs, us) for a single problem with a proximal sigma on Quu, Qxx(t>0), Vxx_T
// (mim_solvers SolverCSQP backward pass, unconstrained case) -> K [T][nu][nx], k [T][nu].  Used only
// by golden KAT-3.  gaps are taken as x0 - xs_0 and xnext_t - xs_{t+1}.
int orc_riccati_sigma(const agx_model* m, const double* refs, const double* dts, int T, const double* x0,
                      const double* xs, const double* us, double sigma, double* out_K, double* out_k,
                      double* out_kkt /* [2]: |Lu + Fu^T lambda|_inf , |gaps|_inf at (xs,us) with lambda from costates */) {
  Fddp s;
  s.init(m, refs, dts, T);
  const int nv = s.nv, n = s.nx;
  for (int i = 0; i < n; ++i) s.x0[i] = x0[i];
  std::memcpy(s.xs.data(), xs, sizeof(double) * (T + 1) * n);
  std::memcpy(s.us.data(), us, sizeof(double) * T * nv);
  s.is_feasible = false;
  s.was_feasible = false;
  if (!s.calc_diff()) return -1;
  // stationarity: lambda_T = Lx_T ; lambda_t = Lx_t + Fx_t^T lambda_{t+1}
  if (out_kkt) {
    std::vector<double> lam(n), nl(n);
    for (int i = 0; i < n; ++i) lam[i] = s.nd[T].Lx[i];
    double ku = 0, kg = 0;
    for (int t = T - 1; t >= 0; --t) {
      const NodeData& d = s.nd[t];
      for (int i = 0; i < nv; ++i) {
        double r = d.Lu[i];
        for (int l = 0; l < n; ++l) r += d.Fu[l * nv + i] * lam[l];
        ku = std::fmax(ku, std::fabs(r));
      }
      for (int i = 0; i < n; ++i) {
        double r = d.Lx[i];
        for (int l = 0; l < n; ++l) r += d.Fx[l * n + i] * lam[l];
        nl[i] = r;
      }
      lam = nl;
    }
    for (size_t i = 0; i < s.fs.size(); ++i) kg = std::fmax(kg, std::fabs(s.fs[i]));
    out_kkt[0] = ku;
    out_kkt[1] = kg;
  }
  // sigma-regularised sweep: emulate by adding sigma to the diagonals through xreg/ureg,
  // except that Qxx at t = 0 is not regularised (irrelevant for K).
  s.xreg = sigma;
  s.ureg = sigma;
  if (!s.backward_pass()) return -2;
  std::memcpy(out_K, s.K.data(), sizeof(double) * T * nv * n);
  std::memcpy(out_k, s.k.data(), sizeof(double) * T * nv);
  return 0;
}

void agx_sqp_opts_default(agx_sqp_opts* o) {
  o->sigma = 1e-6; o->reg = 1e-9; o->mu = 10.0; o->termination_tolerance = 1e-3; o->n_alphas = 10; o->eager_exit = 0; o->max_solve_time = 0.0;
}

// SolverCSQP.solve (unconstrained) over B problems; one problem per OpenMP thread
int orc_solve_sqp(const agx_model* models, int n_models, const double* refs, const double* dts, int B, int T,
                  const double* x0, const double* xs_ws, const double* us_ws, int max_iter, const agx_sqp_opts* opts,
                  double* out_xs, double* out_us, double* out_K, double* out_k, double* out_cost, int32_t* out_iters,
                  int32_t* out_status, double* out_stop, int nthreads) {
  const int nv = models[0].nv, nx = 2 * nv, rs = agx_ref_size(nv);
#ifdef _OPENMP
  if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
  nthreads = 1;
#endif
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
  for (int b = 0; b < B; ++b) {
    Sqp s;
    s.init(&model_of(models, n_models, b), refs + (size_t)b * (T + 1) * rs, dts, T);
    int iters = 0;
    const int st = s.solve_sqp(x0 + (size_t)b * nx, xs_ws + (size_t)b * (T + 1) * nx, us_ws + (size_t)b * T * nv,
                               max_iter, *opts, &iters);
    std::memcpy(out_xs + (size_t)b * (T + 1) * nx, s.xs.data(), sizeof(double) * (T + 1) * nx);
    std::memcpy(out_us + (size_t)b * T * nv, s.us.data(), sizeof(double) * T * nv);
    if (out_K) std::memcpy(out_K + (size_t)b * T * nv * nx, s.K.data(), sizeof(double) * T * nv * nx);
    if (out_k) std::memcpy(out_k + (size_t)b * T * nv, s.k.data(), sizeof(double) * T * nv);
    out_cost[b] = s.cost;
    out_iters[b] = iters;
    out_status[b] = st;
    if (out_stop) out_stop[b] = s.stop;
  }
  return 0;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
