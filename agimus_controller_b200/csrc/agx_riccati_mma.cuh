// agx_riccati_mma.cuh — the backward Riccati sweep on the FP64 tensor cores (DMMA m8n8k4), one WARP per problem.
//
// Same algorithm and outputs as the octet sweep (SolverFDDP::backwardPass / computeGains /
// updateExpectedImprovement), different mapping: the 14-dimensional state is padded to two
// 8-blocks [q 0..6, pad, v 0..6, pad], every matrix lives in 8x8 tiles, and the products
//   Z = S^T V', Vs = Z S, W = Vs G + [Zq 0], Qxx = Lxx + [V'qq 0; 0 0] + G^T W + [Zq^T G; 0],
//   Qux = N^T W, Quu = Luu + N^T Vs N, Vxx = Qxx - Qux^T K
// become 34 tile products D(8x8) += A(8x4) B(4x8) per node (mma.sync.aligned.m8n8k4.f64): one instruction does
// 256 FMAs, so the sweep issues ~8x fewer FP64 instructions than with DFMA and, with one warp per problem,
// 4x more warps are in flight.  The vectors ride in the pad column 15: W[:,15] = S^T v', so that
// (G^T W)[:,15] = G^T S^T v' (-> Qx), (N^T W)[:,15] = N^T S^T v' (-> Qu), K[:,15] = Quu^-1 Qu = k and
// (Qxx - Qux^T K)[:,15] = Qx - Qxu k = Vx come out of the same tile products.
//
// Fragment layouts (lane T, g = T / 4, q = T % 4): A-operand a = A[g][4 kc + q]; B-operand b = B[4 kc + q][g];
// accumulator c[e] = C[g][2 q + e].  B-operands of G and N are read straight from the dynamics record with
// coalesced 256-byte loads (the record stores element [i][j] at (field + i) * 8 + j); products of products go
// through a per-warp shared-memory board.  The 7x7 Cholesky of Quu runs redundantly in registers on every lane
// (no barrier inside); lanes 0..15 each solve one column of [Qux | Qu].
#ifndef AGX_RICCATI_MMA_CUH_
#define AGX_RICCATI_MMA_CUH_

namespace agx {

constexpr int MS8 = 12;    // row stride (doubles) of the 8-column boards: conflict-free operand loads
constexpr int MS16 = 20;   // row stride of the 16-column boards
constexpr int MSV = 17;    // row stride of the 16x16 board
constexpr int MB_VS = 0;        // [8][12]  Vs, later Quu (handed to the factorisation)
constexpr int MB_Z0 = 96;       // [8][12]  Zq
constexpr int MB_VN = 192;      // [8][12]  Vs N
constexpr int MB_W = 288;       // [8][20]  W, later K
constexpr int MB_QUX = 448;     // [8][20]  [Qux | Qu]
constexpr int MB_V = 608;       // [16][17] unsymmetrised Vxx
constexpr int MB_FS = 880;      // [16]     gap of the node (padded indexing)
constexpr int MB_ST = 896;      // [15][32] next node's per-lane operands, staged by cp.async
constexpr int MB_SIZE = 896 + 15 * 32;

AGX_DEV double warp_sum(double x) {
  x += __shfl_xor_sync(0xffffffffu, x, 16, 32);
  x += __shfl_xor_sync(0xffffffffu, x, 8, 32);
  x += __shfl_xor_sync(0xffffffffu, x, 4, 32);
  x += __shfl_xor_sync(0xffffffffu, x, 2, 32);
  x += __shfl_xor_sync(0xffffffffu, x, 1, 32);
  return x;
}

// in-register Cholesky as chol7_registers, reading M[i][k] at sm_M[i * stride + k]
AGX_DEV bool chol7_registers_strided(const double* sm_M, int stride, double* A /*28*/, double* rinv /*7*/) {
#pragma unroll
  for (int k = 0; k < NJ; ++k)
#pragma unroll
    for (int i = 0; i < NJ; ++i)
      if (i >= k) A[lidx_(i, k)] = sm_M[i * stride + k];
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

// per-lane operands of one node
struct MmaNodeIn {
  double gB[2][2];   // G[4 kc + q][8 tc + g]  (tc = 0: dt aq, tc = 1: I + dt av)
  double nB[2];      // N[4 kc + q][g]         (dt Minv)
  double lqq[2];     // Lqq[g][2 q + e]
  double lvv, luu, lq, lv, lu;  // entries g of the diagonals / gradients
  double fs0, fs1;   // gap entries g and 7 + g
  double h;
};

// The 15 per-lane operands of a node are copied global -> shared (8-byte
// cp.async, no registers) while the previous node is processed; every lane reads back only its own slots.
AGX_DEV void mma_stage(double* st, const double* __restrict__ R, const double* __restrict__ C,
                       const double* __restrict__ fs, bool gaps, int g, int q, int lane) {
  const int gg = g < NJ ? g : NJ - 1;
#pragma unroll
  for (int kc = 0; kc < 2; ++kc) {
    const int r = 4 * kc + q, rr = r < NJ ? r : NJ - 1;
    AGX_CP_ASYNC8(st + (0 + kc) * 32 + lane, R + (RK_AQ + rr) * 8 + gg);
    AGX_CP_ASYNC8(st + (2 + kc) * 32 + lane, R + (RK_AV + rr) * 8 + gg);
    AGX_CP_ASYNC8(st + (4 + kc) * 32 + lane, R + (RK_MI + rr) * 8 + gg);
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int c = 2 * q + e < NJ ? 2 * q + e : NJ - 1;
    AGX_CP_ASYNC8(st + (6 + e) * 32 + lane, C + CK_LQQ + (gg >= c ? lidx_(gg, c) : lidx_(c, gg)));
  }
  AGX_CP_ASYNC8(st + 8 * 32 + lane, C + CK_LVV + gg);
  AGX_CP_ASYNC8(st + 9 * 32 + lane, C + CK_LUU + gg);
  AGX_CP_ASYNC8(st + 10 * 32 + lane, C + CK_LQ + gg);
  AGX_CP_ASYNC8(st + 11 * 32 + lane, C + CK_LV + gg);
  AGX_CP_ASYNC8(st + 12 * 32 + lane, C + CK_LU + gg);
  if (gaps) {
    AGX_CP_ASYNC8(st + 13 * 32 + lane, fs + gg);
    AGX_CP_ASYNC8(st + 14 * 32 + lane, fs + NJ + gg);
  }
  AGX_CP_ASYNC_COMMIT();
}
AGX_DEV void mma_unstage(MmaNodeIn& n, const double* st, double h, bool gaps, int g, int q, int lane) {
  AGX_CP_ASYNC_WAIT_ALL();
  const bool gl = g < NJ;
#pragma unroll
  for (int kc = 0; kc < 2; ++kc) {
    const int r = 4 * kc + q;
    const bool ok = gl && r < NJ;
    n.gB[0][kc] = ok ? st[(0 + kc) * 32 + lane] : 0.0;
    n.gB[1][kc] = ok ? st[(2 + kc) * 32 + lane] + ((r == g) ? 1.0 : 0.0) : 0.0;
    n.nB[kc] = ok ? st[(4 + kc) * 32 + lane] : 0.0;
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) n.lqq[e] = (gl && 2 * q + e < NJ) ? st[(6 + e) * 32 + lane] : 0.0;
  n.lvv = gl ? st[8 * 32 + lane] : 0.0;
  n.luu = gl ? st[9 * 32 + lane] : 0.0;
  n.lq = gl ? st[10 * 32 + lane] : 0.0;
  n.lv = gl ? st[11 * 32 + lane] : 0.0;
  n.lu = gl ? st[12 * 32 + lane] : 0.0;
  n.fs0 = (gl && gaps) ? st[13 * 32 + lane] : 0.0;
  n.fs1 = (gl && gaps) ? st[14 * 32 + lane] : 0.0;
  n.h = h;
}

#ifndef AGX_BWM_MINB
#define AGX_BWM_MINB 16
#endif
__global__ void __launch_bounds__(32, AGX_BWM_MINB) backward_mma_kernel(Problem P, Work W, SolverState S, FddpOpts O) {
  AGX_SMEM(smem);
  const int lane = (int)(threadIdx.x & 31u), wrp = (int)(threadIdx.x >> 5);
  const int g = lane >> 2, q = lane & 3;
  const int b = (int)blockIdx.x * (int)(blockDim.x >> 5) + wrp;
  if (b >= P.B) return;
  if (S.done[b] || S.pending[b]) return;  // pending: the candidate did not change, its sweep is still valid
  double* sm = smem + wrp * MB_SIZE;
  const int T = P.T, T1 = T + 1;
  const bool gl = g < NJ;
  const bool vlane = q == 3;  // lanes that hold column 15 (the vectors) as element e = 1 of tile column 1
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
  double cost;
  {
    double part = 0.0;
    for (int t = lane; t <= T; t += 32) part += crec0[(size_t)t * CREC_SIZE + CK_COST];
    cost = warp_sum(part);
  }
  if (!feasible && lane < NX) {
    const int c = lane, jj = c < NJ ? c : c - NJ;
    const int rk = c < NJ ? RK_QN : RK_VN;
    fsb[c] = W.x0[(size_t)b * NX + c] - xs[c];
    for (int t = 0; t < T; ++t) fsb[(t + 1) * NX + c] = rec0[(size_t)t * REC_SIZE + rk * 8 + jj] - xs[(t + 1) * NX + c];
  }
  __syncwarp();

  bool failed = !(cost == cost);  // a NaN node cost marks a failed calcDiff
  double dg = 0.0, dq = 0.0;
  for (;;) {
    bool ok = !failed;
    double Vt[2][2][2];   // V' tiles [tile row][tile col][e]
    double vx[2] = {0.0, 0.0};  // V'x entries g and 8 + g (meaningful on the q == 3 lanes)
    double dgp = 0.0, dqp = 0.0;
    MmaNodeIn cur;
    if (ok) {
      // the first running node's operands start flowing into shared memory while the terminal node is handled
      mma_stage(sm + MB_ST, rec0 + (size_t)(T - 1) * REC_SIZE, crec0 + (size_t)(T - 1) * CREC_SIZE,
                fsb + (size_t)(T - 1) * NX, !feasible, g, q, lane);
      // ---- terminal node: Vxx = Lxx (+ xreg), Vx = Lx (+ Vxx fs)
      const double* C = crec0 + (size_t)T * CREC_SIZE;
      const double lvvT = gl ? C[CK_LVV + g] : 0.0;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 2 * q + e;
        const bool in = gl && c < NJ;
        const double dgn = (in && c == g) ? 1.0 : 0.0;
        Vt[0][0][e] = in ? C[CK_LQQ + (g >= c ? lidx_(g, c) : lidx_(c, g))] + dgn * xreg : 0.0;
        Vt[1][1][e] = dgn * (lvvT + xreg);
        Vt[0][1][e] = 0.0;
        Vt[1][0][e] = 0.0;
      }
      vx[0] = gl ? C[CK_LQ + g] : 0.0;
      vx[1] = gl ? C[CK_LV + g] : 0.0;
      if (!feasible) {
        const double f0 = gl ? fsb[T * NX + g] : 0.0, f1 = gl ? fsb[T * NX + NJ + g] : 0.0;
        if (q == 0) { sm[MB_FS + g] = f0; sm[MB_FS + 8 + g] = f1; }
        __syncwarp();
        double gv[2];
#pragma unroll
        for (int tr = 0; tr < 2; ++tr) {
          double p = 0.0;
#pragma unroll
          for (int tc = 0; tc < 2; ++tc)
#pragma unroll
            for (int e = 0; e < 2; ++e) p += Vt[tr][tc][e] * sm[MB_FS + tc * 8 + 2 * q + e];
          p += __shfl_xor_sync(0xffffffffu, p, 1, 32);
          p += __shfl_xor_sync(0xffffffffu, p, 2, 32);
          gv[tr] = p;
        }
        vx[0] += gv[0]; vx[1] += gv[1];
        if (q == 0 && gl) { gvb[T * NX + g] = gv[0]; gvb[T * NX + NJ + g] = gv[1]; }
        if (vlane) {
          dgp -= vx[0] * f0 + vx[1] * f1;
          dqp += gv[0] * f0 + gv[1] * f1;
        }
        __syncwarp();
      }
    }
    // ---- running nodes
    for (int t = T - 1; ok && t >= 0; --t) {
      // this node's operands were staged during the previous node; the next node's follow behind the arithmetic
      mma_unstage(cur, sm + MB_ST, P.dts[t], !feasible, g, q, lane);
      if (t > 0)
        mma_stage(sm + MB_ST, rec0 + (size_t)(t - 1) * REC_SIZE, crec0 + (size_t)(t - 1) * CREC_SIZE,
                  fsb + (size_t)(t - 1) * NX, !feasible, g, q, lane);
      const double h = cur.h;
      // (1) Z = S^T V' (two tiles), Vs = Z S, sv = S^T v'
      double Zt[2][2], Vs[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        Zt[0][e] = h * Vt[0][0][e] + Vt[1][0][e];
        Zt[1][e] = h * Vt[0][1][e] + Vt[1][1][e];
        Vs[e] = h * Zt[0][e] + Zt[1][e];
      }
      const double sv = h * vx[0] + vx[1];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sm[MB_VS + g * MS8 + 2 * q + e] = Vs[e];
        sm[MB_Z0 + g * MS8 + 2 * q + e] = Zt[0][e];
      }
      if (!feasible && q == 0) { sm[MB_FS + g] = cur.fs0; sm[MB_FS + 8 + g] = cur.fs1; }
      __syncwarp();
      // (2) W = Vs G + [Zq 0] (+ sv in column 15), VN = Vs N
      double aV[2];
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) aV[kc] = sm[MB_VS + g * MS8 + 4 * kc + q];
      double Wt[2][2], VN[2] = {0.0, 0.0};
      Wt[0][0] = Zt[0][0]; Wt[0][1] = Zt[0][1];
      Wt[1][0] = 0.0; Wt[1][1] = vlane ? sv : 0.0;
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        AGX_DMMA(Wt[0][0], Wt[0][1], aV[kc], cur.gB[0][kc], Wt[0][0], Wt[0][1]);
        AGX_DMMA(Wt[1][0], Wt[1][1], aV[kc], cur.gB[1][kc], Wt[1][0], Wt[1][1]);
        AGX_DMMA(VN[0], VN[1], aV[kc], cur.nB[kc], VN[0], VN[1]);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sm[MB_W + g * MS16 + 2 * q + e] = Wt[0][e];
        sm[MB_W + g * MS16 + 8 + 2 * q + e] = Wt[1][e];
        sm[MB_VN + g * MS8 + 2 * q + e] = VN[e];
      }
      __syncwarp();
      // (3) Qxx = Lxx + [V'qq 0; 0 0] + G^T W + [Zq^T G; 0]   (column 15: Qx = Lx + [v'q; 0] + G^T S^T v')
      double bW[2][2], aZ[2], bVN[2];
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        bW[0][kc] = sm[MB_W + (4 * kc + q) * MS16 + g];
        bW[1][kc] = sm[MB_W + (4 * kc + q) * MS16 + 8 + g];
        aZ[kc] = sm[MB_Z0 + (4 * kc + q) * MS8 + g];
        bVN[kc] = sm[MB_VN + (4 * kc + q) * MS8 + g];
      }
      double Qt[2][2][2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = 2 * q + e;
        Qt[0][0][e] = cur.lqq[e] + Vt[0][0][e];
        Qt[0][1][e] = 0.0;
        Qt[1][0][e] = 0.0;
        Qt[1][1][e] = (gl && c == g) ? cur.lvv : 0.0;
      }
      if (vlane) { Qt[0][1][1] += cur.lq + vx[0]; Qt[1][1][1] += cur.lv; }
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
#pragma unroll
        for (int tr = 0; tr < 2; ++tr)
#pragma unroll
          for (int tc = 0; tc < 2; ++tc)
            AGX_DMMA(Qt[tr][tc][0], Qt[tr][tc][1], cur.gB[tr][kc], bW[tc][kc], Qt[tr][tc][0], Qt[tr][tc][1]);
#pragma unroll
        for (int tc = 0; tc < 2; ++tc)
          AGX_DMMA(Qt[0][tc][0], Qt[0][tc][1], aZ[kc], cur.gB[tc][kc], Qt[0][tc][0], Qt[0][tc][1]);
      }
      // (4) [Qux | Qu] = N^T W (+ Lu in column 15), Quu = Luu + N^T VN (+ ureg)
      double Ut[2][2] = {{0.0, 0.0}, {0.0, 0.0}}, Quu[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) Quu[e] = (gl && 2 * q + e == g) ? cur.luu + xreg : 0.0;
      if (vlane) Ut[1][1] = cur.lu;
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        AGX_DMMA(Ut[0][0], Ut[0][1], cur.nB[kc], bW[0][kc], Ut[0][0], Ut[0][1]);
        AGX_DMMA(Ut[1][0], Ut[1][1], cur.nB[kc], bW[1][kc], Ut[1][0], Ut[1][1]);
        AGX_DMMA(Quu[0], Quu[1], cur.nB[kc], bVN[kc], Quu[0], Quu[1]);
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sm[MB_QUX + g * MS16 + 2 * q + e] = Ut[0][e];
        sm[MB_QUX + g * MS16 + 8 + 2 * q + e] = Ut[1][e];
        sm[MB_VS + g * MS8 + 2 * q + e] = Quu[e];
      }
      __syncwarp();
      // (5) computeGains: every lane factors Quu in registers; lanes 0..15 solve one column of [Qux | Qu] each
      {
        double L[28], rinv[NJ];
        ok = chol7_registers_strided(sm + MB_VS, MS8, L, rinv);
        if (!ok) break;
        const int c = lane & 15;
        double col[NJ];
#pragma unroll
        for (int i = 0; i < NJ; ++i) col[i] = sm[MB_QUX + i * MS16 + c];
        chol_solve7(L, rinv, col);
        if (lane < 16) {
#pragma unroll
          for (int i = 0; i < NJ; ++i) sm[MB_W + i * MS16 + c] = col[i];
          sm[MB_W + 7 * MS16 + c] = 0.0;
          // gains out: padded column c -> state index (c < 7: q block, 8..14: v block), column 15 = k
          if (c != 7) {
            if (c < 15) {
              const int s14 = c < NJ ? c : c - 1;
#pragma unroll
              for (int i = 0; i < NJ; ++i) Kb[(t * NJ + i) * NX + s14] = col[i];
            } else {
              double qk = 0.0;
#pragma unroll
              for (int i = 0; i < NJ; ++i) { kb[t * NJ + i] = col[i]; qk += sm[MB_QUX + i * MS16 + 15] * col[i]; }
              // expected improvement: Qu.k and k.Quu.k (= Qu.k, Quu k = Qu)
              dgp += qk;
              dqp -= qk;
            }
          }
        }
      }
      __syncwarp();
      // (6) Vxx = Qxx - Qux^T K  (column 15: Vx = Qx - Qxu k)
      {
        double aU[2][2], bK[2][2];
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          aU[0][kc] = -sm[MB_QUX + (4 * kc + q) * MS16 + g];
          aU[1][kc] = -sm[MB_QUX + (4 * kc + q) * MS16 + 8 + g];
          bK[0][kc] = sm[MB_W + (4 * kc + q) * MS16 + g];
          bK[1][kc] = sm[MB_W + (4 * kc + q) * MS16 + 8 + g];
        }
#pragma unroll
        for (int kc = 0; kc < 2; ++kc)
#pragma unroll
          for (int tr = 0; tr < 2; ++tr)
#pragma unroll
            for (int tc = 0; tc < 2; ++tc)
              AGX_DMMA(Qt[tr][tc][0], Qt[tr][tc][1], aU[tr][kc], bK[tc][kc], Qt[tr][tc][0], Qt[tr][tc][1]);
      }
      // (7) Vx out of column 15, symmetrise, regularise, clear the pads
      vx[0] = Qt[0][1][1];
      vx[1] = Qt[1][1][1];
#pragma unroll
      for (int tr = 0; tr < 2; ++tr)
#pragma unroll
        for (int tc = 0; tc < 2; ++tc)
#pragma unroll
          for (int e = 0; e < 2; ++e) sm[MB_V + (tr * 8 + g) * MSV + tc * 8 + 2 * q + e] = Qt[tr][tc][e];
      __syncwarp();
#pragma unroll
      for (int tr = 0; tr < 2; ++tr)
#pragma unroll
        for (int tc = 0; tc < 2; ++tc)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 2 * q + e;
            const double tv = sm[MB_V + (tc * 8 + c) * MSV + tr * 8 + g];
            double v = 0.5 * (Qt[tr][tc][e] + tv);
            if (tr == tc && c == g) v += xreg;
            Vt[tr][tc][e] = (gl && c < NJ) ? v : 0.0;
          }
      // (8) gap terms: Vx += Vxx fs, expected-improvement pieces
      if (!feasible) {
        double gv[2];
#pragma unroll
        for (int tr = 0; tr < 2; ++tr) {
          double p = 0.0;
#pragma unroll
          for (int tc = 0; tc < 2; ++tc)
#pragma unroll
            for (int e = 0; e < 2; ++e) p += Vt[tr][tc][e] * sm[MB_FS + tc * 8 + 2 * q + e];
          p += __shfl_xor_sync(0xffffffffu, p, 1, 32);
          p += __shfl_xor_sync(0xffffffffu, p, 2, 32);
          gv[tr] = p;
        }
        vx[0] += gv[0]; vx[1] += gv[1];
        if (q == 0 && gl) { gvb[t * NX + g] = gv[0]; gvb[t * NX + NJ + g] = gv[1]; }
        if (vlane) {
          dgp -= vx[0] * cur.fs0 + vx[1] * cur.fs1;
          dqp += gv[0] * cur.fs0 + gv[1] * cur.fs1;
        }
      }
      if (!gl) { vx[0] = 0.0; vx[1] = 0.0; }
      __syncwarp();
    }
    AGX_CP_ASYNC_WAIT_ALL();  // nothing may still be in flight when the sweep is abandoned or restarted
    if (ok) {
      // non-finite value function = failed sweep (SolverDDP::backwardPass raises on NaN)
      double chk = vlane ? vx[0] + vx[1] : 0.0;
#pragma unroll
      for (int tr = 0; tr < 2; ++tr)
#pragma unroll
        for (int tc = 0; tc < 2; ++tc) chk += Vt[tr][tc][0] + Vt[tr][tc][1];
      chk = warp_sum(chk);
      if (!(chk - chk == 0.0)) ok = false;
    }
    __syncwarp();
    if (ok) {
      dg = warp_sum(dgp);
      dq = warp_sum(dqp);
      break;
    }
    // increaseRegularization and retry without recalc
    failed = false;
    xreg *= O.reg_incfactor;
    if (xreg > O.reg_max) xreg = O.reg_max;
    if (xreg == O.reg_max) {
      if (lane == 0) { S.status[b] = 2; S.done[b] = 1; }
      break;
    }
  }
  if (lane == 0) {
    S.xreg[b] = xreg;
    S.cost[b] = cost;
    S.dg[b] = dg;
    S.dq[b] = dq;
  }
}

}  // namespace agx
#endif  // AGX_RICCATI_MMA_CUH_
