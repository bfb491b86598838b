// Kernel 1 -- per-factor residual + Jacobian arithmetic (FP64, one factor per thread).
//
// Replaces, for the graph /root/reference/batch.py:270-305 builds, what gtsam's
// NonlinearFactorGraph::linearize / ::error evaluate inside batch.py:337:
//   PriorFactor<Pose3>/<Vector> (batch.py:281-282), BetweenFactor<Pose3> (north_star),
//   the DVL CustomFactor (batch.py:196-233, residual only -- Jacobians are the analytic ones),
//   GenericStereoFactor3D (batch.py:300-305) and ImuFactor (batch.py:238), with the noise-model
//   whitening (batch.py:95-98, :118) fused in.
//
// Storage is structure-of-arrays, component-major: value[c * n + i].  A warp therefore reads 32
// consecutive factors' component c in one coalesced 256 B request, and consecutive chain factors
// gather consecutive poses.
// Whitened Jacobians are written in NODE order (pose columns before velocity columns):
//   prior_pose 6x6 | prior_vel 3x3 | between 6x12 [H1|H2] | dvl 3x9 [Hx|Hv] | stereo 3x9 [Hpose|Hlm]
//   imu 9x24 [Hxi Hvi | Hxj Hvj | Hbias(acc,gyro)]
// Element (row, col) of factor f lives at J[(row * ncols + col) * n + f].
#pragma once
#include "lie.cuh"

namespace vus {

struct ValuesView {
  const double* pose; long nx;   // [12][nx]  R row-major (9) then t (3)
  const double* vel;  long nv;   // [3][nv]
  const double* bias; long nb;   // [6][nb]   acc, gyro
  const double* lm;   long nl;   // [3][nl]
};

struct FactorView {
  long n;
  const int* idx;        // [slots][n]
  const double* meas;    // [meas_dim][n]
  const double* sinfo;   // [info_dim][n]
};

struct LinOut {
  double* r;   // [m][n] whitened residual (may be null)
  double* J;   // [m*ncols][n] whitened Jacobian (may be null)
  double* e2;  // [n] 0.5*||r||^2 (may be null)
  // stereo only (fused assembly products, may be null):
  double* sE;  // [n][18]  E_o = Jp^T Jl (6x3 row-major), one 144-byte record per observation
  double* sPp; // [n][28]  per-observation pose products: 21 unique Jp^T Jp (a<=b) | 6 Jp^T r | pad
  double* sPl; // [n][12]  per-observation landmark products: 6 unique Jl^T Jl | 3 Jl^T r | pad
};

VUS_HD void load_pose(const double* P, long n, long i, double* R, double* t) {
#pragma unroll
  for (int c = 0; c < 9; ++c) R[c] = P[c * n + i];
#pragma unroll
  for (int c = 0; c < 3; ++c) t[c] = P[(9 + c) * n + i];
}
VUS_HD void load3(const double* P, long n, long i, double* v) {
  v[0] = P[i]; v[1] = P[n + i]; v[2] = P[2 * n + i];
}

static constexpr int kFactorM[VUS_F_NTYPES] = {6, 3, 6, 3, 3, 9};          // residual dim
static constexpr int kFactorCols[VUS_F_NTYPES] = {6, 3, 12, 9, 9, 24};     // Jacobian columns
static constexpr int kFactorSlots[VUS_F_NTYPES] = {1, 1, 2, 2, 2, 5};
static constexpr int kFactorMeas[VUS_F_NTYPES] = {12, 3, 12, 3, 3, 67};
static constexpr int kFactorInfo[VUS_F_NTYPES] = {6, 3, 6, 3, 3, 45};

// ------------------------------------------------------------------ PriorFactor<Pose3>
template <bool WJ>
VUS_HD void f_prior_pose(const ValuesView& V, const FactorView& F, const LinOut& O, long f) {
  const long n = F.n;
  double R[9], t[3], Rm[9], tm[3];
  load_pose(V.pose, V.nx, F.idx[f], R, t);
  load_pose(F.meas, n, f, Rm, tm);
  double Rb[9], d[3], tb[3], xi[6];
  m3_Tmul(R, Rm, Rb);
  d[0] = tm[0] - t[0]; d[1] = tm[1] - t[1]; d[2] = tm[2] - t[2];
  m3_Tvec(R, d, tb);
  pose_log(Rb, tb, xi);
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double s = F.sinfo[k * n + f];
    const double rk = -s * xi[k];
    acc += rk * rk;
    if (O.r) O.r[k * n + f] = rk;
    if (WJ) {
#pragma unroll
      for (int c = 0; c < 6; ++c) O.J[(k * 6 + c) * n + f] = (c == k) ? s : 0.0;   // gtsam: H = I
    }
  }
  if (O.e2) O.e2[f] = 0.5 * acc;
}

// ------------------------------------------------------------------ PriorFactor<Vector3>
template <bool WJ>
VUS_HD void f_prior_vel(const ValuesView& V, const FactorView& F, const LinOut& O, long f) {
  const long n = F.n;
  double v[3];
  load3(V.vel, V.nv, F.idx[f], v);
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double s = F.sinfo[k * n + f];
    const double rk = s * (v[k] - F.meas[k * n + f]);
    acc += rk * rk;
    if (O.r) O.r[k * n + f] = rk;
    if (WJ) {
#pragma unroll
      for (int c = 0; c < 3; ++c) O.J[(k * 3 + c) * n + f] = (c == k) ? s : 0.0;
    }
  }
  if (O.e2) O.e2[f] = 0.5 * acc;
}

// ------------------------------------------------------------------ BetweenFactor<Pose3>
// slow_but_correct = gtsam's GTSAM_SLOW_BUT_CORRECT_BETWEENFACTOR build: H1 = -dLog(e) Ad(hx^-1), H2 = dLog(e).  The default
// gtsam build (4.1 / 4.2 wheels) leaves the derivative of Local() out: H1 = -Ad(hx^-1), H2 = I (BetweenFactor.h).
template <bool WJ>
VUS_HD void f_between(const ValuesView& V, const FactorView& F, const LinOut& O, long f, bool slow_but_correct) {
  const long n = F.n;
  double R1[9], t1[3], R2[9], t2[3], Rm[9], tm[3];
  load_pose(V.pose, V.nx, F.idx[f], R1, t1);
  load_pose(V.pose, V.nx, F.idx[n + f], R2, t2);
  load_pose(F.meas, n, f, Rm, tm);
  double Rh[9], th[3], Re[9], te[3], d[3], xi[6];
  m3_Tmul(R1, R2, Rh);
  d[0] = t2[0] - t1[0]; d[1] = t2[1] - t1[1]; d[2] = t2[2] - t1[2];
  m3_Tvec(R1, d, th);
  m3_Tmul(Rm, Rh, Re);
  d[0] = th[0] - tm[0]; d[1] = th[1] - tm[1]; d[2] = th[2] - tm[2];
  m3_Tvec(Rm, d, te);
  pose_log(Re, te, xi);
  double s[6], acc = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    s[k] = F.sinfo[k * n + f];
    const double rk = s[k] * xi[k];
    acc += rk * rk;
    if (O.r) O.r[k * n + f] = rk;
  }
  if (O.e2) O.e2[f] = 0.5 * acc;
  if (WJ) {
    double Jw[9], Q2[9];
    if (slow_but_correct) {
      pose_dlog(xi, Jw, Q2);
    } else {
#pragma unroll
      for (int c = 0; c < 9; ++c) { Jw[c] = (c % 4 == 0) ? 1.0 : 0.0; Q2[c] = 0.0; }
    }
    // hx^-1 = (Rh^T, -Rh^T th);  Ad = [[Ri,0],[[ti]x Ri, Ri]]
    double ti[3], JwRi[9], Q2Ri[9], Tx[9], TxRi[9], JwTxRi[9];
    m3_Tvec(Rh, th, ti);
    ti[0] = -ti[0]; ti[1] = -ti[1]; ti[2] = -ti[2];
    m3_mulT(Jw, Rh, JwRi);          // Jw Rh^T
    m3_mulT(Q2, Rh, Q2Ri);
    skew(ti, Tx);
    m3_mulT(Tx, Rh, TxRi);
    m3_mul(Jw, TxRi, JwTxRi);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int rc = 3 * r + c;
        // H1 = -D Ad
        O.J[((r) * 12 + c) * n + f] = -s[r] * JwRi[rc];
        O.J[((r) * 12 + 3 + c) * n + f] = 0.0;
        O.J[((3 + r) * 12 + c) * n + f] = -s[3 + r] * (Q2Ri[rc] + JwTxRi[rc]);
        O.J[((3 + r) * 12 + 3 + c) * n + f] = -s[3 + r] * JwRi[rc];
        // H2 = D
        O.J[((r) * 12 + 6 + c) * n + f] = s[r] * Jw[rc];
        O.J[((r) * 12 + 9 + c) * n + f] = 0.0;
        O.J[((3 + r) * 12 + 6 + c) * n + f] = s[3 + r] * Q2[rc];
        O.J[((3 + r) * 12 + 9 + c) * n + f] = s[3 + r] * Jw[rc];
      }
  }
}

// ------------------------------------------------------------------ DVL velocity (batch.py:196-233)
template <bool WJ>
VUS_HD void f_dvl(const ValuesView& V, const FactorView& F, const LinOut& O, long f) {
  const long n = F.n;
  double v[3], R[9], m[3], Rm[3];
  load3(V.vel, V.nv, F.idx[f], v);                      // slot 0 = V(i) (batch.py:247)
  const long xi = F.idx[n + f];                         // slot 1 = X(i)
#pragma unroll
  for (int c = 0; c < 9; ++c) R[c] = V.pose[c * V.nx + xi];
  load3(F.meas, n, f, m);
  m3_vec(R, m, Rm);
  double s[3], acc = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    s[k] = F.sinfo[k * n + f];
    const double rk = s[k] * (Rm[k] - v[k]);
    acc += rk * rk;
    if (O.r) O.r[k * n + f] = rk;
  }
  if (O.e2) O.e2[f] = 0.5 * acc;
  if (WJ) {
    double RM[9];
    m3_mul_skew(R, m, RM);                              // R [m]x
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        O.J[(r * 9 + c) * n + f] = -s[r] * RM[3 * r + c];
        O.J[(r * 9 + 3 + c) * n + f] = 0.0;
        O.J[(r * 9 + 6 + c) * n + f] = (r == c) ? -s[r] : 0.0;
      }
  }
}

// ------------------------------------------------------------------ GenericStereoFactor3D
template <bool WJ>
VUS_HD void f_stereo(const ValuesView& V, const FactorView& F, const LinOut& O, long f, const double* K) {
  const long n = F.n;
  double R[9], t[3], l[3], z[3], d3[3], q[3];
  load_pose(V.pose, V.nx, F.idx[f], R, t);
  load3(V.lm, V.nl, F.idx[n + f], l);
  load3(F.meas, n, f, z);
  const double fx = K[0], fy = K[1], u0 = K[3], v0 = K[4], b = K[5];
  d3[0] = l[0] - t[0]; d3[1] = l[1] - t[1]; d3[2] = l[2] - t[2];
  m3_Tvec(R, d3, q);
  double s[3];
  s[0] = F.sinfo[f]; s[1] = F.sinfo[n + f]; s[2] = F.sinfo[2 * n + f];
  if (q[2] <= 0.0) {                                    // cheirality: e = 2 fx [1,1,1], zero Jacobians
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double rk = s[k] * 2.0 * fx;
      acc += rk * rk;
      if (O.r) O.r[k * n + f] = rk;
    }
    if (O.e2) O.e2[f] = 0.5 * acc;
    if (WJ) {
#pragma unroll
      for (int c = 0; c < 27; ++c) O.J[c * n + f] = 0.0;
      if (O.sE) {
#pragma unroll
        for (int c = 0; c < 18; ++c) O.sE[f * 18 + c] = 0.0;
#pragma unroll
        for (int c = 0; c < 28; ++c) O.sPp[f * 28 + c] = 0.0;
#pragma unroll
        for (int c = 0; c < 12; ++c) O.sPl[f * 12 + c] = 0.0;
      }
    }
    return;
  }
  const double d = 1.0 / q[2];
  const double uL = fx * q[0] * d, uR = fx * (q[0] - b) * d, vv = fy * q[1] * d;
  const double e[3] = {u0 + uL - z[0], u0 + uR - z[1], v0 + vv - z[2]};
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const double rk = s[k] * e[k];
    acc += rk * rk;
    if (O.r) O.r[k * n + f] = rk;
  }
  if (O.e2) O.e2[f] = 0.5 * acc;
  if (WJ) {
    const double v1 = vv / fy, v2 = fx * v1, dx = d * q[0];
    const double Hp[18] = {uL * v1, -fx - dx * uL, v2, -d * fx, 0.0, d * uL,
                           uR * v1, -fx - dx * uR, v2, -d * fx, 0.0, d * uR,
                           fy + vv * v1, -dx * vv, -q[0] * d * fy, 0.0, -d * fy, d * vv};
    double Jp[18], Jl[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 6; ++c) { Jp[6 * r + c] = s[r] * Hp[6 * r + c]; O.J[(r * 9 + c) * n + f] = Jp[6 * r + c]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) {                       // column k of H_lm
      Jl[k] = s[0] * d * (fx * R[3 * k] - R[3 * k + 2] * uL);
      Jl[3 + k] = s[1] * d * (fx * R[3 * k] - R[3 * k + 2] * uR);
      Jl[6 + k] = s[2] * d * (fy * R[3 * k + 1] - R[3 * k + 2] * vv);
      O.J[(0 * 9 + 6 + k) * n + f] = Jl[k];
      O.J[(1 * 9 + 6 + k) * n + f] = Jl[3 + k];
      O.J[(2 * 9 + 6 + k) * n + f] = Jl[6 + k];
    }
    if (O.sE) {                                         // fused assembly products (kernel 2 inputs never re-read J)
      const double rs[3] = {s[0] * e[0], s[1] * e[1], s[2] * e[2]};
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c)
          O.sE[f * 18 + a * 3 + c] = Jp[a] * Jl[c] + Jp[6 + a] * Jl[3 + c] + Jp[12 + a] * Jl[6 + c];
      double* Pp = O.sPp + f * 28;
      int q = 0;
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) Pp[q++] = Jp[a] * Jp[b] + Jp[6 + a] * Jp[6 + b] + Jp[12 + a] * Jp[12 + b];
#pragma unroll
      for (int a = 0; a < 6; ++a) Pp[21 + a] = Jp[a] * rs[0] + Jp[6 + a] * rs[1] + Jp[12 + a] * rs[2];
      Pp[27] = 0.0;
      double* Pl = O.sPl + f * 12;
      q = 0;
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b) Pl[q++] = Jl[a] * Jl[b] + Jl[3 + a] * Jl[3 + b] + Jl[6 + a] * Jl[6 + b];
#pragma unroll
      for (int a = 0; a < 3; ++a) Pl[6 + a] = Jl[a] * rs[0] + Jl[3 + a] * rs[1] + Jl[6 + a] * rs[2];
      Pl[9] = Pl[10] = Pl[11] = 0.0;
    }
  }
}

// ------------------------------------------------------------------ ImuFactor (both gtsam preintegration builds)
// packed PIM row: dR 9 (tangent: theta 3, then 6 unused) | dP 3 | dV 3 | dt 1 | bhat 6 | JRg 9 | JPa 9 | JPg 9 | JVa 9 | JVg 9
VUS_HD void imu_whiten9(const double* W, const double* e, double* y) {   // y = W e, W upper-tri packed row-major
  int p = 0;
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    double a = 0.0;
#pragma unroll
    for (int c = r; c < 9; ++c) a += W[p++] * e[c];
    y[r] = a;
  }
}
// store the whitened 9x3 block `blk` (rows: rot 0-2 / pos 3-5 / vel 6-8; row-major 9x3) at columns col0..col0+2
VUS_HD void imu_emit(double* J, long n, long f, int col0, const double* W, const double* blk) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double col[9], y[9];
#pragma unroll
    for (int r = 0; r < 9; ++r) col[r] = blk[3 * r + c];
    imu_whiten9(W, col, y);
#pragma unroll
    for (int r = 0; r < 9; ++r) J[(r * 24 + col0 + c) * n + f] = y[r];
  }
}

// tangent = gtsam's GTSAM_TANGENT_PREINTEGRATION build (TangentPreintegration::biasCorrectedDelta): the packed row holds
// theta (columns 0..2) and d theta / d bg (22..30); theta_c = theta + (d theta / d bg) dbg and dR_c = Exp(theta_c).
template <bool WJ>
VUS_HD void f_imu(const ValuesView& V, const FactorView& F, const LinOut& O, long f, const double* g, bool tangent) {
  const long n = F.n;
  double Ri[9], ti[3], vi[3], Rj[9], tj[3], vj[3], bias[6];
  load_pose(V.pose, V.nx, F.idx[f], Ri, ti);
  load3(V.vel, V.nv, F.idx[n + f], vi);
  load_pose(V.pose, V.nx, F.idx[2 * n + f], Rj, tj);
  load3(V.vel, V.nv, F.idx[3 * n + f], vj);
  const long bi = F.idx[4 * n + f];
#pragma unroll
  for (int c = 0; c < 6; ++c) bias[c] = V.bias[c * V.nb + bi];
  const double* M = F.meas;
#define PIM(c) M[(c) * n + f]
  double dR[9], dP[3], dV[3], dba[3], dbg[3];
#pragma unroll
  for (int c = 0; c < 9; ++c) dR[c] = PIM(c);
#pragma unroll
  for (int c = 0; c < 3; ++c) { dP[c] = PIM(9 + c); dV[c] = PIM(12 + c); }
  const double dt = PIM(15);
#pragma unroll
  for (int c = 0; c < 3; ++c) { dba[c] = bias[c] - PIM(16 + c); dbg[c] = bias[3 + c] - PIM(19 + c); }
  double JRg[9];
#pragma unroll
  for (int c = 0; c < 9; ++c) JRg[c] = PIM(22 + c);
  // bias-corrected deltas
  double corr[3], Ec[9], dRc[9], pc[3], vc[3];
  m3_vec(JRg, dbg, corr);
  if (tangent) {
#pragma unroll
    for (int c = 0; c < 3; ++c) corr[c] += dR[c];      // theta_c
    so3_exp(corr, dRc);
  } else {
    so3_exp(corr, Ec);
    m3_mul(dR, Ec, dRc);
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double ap = dP[r], av = dV[r];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ap += PIM(31 + 3 * r + c) * dba[c] + PIM(40 + 3 * r + c) * dbg[c];
      av += PIM(49 + 3 * r + c) * dba[c] + PIM(58 + 3 * r + c) * dbg[c];
    }
    pc[r] = ap; vc[r] = av;
  }
  double A[9], E[9], e[9], tmp[3], tmp2[3];
  m3_Tmul(Rj, Ri, A);                // A = Rj^T Ri
  m3_mul(A, dRc, E);
  so3_log(E, e);
  m3_vec(Ri, pc, tmp);
#pragma unroll
  for (int k = 0; k < 3; ++k) tmp2[k] = ti[k] + vi[k] * dt + 0.5 * g[k] * dt * dt + tmp[k] - tj[k];
  m3_Tvec(Rj, tmp2, e + 3);
  m3_vec(Ri, vc, tmp);
#pragma unroll
  for (int k = 0; k < 3; ++k) tmp2[k] = vi[k] + g[k] * dt + tmp[k] - vj[k];
  m3_Tvec(Rj, tmp2, e + 6);
  double W[45];
#pragma unroll
  for (int c = 0; c < 45; ++c) W[c] = F.sinfo[c * n + f];
  double y[9], acc = 0.0;
  imu_whiten9(W, e, y);
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    acc += y[k] * y[k];
    if (O.r) O.r[k * n + f] = y[k];
  }
  if (O.e2) O.e2[f] = 0.5 * acc;
  if (WJ) {
    double dlog[9], blk[27], T[9];
    so3_dlog(e, dlog);
    // --- X_i rotation columns: [dlog dRc^T ; -A [pc]x ; -A [vc]x]
    m3_mulT(dlog, dRc, T);
#pragma unroll
    for (int k = 0; k < 9; ++k) blk[k] = T[k];
    m3_mul_skew(A, pc, T);
#pragma unroll
    for (int k = 0; k < 9; ++k) blk[9 + k] = -T[k];
    m3_mul_skew(A, vc, T);
#pragma unroll
    for (int k = 0; k < 9; ++k) blk[18 + k] = -T[k];
    imu_emit(O.J, n, f, 0, W, blk);
    // --- X_i translation columns: [0 ; A ; 0]
#pragma unroll
    for (int k = 0; k < 9; ++k) { blk[k] = 0.0; blk[9 + k] = A[k]; blk[18 + k] = 0.0; }
    imu_emit(O.J, n, f, 3, W, blk);
    // --- V_i: [0 ; Rj^T dt ; Rj^T]
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) { blk[3 * r + c] = 0.0; blk[9 + 3 * r + c] = Rj[3 * c + r] * dt; blk[18 + 3 * r + c] = Rj[3 * c + r]; }
    imu_emit(O.J, n, f, 6, W, blk);
    // --- X_j rotation: [-dlog E^T ; [rp]x ; [rv]x]
    m3_mulT(dlog, E, T);
#pragma unroll
    for (int k = 0; k < 9; ++k) blk[k] = -T[k];
    skew(e + 3, blk + 9);
    skew(e + 6, blk + 18);
    imu_emit(O.J, n, f, 9, W, blk);
    // --- X_j translation: [0 ; -I ; 0]
#pragma unroll
    for (int k = 0; k < 27; ++k) blk[k] = 0.0;
    blk[9] = blk[13] = blk[17] = -1.0;
    imu_emit(O.J, n, f, 12, W, blk);
    // --- V_j: [0 ; 0 ; -Rj^T]
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) { blk[3 * r + c] = 0.0; blk[9 + 3 * r + c] = 0.0; blk[18 + 3 * r + c] = -Rj[3 * c + r]; }
    imu_emit(O.J, n, f, 15, W, blk);
    // --- bias acc: [0 ; A JPa ; A JVa]
    double Jm[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { blk[k] = 0.0; Jm[k] = PIM(31 + k); }
    m3_mul(A, Jm, blk + 9);
#pragma unroll
    for (int k = 0; k < 9; ++k) Jm[k] = PIM(49 + k);
    m3_mul(A, Jm, blk + 18);
    imu_emit(O.J, n, f, 18, W, blk);
    // --- bias gyro: [dlog Jr(corr) JRg ; A JPg ; A JVg]   (tangent: corr = theta_c, JRg = d theta / d bg)
    double Jr[9];
    so3_dexp(corr, Jr);
    m3_mul(dlog, Jr, T);
    m3_mul(T, JRg, blk);
#pragma unroll
    for (int k = 0; k < 9; ++k) Jm[k] = PIM(40 + k);
    m3_mul(A, Jm, blk + 9);
#pragma unroll
    for (int k = 0; k < 9; ++k) Jm[k] = PIM(58 + k);
    m3_mul(A, Jm, blk + 18);
    imu_emit(O.J, n, f, 21, W, blk);
  }
#undef PIM
}

}  // namespace vus
