// "Next" rows of the hot-path scope: the two per-keyframe loops that feed the graph batch.py
// builds, moved to the device so raw 200 Hz IMU streams and stereo feature tracks can be ingested without a host pass.
//
//   PreintBody   gtsam.PreintegratedImuMeasurements.integrateMeasurement x k + resetIntegration (batch.py:289-293):
//                one thread per ImuFactor runs the k samples of its keyframe interval -- tangent (gtsam
//                TangentPreintegration.cpp, the 4.0-4.2 wheel default) or manifold (Forster et al.) preintegration
//                (vus_set_gtsam_build), bias Jacobians, first-order 9x9 covariance -- and emits the packed
//                PIM row (67) and the upper sqrt-information (45) the ImuFactor table of include/vus.h takes.
//   BackprojBody the disparity back-projection of get_landmarks (batch.py:144-176) in gtsam's StereoCamera convention
//                (uR = fx (x - b) / z): landmark = T_i * backproject(uL, uR, v).
#pragma once
#include "factors.cuh"

namespace vus {

struct PreintArgs {
  long n; int k;
  const double* acc; const double* gyro;     // [n][k][3] row-major
  double dt;
  double bhat[6];                            // accelerometer, gyro
  double aC[9], wC[9], iC[9];                // accelerometer / gyroscope / integration covariance (continuous-time, batch.py:183-185)
  double* pim;                               // [67][n] component-major
  double* sinfo;                             // [45][n]
  int* fail;
  bool tangent;                              // TangentPreintegration (state = 9-vector [theta, p, v]) instead of ManifoldPreintegration
};

VUS_HD void m9_mul(const double* A, const double* B, double* C, bool tb) {   // C = A B  or  A B^T   (9x9)
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      double s = 0.0;
      for (int l = 0; l < 9; ++l) s += A[9 * i + l] * (tb ? B[9 * j + l] : B[9 * l + j]);
      C[9 * i + j] = s;
    }
}

// d/dw [ Jr(w) c ] at fixed c  (gtsam so3::DexpFunctor::applyDexp, H1):
//   Jr(w) c = c - A (w x c) + B (w x (w x c)),  A = (1 - cos th) / th^2,  B = (th - sin th) / th^3
VUS_HD void so3_apply_dexp_deriv(const double* w, const double* c, double* D) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double Cx[9];
  skew(c, Cx);
  if (th2 <= VUS_EPS) {
    for (int e = 0; e < 9; ++e) D[e] = 0.5 * Cx[e];
    return;
  }
  const double th = sqrt(th2), sn = sin(th), s2 = sin(0.5 * th), omc = 2.0 * s2 * s2;
  const double A = omc / th2, B = (th - sn) / (th2 * th);
  const double dA = (th * sn - 2.0 * omc) / (th2 * th) / th;            // (dA/dth) / th
  const double dB = (omc * th - 3.0 * (th - sn)) / (th2 * th2) / th;
  const double wxc[3] = {w[1] * c[2] - w[2] * c[1], w[2] * c[0] - w[0] * c[2], w[0] * c[1] - w[1] * c[0]};
  const double wwc[3] = {w[1] * wxc[2] - w[2] * wxc[1], w[2] * wxc[0] - w[0] * wxc[2], w[0] * wxc[1] - w[1] * wxc[0]};
  const double wc = w[0] * c[0] + w[1] * c[1] + w[2] * c[2];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      D[3 * i + j] = -dA * wxc[i] * w[j] + A * Cx[3 * i + j] + dB * wwc[i] * w[j]
                     + B * ((i == j ? wc : 0.0) + w[i] * c[j] - 2.0 * c[i] * w[j]);
}

// One sample of TangentPreintegration::update (+ the covariance line of PreintegratedImuMeasurements::integrateMeasurement):
//   theta += Jr(theta)^-1 w h ;  p += v h + Exp(theta) a h^2 / 2 ;  v += Exp(theta) a h
//   H_biasAcc <- A H_biasAcc - B ;  H_biasOmega <- A H_biasOmega - C ;  cov <- A cov A^T + B (aC/h) B^T + C (wC/h) C^T, cov[3:6,3:6] += iC h
// x[9] state, Ha / Hg [9][3] row-major, cov [81]; A9 / tmp: 81-double scratch.
VUS_HD void preint_tangent_step(double* x, double* Ha, double* Hg, double* cov, const double* a, const double* w, double h,
                                const double* aC, const double* wC, const double* iC, double* A9, double* tmp) {
  const double q = 0.5 * h * h;
  double R[9], Jr[9], inv[9], c[3], D[9], wH[9], ax[9], Rax[9], aH[9];
  so3_exp(x, R);
  so3_dexp(x, Jr);
  so3_dlog(x, inv);                                  // Jr^-1
  m3_vec(inv, w, c);
  so3_apply_dexp_deriv(x, c, D);
  m3_mul(inv, D, wH);                                // -(d c / d theta)
  skew(a, ax);
  m3_mul(R, ax, Rax);
  m3_mul(Rax, Jr, aH);                               // -(d (R a) / d theta)
  for (int e = 0; e < 81; ++e) A9[e] = 0.0;
  for (int i = 0; i < 9; ++i) A9[10 * i] = 1.0;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      A9[9 * i + j] -= wH[3 * i + j] * h;
      A9[9 * (3 + i) + j] = -aH[3 * i + j] * q;
      A9[9 * (6 + i) + j] = -aH[3 * i + j] * h;
    }
  for (int i = 0; i < 3; ++i) A9[9 * (3 + i) + 6 + i] = h;
  // bias Jacobians: H <- A H - [B | C]   (B = [0; R q; R h], C = [Jr^-1 h; 0; 0])
  double nHa[27], nHg[27];
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 3; ++j) {
      double sa = 0.0, sg = 0.0;
      for (int l = 0; l < 9; ++l) { sa += A9[9 * i + l] * Ha[3 * l + j]; sg += A9[9 * i + l] * Hg[3 * l + j]; }
      if (i < 3) sg -= inv[3 * i + j] * h;
      else if (i < 6) sa -= R[3 * (i - 3) + j] * q;
      else sa -= R[3 * (i - 6) + j] * h;
      nHa[3 * i + j] = sa; nHg[3 * i + j] = sg;
    }
  for (int e = 0; e < 27; ++e) { Ha[e] = nHa[e]; Hg[e] = nHg[e]; }
  // covariance
  m9_mul(A9, cov, tmp, false);
  m9_mul(tmp, A9, cov, true);
  {
    double Bm[18], t1[18];                           // rows 3..8 of B
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) { Bm[3 * i + j] = R[3 * i + j] * q; Bm[3 * (3 + i) + j] = R[3 * i + j] * h; }
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 3; ++j) t1[3 * i + j] = (Bm[3 * i] * aC[j] + Bm[3 * i + 1] * aC[3 + j] + Bm[3 * i + 2] * aC[6 + j]) / h;
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) cov[9 * (3 + i) + 3 + j] += t1[3 * i] * Bm[3 * j] + t1[3 * i + 1] * Bm[3 * j + 1] + t1[3 * i + 2] * Bm[3 * j + 2];
    double Cm[9], t2[9];
    for (int e = 0; e < 9; ++e) Cm[e] = inv[e] * h;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) t2[3 * i + j] = (Cm[3 * i] * wC[j] + Cm[3 * i + 1] * wC[3 + j] + Cm[3 * i + 2] * wC[6 + j]) / h;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        cov[9 * i + j] += t2[3 * i] * Cm[3 * j] + t2[3 * i + 1] * Cm[3 * j + 1] + t2[3 * i + 2] * Cm[3 * j + 2];
        cov[9 * (3 + i) + 3 + j] += iC[3 * i + j] * h;
      }
  }
  // state
  double Ra[3];
  m3_vec(R, a, Ra);
  for (int k = 0; k < 3; ++k) {
    x[3 + k] += x[6 + k] * h + Ra[k] * q;
    x[6 + k] += Ra[k] * h;
    x[k] += c[k] * h;
  }
}

struct PreintBody {
  static VUS_DEV void run(const PreintArgs& P, long f) {
    double dR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, dP[3] = {0, 0, 0}, dV[3] = {0, 0, 0}, T = 0.0;
    double JRg[9] = {0}, JPa[9] = {0}, JPg[9] = {0}, JVa[9] = {0}, JVg[9] = {0};
    double cov[81], A[81], tmp[81];
    for (int e = 0; e < 81; ++e) cov[e] = 0.0;
    const double h = P.dt, q = 0.5 * h * h;
    if (P.tangent) {
      double x[9] = {0}, Ha[27] = {0}, Hg[27] = {0};
      for (int s = 0; s < P.k; ++s) {
        const double* am = P.acc + ((long)f * P.k + s) * 3;
        const double* wm = P.gyro + ((long)f * P.k + s) * 3;
        const double a[3] = {am[0] - P.bhat[0], am[1] - P.bhat[1], am[2] - P.bhat[2]};
        const double w[3] = {wm[0] - P.bhat[3], wm[1] - P.bhat[4], wm[2] - P.bhat[5]};
        preint_tangent_step(x, Ha, Hg, cov, a, w, h, P.aC, P.wC, P.iC, A, tmp);
        T += h;
      }
      for (int e = 0; e < 9; ++e) dR[e] = e < 3 ? x[e] : 0.0;        // packed row: theta in the first three slots
      for (int c = 0; c < 3; ++c) { dP[c] = x[3 + c]; dV[c] = x[6 + c]; }
      for (int e = 0; e < 9; ++e) { JRg[e] = Hg[e]; JPa[e] = Ha[9 + e]; JPg[e] = Hg[9 + e]; JVa[e] = Ha[18 + e]; JVg[e] = Hg[18 + e]; }
    }
    for (int s = 0; s < (P.tangent ? 0 : P.k); ++s) {
      const double* am = P.acc + ((long)f * P.k + s) * 3;
      const double* wm = P.gyro + ((long)f * P.k + s) * 3;
      const double a[3] = {am[0] - P.bhat[0], am[1] - P.bhat[1], am[2] - P.bhat[2]};
      const double w[3] = {(wm[0] - P.bhat[3]) * h, (wm[1] - P.bhat[4]) * h, (wm[2] - P.bhat[5]) * h};
      double inc[9], Jr[9], incT[9], ax[9], iTa[9];
      so3_exp(w, inc);
      so3_dexp(w, Jr);
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) incT[3 * i + j] = inc[3 * j + i];
      skew(a, ax);
      m3_mul(incT, ax, iTa);
      // ---- covariance: cov = A cov A^T + B (aC / h) B^T ; cov[0:3,0:3] += C (wC / h) C^T ; cov[3:6,3:6] += iC h
      for (int e = 0; e < 81; ++e) A[e] = 0.0;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          const double it = incT[3 * i + j], ia = iTa[3 * i + j];
          A[9 * i + j] = it;
          A[9 * (3 + i) + j] = -ia * q;  A[9 * (3 + i) + 3 + j] = it;  A[9 * (3 + i) + 6 + j] = it * h;
          A[9 * (6 + i) + j] = -ia * h;  A[9 * (6 + i) + 6 + j] = it;
        }
      m9_mul(A, cov, tmp, false);
      m9_mul(tmp, A, cov, true);
      {
        double Bm[18], t1[18];                   // rows 3..8 of B = [incT q ; incT h]  (6 x 3); rows 0..2 are zero
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) { Bm[3 * i + j] = incT[3 * i + j] * q; Bm[3 * (3 + i) + j] = incT[3 * i + j] * h; }
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 3; ++j) t1[3 * i + j] = (Bm[3 * i] * P.aC[j] + Bm[3 * i + 1] * P.aC[3 + j] + Bm[3 * i + 2] * P.aC[6 + j]) / h;
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) cov[9 * (3 + i) + 3 + j] += t1[3 * i] * Bm[3 * j] + t1[3 * i + 1] * Bm[3 * j + 1] + t1[3 * i + 2] * Bm[3 * j + 2];
        double Cm[9], t2[9];
        for (int e = 0; e < 9; ++e) Cm[e] = Jr[e] * h;
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) t2[3 * i + j] = (Cm[3 * i] * P.wC[j] + Cm[3 * i + 1] * P.wC[3 + j] + Cm[3 * i + 2] * P.wC[6 + j]) / h;
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) {
            cov[9 * i + j] += t2[3 * i] * Cm[3 * j] + t2[3 * i + 1] * Cm[3 * j + 1] + t2[3 * i + 2] * Cm[3 * j + 2];
            cov[9 * (3 + i) + 3 + j] += P.iC[3 * i + j] * h;
          }
      }
      // ---- bias Jacobians (with the pre-update dR, dR/dbg)
      double Rax[9], dacc[9];
      m3_mul(dR, ax, Rax);
      m3_mul(Rax, JRg, dacc);
      for (int e = 0; e < 9; ++e) {
        dacc[e] = -dacc[e];
        JPa[e] += JVa[e] * h - q * dR[e];
        JPg[e] += JVg[e] * h + q * dacc[e];
        JVa[e] -= dR[e] * h;
        JVg[e] += dacc[e] * h;
      }
      double nJ[9];
      m3_mul(incT, JRg, nJ);
      for (int e = 0; e < 9; ++e) JRg[e] = nJ[e] - Jr[e] * h;
      // ---- deltas
      double Ra[3], nR[9];
      m3_vec(dR, a, Ra);
      for (int c = 0; c < 3; ++c) { dP[c] += dV[c] * h + Ra[c] * q; dV[c] += Ra[c] * h; }
      m3_mul(dR, inc, nR);
      for (int e = 0; e < 9; ++e) dR[e] = nR[e];
      T += h;
    }
    // ---- packed PIM row
    const long n = P.n;
    for (int e = 0; e < 9; ++e) P.pim[(long)e * n + f] = dR[e];
    for (int c = 0; c < 3; ++c) { P.pim[(long)(9 + c) * n + f] = dP[c]; P.pim[(long)(12 + c) * n + f] = dV[c]; }
    P.pim[(long)15 * n + f] = T;
    for (int c = 0; c < 6; ++c) P.pim[(long)(16 + c) * n + f] = P.bhat[c];
    for (int e = 0; e < 9; ++e) {
      P.pim[(long)(22 + e) * n + f] = JRg[e]; P.pim[(long)(31 + e) * n + f] = JPa[e]; P.pim[(long)(40 + e) * n + f] = JPg[e];
      P.pim[(long)(49 + e) * n + f] = JVa[e]; P.pim[(long)(58 + e) * n + f] = JVg[e];
    }
    // ---- sqrt information: R upper, R^T R = cov^-1  (noiseModel::Gaussian::Covariance)
    // cov = L L^T ; cov^-1 = L^-T L^-1 =: M ; M = Lm Lm^T ; R = Lm^T
    double* L = A;
    for (int e = 0; e < 81; ++e) L[e] = 0.0;
    bool ok = true;
    for (int j = 0; j < 9; ++j) {
      double d = cov[9 * j + j];
      for (int l = 0; l < j; ++l) d -= L[9 * j + l] * L[9 * j + l];
      if (!(d > 0.0)) { ok = false; d = 1.0; }
      const double dj = sqrt(d);
      L[9 * j + j] = dj;
      for (int i = j + 1; i < 9; ++i) {
        double v = 0.5 * (cov[9 * i + j] + cov[9 * j + i]);
        for (int l = 0; l < j; ++l) v -= L[9 * i + l] * L[9 * j + l];
        L[9 * i + j] = v / dj;
      }
    }
    double* Li = tmp;                            // L^-1 (lower)
    for (int e = 0; e < 81; ++e) Li[e] = 0.0;
    for (int c = 0; c < 9; ++c) {
      Li[9 * c + c] = 1.0 / L[9 * c + c];
      for (int i = c + 1; i < 9; ++i) {
        double v = 0.0;
        for (int l = c; l < i; ++l) v -= L[9 * i + l] * Li[9 * l + c];
        Li[9 * i + c] = v / L[9 * i + i];
      }
    }
    double* M = cov;                             // M = Li^T Li
    for (int i = 0; i < 9; ++i)
      for (int j = 0; j < 9; ++j) {
        double v = 0.0;
        for (int l = (i > j ? i : j); l < 9; ++l) v += Li[9 * l + i] * Li[9 * l + j];
        M[9 * i + j] = v;
      }
    for (int e = 0; e < 81; ++e) L[e] = 0.0;     // Lm
    for (int j = 0; j < 9; ++j) {
      double d = M[9 * j + j];
      for (int l = 0; l < j; ++l) d -= L[9 * j + l] * L[9 * j + l];
      if (!(d > 0.0)) { ok = false; d = 1.0; }
      const double dj = sqrt(d);
      L[9 * j + j] = dj;
      for (int i = j + 1; i < 9; ++i) {
        double v = M[9 * i + j];
        for (int l = 0; l < j; ++l) v -= L[9 * i + l] * L[9 * j + l];
        L[9 * i + j] = v / dj;
      }
    }
    int p = 0;
    for (int r = 0; r < 9; ++r)
      for (int c = r; c < 9; ++c) P.sinfo[(long)(p++) * n + f] = L[9 * c + r];   // R[r][c] = Lm[c][r]
    if (!ok) *P.fail = 1;
  }
};

struct BackprojArgs {
  long n;
  const int* pose_idx;             // [n] pose of each observation
  const double* pose; long nx;     // [12][nx]
  const double* meas;              // [3][n]  uL, uR, v
  double K[6];
  double* out;                     // [3][n] world points
  int* fail;
};
struct BackprojBody {
  static VUS_DEV void run(const BackprojArgs& A, long o) {
    double R[9], t[3], q[3], w[3];
    load_pose(A.pose, A.nx, A.pose_idx[o], R, t);
    const double uL = A.meas[o], uR = A.meas[A.n + o], v = A.meas[2 * A.n + o];
    const double fx = A.K[0], fy = A.K[1], u0 = A.K[3], v0 = A.K[4], b = A.K[5];
    const double disp = uL - uR;
    if (!(disp > 0.0)) *A.fail = 1;                      // point at or behind infinity (cheirality)
    const double z = fx * b / disp;
    q[0] = (uL - u0) * z / fx; q[1] = (v - v0) * z / fy; q[2] = z;
    m3_vec(R, q, w);
    A.out[o] = w[0] + t[0]; A.out[A.n + o] = w[1] + t[1]; A.out[2 * A.n + o] = w[2] + t[2];
  }
};

}  // namespace vus
