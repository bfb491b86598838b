// FP64 SO(3)/SE(3) device library (per-thread, register resident).
//
// Same semantics as the functions gtsam evaluates under LevenbergMarquardtOptimizer::optimize
// (reference call site /root/reference/batch.py:337): SO3 Expmap/Logmap/ExpmapDerivative/
// LogmapDerivative and Pose3 Expmap/Logmap/AdjointMap/LogmapDerivative, with
// gtsam's near-zero thresholds.  3x3 matrices are row-major double[9]; pose tangent is
// [omega; v]; retraction is T * Exp(xi).
#pragma once
#include "vus_common.h"

namespace vus {

VUS_HD void m3_mul(const double* A, const double* B, double* C) {   // C = A B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
VUS_HD void m3_mulT(const double* A, const double* B, double* C) {  // C = A B^T
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[3 * j] + A[3 * i + 1] * B[3 * j + 1] + A[3 * i + 2] * B[3 * j + 2];
}
VUS_HD void m3_Tmul(const double* A, const double* B, double* C) {  // C = A^T B
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}
VUS_HD void m3_vec(const double* A, const double* x, double* y) {   // y = A x
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = A[3 * i] * x[0] + A[3 * i + 1] * x[1] + A[3 * i + 2] * x[2];
}
VUS_HD void m3_Tvec(const double* A, const double* x, double* y) {  // y = A^T x
#pragma unroll
  for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[3 + i] * x[1] + A[6 + i] * x[2];
}
VUS_HD void skew(const double* w, double* W) {
  W[0] = 0.0; W[1] = -w[2]; W[2] = w[1];
  W[3] = w[2]; W[4] = 0.0; W[5] = -w[0];
  W[6] = -w[1]; W[7] = w[0]; W[8] = 0.0;
}
// C = A [w]x   (column k of C = A (e_k-th column of [w]x))
VUS_HD void m3_mul_skew(const double* A, const double* w, double* C) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double a0 = A[3 * i], a1 = A[3 * i + 1], a2 = A[3 * i + 2];
    C[3 * i] = a1 * w[2] - a2 * w[1];
    C[3 * i + 1] = a2 * w[0] - a0 * w[2];
    C[3 * i + 2] = a0 * w[1] - a1 * w[0];
  }
}

// SO3::Expmap: Rodrigues; theta^2 <= eps -> I + W.
VUS_HD void so3_exp(const double* w, double* R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (th2 <= VUS_EPS) {
    R[0] = 1.0; R[1] = -w[2]; R[2] = w[1];
    R[3] = w[2]; R[4] = 1.0; R[5] = -w[0];
    R[6] = -w[1]; R[7] = w[0]; R[8] = 1.0;
    return;
  }
  const double th = sqrt(th2), inv = 1.0 / th;
  const double k[3] = {w[0] * inv, w[1] * inv, w[2] * inv};
  const double s = sin(th), s2 = sin(0.5 * th), c1 = 2.0 * s2 * s2;
  double K[9], KK[9];
  skew(k, K);
  m3_mul(K, K, KK);
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = s * K[i] + c1 * KK[i];
  R[0] += 1.0; R[4] += 1.0; R[8] += 1.0;
}

// SO3::Logmap (gtsam 4.1 thresholds; see oracle/lie.py::so3_log).
VUS_HD void so3_log(const double* R, double* w) {
  const double tr = R[0] + R[4] + R[8];
  if (fabs(tr + 1.0) < 1e-10) {
    if (fabs(R[8] + 1.0) > 1e-5) {
      const double m = M_PI / sqrt(2.0 + 2.0 * R[8]);
      w[0] = m * R[2]; w[1] = m * R[5]; w[2] = m * (1.0 + R[8]);
    } else if (fabs(R[4] + 1.0) > 1e-5) {
      const double m = M_PI / sqrt(2.0 + 2.0 * R[4]);
      w[0] = m * R[1]; w[1] = m * (1.0 + R[4]); w[2] = m * R[7];
    } else {
      const double m = M_PI / sqrt(2.0 + 2.0 * R[0]);
      w[0] = m * (1.0 + R[0]); w[1] = m * R[3]; w[2] = m * R[6];
    }
    return;
  }
  const double tr3 = tr - 3.0;
  double mag;
  if (tr3 < -1e-6) {
    double c = (tr - 1.0) * 0.5;
    c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
    const double th = acos(c);
    mag = th / (2.0 * sin(th));
  } else {
    mag = 0.5 - tr3 / 12.0 + tr3 * tr3 / 60.0;
  }
  w[0] = mag * (R[7] - R[5]);
  w[1] = mag * (R[2] - R[6]);
  w[2] = mag * (R[3] - R[1]);
}

// SO3::ExpmapDerivative (right Jacobian).
VUS_HD void so3_dexp(const double* w, double* J) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double W[9];
  if (th2 <= VUS_EPS) {
    skew(w, W);
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = -0.5 * W[i];
    J[0] += 1.0; J[4] += 1.0; J[8] += 1.0;
    return;
  }
  const double th = sqrt(th2), inv = 1.0 / th;
  const double k[3] = {w[0] * inv, w[1] * inv, w[2] * inv};
  double KK[9];
  skew(k, W);
  m3_mul(W, W, KK);
  const double s2 = sin(0.5 * th);
  const double a = 2.0 * s2 * s2 / th, b = 1.0 - sin(th) / th;
#pragma unroll
  for (int i = 0; i < 9; ++i) J[i] = -a * W[i] + b * KK[i];
  J[0] += 1.0; J[4] += 1.0; J[8] += 1.0;
}

// SO3::LogmapDerivative (inverse right Jacobian); theta^2 <= eps -> I.
VUS_HD void so3_dlog(const double* w, double* J) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (th2 <= VUS_EPS) {
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = 0.0;
    J[0] = J[4] = J[8] = 1.0;
    return;
  }
  const double th = sqrt(th2);
  double W[9], WW[9];
  skew(w, W);
  m3_mul(W, W, WW);
  const double coef = 1.0 / th2 - (1.0 + cos(th)) / (2.0 * th * sin(th));
#pragma unroll
  for (int i = 0; i < 9; ++i) J[i] = 0.5 * W[i] + coef * WW[i];
  J[0] += 1.0; J[4] += 1.0; J[8] += 1.0;
}

// Pose3::Expmap.
VUS_HD void pose_exp(const double* xi, double* R, double* t) {
  const double* w = xi;
  const double* v = xi + 3;
  so3_exp(w, R);
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  if (th2 > VUS_EPS) {
    const double wv = w[0] * v[0] + w[1] * v[1] + w[2] * v[2];
    const double c[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
    double Rc[3];
    m3_vec(R, c, Rc);
    const double inv = 1.0 / th2;
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = (c[i] - Rc[i] + w[i] * wv) * inv;
  } else {
    t[0] = v[0]; t[1] = v[1]; t[2] = v[2];
  }
}

// Pose3::Logmap.
VUS_HD void pose_log(const double* R, const double* t, double* xi) {
  so3_log(R, xi);
  const double th = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
  if (th < 1e-10) {
    xi[3] = t[0]; xi[4] = t[1]; xi[5] = t[2];
    return;
  }
  const double inv = 1.0 / th;
  const double k[3] = {xi[0] * inv, xi[1] * inv, xi[2] * inv};
  double W[9], WT[3], WWT[3];
  skew(k, W);
  m3_vec(W, t, WT);
  m3_vec(W, WT, WWT);
  const double a = 0.5 * th, b = 1.0 - th / (2.0 * tan(0.5 * th));
#pragma unroll
  for (int i = 0; i < 3; ++i) xi[3 + i] = t[i] - a * WT[i] + b * WWT[i];
}

// Pose3::computeQforExpmapDerivative.
VUS_HD void pose_Q(const double* xi, double* Q) {
  double V[9], W[9], WV[9], VW[9], WVW[9], WW[9], WWV[9], VWW[9], A[9], Bm[9];
  skew(xi + 3, V);
  skew(xi, W);
  m3_mul(W, V, WV);
  m3_mul(V, W, VW);
  m3_mul(WV, W, WVW);
  m3_mul(W, W, WW);
  m3_mul(WW, V, WWV);
  m3_mul(V, WW, VWW);
  m3_mul(WVW, W, A);   // WVW W
  m3_mul(W, WVW, Bm);  // W WVW
  const double phi = sqrt(xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]);
  double c1, c2, c3;
  if (phi > 1e-5) {
    const double s = sin(phi), c = cos(phi);
    const double p2 = phi * phi, p3 = p2 * phi, p4 = p3 * phi, p5 = p4 * phi;
    c1 = (phi - s) / p3;
    c2 = (1.0 - p2 / 2.0 - c) / p4;
    c3 = -0.5 * (c2 - 3.0 * (phi - s - p3 / 6.0) / p5);
  } else {
    c1 = 1.0 / 6.0;
    c2 = -1.0 / 24.0;
    c3 = 0.5 * (1.0 / 24.0 + 3.0 / 120.0);
  }
#pragma unroll
  for (int i = 0; i < 9; ++i)
    Q[i] = -0.5 * V[i] + c1 * (WV[i] + VW[i] - WVW[i]) + c2 * (WWV[i] + VWW[i] - 3.0 * WVW[i]) + c3 * (A[i] + Bm[i]);
}

// Pose3::LogmapDerivative at xi = Logmap(pose): J = [[Jw,0],[Q2,Jw]], Q2 = -Jw Q Jw.  Returns Jw and Q2.
VUS_HD void pose_dlog(const double* xi, double* Jw, double* Q2) {
  double Q[9], T[9];
  so3_dlog(xi, Jw);
  pose_Q(xi, Q);
  m3_mul(Jw, Q, T);
  m3_mul(T, Jw, Q2);
#pragma unroll
  for (int i = 0; i < 9; ++i) Q2[i] = -Q2[i];
}

}  // namespace vus
