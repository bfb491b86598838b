// Kernel bodies for the batch-LM hot path (kernels 1-4 of BASELINE.json north_star), single source
// for the sm_100a build and the test-only host emulation (see vus_common.h).
//
//   1. linearize / error / linearised-error  : one factor per thread, SoA, FP64       (factors.cuh)
//   2. normal-equation assembly               : J^T J / J^T b scatter-add into the block-banded
//      Hessian (supernode-dense block-tridiagonal + off-band remainder blocks + dense bias border),
//      landmark blocks and the landmark Schur complement
//   3. linear solve                           : PCG on the reduced camera system; operator = banded
//      block matvec + remainder block-SpMV + border; preconditioner = block cyclic reduction (BCR)
//      of the band with the bias border eliminated exactly
//   4. LM step                                : retract, error, controller scalars
#pragma once
#include "factors.cuh"

namespace vus {

#ifdef VUS_EMU
VUS_DEV void atomic_add(double* p, double v) { *p += v; }
#else
VUS_DEV void atomic_add(double* p, double v) { atomicAdd(p, v); }
#endif

// =====================================================================================
// Kernel 1: linearize / error
// =====================================================================================
struct LinArgs {
  ValuesView V;
  FactorView F;
  LinOut O;
  double K[6];
  double g[3];
  int type;
};

template <int TYPE, bool WJ>
struct LinBody {
  static VUS_DEV void run(const LinArgs& a, long f) {
    if (TYPE == VUS_F_PRIOR_POSE) f_prior_pose<WJ>(a.V, a.F, a.O, f);
    else if (TYPE == VUS_F_PRIOR_VEL) f_prior_vel<WJ>(a.V, a.F, a.O, f);
    else if (TYPE == VUS_F_BETWEEN) f_between<WJ>(a.V, a.F, a.O, f);
    else if (TYPE == VUS_F_DVL) f_dvl<WJ>(a.V, a.F, a.O, f);
    else if (TYPE == VUS_F_STEREO) f_stereo<WJ>(a.V, a.F, a.O, f, a.K);
    else f_imu<WJ>(a.V, a.F, a.O, f, a.g);
  }
};

// delta layout: camera part xc[node * D + dof] (pose dofs 0-5, velocity dofs 6-8), bias xb[6], landmarks xl[c * nl + l]
struct DeltaView {
  const double* xc; const double* xb; const double* xl; long nl; int D;
};

// linear.error(delta) per factor: 0.5 || J delta + r ||^2  (gtsam: b = -r)
struct LinErrArgs {
  FactorView F;
  const double* r; const double* J;
  DeltaView X;
  double* out;     // [n]
  int type;
};
template <int TYPE>
struct LinErrBody {
  static VUS_DEV void run(const LinErrArgs& a, long f) {
    constexpr int M = kFactorM[TYPE], C = kFactorCols[TYPE];
    const long n = a.F.n;
    const int D = a.X.D;
    double d[C];
    const int* ix = a.F.idx;
    if (TYPE == VUS_F_PRIOR_POSE) {
      const long x = ix[f];
#pragma unroll
      for (int c = 0; c < 6; ++c) d[c] = a.X.xc[x * D + c];
    } else if (TYPE == VUS_F_PRIOR_VEL) {
      const long v = ix[f];
#pragma unroll
      for (int c = 0; c < 3; ++c) d[c] = a.X.xc[v * D + 6 + c];
    } else if (TYPE == VUS_F_BETWEEN) {
      const long x1 = ix[f], x2 = ix[n + f];
#pragma unroll
      for (int c = 0; c < 6; ++c) { d[c] = a.X.xc[x1 * D + c]; d[6 + c] = a.X.xc[x2 * D + c]; }
    } else if (TYPE == VUS_F_DVL) {
      const long v = ix[f], x = ix[n + f];
#pragma unroll
      for (int c = 0; c < 6; ++c) d[c] = a.X.xc[x * D + c];
#pragma unroll
      for (int c = 0; c < 3; ++c) d[6 + c] = a.X.xc[v * D + 6 + c];
    } else if (TYPE == VUS_F_STEREO) {
      const long x = ix[f], l = ix[n + f];
#pragma unroll
      for (int c = 0; c < 6; ++c) d[c] = a.X.xc[x * D + c];
#pragma unroll
      for (int c = 0; c < 3; ++c) d[6 + c] = a.X.xl[c * a.X.nl + l];
    } else {
      const long xi = ix[f], vi = ix[n + f], xj = ix[2 * n + f], vj = ix[3 * n + f];
#pragma unroll
      for (int c = 0; c < 6; ++c) { d[c] = a.X.xc[xi * D + c]; d[9 + c] = a.X.xc[xj * D + c]; d[18 + c] = a.X.xb[c]; }
#pragma unroll
      for (int c = 0; c < 3; ++c) { d[6 + c] = a.X.xc[vi * D + 6 + c]; d[15 + c] = a.X.xc[vj * D + 6 + c]; }
    }
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < M; ++r) {
      double s = a.r[r * n + f];
#pragma unroll
      for (int c = 0; c < C; ++c) s += a.J[(r * C + c) * n + f] * d[c];
      acc += s * s;
    }
    a.out[f] = 0.5 * acc;
  }
};

// =====================================================================================
// deterministic two-stage sum reduction with a scalar post-op
// =====================================================================================
enum { RED_STORE = 0, RED_PAP = 1, RED_RZ = 2, RED_RZ0 = 3 };
// scalar slots
enum { S_RZ = 0, S_PAP = 1, S_ALPHA = 2, S_BETA = 3, S_RR = 4, S_TMP = 5, S_NEG_ALPHA = 6, S_COUNT = 16 };

struct RedArgs {
  const double* a; const double* b;   // sum a[i]*b[i] (b null -> sum a[i])
  long n;
  double* partials;                   // [grid]
  int grid;
};
struct Red1Body {
  static VUS_DEV void run(const RedArgs& A, int bid, int tid, int nthr, double* sm) {
    const long chunk = (A.n + A.grid - 1) / A.grid;
    const long i0 = (long)bid * chunk;
    long i1 = i0 + chunk;
    if (i1 > A.n) i1 = A.n;
    double acc = 0.0;
    if (A.b) for (long i = i0 + tid; i < i1; i += nthr) acc += A.a[i] * A.b[i];
    else for (long i = i0 + tid; i < i1; i += nthr) acc += A.a[i];
    sm[tid] = acc;
    VUS_SYNC();
    for (int s = nthr >> 1; s > 0; s >>= 1) {
      for (int t = tid; t < s; t += nthr) sm[t] += sm[t + s];
      VUS_SYNC();
    }
    if (tid == 0) A.partials[bid] = sm[0];
  }
};
struct Red2Args {
  const double* partials; int grid;
  double* scal;      // scalar block
  int slot; int op;
};
struct Red2Body {
  static VUS_DEV void run(const Red2Args& A, int, int tid, int nthr, double* sm) {
    double acc = 0.0;
    for (int i = tid; i < A.grid; i += nthr) acc += A.partials[i];
    sm[tid] = acc;
    VUS_SYNC();
    for (int s = nthr >> 1; s > 0; s >>= 1) {
      for (int t = tid; t < s; t += nthr) sm[t] += sm[t + s];
      VUS_SYNC();
    }
    if (tid == 0) {
      const double v = sm[0];
      double* s = A.scal;
      if (A.op == RED_STORE) s[A.slot] = v;
      else if (A.op == RED_PAP) { s[S_PAP] = v; const double al = (v != 0.0) ? s[S_RZ] / v : 0.0; s[S_ALPHA] = al; s[S_NEG_ALPHA] = -al; }
      else if (A.op == RED_RZ) { const double old = s[S_RZ]; s[S_BETA] = (old != 0.0) ? v / old : 0.0; s[S_RZ] = v; }
      else if (A.op == RED_RZ0) { s[S_RZ] = v; s[S_BETA] = 0.0; }
    }
  }
};

// =====================================================================================
// Kernel 2: assembly of the non-stereo factors (atomic scatter-add)
// =====================================================================================
// Block-banded Hessian storage (all FP64, row-major):
//   SD[Ns][B][B]   diagonal supernode blocks (B = k*D, full symmetric storage)
//   SU[Ns-1][B][B] coupling (I, I+1)
//   REM[nrem][D][D] off-band node blocks, both (p,q) and (q,p) stored
//   F[Npad*D][6]   dense bias border, Hbb[36], g[Npad*D], gb[6]
struct PairDst {        // where block (p,q) of a two-node factor goes (built by vus_analyze)
  long off;             // element (0,0) of the primary destination inside Hval
  long moff;            // mirror (q,p) destination, -1 if none
  int ld, mld;          // row strides
  int transposed;       // primary stores the block transposed (p is the higher supernode)
  int pad;
};

struct AsmArgs {
  int type; long n;
  const int* idx; const double* J; const double* r;
  int D, k, B;
  double* Hval;         // base of SD | SU | REM
  double* g; double* F; double* Hbb; double* gb;
  const PairDst* pair;  // [n] for between / imu
};

VUS_HD long diag_off(long node, int la, int lb, int D, int k, int B) {
  const long I = node / k;
  const int rp = (int)(node % k);
  return I * (long)B * B + (long)(rp * D + la) * B + rp * D + lb;
}

template <int TYPE>
struct AsmBody {
  // work item = (e, f): e < C*C -> Hessian entry (a,b); e >= C*C -> gradient entry a
  static VUS_DEV void run(const AsmArgs& A, long w) {
    constexpr int M = kFactorM[TYPE], C = kFactorCols[TYPE];
    const long n = A.n;
    const long f = w % n;
    const int e = (int)(w / n);
    const int D = A.D;
    long p = 0, q = 0;
    if (TYPE == VUS_F_PRIOR_POSE || TYPE == VUS_F_PRIOR_VEL) p = A.idx[f];
    else if (TYPE == VUS_F_BETWEEN) { p = A.idx[f]; q = A.idx[n + f]; }
    else if (TYPE == VUS_F_DVL) p = A.idx[n + f];
    else { p = A.idx[f]; q = A.idx[2 * n + f]; }
    // column -> (group, local dof)
    auto grp = [](int c, int& g, int& l) {
      if (TYPE == VUS_F_PRIOR_POSE || TYPE == VUS_F_DVL) { g = 0; l = c; }
      else if (TYPE == VUS_F_PRIOR_VEL) { g = 0; l = 6 + c; }
      else if (TYPE == VUS_F_BETWEEN) { g = c < 6 ? 0 : 1; l = c < 6 ? c : c - 6; }
      else { g = c < 9 ? 0 : (c < 18 ? 1 : 2); l = c < 9 ? c : (c < 18 ? c - 9 : c - 18); }
    };
    if (e >= C * C) {                              // gradient: g_a -= sum_r J[r][a] r[r]
      const int a = e - C * C;
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < M; ++r) s += A.J[(r * C + a) * n + f] * A.r[r * n + f];
      int ga, la;
      grp(a, ga, la);
      if (ga == 2) return;                           // bias gradient: ImuBiasBody (block reduction, no atomics)
      atomic_add(&A.g[(ga == 0 ? p : q) * D + la], -s);
      return;
    }
    const int a = e / C, b = e % C;
    int ga, la, gb, lb;
    grp(a, ga, la);
    grp(b, gb, lb);
    if (ga > gb) return;                           // (q,p), (bias,p), (bias,q): written by the mirrored item
    if (ga == 2 && gb == 2) return;                // bias-bias block: ImuBiasBody
    double h = 0.0;
#pragma unroll
    for (int r = 0; r < M; ++r) h += A.J[(r * C + a) * n + f] * A.J[(r * C + b) * n + f];
    if (ga == gb) {
      if (ga == 2) return;                           // bias-bias block: ImuBiasBody
      atomic_add(&A.Hval[diag_off(ga == 0 ? p : q, la, lb, D, A.k, A.B)], h);
    } else if (gb == 2) {                          // (node, bias) border
      atomic_add(&A.F[((ga == 0 ? p : q) * D + la) * 6 + lb], h);
    } else {                                       // (p, q) coupling
      const PairDst d = A.pair[f];
      if (d.transposed) atomic_add(&A.Hval[d.off + (long)lb * d.ld + la], h);
      else atomic_add(&A.Hval[d.off + (long)la * d.ld + lb], h);
      if (d.moff >= 0) atomic_add(&A.Hval[d.moff + (long)lb * d.mld + la], h);
    }
  }
};

// shared-bias block: Hbb = sum_f Jb^T Jb (36) and gb = -sum_f Jb^T r (6) over ALL imu factors -- every factor hits the
// same 42 addresses, so this is a two-stage block reduction instead of atomics.  partials [grid][42]
struct ImuBiasArgs { long n; const double* J; const double* r; double* partials; int grid; double* Hbb; double* gb; };
struct ImuBias1Body {
  static VUS_DEV void run(const ImuBiasArgs& A, int bid, int tid, int nthr, double* sm) {
    const long chunk = (A.n + A.grid - 1) / A.grid;
    const long f0 = (long)bid * chunk;
    long f1 = f0 + chunk;
    if (f1 > A.n) f1 = A.n;
    for (int e = 0; e < 42; ++e) {
      double acc = 0.0;
      if (e < 36) {
        const int a = 18 + e / 6, b = 18 + e % 6;
        for (long f = f0 + tid; f < f1; f += nthr)
          for (int r = 0; r < 9; ++r) acc += A.J[(r * 24 + a) * A.n + f] * A.J[(r * 24 + b) * A.n + f];
      } else {
        const int a = 18 + e - 36;
        for (long f = f0 + tid; f < f1; f += nthr)
          for (int r = 0; r < 9; ++r) acc -= A.J[(r * 24 + a) * A.n + f] * A.r[r * A.n + f];
      }
      sm[tid] = acc;
      VUS_SYNC();
      for (int s = nthr >> 1; s > 0; s >>= 1) {
        for (int t = tid; t < s; t += nthr) sm[t] += sm[t + s];
        VUS_SYNC();
      }
      if (tid == 0) A.partials[(long)bid * 42 + e] = sm[0];
      VUS_SYNC();
    }
  }
};
struct ImuBias2Body {
  static VUS_DEV void run(const ImuBiasArgs& A, long e) {
    double s = 0.0;
    for (int b = 0; b < A.grid; ++b) s += A.partials[(long)b * 42 + e];
    if (e < 36) A.Hbb[e] = s; else A.gb[e - 36] = s;
  }
};

// ------------------------------------------------------------------ stereo: gathers (no atomics)
struct StereoAsmArgs {
  long n;                      // #observations
  const int* idx;              // [2][n] pose, landmark
  const double* J; const double* r;      // 3x9 / 3
  int D, k, B;
  double* SD; double* g;       // camera diag blocks / gradient
  double* C; double* gl;       // [9][nl], [3][nl]
  double* E;                   // [18][n]   E_o = Jp^T Jl (6x3 row-major)
  const double* Pp; const double* Pl;    // per-observation products written by the stereo linearize kernel
  long nl;
  const int* pose_ptr; const int* pose_obs; const int* pose_ids; long nposes_obs;   // CSR pose -> obs
  const int* lm_ptr; const int* lm_obs;                                              // CSR landmark -> obs
};
// per (pose-with-observations, e<28): sums the per-observation products the stereo linearize kernel wrote
// (P_pose[o][e], contiguous in e -> coalesced) over the pose's observations.  e<21: unique B_ii entry (a<=b), 21..26: gradient
struct StereoPoseBody {
  static VUS_DEV void run(const StereoAsmArgs& A, long w) {
    const long pi = w / 28;
    const int e = (int)(w - pi * 28);
    if (e >= 27) return;
    const long node = A.pose_ids[pi];
    double s = 0.0;
    for (int t = A.pose_ptr[pi]; t < A.pose_ptr[pi + 1]; ++t) s += A.Pp[(long)A.pose_obs[t] * 28 + e];
    if (e < 21) {
      int a = 0, rem = e;
      while (rem >= 6 - a) { rem -= 6 - a; ++a; }
      const int b = a + rem;
      A.SD[diag_off(node, a, b, A.D, A.k, A.B)] += s;
      if (a != b) A.SD[diag_off(node, b, a, A.D, A.k, A.B)] += s;
    } else {
      A.g[node * A.D + (e - 21)] -= s;
    }
  }
};
// per (landmark, e<12): e<6 unique C entry (a<=b), 6..8: g_l
struct StereoLmBody {
  static VUS_DEV void run(const StereoAsmArgs& A, long w) {
    const long l = w / 12;
    const int e = (int)(w - l * 12);
    if (e >= 9) return;
    double s = 0.0;
    for (int t = A.lm_ptr[l]; t < A.lm_ptr[l + 1]; ++t) s += A.Pl[(long)A.lm_obs[t] * 12 + e];
    if (e < 6) {
      int a = 0, rem = e;
      while (rem >= 3 - a) { rem -= 3 - a; ++a; }
      const int b = a + rem;
      A.C[(a * 3 + b) * A.nl + l] = s;
      if (a != b) A.C[(b * 3 + a) * A.nl + l] = s;
    } else {
      A.gl[(e - 6) * A.nl + l] = -s;
    }
  }
};

// ------------------------------------------------------------------ per-lambda landmark Schur complement
struct SchurArgs {
  long n; long nl;
  const int* idx;                    // [2][n]
  const double* C; const double* gl; // undamped landmark blocks
  double* Cinv;                      // [9][nl]  (C + lambda I)^-1
  const double* E; double* W;        // [18][n]  W_o = E_o Cinv
  double lambda;
  int D, k, B;
  double* Hval;                      // damped system being formed
  double* gs; const double* g;       // reduced gradient
  const int* pose_ptr; const int* pose_obs; const int* pose_ids; long nposes_obs;
  const int* lm_ptr; const int* lm_obs;
  // destination lists
  long ndst; const PairDst* dst; const int* dst_ptr; const int* term_a; const int* term_b;
  int* fail;
  // back-substitution
  const double* xc; double* xl;
};
struct LmInvertBody {    // per landmark
  static VUS_DEV void run(const SchurArgs& A, long l) {
    double c[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) c[e] = A.C[e * A.nl + l];
    c[0] += A.lambda; c[4] += A.lambda; c[8] += A.lambda;
    const double c00 = c[4] * c[8] - c[5] * c[7], c01 = c[5] * c[6] - c[3] * c[8], c02 = c[3] * c[7] - c[4] * c[6];
    const double det = c[0] * c00 + c[1] * c01 + c[2] * c02;
    if (!(det > 0.0)) { *A.fail = 1; }
    const double id = 1.0 / det;
    double inv[9];
    inv[0] = c00 * id; inv[1] = (c[2] * c[7] - c[1] * c[8]) * id; inv[2] = (c[1] * c[5] - c[2] * c[4]) * id;
    inv[3] = c01 * id; inv[4] = (c[0] * c[8] - c[2] * c[6]) * id; inv[5] = (c[2] * c[3] - c[0] * c[5]) * id;
    inv[6] = c02 * id; inv[7] = (c[1] * c[6] - c[0] * c[7]) * id; inv[8] = (c[0] * c[4] - c[1] * c[3]) * id;
#pragma unroll
    for (int e = 0; e < 9; ++e) A.Cinv[e * A.nl + l] = inv[e];
  }
};
struct StereoWBody {     // per (obs, e<18): W_o[a][c] = sum_d E_o[a][d] Cinv[d][c]
  static VUS_DEV void run(const SchurArgs& A, long w) {
    const long o = w % A.n;
    const int e = (int)(w / A.n);
    const int a = e / 3, c = e % 3;
    const long l = A.idx[A.n + o];
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) s += A.E[(a * 3 + d) * A.n + o] * A.Cinv[(d * 3 + c) * A.nl + l];
    A.W[e * A.n + o] = s;
  }
};
struct SchurGradBody {   // per (pose-with-obs, a<6): gs -= sum_o W_o g_l
  static VUS_DEV void run(const SchurArgs& A, long w) {
    const long pi = w % A.nposes_obs;
    const int a = (int)(w / A.nposes_obs);
    const long node = A.pose_ids[pi];
    double s = 0.0;
    for (int t = A.pose_ptr[pi]; t < A.pose_ptr[pi + 1]; ++t) {
      const long o = A.pose_obs[t];
      const long l = A.idx[A.n + o];
#pragma unroll
      for (int c = 0; c < 3; ++c) s += A.W[(a * 3 + c) * A.n + o] * A.gl[c * A.nl + l];
    }
    A.gs[node * A.D + a] -= s;
  }
};
struct SchurBlockBody {  // per destination (i<=j): S_(i,j) -= sum_terms W_a E_b^T   (6x6, all 36 outputs in registers)
  static VUS_DEV void run(const SchurArgs& A, long d) {
    double acc[36];
#pragma unroll
    for (int e = 0; e < 36; ++e) acc[e] = 0.0;
    for (int t = A.dst_ptr[d]; t < A.dst_ptr[d + 1]; ++t) {
      const long oa = A.term_a[t], ob = A.term_b[t];
      double w[18], eb[18];
#pragma unroll
      for (int e = 0; e < 18; ++e) { w[e] = A.W[e * A.n + oa]; eb[e] = A.E[e * A.n + ob]; }
#pragma unroll
      for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int s = 0; s < 6; ++s)
          acc[r * 6 + s] += w[r * 3] * eb[s * 3] + w[r * 3 + 1] * eb[s * 3 + 1] + w[r * 3 + 2] * eb[s * 3 + 2];
    }
    const PairDst D = A.dst[d];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int s = 0; s < 6; ++s) {
        const double v = acc[r * 6 + s];
        if (D.transposed) A.Hval[D.off + (long)s * D.ld + r] -= v;
        else A.Hval[D.off + (long)r * D.ld + s] -= v;
        if (D.moff >= 0) A.Hval[D.moff + (long)s * D.mld + r] -= v;
      }
  }
};
struct LmBacksubBody {   // per landmark: xl = Cinv (gl - sum_o E_o^T xc[pose_o])
  static VUS_DEV void run(const SchurArgs& A, long l) {
    double t[3] = {A.gl[l], A.gl[A.nl + l], A.gl[2 * A.nl + l]};
    for (int q = A.lm_ptr[l]; q < A.lm_ptr[l + 1]; ++q) {
      const long o = A.lm_obs[q];
      const long node = A.idx[o];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        const double x = A.xc[node * A.D + a];
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] -= A.E[(a * 3 + c) * A.n + o] * x;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      A.xl[c * A.nl + l] = A.Cinv[(c * 3) * A.nl + l] * t[0] + A.Cinv[(c * 3 + 1) * A.nl + l] * t[1] + A.Cinv[(c * 3 + 2) * A.nl + l] * t[2];
  }
};

// add lambda (and identity on padding dofs) to the diagonal of SD and Hbb
struct DampArgs { double* SD; double* Hbb; long ndof; long nreal; int B; double lambda; };
struct DampBody {
  static VUS_DEV void run(const DampArgs& A, long i) {
    if (i < A.ndof) {
      const long I = i / A.B;
      const int r = (int)(i % A.B);
      A.SD[I * (long)A.B * A.B + (long)r * A.B + r] += (i < A.nreal) ? A.lambda : 1.0;
    } else {
      const int c = (int)(i - A.ndof);
      A.Hbb[c * 6 + c] += A.lambda;
    }
  }
};

// =====================================================================================
// Kernel 3a: system operator  y = A x  on the reduced camera system
// =====================================================================================
struct MatvecArgs {
  const double* SD; const double* SU; long Ns; int B;
  const double* x; double* y;        // camera parts, length Ns*B
  // remainder
  const int* rem_ptr; const int* rem_col; const double* rem_val; long nnodes; int D;
  // border
  const double* F; const double* Hbb; const double* xb; double* yb; int has_bias;
};
// one CTA per supernode: y_I = SD_I x_I + SU_I x_{I+1} + SU_{I-1}^T x_{I-1}.  Blocks are staged in
// shared memory with coalesced loads, then each thread owns one output row.
struct BandMatvecBody {
  static VUS_DEV void run(const MatvecArgs& A, int I, int tid, int nthr, double* sm) {
    const int B = A.B;
    const long BB = (long)B * B;
    double* M = sm;             // [B*B]
    double* xs = sm + BB;       // [B]
    double* acc = xs + B;       // [B]
    for (int i = tid; i < B; i += nthr) acc[i] = 0.0;
    // diagonal block (symmetric)
    for (long i = tid; i < BB; i += nthr) M[i] = A.SD[I * BB + i];
    for (int i = tid; i < B; i += nthr) xs[i] = A.x[(long)I * B + i];
    VUS_SYNC();
    for (int r = tid; r < B; r += nthr) {
      double s = 0.0;
      for (int c = 0; c < B; ++c) s += M[(long)c * B + r] * xs[c];     // column r of a symmetric block
      acc[r] += s;
    }
    VUS_SYNC();
    if (I + 1 < A.Ns) {
      for (long i = tid; i < BB; i += nthr) M[i] = A.SU[I * BB + i];
      for (int i = tid; i < B; i += nthr) xs[i] = A.x[(long)(I + 1) * B + i];
      VUS_SYNC();
      for (int r = tid; r < B; r += nthr) {
        double s = 0.0;
        for (int c = 0; c < B; ++c) s += M[(long)r * B + c] * xs[c];
        acc[r] += s;
      }
      VUS_SYNC();
    }
    if (I > 0) {
      for (long i = tid; i < BB; i += nthr) M[i] = A.SU[(I - 1) * BB + i];
      for (int i = tid; i < B; i += nthr) xs[i] = A.x[(long)(I - 1) * B + i];
      VUS_SYNC();
      for (int r = tid; r < B; r += nthr) {
        double s = 0.0;
        for (int c = 0; c < B; ++c) s += M[(long)c * B + r] * xs[c];
        acc[r] += s;
      }
      VUS_SYNC();
    }
    for (int r = tid; r < B; r += nthr) A.y[(long)I * B + r] = acc[r];
  }
};
// per (node, row): remainder blocks + border column
struct RemBorderMatvecBody {
  static VUS_DEV void run(const MatvecArgs& A, long w) {
    const int D = A.D;
    const long node = w / D;
    const int r = (int)(w % D);
    double s = 0.0;
    if (A.rem_ptr) {
      for (int t = A.rem_ptr[node]; t < A.rem_ptr[node + 1]; ++t) {
        const long col = A.rem_col[t];
        const double* v = A.rem_val + (long)t * D * D + r * D;
        for (int c = 0; c < D; ++c) s += v[c] * A.x[col * D + c];
      }
    }
    if (A.has_bias) {
#pragma unroll
      for (int c = 0; c < 6; ++c) s += A.F[w * 6 + c] * A.xb[c];
    }
    A.y[w] += s;
  }
};
// out[v][c] = sum_i F[i][c] * Y[v][i]  (F^T Y) for nv vectors, two-stage; stage 1 partials [grid][nv*6]
struct BorderDotArgs { const double* F; const double* Y; long len; long ystride; int nv; double* partials; int grid; };
struct BorderDot1Body {
  static VUS_DEV void run(const BorderDotArgs& A, int bid, int tid, int nthr, double* sm) {
    const long chunk = (A.len + A.grid - 1) / A.grid;
    const long i0 = (long)bid * chunk;
    long i1 = i0 + chunk;
    if (i1 > A.len) i1 = A.len;
    for (int v = 0; v < A.nv; ++v)
      for (int c = 0; c < 6; ++c) {
        double acc = 0.0;
        for (long i = i0 + tid; i < i1; i += nthr) acc += A.F[i * 6 + c] * A.Y[(long)v * A.ystride + i];
        sm[tid] = acc;
        VUS_SYNC();
        for (int s = nthr >> 1; s > 0; s >>= 1) {
          for (int t = tid; t < s; t += nthr) sm[t] += sm[t + s];
          VUS_SYNC();
        }
        if (tid == 0) A.partials[(long)bid * (A.nv * 6) + v * 6 + c] = sm[0];
        VUS_SYNC();
      }
  }
};

// =====================================================================================
// Kernel 3b: block cyclic reduction of the supernode block-tridiagonal band
// =====================================================================================
// CTA-level dense helpers on BxB blocks (B <= 96), operands staged whole in shared memory.
//   "A" operands: row stride B, bcr_rows_a(B) rows allocated (rows >= B are never initialised: they only feed
//                 register-tile lanes whose results are discarded);
//   "B" operands: row stride bcr_ldb(B) (multiple of 4 -> 16/32-byte aligned vector loads), B rows.
VUS_HD int bcr_ldb(int B) { return (B + 3) & ~3; }
VUS_HD int bcr_rows_a(int B) { return (B + 7) & ~7; }
VUS_HD long bcr_smem_doubles(int B) { return (long)bcr_rows_a(B) * B + (long)B * bcr_ldb(B) + 4 * B + 8; }

// In-place Gauss-Jordan inverse of an SPD block held in shared memory with row stride ld (no pivoting).
#ifdef VUS_EMU
// host emulation: plain sweep over the block in memory
VUS_DEV void cta_spd_inverse(double* M, int ld, double* rowp, double* colp, int B, int tid, int nthr, int* fail) {
  for (int p = 0; p < B; ++p) {
    for (int i = tid; i < B; i += nthr) { colp[i] = M[(long)i * ld + p]; rowp[i] = M[(long)p * ld + i]; }
    VUS_SYNC();
    const double piv = rowp[p];
    if (tid == 0 && !(piv > 0.0)) *fail = 1;
    const double d = 1.0 / piv;
    for (int i = tid; i < B; i += nthr) {
      double* row = M + (long)i * ld;
      if (i == p) { for (int j = 0; j < B; ++j) row[j] = (j == p) ? d : rowp[j] * d; }
      else { const double ci = colp[i] * d; for (int j = 0; j < B; ++j) row[j] = (j == p) ? -ci : row[j] - ci * rowp[j]; }
    }
    VUS_SYNC();
  }
}
#else
// sm_100a: the block lives in REGISTERS for all B sweeps (thread (tx,ty) of a 256-thread CTA owns rows ty+8a, columns
// tx+32b, a<12, b<3); only the pivot row / column travel through shared memory, double-buffered -> one barrier per sweep.
VUS_DEV void cta_spd_inverse(double* M, int ld, double* rowp, double* colp, int B, int tid, int nthr, int* fail) {
  const int tx = tid & 31, ty = tid >> 5;                  // requires nthr == 256 and rowp/colp of 2*B doubles each
  double m[12][3];
#pragma unroll
  for (int a = 0; a < 12; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int i = ty + 8 * a, j = tx + 32 * b;
      m[a][b] = (i < B && j < B) ? M[(long)i * ld + j] : 0.0;
    }
  for (int p = 0; p < B; ++p) {
    double* rp = rowp + (p & 1) * B;
    double* cp = colp + (p & 1) * B;
    const int pa = p >> 3, pb = p >> 5;
    if (ty == (p & 7)) {
#pragma unroll
      for (int a = 0; a < 12; ++a)
        if (a == pa) {
#pragma unroll
          for (int b = 0; b < 3; ++b) { const int j = tx + 32 * b; if (j < B) rp[j] = m[a][b]; }
        }
    }
    if (tx == (p & 31)) {
#pragma unroll
      for (int b = 0; b < 3; ++b)
        if (b == pb) {
#pragma unroll
          for (int a = 0; a < 12; ++a) { const int i = ty + 8 * a; if (i < B) cp[i] = m[a][b]; }
        }
    }
    __syncthreads();
    const double piv = rp[p];
    if (tid == 0 && !(piv > 0.0)) *fail = 1;
    const double d = 1.0 / piv;
    double r[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) { const int j = tx + 32 * b; r[b] = (j < B) ? rp[j] : 0.0; }
#pragma unroll
    for (int a = 0; a < 12; ++a) {
      const int i = ty + 8 * a;
      if (i < B) {
        const double ci = cp[i] * d;
        if (i == p) {
#pragma unroll
          for (int b = 0; b < 3; ++b) m[a][b] = (tx + 32 * b == p) ? d : r[b] * d;
        } else {
#pragma unroll
          for (int b = 0; b < 3; ++b) m[a][b] = (tx + 32 * b == p) ? -ci : m[a][b] - ci * r[b];
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 12; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int i = ty + 8 * a, j = tx + 32 * b;
      if (i < B && j < B) M[(long)i * ld + j] = m[a][b];
    }
  __syncthreads();
}
#endif
// C(global, row stride B) = beta*C + alpha * sA * sB, 8x4 register tiles, FP64 FMA pipe bound.
// If CT != null also writes the transpose of the result.
VUS_DEV void cta_gemm_ss(double* C, double* CT, const double* sA, const double* sB, int B, double alpha, double beta, int tid, int nthr) {
  const int ldb = bcr_ldb(B);
  const int TI = (B + 7) / 8, TJ = (B + 3) / 4;
  for (int tile = tid; tile < TI * TJ; tile += nthr) {
    const int i0 = (tile / TJ) * 8, j0 = (tile % TJ) * 4;
    double acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const double* pa = sA + (long)i0 * B;
    const double* pb = sB + j0;
#pragma unroll 3
    for (int k = 0; k < B; ++k) {
      double bv[4];
#ifdef VUS_EMU
      for (int b = 0; b < 4; ++b) bv[b] = pb[(long)k * ldb + b];
#else
      const double2 b01 = *reinterpret_cast<const double2*>(pb + (long)k * ldb);
      const double2 b23 = *reinterpret_cast<const double2*>(pb + (long)k * ldb + 2);
      bv[0] = b01.x; bv[1] = b01.y; bv[2] = b23.x; bv[3] = b23.y;
#endif
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const double av = pa[(long)a * B + k];
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] += av * bv[b];
      }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = i0 + a, j = j0 + b;
        if (i < B && j < B) {
          const long o = (long)i * B + j;
          const double v = alpha * acc[a][b] + (beta != 0.0 ? beta * C[o] : 0.0);
          C[o] = v;
          if (CT) CT[(long)j * B + i] = v;
        }
      }
  }
}
// stage a global BxB row-major block into shared memory with row stride ld, optionally transposed
VUS_DEV void cta_load(double* s, int ld, const double* g, int B, bool transpose, int tid, int nthr) {
  const int BB = B * B;
  if (!transpose) for (int e = tid; e < BB; e += nthr) { const int i = e / B, j = e - i * B; s[(long)i * ld + j] = g[e]; }
  else for (int e = tid; e < BB; e += nthr) { const int i = e / B, j = e - i * B; s[(long)j * ld + i] = g[e]; }
}

struct BcrArgs {
  long Ns; int B; long s;          // level stride
  double* Dw;                      // working diagonal blocks [Ns]
  const double* Ucur; double* Unext;   // couplings at this level / next level, indexed by node id
  double* Dinv; double* Gl; double* Gr; double* GlT; double* GrT;   // per eliminated node
  int* fail;
  // solve
  double* X; long xstride; int nrhs;
};
// per eliminated node j = s*(2m+1): Dinv_j, Gl_j = U[j-s] Dinv_j, Gr_j = U[j]^T Dinv_j  (+ transposed copies)
struct BcrElimBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int nthr, double* sm) {
    const int B = A.B, ldb = bcr_ldb(B);
    const long BB = (long)B * B;
    const long j = A.s * (2L * m + 1);
    double* sA = sm;                                   // "A" operand (row stride B)
    double* sX = sm + (long)bcr_rows_a(B) * B;         // Dinv as "B" operand (row stride ldb)
    double* rowp = sX + (long)B * ldb;
    double* colp = rowp + 2 * B;
    cta_load(sX, ldb, A.Dw + j * BB, B, false, tid, nthr);
    VUS_SYNC();
    cta_spd_inverse(sX, ldb, rowp, colp, B, tid, nthr, A.fail);
    for (long e = tid; e < BB; e += nthr) { const int i = (int)(e / B), c = (int)(e - (long)i * B); A.Dinv[j * BB + e] = sX[(long)i * ldb + c]; }
    cta_load(sA, B, A.Ucur + (j - A.s) * BB, B, false, tid, nthr);
    VUS_SYNC();
    cta_gemm_ss(A.Gl + j * BB, A.GlT + j * BB, sA, sX, B, 1.0, 0.0, tid, nthr);
    VUS_SYNC();
    if (j + A.s < A.Ns) {
      cta_load(sA, B, A.Ucur + j * BB, B, true, tid, nthr);
      VUS_SYNC();
      cta_gemm_ss(A.Gr + j * BB, A.GrT + j * BB, sA, sX, B, 1.0, 0.0, tid, nthr);
    }
  }
};
// per surviving node c = 2*m*s: Dw_c -= Gr_{c-s} U_{c-s} + Gl_{c+s} U_c^T ; Unext_c = -Gl_{c+s} U_{c+s}
struct BcrUpdateBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int nthr, double* sm) {
    const int B = A.B, ldb = bcr_ldb(B);
    const long BB = (long)B * B;
    const long c = 2L * m * A.s;
    double* sA = sm;
    double* sB = sm + (long)bcr_rows_a(B) * B;
    if (c - A.s >= 0) {
      const long j = c - A.s;
      cta_load(sA, B, A.Gr + j * BB, B, false, tid, nthr);
      cta_load(sB, ldb, A.Ucur + j * BB, B, false, tid, nthr);
      VUS_SYNC();
      cta_gemm_ss(A.Dw + c * BB, nullptr, sA, sB, B, -1.0, 1.0, tid, nthr);
      VUS_SYNC();
    }
    if (c + A.s < A.Ns) {
      const long j = c + A.s;
      cta_load(sA, B, A.Gl + j * BB, B, false, tid, nthr);
      cta_load(sB, ldb, A.Ucur + c * BB, B, true, tid, nthr);
      VUS_SYNC();
      cta_gemm_ss(A.Dw + c * BB, nullptr, sA, sB, B, -1.0, 1.0, tid, nthr);
      VUS_SYNC();
      if (j + A.s < A.Ns) {
        cta_load(sB, ldb, A.Ucur + j * BB, B, false, tid, nthr);
        VUS_SYNC();
        cta_gemm_ss(A.Unext + c * BB, nullptr, sA, sB, B, -1.0, 0.0, tid, nthr);
      }
    }
  }
};
// root: Dinv_0 = inv(Dw_0)
struct BcrRootBody {
  static VUS_DEV void run(const BcrArgs& A, int, int tid, int nthr, double* sm) {
    const int B = A.B, ldb = bcr_ldb(B);
    const long BB = (long)B * B;
    double* sX = sm + (long)bcr_rows_a(B) * B;
    double* rowp = sX + (long)B * ldb;
    double* colp = rowp + 2 * B;
    cta_load(sX, ldb, A.Dw, B, false, tid, nthr);
    VUS_SYNC();
    cta_spd_inverse(sX, ldb, rowp, colp, B, tid, nthr, A.fail);
    for (long e = tid; e < BB; e += nthr) { const int i = (int)(e / B), c = (int)(e - (long)i * B); A.Dinv[e] = sX[(long)i * ldb + c]; }
  }
};
// forward sweep, per surviving node c: b_c -= Gr_{c-s} b_{c-s} + Gl_{c+s} b_{c+s}   (reads the transposed copies: coalesced)
struct BcrFwdBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int nthr, double* sm) {
    const int B = A.B;
    const long BB = (long)B * B;
    const long c = 2L * m * A.s;
    double* xs = sm;   // [nrhs][B] neighbour rhs
    for (int side = 0; side < 2; ++side) {
      const long j = side == 0 ? c - A.s : c + A.s;
      if (j < 0 || j >= A.Ns) continue;
      const double* GT = (side == 0 ? A.GrT : A.GlT) + j * BB;
      for (int e = tid; e < A.nrhs * B; e += nthr) xs[e] = A.X[(long)(e / B) * A.xstride + j * B + (e % B)];
      VUS_SYNC();
      for (int e = tid; e < A.nrhs * B; e += nthr) {
        const int v = e / B, r = e % B;
        double s = 0.0;
        for (int k = 0; k < B; ++k) s += GT[(long)k * B + r] * xs[v * B + k];
        A.X[(long)v * A.xstride + c * B + r] -= s;
      }
      VUS_SYNC();
    }
  }
};
// backward sweep, per eliminated node j: x_j = Dinv_j b_j - Gl_j^T x_{j-s} - Gr_j^T x_{j+s}
struct BcrBwdBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int nthr, double* sm) {
    const int B = A.B;
    const long BB = (long)B * B;
    const long j = A.s * (2L * m + 1);
    double* xs = sm;                 // [nrhs][B]
    double* out = sm + A.nrhs * B;   // [nrhs][B]
    for (int e = tid; e < A.nrhs * B; e += nthr) xs[e] = A.X[(long)(e / B) * A.xstride + j * B + (e % B)];
    VUS_SYNC();
    for (int e = tid; e < A.nrhs * B; e += nthr) {
      const int v = e / B, r = e % B;
      double s = 0.0;
      const double* Di = A.Dinv + j * BB;
      for (int k = 0; k < B; ++k) s += Di[(long)k * B + r] * xs[v * B + k];     // symmetric
      out[e] = s;
    }
    VUS_SYNC();
    for (int side = 0; side < 2; ++side) {
      const long nb = side == 0 ? j - A.s : j + A.s;
      if (nb < 0 || nb >= A.Ns) continue;
      const double* G = (side == 0 ? A.Gl : A.Gr) + j * BB;
      for (int e = tid; e < A.nrhs * B; e += nthr) xs[e] = A.X[(long)(e / B) * A.xstride + nb * B + (e % B)];
      VUS_SYNC();
      for (int e = tid; e < A.nrhs * B; e += nthr) {
        const int v = e / B, r = e % B;
        double s = 0.0;
        for (int k = 0; k < B; ++k) s += G[(long)k * B + r] * xs[v * B + k];    // (G^T x)[r]
        out[e] -= s;
      }
      VUS_SYNC();
    }
    for (int e = tid; e < A.nrhs * B; e += nthr) A.X[(long)(e / B) * A.xstride + j * B + (e % B)] = out[e];
  }
};
// root solve: x_0 = Dinv_0 b_0
struct BcrRootSolveBody {
  static VUS_DEV void run(const BcrArgs& A, int, int tid, int nthr, double* sm) {
    const int B = A.B;
    double* xs = sm;
    for (int e = tid; e < A.nrhs * B; e += nthr) xs[e] = A.X[(long)(e / B) * A.xstride + (e % B)];
    VUS_SYNC();
    for (int e = tid; e < A.nrhs * B; e += nthr) {
      const int v = e / B, r = e % B;
      double s = 0.0;
      for (int k = 0; k < B; ++k) s += A.Dinv[(long)k * B + r] * xs[v * B + k];
      A.X[(long)v * A.xstride + r] = s;
    }
  }
};

// =====================================================================================
// small dense bias-border algebra (one thread)
// =====================================================================================
// Sb = Hbb - F^T Z (from partials of BorderDot over the 6 Z columns), SbInv = Sb^-1
struct BorderSchurArgs { const double* Hbb; const double* partials; int grid; int nv; double* SbInv; int* fail; };
struct BorderSchurBody {
  static VUS_DEV void run(const BorderSchurArgs& A, long) {
    double S[36];
    for (int e = 0; e < 36; ++e) S[e] = A.Hbb[e];
    // partial layout [grid][nv*6]: entry (v,c) = sum_i F[i][c] Z_v[i] = (F^T Z)[c][v]
    for (int v = 0; v < 6; ++v)
      for (int c = 0; c < 6; ++c) {
        double s = 0.0;
        for (int b = 0; b < A.grid; ++b) s += A.partials[(long)b * (A.nv * 6) + v * 6 + c];
        S[c * 6 + v] -= s;
      }
    for (int p = 0; p < 6; ++p) {            // Gauss-Jordan
      const double piv = S[p * 6 + p];
      if (!(piv > 0.0)) *A.fail = 1;
      const double d = 1.0 / piv;
      double rowp[6], colp[6];
      for (int i = 0; i < 6; ++i) { rowp[i] = S[p * 6 + i]; colp[i] = S[i * 6 + p]; }
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
          double v;
          if (i == p) v = (j == p) ? d : rowp[j] * d;
          else if (j == p) v = -colp[i] * d;
          else v = S[i * 6 + j] - colp[i] * rowp[j] * d;
          S[i * 6 + j] = v;
        }
    }
    for (int e = 0; e < 36; ++e) A.SbInv[e] = S[e];
  }
};
// xb = SbInv (rb - F^T y)   with F^T y from partials (nv = 1)
struct BorderSolveArgs { const double* SbInv; const double* rb; const double* partials; int grid; double* xb; };
struct BorderSolveBody {
  static VUS_DEV void run(const BorderSolveArgs& A, long) {
    double t[6];
    for (int c = 0; c < 6; ++c) {
      double s = 0.0;
      for (int b = 0; b < A.grid; ++b) s += A.partials[(long)b * 6 + c];
      t[c] = A.rb[c] - s;
    }
    for (int r = 0; r < 6; ++r) {
      double s = 0.0;
      for (int c = 0; c < 6; ++c) s += A.SbInv[r * 6 + c] * t[c];
      A.xb[r] = s;
    }
  }
};
// yb = F^T x (partials) + Hbb xb
struct BorderRowArgs { const double* Hbb; const double* xb; const double* partials; int grid; double* yb; };
struct BorderRowBody {
  static VUS_DEV void run(const BorderRowArgs& A, long) {
    for (int r = 0; r < 6; ++r) {
      double s = 0.0;
      for (int b = 0; b < A.grid; ++b) s += A.partials[(long)b * 6 + r];
      for (int c = 0; c < 6; ++c) s += A.Hbb[r * 6 + c] * A.xb[c];
      A.yb[r] = s;
    }
  }
};

// =====================================================================================
// vector kernels with device-resident scalars
// =====================================================================================
struct VecArgs { double* y; const double* x; const double* z; const double* scal; int slot; long n; const double* Z; const double* xb; long zstride; };
struct AxpyBody {       // y += scal[slot] * x
  static VUS_DEV void run(const VecArgs& A, long i) { A.y[i] += A.scal[A.slot] * A.x[i]; }
};
struct XpbyBody {       // y = x + scal[slot] * y
  static VUS_DEV void run(const VecArgs& A, long i) { A.y[i] = A.x[i] + A.scal[A.slot] * A.y[i]; }
};
struct SubZxbBody {     // y[i] = x[i] - sum_c Z_c[i] xb[c]
  static VUS_DEV void run(const VecArgs& A, long i) {
    double s = A.x[i];
#pragma unroll
    for (int c = 0; c < 6; ++c) s -= A.Z[(long)c * A.zstride + i] * A.xb[c];
    A.y[i] = s;
  }
};
// Z_c[i] = F[i][c]  (border columns as right-hand sides)
struct BorderColsArgs { const double* F; double* Z; long len; long zstride; };
struct BorderColsBody {
  static VUS_DEV void run(const BorderColsArgs& A, long w) {
    const long i = w % A.len;
    const int c = (int)(w / A.len);
    A.Z[(long)c * A.zstride + i] = A.F[i * 6 + c];
  }
};

// =====================================================================================
// Kernel 4: retract  x (+) delta  into the trial buffers
// =====================================================================================
struct RetractArgs {
  const double* pose; double* pose_out; long nx;
  const double* vel; double* vel_out; long nv;
  const double* bias; double* bias_out; long nb;
  const double* lm; double* lm_out; long nl;
  const double* xc; const double* xb; const double* xl; int D;
};
struct RetractBody {    // work items: nx poses, then nv velocities, then nl landmarks, then nb biases
  static VUS_DEV void run(const RetractArgs& A, long w) {
    if (w < A.nx) {
      double R[9], t[3], dR[9], dt[3], xi[6], Rn[9], tn[3];
      load_pose(A.pose, A.nx, w, R, t);
#pragma unroll
      for (int c = 0; c < 6; ++c) xi[c] = A.xc[w * A.D + c];
      pose_exp(xi, dR, dt);
      m3_mul(R, dR, Rn);
      m3_vec(R, dt, tn);
#pragma unroll
      for (int c = 0; c < 9; ++c) A.pose_out[c * A.nx + w] = Rn[c];
#pragma unroll
      for (int c = 0; c < 3; ++c) A.pose_out[(9 + c) * A.nx + w] = t[c] + tn[c];
      return;
    }
    w -= A.nx;
    if (w < A.nv) {
#pragma unroll
      for (int c = 0; c < 3; ++c) A.vel_out[c * A.nv + w] = A.vel[c * A.nv + w] + A.xc[w * A.D + 6 + c];
      return;
    }
    w -= A.nv;
    if (w < A.nl) {
#pragma unroll
      for (int c = 0; c < 3; ++c) A.lm_out[c * A.nl + w] = A.lm[c * A.nl + w] + A.xl[c * A.nl + w];
      return;
    }
    w -= A.nl;
#pragma unroll
    for (int c = 0; c < 6; ++c) A.bias_out[c * A.nb + w] = A.bias[c * A.nb + w] + A.xb[c];
  }
};

}  // namespace vus
