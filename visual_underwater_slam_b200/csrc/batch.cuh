// Batched mode: many INDEPENDENT trajectories in one handle (BASELINE.json config 4: 4096 x 500-pose graphs, a block of
// them per GPU).  The reference solves one graph per gtsam.LevenbergMarquardtOptimizer (/root/reference/batch.py:337);
// a 500-pose graph cannot fill a B200 (its solve is launch-latency bound), so the trajectories of a shard are
// concatenated into ONE block-diagonal system -- component c owns the contiguous node range [node_start[c],
// node_start[c+1]) and its own bias B(c) -- and every quantity gtsam keeps per optimizer is kept per component on the
// device: lambda, graph error, linearised error, the PCG scalars (alpha, beta, residual norms) and the 6 x 6 bias
// border.  Linearization, assembly, the band factorization / solve and the band operator are the single-graph kernels
// (a block-diagonal band is a band); the kernels here are the per-component reductions and updates around them.
// One CTA per component for the reductions (a component is a few thousand dofs), elementwise kernels look the
// component of a dof up in node_comp.
#pragma once

namespace vus {

// per-component scalar block: the single-graph slots (kernels.cuh) plus the convergence bookkeeping
enum { SB_RR0 = 9, SB_TOL2 = 10, SB_STRIDE = 16 };
enum { BOP_STORE = 0, BOP_PAP = 1, BOP_RZ = 2, BOP_RZ0 = 3, BOP_RR0 = 4 };

struct BCtx {                       // where the components live
  const int* node_comp;             // [Npad]
  const long* seg;                  // [ncomp + 1] first camera dof of every component (node_start * D)
  int ncomp, D, has_bias;
  long Lc;
};
VUS_DEV int comp_of_dof(const BCtx& C, long i) { return i < C.Lc ? C.node_comp[i / C.D] : (int)((i - C.Lc) / 6); }

// block sum of one value per thread -> sm[0]
VUS_DEV void cta_sum1(double v, int tid, int nthr, double* sm) {
  sm[tid] = v;
  VUS_SYNC();
  for (int s = nthr >> 1; s > 0; s >>= 1) {
    for (int t = tid; t < s; t += nthr) sm[t] += sm[t + s];
    VUS_SYNC();
  }
}

// ---- dot products with the PCG post-op, one CTA per component (nthr a power of two)
struct BDotArgs { BCtx C; const double* a; const double* b; double* scal; int slot, op; double tol; };
struct BDotBody {
  static VUS_DEV void run(const BDotArgs& A, int c, int tid, int nthr, double* sm) {
    double acc = 0.0;
    for (long i = A.C.seg[c] + tid; i < A.C.seg[c + 1]; i += nthr) acc += A.a[i] * A.b[i];
    if (A.C.has_bias)
      for (int j = tid; j < 6; j += nthr) acc += A.a[A.C.Lc + 6 * c + j] * A.b[A.C.Lc + 6 * c + j];
    cta_sum1(acc, tid, nthr, sm);
    if (tid != 0) return;
    const double v = sm[0];
    double* s = A.scal + (long)c * SB_STRIDE;
    const bool live = s[S_RR] > s[SB_TOL2];          // a component whose residual met the tolerance stops moving
    if (A.op == BOP_STORE) s[A.slot] = v;
    else if (A.op == BOP_RR0) { s[S_RR] = v; s[SB_RR0] = v; s[SB_TOL2] = A.tol * A.tol * v; }
    else if (A.op == BOP_PAP) { s[S_PAP] = v; const double al = (live && v != 0.0) ? s[S_RZ] / v : 0.0; s[S_ALPHA] = al; s[S_NEG_ALPHA] = -al; }
    else if (A.op == BOP_RZ) { const double old = s[S_RZ]; s[S_BETA] = (live && old != 0.0) ? v / old : 0.0; s[S_RZ] = v; }
    else if (A.op == BOP_RZ0) { s[S_RZ] = v; s[S_BETA] = 0.0; }
  }
};
// y += scal[comp][slot] * x      /      y = x + scal[comp][slot] * y
struct BVecArgs { BCtx C; double* y; const double* x; const double* scal; int slot; const double* Z; const double* xb; long zstride; };
struct BAxpyBody {
  static VUS_DEV void run(const BVecArgs& A, long i) { A.y[i] += A.scal[(long)comp_of_dof(A.C, i) * SB_STRIDE + A.slot] * A.x[i]; }
};
struct BXpbyBody {
  static VUS_DEV void run(const BVecArgs& A, long i) { A.y[i] = A.x[i] + A.scal[(long)comp_of_dof(A.C, i) * SB_STRIDE + A.slot] * A.y[i]; }
};
// y[i] = x[i] - sum_k Z_k[i] xb[comp(i)][k]      (camera dofs)
struct BSubZxbBody {
  static VUS_DEV void run(const BVecArgs& A, long i) {
    const double* xb = A.xb + 6 * (long)A.C.node_comp[i / A.C.D];
    double s = A.x[i];
#pragma unroll
    for (int k = 0; k < 6; ++k) s -= A.Z[(long)k * A.zstride + i] * xb[k];
    A.y[i] = s;
  }
};

// zero the entries of the components whose factorization failed (mask[c] != 0): with a zero right-hand side their PCG
// recursion never starts (rr0 = 0), so a bad pivot in one trajectory leaves every other trajectory's solve untouched
struct BZeroCompArgs { BCtx C; const int* mask; double* v; };
struct BZeroCompBody {
  static VUS_DEV void run(const BZeroCompArgs& A, long i) { if (A.mask[comp_of_dof(A.C, i)]) A.v[i] = 0.0; }
};

// ---- damping: lambda of the dof's component on the real dofs, identity on the padding, lambda on every bias diagonal
struct BDampArgs { BCtx C; double* SD; double* Hbb; long nreal; int B; const double* lam; int ld; long bs; };
struct BDampBody {
  static VUS_DEV void run(const BDampArgs& A, long i) {
    if (i < A.C.Lc) {
      const long I = i / A.B;
      const int r = (int)(i % A.B);
      A.SD[I * A.bs + (long)r * A.ld + r] += (i < A.nreal) ? A.lam[A.C.node_comp[i / A.C.D]] : 1.0;
    } else {
      const long e = i - A.C.Lc;
      const int c = (int)(e / 6), j = (int)(e % 6);
      A.Hbb[(long)c * 36 + j * 7] += A.lam[c];
    }
  }
};

// ---- bias block of every component: Hbb_c = sum Jb^T Jb, gb_c = -sum Jb^T r over the component's IMU factors
struct BImuBiasArgs { long n; const double* J; const double* r; const int* ptr; const int* list; double* Hbb; double* gb; };
struct BImuBiasBody {
  static VUS_DEV void run(const BImuBiasArgs& A, int c, int tid, int nthr, double* sm) {
    for (int e = 0; e < 42; ++e) {
      double acc = 0.0;
      if (e < 36) {
        const int a = 18 + e / 6, b = 18 + e % 6;
        for (int q = A.ptr[c] + tid; q < A.ptr[c + 1]; q += nthr) {
          const long f = A.list[q];
          for (int r = 0; r < 9; ++r) acc += A.J[(r * 24 + a) * A.n + f] * A.J[(r * 24 + b) * A.n + f];
        }
      } else {
        const int a = 18 + e - 36;
        for (int q = A.ptr[c] + tid; q < A.ptr[c + 1]; q += nthr) {
          const long f = A.list[q];
          for (int r = 0; r < 9; ++r) acc -= A.J[(r * 24 + a) * A.n + f] * A.r[r * A.n + f];
        }
      }
      cta_sum1(acc, tid, nthr, sm);
      if (tid == 0) { if (e < 36) A.Hbb[(long)c * 36 + e] = sm[0]; else A.gb[(long)c * 6 + e - 36] = sm[0]; }
      VUS_SYNC();
    }
  }
};

// ---- border products per component:  out[c][v*6 + k] = sum_{i in c} F[i][k] Y_v[i]   (needs 6 * nthr doubles)
struct BBorderDotArgs { BCtx C; const double* F; const double* Y; long ystride; int nv; double* out; };
struct BBorderDotBody {
  static VUS_DEV void run(const BBorderDotArgs& A, int c, int tid, int nthr, double* sm) {
    double acc[6][6];
    for (int v = 0; v < 6; ++v)
      for (int k = 0; k < 6; ++k) acc[v][k] = 0.0;
    for (long i = A.C.seg[c] + tid; i < A.C.seg[c + 1]; i += nthr) {
      double f[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) f[k] = A.F[i * 6 + k];
#pragma unroll
      for (int v = 0; v < 6; ++v)
        if (v < A.nv) {
          const double y = A.Y[(long)v * A.ystride + i];
#pragma unroll
          for (int k = 0; k < 6; ++k) acc[v][k] += f[k] * y;
        }
    }
#pragma unroll
    for (int v = 0; v < 6; ++v)
      if (v < A.nv) cta_sum6(acc[v], A.out + (long)c * (A.nv * 6) + v * 6, tid, nthr, sm);
  }
};
// out[c][a*6 + b] = sum_{i in c} Z_a[i] R_b[i]
struct BColDotArgs { BCtx C; const double* Z; const double* R; long stride; double* out; };
struct BColDotBody {
  static VUS_DEV void run(const BColDotArgs& A, int c, int tid, int nthr, double* sm) {
    double acc[6][6];
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) acc[a][b] = 0.0;
    for (long i = A.C.seg[c] + tid; i < A.C.seg[c + 1]; i += nthr) {
      double z[6], r[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) { z[k] = A.Z[(long)k * A.stride + i]; r[k] = A.R[(long)k * A.stride + i]; }
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) acc[a][b] += z[a] * r[b];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) cta_sum6(acc[a], A.out + (long)c * 36 + a * 6, tid, nthr, sm);
  }
};
// SbInv_c = (Hbb_c - F_c^T Z_c - sym(Z_c^T R_c))^-1, one thread per component (6 x 6 Gauss-Jordan, as BorderSchurBody)
struct BBorderSchurArgs { int ncomp; const double* Hbb; const double* ftz; const double* ztr; double* SbInv; int* fail; };
struct BBorderSchurBody {
  static VUS_DEV void run(const BBorderSchurArgs& A, long c) {
    double S[36];
    bool failed = false;                                 // a failed complement is returned as zero: the component is frozen anyway
    for (int e = 0; e < 36; ++e) S[e] = A.Hbb[c * 36 + e];
    for (int v = 0; v < 6; ++v)
      for (int k = 0; k < 6; ++k) S[k * 6 + v] -= A.ftz[c * 36 + v * 6 + k];
    for (int e = 0; e < 36; ++e) {
      const int a = e / 6, k = e - a * 6;
      S[a * 6 + k] -= 0.5 * A.ztr[c * 36 + e];
      S[k * 6 + a] -= 0.5 * A.ztr[c * 36 + e];
    }
    for (int p = 0; p < 6; ++p) {
      double piv = S[p * 6 + p];
      if (!(piv > 0.0)) { A.fail[c] = 1; failed = true; piv = 1.0; }
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
    for (int e = 0; e < 36; ++e) A.SbInv[c * 36 + e] = failed ? 0.0 : S[e];
  }
};
// work item (c, r):  xb[c][r] = sum_k SbInv_c[r][k] (rb[c][k] - fty[c][k])
struct BBorderSmallArgs { const double* M; const double* rb; const double* dots; double* out; };
struct BBorderSolveBody {
  static VUS_DEV void run(const BBorderSmallArgs& A, long w) {
    const long c = w / 6;
    const int r = (int)(w % 6);
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s += A.M[c * 36 + r * 6 + k] * (A.rb[c * 6 + k] - A.dots[c * 6 + k]);
    A.out[w] = s;
  }
};
// work item (c, r):  yb[c][r] = ftx[c][r] + sum_k Hbb_c[r][k] xb[c][k]
struct BBorderRowBody {
  static VUS_DEV void run(const BBorderSmallArgs& A, long w) {
    const long c = w / 6;
    const int r = (int)(w % 6);
    double s = A.dots[w];
    for (int k = 0; k < 6; ++k) s += A.M[c * 36 + r * 6 + k] * A.rb[c * 6 + k];
    A.out[w] = s;
  }
};
// per (node, row): off-band remainder blocks + the border column of the node's component  (RemBorderMatvecBody)
struct BRemBorderArgs { BCtx C; const int* rem_ptr; const int* rem_col; const double* rem_val; const double* F; const double* x; const double* xb; double* y; };
struct BRemBorderBody {
  static VUS_DEV void run(const BRemBorderArgs& A, long w) {
    const int D = A.C.D;
    const long node = w / D;
    const int r = (int)(w % D);
    double s = 0.0;
    if (A.rem_ptr) {
      for (int t = A.rem_ptr[node]; t < A.rem_ptr[node + 1]; ++t) {
        const long col = A.rem_col[t];
        const double* v = A.rem_val + (long)t * D * D + r * D;
        for (int k = 0; k < D; ++k) s += v[k] * A.x[col * D + k];
      }
    }
    if (A.C.has_bias) {
      const double* xb = A.xb + 6 * (long)A.C.node_comp[node];
#pragma unroll
      for (int k = 0; k < 6; ++k) s += A.F[w * 6 + k] * xb[k];
    }
    A.y[w] += s;
  }
};

// =====================================================================================  loop closures by capacitance
// A trajectory of config 4 has a handful of loop closures.  Their off-band blocks W = U S U^T (U selects the 2 x 6 pose
// dofs of every closure, S = [[0, H_pq], [H_qp, 0]] per closure) are a rank-12 update per closure of the exactly
// factored part P (band + bias border), and plain PCG needs ~rank iterations per solve -- ~100 for the slowest of 512
// components.  With R = 12 * (max closures per component) <= 96 columns the update is inverted exactly instead
// (Woodbury):  A^-1 = P^-1 - Y (I + S K)^-1 S U^T P^-1,  Y = P^-1 U,  K = U^T Y.  The R columns of Y are R / 6
// six-vector applications of P^-1 for the WHOLE batch (column j of every component rides in the same vector), the
// R x R capacitance matrix of a component is inverted by one CTA, and PCG on top needs 2-3 iterations of refinement.
struct BWbCtx {
  const int* node;      // [ncomp][2 * lmax]  end nodes (p, q) of closure l of the component, -1: unused slot
  const long* blk;      // [ncomp][lmax]      offset of the REM block (p, q) in Hval
  int lmax, R;
};
// unit right-hand sides: work item (c, v) sets column j0 + v of component c
struct BWbUnitArgs { BCtx C; BWbCtx W; double* Y; long L; int j0; };
struct BWbUnitBody {
  static VUS_DEV void run(const BWbUnitArgs& A, long w) {
    const int v = (int)(w % 6);
    const long c = w / 6;
    const int col = A.j0 + v, l = col / 12, side = (col % 12) / 6, a = col % 6;
    const int node = A.W.node[(c * A.W.lmax + l) * 2 + side];
    if (node >= 0) A.Y[(long)v * A.L + (long)node * A.C.D + a] = 1.0;
  }
};
// multi-vector border part of P^-1 for right-hand sides with a ZERO bias part:  xb[c][v] = -SbInv_c (F_c^T y_v)
struct BBorderMultiArgs { BCtx C; const double* SbInv; const double* dots; double* Y; long L; int nv; const double* Z; long zstride; };
struct BBorderSolveMultiBody {      // work item (c, v, r)
  static VUS_DEV void run(const BBorderMultiArgs& A, long w) {
    const int r = (int)(w % 6), v = (int)((w / 6) % A.nv);
    const long c = w / (6 * A.nv);
    double s = 0.0;
    for (int k = 0; k < 6; ++k) s -= A.SbInv[c * 36 + r * 6 + k] * A.dots[c * (A.nv * 6) + v * 6 + k];
    A.Y[(long)v * A.L + A.C.Lc + 6 * c + r] = s;
  }
};
struct BSubZxbMultiBody {           // work item (v, i):  y_v[i] -= sum_k Z_k[i] xb[comp(i)][v][k]
  static VUS_DEV void run(const BBorderMultiArgs& A, long w) {
    const long i = w % A.C.Lc;
    const int v = (int)(w / A.C.Lc);
    const double* xb = A.Y + (long)v * A.L + A.C.Lc + 6 * (long)A.C.node_comp[i / A.C.D];
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s += A.Z[(long)k * A.zstride + i] * xb[k];
    A.Y[(long)v * A.L + i] -= s;
  }
};
// dof (in the camera vector) of capacitance row m of component c, -1 for an unused slot
VUS_DEV long wb_dof(const BWbCtx& W, int D, long c, int m) {
  const int l = m / 12, side = (m % 12) / 6, a = m % 6;
  const int node = W.node[(c * W.lmax + l) * 2 + side];
  return node < 0 ? -1 : (long)node * D + a;
}
// (S t)[i] for row i = (l, side, a): side p -> sum_b H_pq[a][b] t[(l, q, b)] ; side q -> sum_b H_pq[b][a] t[(l, p, b)]
VUS_DEV double wb_S_row(const BWbCtx& W, int D, const double* Hval, long c, int i, const double* t, int tstride) {
  const int l = i / 12, side = (i % 12) / 6, a = i % 6;
  if (W.node[(c * W.lmax + l) * 2] < 0) return 0.0;
  const double* H = Hval + W.blk[c * W.lmax + l];
  double s = 0.0;
  for (int b = 0; b < 6; ++b)
    s += (side == 0 ? H[a * D + b] * t[(long)(l * 12 + 6 + b) * tstride] : H[b * D + a] * t[(long)(l * 12 + b) * tstride]);
  return s;
}
// capacitance matrix of every component, inverted in shared memory: CapInv_c = (I + S_c K_c)^-1, K_c[m][j] = Y_j[dof(c, m)].
// Gauss-Jordan with partial pivoting on the augmented [Cap | I] (R x 2R doubles of shared memory + R for K gathering is
// avoided by building Cap row by row from global Y).  One CTA per component.
struct BWbCapArgs { BCtx C; BWbCtx W; const double* Y; long L; const double* Hval; double* CapInv; int* fail; };
struct BWbCapBody {
  static VUS_DEV void run(const BWbCapArgs& A, int c, int tid, int nthr, double* sm) {
    const int R = A.W.R, R2 = 2 * R, D = A.C.D;
    double* M = sm;                   // [R][2R]
    double* aux = sm + (long)R * R2;  // [R] multipliers | [1] pivot row
    // K into the right half first (as scratch), then Cap = I + S K into the left half
    for (int e = tid; e < R * R; e += nthr) {
      const int m = e / R, j = e % R;
      const long d = wb_dof(A.W, D, c, m);
      M[m * R2 + R + j] = d < 0 ? 0.0 : A.Y[(long)j * A.L + d];
    }
    VUS_SYNC();
    for (int e = tid; e < R * R; e += nthr) {
      const int i = e / R, j = e % R;
      M[i * R2 + j] = (i == j ? 1.0 : 0.0) + wb_S_row(A.W, D, A.Hval, c, i, M + R + j, R2);
    }
    VUS_SYNC();
    for (int e = tid; e < R * R; e += nthr) { const int i = e / R, j = e % R; M[i * R2 + R + j] = i == j ? 1.0 : 0.0; }
    VUS_SYNC();
    for (int p = 0; p < R; ++p) {
      if (tid == 0) {                                   // partial pivoting: largest |entry| of column p at or below the diagonal
        int best = p;
        double bv = fabs(M[p * R2 + p]);
        for (int i = p + 1; i < R; ++i) { const double v = fabs(M[i * R2 + p]); if (v > bv) { bv = v; best = i; } }
        if (!(bv > 0.0)) { A.fail[c] = 1; M[best * R2 + p] = 1.0; }
        aux[R] = (double)best;
      }
      VUS_SYNC();
      const int best = (int)aux[R];
      if (best != p)
        for (int j = tid; j < R2; j += nthr) { const double v = M[p * R2 + j]; M[p * R2 + j] = M[best * R2 + j]; M[best * R2 + j] = v; }
      VUS_SYNC();
      for (int i = tid; i < R; i += nthr) aux[i] = M[i * R2 + p];
      VUS_SYNC();
      const double d = 1.0 / aux[p];
      for (int j = tid; j < R2; j += nthr) M[p * R2 + j] *= d;
      VUS_SYNC();
      for (int e = tid; e < R * R2; e += nthr) {
        const int i = e / R2, j = e % R2;
        if (i != p) M[i * R2 + j] -= aux[i] * M[p * R2 + j];
      }
      VUS_SYNC();
    }
    for (int e = tid; e < R * R; e += nthr) { const int i = e / R, j = e % R; A.CapInv[((long)c * R + i) * R + j] = M[i * R2 + R + j]; }
  }
};
// w_c = CapInv_c S_c U^T z   (one CTA per component; 2R doubles of shared memory)
struct BWbSmallArgs { BCtx C; BWbCtx W; const double* z; const double* Hval; const double* CapInv; double* wout; };
struct BWbSmallBody {
  static VUS_DEV void run(const BWbSmallArgs& A, int c, int tid, int nthr, double* sm) {
    const int R = A.W.R, D = A.C.D;
    double* t = sm;
    double* st = sm + R;
    for (int m = tid; m < R; m += nthr) { const long d = wb_dof(A.W, D, c, m); t[m] = d < 0 ? 0.0 : A.z[d]; }
    VUS_SYNC();
    for (int i = tid; i < R; i += nthr) st[i] = wb_S_row(A.W, D, A.Hval, c, i, t, 1);
    VUS_SYNC();
    for (int j = tid; j < R; j += nthr) {
      const double* row = A.CapInv + ((long)c * R + j) * R;
      double s = 0.0;
      for (int i = 0; i < R; ++i) s += row[i] * st[i];
      A.wout[(long)c * R + j] = s;
    }
  }
};
// z[i] -= sum_j Y_j[i] w[comp(i)][j]     (all dofs, camera and bias)
struct BWbSubArgs { BCtx C; const double* Y; long L; const double* w; int R; double* z; };
struct BWbSubBody {
  static VUS_DEV void run(const BWbSubArgs& A, long i) {
    const double* w = A.w + (long)comp_of_dof(A.C, i) * A.R;
    double s = 0.0;
    for (int j = 0; j < A.R; ++j) s += A.Y[(long)j * A.L + i] * w[j];
    A.z[i] -= s;
  }
};

// ---- per-component sums of per-factor values (graph error, linearised error): out[c] = sum_{q in c} e[list[q]]
struct BErrSumArgs { const double* e; const int* ptr; const int* list; double* out; };
struct BErrSumBody {
  static VUS_DEV void run(const BErrSumArgs& A, int c, int tid, int nthr, double* sm) {
    double acc = 0.0;
    for (int q = A.ptr[c] + tid; q < A.ptr[c + 1]; q += nthr) acc += A.e[A.list[q]];
    cta_sum1(acc, tid, nthr, sm);
    if (tid == 0) A.out[c] = sm[0];
  }
};
// ---- accepted components take the trial values (work items: nx poses, nv velocities, nb biases)
struct BCommitArgs { const int* node_comp; const double* accept; long nx, nv, nb;
                     double* pose; const double* pose_t; double* vel; const double* vel_t; double* bias; const double* bias_t; };
struct BCommitBody {
  static VUS_DEV void run(const BCommitArgs& A, long w) {
    if (w < A.nx) {
      if (A.accept[A.node_comp[w]] != 0.0)
        for (int k = 0; k < 12; ++k) A.pose[k * A.nx + w] = A.pose_t[k * A.nx + w];
      return;
    }
    w -= A.nx;
    if (w < A.nv) {
      if (A.accept[A.node_comp[w]] != 0.0)
        for (int k = 0; k < 3; ++k) A.vel[k * A.nv + w] = A.vel_t[k * A.nv + w];
      return;
    }
    w -= A.nv;
    if (A.accept[w] != 0.0)
      for (int k = 0; k < 6; ++k) A.bias[k * A.nb + w] = A.bias_t[k * A.nb + w];
  }
};

}  // namespace vus
