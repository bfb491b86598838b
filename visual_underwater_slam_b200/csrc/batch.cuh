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
    for (int e = 0; e < 36; ++e) S[e] = A.Hbb[c * 36 + e];
    for (int v = 0; v < 6; ++v)
      for (int k = 0; k < 6; ++k) S[k * 6 + v] -= A.ftz[c * 36 + v * 6 + k];
    for (int e = 0; e < 36; ++e) {
      const int a = e / 6, k = e - a * 6;
      S[a * 6 + k] -= 0.5 * A.ztr[c * 36 + e];
      S[k * 6 + a] -= 0.5 * A.ztr[c * 36 + e];
    }
    for (int p = 0; p < 6; ++p) {
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
    for (int e = 0; e < 36; ++e) A.SbInv[c * 36 + e] = S[e];
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
