// Kernel 3c -- the band preconditioner of a pose graph that is split across ranks by pose range (BASELINE config 5).
//
// Every rank factors the band of the poses it owns (bcr.cuh).  Used alone that is block-Jacobi ACROSS ranks: the couplings
// between the last poses of rank r and the first poses of rank r + 1 are dropped, a cut-off piece of the chain has no anchor
// (the prior sits on pose 0), PCG stalls and the LM path leaves the one-rank path (measured: 100 LM iterations instead of 4
// on a 20 000-pose chain over 8 ranks).  Here the rank bands are tied together exactly, the way chunk.cuh ties the chunks
// of one GPU together -- a partitioned ("spike") solve of the global block-tridiagonal band M.  With k poses per supernode
// (K = 6 k dofs), t_r / b_r the first / last K dofs of rank r's part x_r of the solution:
//
//     M_r x_r + E_last A(last k poses of r, first k of r + 1) t_{r+1} + E_first A(first k of r, last k of r - 1) b_{r-1} = f_r
//     V_r = M_r^-1 E_last A(last, next),   W_r = M_r^-1 E_first A(first, prev)            (set-up: 2 K band-solve columns)
//     x_r = y_r - V_r t_{r+1} - W_r b_{r-1},   y_r = M_r^-1 f_r                            (every application)
//
// Restricting the last line to the first / last K dofs of every rank gives a dense system in the 2 K P unknowns (t_r, b_r):
//     t_r + Vf_r t_{r+1} + Wf_r b_{r-1} = yf_r ,      b_r + Vl_r t_{r+1} + Wl_r b_{r-1} = yl_r
// whose four K x K blocks per rank are shared by ONE all-reduce of a zero-padded table at set-up and inverted by every rank
// (2 K P <= 384: one CTA); an application shares the 2 K numbers (yf_r, yl_r) of every rank by one more all-reduce.  The
// result is the band solve of the undivided chain up to rounding, so the partitioned PCG takes the one-rank iterations.
#pragma once
#include "kernels.cuh"

namespace vus {

struct SpikeArgs {
  int P, rank, k, K;           // ranks, this rank, poses per supernode, K = 6 k
  long Ls;                     // column stride of Z (rows the band solve covers)
  long n_owned;
  const PairDst* pairs;        // [2][k][k]: A(last pose p, next pose j) | A(first pose p, prev pose jj); off < 0: no such block
  const double* Hval;
  double* Z;                   // [2 K][Ls]: V (K columns) | W (K columns)
  double* G;                   // [P][4][K K]: Vf, Vl, Wf, Wl of every rank
  double* R;                   // [2 K P][2 K P] scratch
  double* Rinv;                // [2 K P][2 K P]
  double* Y;                   // [P][2 K]: (yf, yl) of every rank
  double* coef;                // [2 K]: t_{rank + 1} | b_{rank - 1}
  double* z;                   // the vector being preconditioned
  int* fail;
};
VUS_DEV double spike_blk(const double* H, const PairDst& d, int a, int b) {
  return d.transposed ? H[d.off + (long)b * d.ld + a] : H[d.off + (long)a * d.ld + b];
}
// right-hand sides of the 2 K spike columns (Z zeroed before); w < 2 K * K:  (column, pose p of the interface supernode, dof i)
struct SpikeRhsBody {
  static VUS_DEV void run(const SpikeArgs& A, long w) {
    const int k = A.k, K = A.K;
    const int col = (int)(w / K), row = (int)(w % K), p = row / 6, i = row % 6;
    const int side = col >= K, cc = side ? col - K : col, j = cc / 6, c = cc % 6;
    const PairDst d = A.pairs[(side * k + p) * k + j];
    if (d.off < 0) return;
    const long node = side ? p : A.n_owned - k + p;
    A.Z[(long)col * A.Ls + node * 6 + i] = spike_blk(A.Hval, d, i, c);
  }
};
// this rank's four blocks into its slot of G (zeroed before); w < 4 K K:  block 0 Vf, 1 Vl, 2 Wf, 3 Wl
struct SpikeGatherBody {
  static VUS_DEV void run(const SpikeArgs& A, long w) {
    const int K = A.K, KK = K * K;
    const int b = (int)(w / KK), i = (int)((w % KK) / K), c = (int)(w % K);
    const long row0 = (b & 1) ? (A.n_owned - A.k) * 6 : 0;
    const int col = b < 2 ? c : K + c;
    A.G[((long)A.rank * 4 + b) * KK + i * K + c] = A.Z[(long)col * A.Ls + row0 + i];
  }
};
// reduced matrix from the all-reduced G, inverted in place by Gauss-Jordan with partial pivoting; one CTA, sm >= 2 K P + 2 doubles
struct SpikeInvertBody {
  static VUS_DEV void run(const SpikeArgs& A, int, int tid, int nthr, double* sm) {
    const int K = A.K, K2 = 2 * K, KK = K * K, n = K2 * A.P;
    for (int e = tid; e < n * n; e += nthr) {
      const int row = e / n, col = e - row * n;
      const int r = row / K2, ri = row % K2;           // ri < K: t_r equation, else b_r equation
      const int q = col / K2, qi = col % K2;
      double v = row == col ? 1.0 : 0.0;
      if (q == r + 1 && qi < K) v += A.G[((long)r * 4 + (ri < K ? 0 : 1)) * KK + (ri % K) * K + qi];
      if (q == r - 1 && qi >= K) v += A.G[((long)r * 4 + (ri < K ? 2 : 3)) * KK + (ri % K) * K + (qi - K)];
      A.R[e] = v;
      A.Rinv[e] = row == col ? 1.0 : 0.0;
    }
    VUS_SYNC();
    double* fac = sm;
    for (int col = 0; col < n; ++col) {
      if (tid == 0) {
        int p = col;
        double best = A.R[col * n + col] < 0 ? -A.R[col * n + col] : A.R[col * n + col];
        for (int row = col + 1; row < n; ++row) {
          const double v = A.R[row * n + col] < 0 ? -A.R[row * n + col] : A.R[row * n + col];
          if (v > best) { best = v; p = row; }
        }
        if (!(best > 0.0)) { *A.fail = 1; A.R[p * n + col] = 1.0; }
        sm[n] = (double)p;
      }
      VUS_SYNC();
      const int p = (int)sm[n];
      if (p != col) {
        for (int j = tid; j < 2 * n; j += nthr) {
          double* M = j < n ? A.R : A.Rinv;
          const int jj = j < n ? j : j - n;
          const double t = M[col * n + jj]; M[col * n + jj] = M[p * n + jj]; M[p * n + jj] = t;
        }
      }
      VUS_SYNC();
      if (tid == 0) sm[n + 1] = 1.0 / A.R[col * n + col];
      for (int row = tid; row < n; row += nthr) fac[row] = row == col ? 0.0 : A.R[row * n + col];
      VUS_SYNC();
      const double inv = sm[n + 1];
      for (int j = tid; j < 2 * n; j += nthr) {
        double* M = j < n ? A.R : A.Rinv;
        const int jj = j < n ? j : j - n;
        M[col * n + jj] *= inv;
      }
      VUS_SYNC();
      for (int e = tid; e < 2 * n * n; e += nthr) {
        const int row = e / (2 * n), j = e - row * 2 * n;
        if (row == col) continue;
        double* M = j < n ? A.R : A.Rinv;
        const int jj = j < n ? j : j - n;
        M[row * n + jj] -= fac[row] * M[col * n + jj];
      }
      VUS_SYNC();
    }
  }
};
// this rank's (yf, yl) into its slot of Y (zeroed before); w < 2 K
struct SpikeYBody {
  static VUS_DEV void run(const SpikeArgs& A, long w) {
    A.Y[(long)A.rank * 2 * A.K + w] = w < A.K ? A.z[w] : A.z[(A.n_owned - A.k) * 6 + (w - A.K)];
  }
};
// the two interface blocks this rank's correction needs: coef[0 .. K) = t_{rank+1}, coef[K .. 2 K) = b_{rank-1}; w < 2 K
struct SpikeCoefBody {
  static VUS_DEV void run(const SpikeArgs& A, long w) {
    const int K = A.K, K2 = 2 * K, n = K2 * A.P;
    long row = -1;
    if (w < K) { if (A.rank + 1 < A.P) row = (long)K2 * (A.rank + 1) + w; }
    else if (A.rank > 0) row = (long)K2 * (A.rank - 1) + w;            // K + (w - K)
    double s = 0.0;
    if (row >= 0) for (int m = 0; m < n; ++m) s += A.Rinv[row * n + m] * A.Y[m];
    A.coef[w] = s;
  }
};
// z_i -= V_i t_next + W_i b_prev over the owned dofs
struct SpikeCorrBody {
  static VUS_DEV void run(const SpikeArgs& A, long i) {
    double s = 0.0;
    for (int c = 0; c < 2 * A.K; ++c) s += A.Z[(long)c * A.Ls + i] * A.coef[c];
    A.z[i] -= s;
  }
};

}  // namespace vus
