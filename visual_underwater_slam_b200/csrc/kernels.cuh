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
  int opts;        // VUS_OPT_* bits: the gtsam build switches that change the arithmetic (vus_set_gtsam_build)
};
enum { VUS_OPT_TANGENT = 1, VUS_OPT_SLOW_BETWEEN = 2 };

template <int TYPE, bool WJ>
struct LinBody {
  static VUS_DEV void run(const LinArgs& a, long f) {
    if (TYPE == VUS_F_PRIOR_POSE) f_prior_pose<WJ>(a.V, a.F, a.O, f);
    else if (TYPE == VUS_F_PRIOR_VEL) f_prior_vel<WJ>(a.V, a.F, a.O, f);
    else if (TYPE == VUS_F_BETWEEN) f_between<WJ>(a.V, a.F, a.O, f, (a.opts & VUS_OPT_SLOW_BETWEEN) != 0);
    else if (TYPE == VUS_F_DVL) f_dvl<WJ>(a.V, a.F, a.O, f);
    else if (TYPE == VUS_F_STEREO) f_stereo<WJ>(a.V, a.F, a.O, f, a.K);
    else f_imu<WJ>(a.V, a.F, a.O, f, a.g, (a.opts & VUS_OPT_TANGENT) != 0);
  }
};

// Stereo linearize with the fused assembly products: one CTA per 128 consecutive observations.  The component-major
// outputs (r, J, e2) are already coalesced; the three per-observation record arrays (E 18, Pp 28, Pl 12 doubles) would
// be written with a 144 / 224 / 96-byte stride between neighbouring threads, so every thread parks its records in
// shared memory (odd row strides: conflict-free) and the CTA streams them out as contiguous runs.
#define VUS_LIN_TILE 128
#define VUS_LIN_ROW 61                 // 19 + 29 + 13: E | Pp | Pl with one padding double each (odd strides)
struct LinStereoTileBody {
  static VUS_DEV void run(const LinArgs& a, int tile, int tid, int nthr, double* sm) {
    const long f0 = (long)tile * VUS_LIN_TILE;
    const long n = a.F.n;
    const int nf = (int)(n - f0 < VUS_LIN_TILE ? n - f0 : VUS_LIN_TILE);
    for (int fl = tid; fl < nf; fl += nthr) {
      double* row = sm + fl * VUS_LIN_ROW;
      const long f = f0 + fl;
      LinOut O = a.O;                                   // f_stereo indexes its record arrays by f: aim them at this row
      O.sE = row - f * 18; O.sPp = row + 19 - f * 28; O.sPl = row + 48 - f * 12;
      f_stereo<true>(a.V, a.F, O, f, a.K);
    }
    VUS_SYNC();
    for (int e = tid; e < nf * 18; e += nthr) { const int fl = e / 18, c = e - fl * 18; a.O.sE[f0 * 18 + e] = sm[fl * VUS_LIN_ROW + c]; }
    for (int e = tid; e < nf * 28; e += nthr) { const int fl = e / 28, c = e - fl * 28; a.O.sPp[f0 * 28 + e] = sm[fl * VUS_LIN_ROW + 19 + c]; }
    for (int e = tid; e < nf * 12; e += nthr) { const int fl = e / 12, c = e - fl * 12; a.O.sPl[f0 * 12 + e] = sm[fl * VUS_LIN_ROW + 48 + c]; }
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
      for (int c = 0; c < 6; ++c) { d[c] = a.X.xc[xi * D + c]; d[9 + c] = a.X.xc[xj * D + c]; d[18 + c] = a.X.xb[6 * (long)ix[4 * n + f] + c]; }
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
enum { RED_STORE = 0, RED_PAP = 1, RED_RZ = 2, RED_RZ0 = 3, RED_RR = 4 };
// scalar slots.  S_TOL2: the PCG recursion is live while S_RR > S_TOL2 -- once the residual meets the tolerance alpha and beta
// are forced to zero ON THE DEVICE, so a block of iterations can be replayed without the host looking at every residual;
// S_NAN counts residual norms that were not finite, S_ITS the iterations that ran live.
enum { S_RZ = 0, S_PAP = 1, S_ALPHA = 2, S_BETA = 3, S_RR = 4, S_TMP = 5, S_NEG_ALPHA = 6, S_TOL2 = 7, S_COMM = 8, S_NAN = 9, S_ITS = 10, S_COUNT = 16 };

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
VUS_DEV void red_post(double* s, int slot, int op, double v) {
  const bool live = s[S_RR] > s[S_TOL2];               // (S_TOL2 = -1 outside a PCG recursion: always live)
  if (op == RED_STORE) s[slot] = v;
  else if (op == RED_PAP) { s[S_PAP] = v; const double al = (live && v != 0.0) ? s[S_RZ] / v : 0.0; s[S_ALPHA] = al; s[S_NEG_ALPHA] = -al; }
  else if (op == RED_RZ) { const double old = s[S_RZ]; s[S_BETA] = (live && old != 0.0) ? v / old : 0.0; s[S_RZ] = v; }
  else if (op == RED_RZ0) { s[S_RZ] = v; s[S_BETA] = 0.0; }
  else if (op == RED_RR) { if (live) { s[S_RR] = v; s[S_ITS] += 1.0; if (!(v == v)) s[S_NAN] += 1.0; } }   // a frozen recursion keeps the norm it stopped at
}
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
    if (tid == 0) red_post(A.scal, A.slot, A.op, sm[0]);
  }
};
// the same post-op on a value that was all-reduced across ranks first (partitioned graphs)
struct RedPostBody {
  static VUS_DEV void run(const Red2Args& A, long) { red_post(A.scal, A.slot, A.op, A.scal[S_COMM]); }
};

// =====================================================================================
// Kernel 2: assembly of the non-stereo factors (node-major / pair-major gathers: no atomics, fixed summation order)
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

// element (la, lb) of the diagonal node block of `node` inside SD; ld / bs = row stride / block stride of the padded
// supernode tiles ([KP][LD], bcr.cuh)
VUS_HD long diag_off(long node, int la, int lb, int D, int k, int ld, long bs) {
  const long I = node / k;
  const int rp = (int)(node % k);
  return I * bs + (long)(rp * D + la) * ld + rp * D + lb;
}

// Every Hessian / gradient entry the chain factors (priors, Between, DVL, IMU) touch has exactly ONE writer, which sums the
// contributions of the factors incident to it in a fixed order (factor type, then insertion index: the lists vus_analyze
// builds) and stores the result: the base system is bit-identical from run to run and needs no zero-fill per linearization.
// (The first version scattered J^T J with FP64 atomicAdd: ~30 M atomics per config-3 linearization, 0.11 of the HBM
// roofline, and a summation order -- hence low-order bits -- that changed from run to run.)
//   list entry: code = type * 2 + side  (side 0: the factor's first node p, 1: its second node q), f = row in the type's table
struct ChainTables {
  long n[VUS_F_NTYPES];
  const double* J[VUS_F_NTYPES];
  const double* r[VUS_F_NTYPES];
};
// Jacobian column of local node dof `a` (pose 0..5, velocity 6..8) on `side` of a factor of `type`; -1: the factor does not touch it
VUS_HD int chain_col(int type, int side, int a) {
  switch (type) {
    case VUS_F_PRIOR_POSE: return a < 6 ? a : -1;
    case VUS_F_PRIOR_VEL: return a >= 6 ? a - 6 : -1;
    case VUS_F_BETWEEN: return a < 6 ? 6 * side + a : -1;
    case VUS_F_DVL: return a;                                  // [Hx | Hv], node order
    case VUS_F_IMU: return 9 * side + a;
    default: return -1;
  }
}
VUS_HD int chain_rows(int type) { return type == VUS_F_IMU ? 9 : ((type == VUS_F_PRIOR_POSE || type == VUS_F_BETWEEN) ? 6 : 3); }
VUS_HD int chain_cols(int type) {
  return type == VUS_F_IMU ? 24 : (type == VUS_F_BETWEEN ? 12 : (type == VUS_F_DVL ? 9 : (type == VUS_F_PRIOR_POSE ? 6 : 3)));
}

struct NodeAsmArgs {
  ChainTables T;
  long nnodes;                 // real nodes
  int D, k, ld; long bs;       // node dof, nodes per supernode, row / block stride of the padded supernode tiles
  int has_bias;
  const int* ptr;              // [nnodes + 1]
  const int* code; const int* fac;     // [nent]
  double* SD; double* g; double* F;
};
// One CTA per tile of 32 consecutive nodes: lane = node, warp = local dof a.  A thread owns ROW a of its node's outputs --
// the diagonal block entries (a, b >= a) with their mirror images, the gradient entry a and the six bias-border entries of row
// a -- and walks the node's incident factors once: column a of the factor's Jacobian goes into registers (<= 9 values) and
// is multiplied with every partner column.  A warp reads the same column of 32 neighbouring factors (component-major J:
// coalesced) and the D warps of a tile re-read the same ~65 KB of Jacobians from the SM's own L1.
// (History: an item-per-thread grid, one output each, spread the outputs of a tile over a dozen SMs and was bound by
// L2 -> L1 traffic, 590 us per launch at config 3; a CTA per tile with one output per thread, 323 us, was bound by
// instruction issue -- two loads and the index arithmetic for every multiply-add.)
#define VUS_ASM_MAXD 9
VUS_DEV void chain_load_col(double (&v)[VUS_ASM_MAXD], const double* J, int M, int C, int c, long n, long f) {
  const double* p = J + (long)c * n + f;
  const long stride = (long)C * n;
#pragma unroll
  for (int q = 0; q < VUS_ASM_MAXD; ++q) { v[q] = q < M ? *p : 0.0; p += stride; }
}
VUS_DEV double chain_dot_col(const double (&v)[VUS_ASM_MAXD], const double* J, int M, int C, int c, long n, long f) {
  const double* p = J + (long)c * n + f;
  const long stride = (long)C * n;
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < VUS_ASM_MAXD; ++q) { if (q < M) s += v[q] * *p; p += stride; }
  return s;
}
struct NodeAsmBody {
  static VUS_DEV void run(const NodeAsmArgs& A, int tile, int tid, int nthr, double*) {
    for (int w = tid; w < 32 * A.D; w += nthr) item(A, tile, w >> 5, w & 31);
  }
  static VUS_DEV void item(const NodeAsmArgs& A, long tile, int a, int lane) {
    const long node = tile * 32 + lane;
    if (node >= A.nnodes) return;
    const int D = A.D;
    double acc[VUS_ASM_MAXD], accF[6], accg = 0.0, va[VUS_ASM_MAXD];
#pragma unroll
    for (int j = 0; j < VUS_ASM_MAXD; ++j) acc[j] = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) accF[j] = 0.0;
    for (int t = A.ptr[node]; t < A.ptr[node + 1]; ++t) {
      const int code = A.code[t], type = code >> 1, side = code & 1;
      const int ca = chain_col(type, side, a);
      if (ca < 0) continue;
      const long n = A.T.n[type], f = A.fac[t];
      const int M = chain_rows(type), C = chain_cols(type);
      const double* J = A.T.J[type];
      chain_load_col(va, J, M, C, ca, n, f);
#pragma unroll
      for (int j = 0; j < VUS_ASM_MAXD; ++j) {
        if (a + j >= D) continue;
        const int cb = j == 0 ? ca : chain_col(type, side, a + j);
        if (cb >= 0) acc[j] += chain_dot_col(va, J, M, C, cb, n, f);
      }
      {
        const double* r = A.T.r[type] + f;
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < VUS_ASM_MAXD; ++q) if (q < M) s += va[q] * r[(long)q * n];
        accg += s;
      }
      if (type == VUS_F_IMU && A.has_bias) {
#pragma unroll
        for (int j = 0; j < 6; ++j) accF[j] += chain_dot_col(va, J, M, C, 18 + j, n, f);
      }
    }
#pragma unroll
    for (int j = 0; j < VUS_ASM_MAXD; ++j) {
      if (a + j >= D) continue;
      A.SD[diag_off(node, a, a + j, D, A.k, A.ld, A.bs)] = acc[j];
      if (j) A.SD[diag_off(node, a + j, a, D, A.k, A.ld, A.bs)] = acc[j];
    }
    A.g[node * D + a] = -accg;
    if (A.has_bias) {
#pragma unroll
      for (int j = 0; j < 6; ++j) A.F[(node * D + a) * 6 + j] = accF[j];
    }
  }
};

// coupling blocks: one group per unordered node pair {lo < hi} with at least one two-node factor; entry (a, b) of block
// (lo, hi) = sum over the group's factors of J[:, col(lo, a)] . J[:, col(hi, b)].   code = type * 2 + flip  (flip: the
// factor's first node is hi).  One CTA per tile of 32 groups, lane = group, warp = row a of the block.
struct PairAsmArgs {
  ChainTables T;
  long ngroups; int D;
  const int* ptr; const int* code; const int* fac;
  const PairDst* dst;          // [ngroups], oriented lo -> hi
  double* Hval;                // base of SD | SU | REM
};
struct PairAsmBody {
  static VUS_DEV void run(const PairAsmArgs& A, int tile, int tid, int nthr, double*) {
    for (int w = tid; w < 32 * A.D; w += nthr) item(A, tile, w >> 5, w & 31);
  }
  static VUS_DEV void item(const PairAsmArgs& A, long tile, int a, int lane) {
    const int D = A.D;
    const long grp = tile * 32 + lane;
    if (grp >= A.ngroups) return;
    double acc[VUS_ASM_MAXD], va[VUS_ASM_MAXD];
#pragma unroll
    for (int j = 0; j < VUS_ASM_MAXD; ++j) acc[j] = 0.0;
    for (int t = A.ptr[grp]; t < A.ptr[grp + 1]; ++t) {
      const int code = A.code[t], type = code >> 1, flip = code & 1;
      const int ca = chain_col(type, flip, a);
      if (ca < 0) continue;
      const long n = A.T.n[type], f = A.fac[t];
      const int M = chain_rows(type), C = chain_cols(type);
      const double* J = A.T.J[type];
      chain_load_col(va, J, M, C, ca, n, f);
#pragma unroll
      for (int b = 0; b < VUS_ASM_MAXD; ++b) {
        if (b >= D) continue;
        const int cb = chain_col(type, 1 - flip, b);
        if (cb >= 0) acc[b] += chain_dot_col(va, J, M, C, cb, n, f);
      }
    }
    const PairDst d = A.dst[grp];
#pragma unroll
    for (int b = 0; b < VUS_ASM_MAXD; ++b) {
      if (b >= D) continue;
      if (d.transposed) A.Hval[d.off + (long)b * d.ld + a] = acc[b];
      else A.Hval[d.off + (long)a * d.ld + b] = acc[b];
      if (d.moff >= 0) A.Hval[d.moff + (long)b * d.mld + a] = acc[b];
    }
  }
};

#ifndef VUS_EMU
namespace rt {     // 32 * D <= 288 threads per CTA: room for the register-resident column and accumulators
template <> struct CoopBounds<NodeAsmBody> { static constexpr int kMaxThreads = 32 * VUS_ASM_MAXD, kMinBlocks = 2; };
template <> struct CoopBounds<PairAsmBody> { static constexpr int kMaxThreads = 32 * VUS_ASM_MAXD, kMinBlocks = 2; };
}  // namespace rt
#endif

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
  int ld; long bs;             // row / block stride of the padded supernode tiles
  double* SD; double* g;       // camera diag blocks / gradient
  double* C; double* gl;       // [9][nl], [3][nl]
  double* E;                   // [n][18]   E_o = Jp^T Jl (6x3 row-major)
  const double* Pp; const double* Pl;    // per-observation products written by the stereo linearize kernel
  long nl;
  const int* pose_ptr; const int* pose_obs; const int* pose_ids; long nposes_obs;   // CSR pose -> obs
  const int* lm_ptr;                                                                 // landmark -> contiguous obs range
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
      A.SD[diag_off(node, a, b, A.D, A.k, A.ld, A.bs)] += s;
      if (a != b) A.SD[diag_off(node, b, a, A.D, A.k, A.ld, A.bs)] += s;
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
    for (int t = A.lm_ptr[l]; t < A.lm_ptr[l + 1]; ++t) s += A.Pl[(long)t * 12 + e];   // observations are stored landmark-major
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
  const int* idx;                    // [2][n] pose, landmark
  const double* C; const double* gl; // undamped landmark blocks [9][nl], [3][nl]
  double* Cinv;                      // [9][nl]  (C + lambda I)^-1
  const double* E;                   // [n][18]  E_o = Jp^T Jl (6x3 row-major)
  double lambda;
  int D, k, B;
  int ld; long bs;                   // row / block stride of the padded supernode tiles
  double* SD; double* SU;            // damped system being formed (Schur complement subtracted in place)
  const double* SD0; const double* SU0;   // the base (undamped, unreduced) system
  double* gs;                        // reduced gradient (in/out)
  const int* pose_ptr; const int* pose_obs; const int* pose_ids; long nposes_obs;
  const int* lm_ptr;
  const int* lm_long;                // 1: the landmark's track does not fit in the band -> its Schur term stays implicit
  int* fail;
  // back-substitution
  const double* xc; double* xl;
  // implicit Schur term of the long-track landmarks (operator only)
  long nlong; const int* long_ids; double* ulong;     // ulong [3][nlong]
  const double* xin; double* yout;
  // SchurBlockBody: partner[t * ndj + dj] = row of the observation of the same landmark from pose i + dj (-1: none), t over
  // the pose-major observation list; ndj = longest short-track span + 1
  int* partner; int ndj;
  // long-track landmarks in the PRECONDITIONER (the operator keeps their exact implicit term): the track is cut into segments
  // that fit in the band, every segment acts as a landmark of its own (obs_seg[o] = segment of observation o, -1: short track;
  // seg_ptr[2 s], seg_ptr[2 s + 1]: observation rows of segment s; CinvSeg [9][nseg]); long_pass = 1: SchurBlockBody subtracts only these terms
  const int* obs_seg; const int* seg_ptr; double* CinvSeg; long nseg; const double* Pl; int long_pass; int has_long;
};
// analysis time, work item (t, dj): the partner of observation pose_obs[t] at pose offset dj.  Tracks are pose-sorted and
// hold every pose at most once (a landmark seen twice from one pose is handled by the implicit path, like a long track).
struct SchurPartnerBody {
  static VUS_DEV void run(const SchurArgs& A, long w) {
    const int dj = (int)(w % A.ndj);
    const long t = w / A.ndj;
    const int o = A.pose_obs[t];
    const int l = A.idx[A.n + o];
    int q = -1;
    // a long track pairs observations only inside one of its segments (the relaxed term of the preconditioner)
    const int end = !A.lm_long[l] ? A.lm_ptr[l + 1] : ((A.obs_seg && A.obs_seg[o] >= 0) ? A.seg_ptr[2 * A.obs_seg[o] + 1] : o);
    const int target = A.idx[o] + dj;
    for (int c = o; c < end; ++c) {
      const int pc = A.idx[c];
      if (pc == target) { q = c; break; }
      if (pc > target) break;
    }
    A.partner[w] = q;
  }
};
// per lambda, work item (pose slot pi, dj, row r): row r of the 6 x 6 block  S(i, i + dj) -= sum_o W_o E_q^T  (W_o = E_o Cinv_l,
// q the partner of o at pose i + dj) accumulated in registers, no shared memory, exactly one thread writes an entry (fixed
// summation order).  The six rows of a block sit in neighbouring lanes: their Cinv / E_q / partner loads are the same
// addresses (one transaction per warp instruction), so splitting a block over six threads costs no extra memory traffic but
// gives three times the loads in flight of the half-block version, which was bound by the latency of its dependent
// partner -> E_q loads (20 observations per pose, one after another: 1.39 ms per launch at config 3, 0.31 of the HBM
// roofline).  History: a thread per block ENTRY with partial blocks in shared memory was bound by L1 sector throughput
// (516 M sectors per launch, profiles/r1_ncu_schur_c3.txt); a thread per half block (round 1) by latency.
// The dj = 0 items also reduce the gradient  gs_i -= sum_o W_o gl_l.
struct SchurBlockBody {
  static VUS_DEV void run(const SchurArgs& A, long w) {
    const int r = (int)(w % 6);
    const int dj = (int)((w / 6) % A.ndj);
    const long pi = (w / 6) / A.ndj;
    const long i = A.pose_ids[pi];
    double acc[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, gacc = 0.0;
    bool any = false;
    const int t0 = A.pose_ptr[pi], t1 = A.pose_ptr[pi + 1];
    int qn = t0 < t1 ? A.partner[(long)t0 * A.ndj + dj] : -1;          // the partner of the next observation is fetched one
    for (int t = t0; t < t1; ++t) {                                    // iteration ahead of the record it points to
      const int q = qn;
      if (t + 1 < t1) qn = A.partner[(long)(t + 1) * A.ndj + dj];
      if (q < 0 && (dj != 0 || A.long_pass)) continue;
      const long o = A.pose_obs[t];
      const long l = A.idx[A.n + o];
      const bool lng = A.has_long && A.lm_long[l] != 0;    // has_long = 0: no extra load on graphs whose tracks all fit in the band
      if (lng != (A.long_pass != 0) && !(lng && dj == 0 && !A.long_pass)) continue;   // pass 0: short tracks (+ the gradient of long ones); pass 1: long tracks
      const double* er = A.E + o * 18 + r * 3;
      const double e0 = er[0], e1 = er[1], e2 = er[2];
      // pass 1 uses the inverse of the SEGMENT's landmark block: the segment is a landmark of its own in the preconditioner
      const double* ci = A.long_pass ? A.CinvSeg + A.obs_seg[o] : A.Cinv + l;
      const long nl = A.long_pass ? A.nseg : A.nl;
      const double w0 = e0 * ci[0] + e1 * ci[3 * nl] + e2 * ci[6 * nl];
      const double w1 = e0 * ci[nl] + e1 * ci[4 * nl] + e2 * ci[7 * nl];
      const double w2 = e0 * ci[2 * nl] + e1 * ci[5 * nl] + e2 * ci[8 * nl];
      if (dj == 0 && !A.long_pass) {
        gacc += w0 * A.gl[l] + w1 * A.gl[A.nl + l] + w2 * A.gl[2 * A.nl + l];
        if (q < 0 || lng) continue;                    // long track: only the (exact) gradient is reduced here (LongSchur*Body)
      }
      any = true;
      const double* eq = A.E + (long)q * 18;
#pragma unroll
      for (int sc = 0; sc < 6; ++sc) acc[sc] += w0 * eq[sc * 3] + w1 * eq[sc * 3 + 1] + w2 * eq[sc * 3 + 2];
    }
    const int D = A.D, k = A.k, B = A.ld;
    if (dj == 0 && !A.long_pass) A.gs[i * D + r] -= gacc;
    if (!any) return;
    const long I = i / k, j = i + dj, J = j / k;
    const int ri = (int)(i - I * k), rj = (int)(j - J * k);
    double* blk = (J == I ? A.SD : A.SU) + I * A.bs;
    // pass 0, dj = 0: the node's diagonal block was refreshed (base system + lambda) just before; dj >= 1: the block is ASSIGNED
    // from the base system, so the damped system needs no full copy per lambda try (form_system: only node and pair blocks are
    // copied).  pass 1 works on a fresh copy of the finished damped system: plain subtraction.
    const double* blk0 = (J == I ? A.SD0 : A.SU0) + I * A.bs;
#pragma unroll
    for (int sc = 0; sc < 6; ++sc) {
      const long o1 = (long)(ri * D + r) * B + rj * D + sc;
      const long o2 = (long)(rj * D + sc) * B + ri * D + r;
      if (dj == 0 || A.long_pass) {
        blk[o1] -= acc[sc];
        if (dj && J == I) blk[o2] -= acc[sc];
        continue;
      }
      blk[o1] = blk0[o1] - acc[sc];
      if (J == I) blk[o2] = blk0[o2] - acc[sc];
    }
  }
};
// per lambda and segment of a long track: CinvSeg = (sum of the segment's J_l^T J_l + lambda I)^-1  (products Pl of the stereo
// linearize kernel, as StereoLmBody / LmInvertBody do for whole landmarks)
struct SegInvertBody {
  static VUS_DEV void run(const SchurArgs& A, long sg) {
    double u[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    for (int t = A.seg_ptr[2 * sg]; t < A.seg_ptr[2 * sg + 1]; ++t)
#pragma unroll
      for (int e = 0; e < 6; ++e) u[e] += A.Pl[(long)t * 12 + e];
    // unique entries in the order (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
    const double c0 = u[0] + A.lambda, c1 = u[1], c2 = u[2], c4 = u[3] + A.lambda, c5 = u[4], c8 = u[5] + A.lambda;
    const double c00 = c4 * c8 - c5 * c5, c01 = c5 * c2 - c1 * c8, c02 = c1 * c5 - c4 * c2;
    const double det = c0 * c00 + c1 * c01 + c2 * c02;
    if (!(det > 0.0)) { *A.fail = 1; }
    const double id = 1.0 / det;
    const long ns = A.nseg;
    double* o = A.CinvSeg + sg;
    o[0] = c00 * id;          o[ns] = c01 * id;                      o[2 * ns] = c02 * id;
    o[3 * ns] = c01 * id;     o[4 * ns] = (c0 * c8 - c2 * c2) * id;  o[5 * ns] = (c2 * c1 - c0 * c5) * id;
    o[6 * ns] = c02 * id;     o[7 * ns] = (c2 * c1 - c0 * c5) * id;  o[8 * ns] = (c0 * c4 - c1 * c1) * id;
  }
};
// damped system <- base system on the blocks that change between linearizations: the diagonal block of every node and the
// coupling block (with its mirror image) of every node pair that shares a chain factor.  Everything else in SD | SU | REM is
// either structurally zero for the life of the graph or assigned by SchurBlockBody.
struct CopyBlocksArgs { const double* H0; double* H; long nnodes, ngroups; int D, k, ld; long bs, sd_off; const PairDst* dst; };
struct CopyBlocksBody {
  static VUS_DEV void run(const CopyBlocksArgs& A, long w) {
    const int DD = A.D * A.D;
    const long blkid = w / DD;
    const int e = (int)(w - blkid * DD), a = e / A.D, b = e - a * A.D;
    if (blkid < A.nnodes) {
      const long o = A.sd_off + diag_off(blkid, a, b, A.D, A.k, A.ld, A.bs);
      A.H[o] = A.H0[o];
      return;
    }
    const PairDst d = A.dst[blkid - A.nnodes];
    const long o = d.off + (long)a * d.ld + b;
    A.H[o] = A.H0[o];
    if (d.moff >= 0) { const long m = d.moff + (long)a * d.mld + b; A.H[m] = A.H0[m]; }
  }
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
// Landmarks whose track is longer than the band cannot be folded into the block-tridiagonal matrix without making it
// indefinite (the Schur complement subtracts from every block it touches).  They are still eliminated EXACTLY, but
// their term  - E_l (C_l + lambda I)^-1 E_l^T  is applied implicitly inside the operator and left out of the band
// preconditioner (which therefore stays an upper bound of the operator in the SPD order).
// pass 1, per long landmark:  u_l = Cinv_l * sum_o E_o^T x[pose_o]
struct LongSchur1Body {
  static VUS_DEV void run(const SchurArgs& A, long q) {
    const long l = A.long_ids[q];
    double t[3] = {0.0, 0.0, 0.0};
    for (int o = A.lm_ptr[l]; o < A.lm_ptr[l + 1]; ++o) {
      const long node = A.idx[o];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        const double x = A.xin[node * A.D + a];
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] += A.E[(long)o * 18 + a * 3 + c] * x;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      A.ulong[c * A.nlong + q] = A.Cinv[(c * 3) * A.nl + l] * t[0] + A.Cinv[(c * 3 + 1) * A.nl + l] * t[1] + A.Cinv[(c * 3 + 2) * A.nl + l] * t[2];
  }
};
// pass 2, per (pose with observations, dof a): y_i[a] -= sum over its observations o of long landmarks of E_o[a][:] u_l
// (gather through the pose -> observation lists: fixed summation order, no atomics); lm_long holds 1 + position in long_ids
struct LongSchur2Body {
  static VUS_DEV void run(const SchurArgs& A, long w) {
    const long pi = w / 6;
    const int a = (int)(w - pi * 6);
    const long node = A.pose_ids[pi];
    double s = 0.0;
    for (int t = A.pose_ptr[pi]; t < A.pose_ptr[pi + 1]; ++t) {
      const long o = A.pose_obs[t];
      const int q = A.lm_long[A.idx[A.n + o]];
      if (!q) continue;
#pragma unroll
      for (int c = 0; c < 3; ++c) s += A.E[o * 18 + a * 3 + c] * A.ulong[c * A.nlong + (q - 1)];
    }
    if (s != 0.0) A.yout[node * A.D + a] -= s;
  }
};
struct LmBacksubBody {   // per landmark: xl = Cinv (gl - sum_o E_o^T xc[pose_o])
  static VUS_DEV void run(const SchurArgs& A, long l) {
    double t[3] = {A.gl[l], A.gl[A.nl + l], A.gl[2 * A.nl + l]};
    for (int q = A.lm_ptr[l]; q < A.lm_ptr[l + 1]; ++q) {
      const long o = q;
      const long node = A.idx[o];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        const double x = A.xc[node * A.D + a];
#pragma unroll
        for (int c = 0; c < 3; ++c) t[c] -= A.E[o * 18 + a * 3 + c] * x;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      A.xl[c * A.nl + l] = A.Cinv[(c * 3) * A.nl + l] * t[0] + A.Cinv[(c * 3 + 1) * A.nl + l] * t[1] + A.Cinv[(c * 3 + 2) * A.nl + l] * t[2];
  }
};

// add lambda (and identity on padding dofs) to the diagonal of SD and Hbb
struct DampArgs { double* SD; double* Hbb; long ndof; long nreal; int B; double lambda; int ld; long bs; double pad_add; };   // pad_add: 1 when SD was just copied from the base system (padding diagonal 0 -> 1), 0 when the padding rows were left alone
struct DampBody {
  static VUS_DEV void run(const DampArgs& A, long i) {
    if (i < A.ndof) {
      const long I = i / A.B;
      const int r = (int)(i % A.B);
      A.SD[I * A.bs + (long)r * A.ld + r] += (i < A.nreal) ? A.lambda : A.pad_add;
    } else {
      const int c = (int)(i - A.ndof);
      A.Hbb[c * 6 + c] += A.lambda;
    }
  }
};

// =====================================================================================
// Kernel 3a: system operator  y = A x  on the reduced camera system
// =====================================================================================
// MatvecArgs / BandMatvecBody (band operator) live in bcr.cuh with the other streaming block kernels
struct MatvecArgs {
  const double* SD; const double* SU; long Ns; int B;
  long Nrows;                        // supernode rows computed (a rank of a partitioned graph only needs the rows of the nodes it owns)
  const double* x; double* y;        // camera parts, length Ns*B (per vector)
  int nv; long xstride, ystride;     // nv vectors: x[v*xstride + i], y[v*ystride + i]
  // remainder
  const int* rem_ptr; const int* rem_col; const double* rem_val; long nnodes; int D;
  // border
  const double* F; const double* Hbb; const double* xb; double* yb; int has_bias;
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
// block reduction of six per-thread values at a time: part[c] (c < 6) summed over the CTA into out[c]
VUS_DEV void cta_sum6(const double* part, double* out, int tid, int nthr, double* sm) {
  for (int c = 0; c < 6; ++c) sm[c * nthr + tid] = part[c];
  VUS_SYNC();
  for (int s = nthr >> 1; s > 0; s >>= 1) {
    for (int t = tid; t < s; t += nthr)
      for (int c = 0; c < 6; ++c) sm[c * nthr + t] += sm[c * nthr + t + s];
    VUS_SYNC();
  }
  if (tid == 0)
    for (int c = 0; c < 6; ++c) out[c] = sm[c * nthr];
  VUS_SYNC();
}
struct BorderDot1Body {       // one pass over F and the nv vectors; needs 6 * nthr doubles of shared memory
  static VUS_DEV void run(const BorderDotArgs& A, int bid, int tid, int nthr, double* sm) {
    const long chunk = (A.len + A.grid - 1) / A.grid;
    const long i0 = (long)bid * chunk;
    long i1 = i0 + chunk;
    if (i1 > A.len) i1 = A.len;
    double acc[6][6];
    for (int v = 0; v < 6; ++v)
      for (int c = 0; c < 6; ++c) acc[v][c] = 0.0;
    for (long i = i0 + tid; i < i1; i += nthr) {
      double f[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) f[c] = A.F[i * 6 + c];
#pragma unroll
      for (int v = 0; v < 6; ++v) {
        if (v < A.nv) {
          const double y = A.Y[(long)v * A.ystride + i];
#pragma unroll
          for (int c = 0; c < 6; ++c) acc[v][c] += f[c] * y;
        }
      }
    }
#pragma unroll
    for (int v = 0; v < 6; ++v)
      if (v < A.nv) cta_sum6(acc[v], A.partials + (long)bid * (A.nv * 6) + v * 6, tid, nthr, sm);
  }
};

// Kernel 3b (block cyclic reduction of the band) lives in bcr.cuh

// =====================================================================================
// small dense bias-border algebra (one thread)
// =====================================================================================
// out[a][b] = sum_i Z_a[i] R_b[i]  (Z^T R, both [6][len] column sets); stage-1 partials [grid][36]
struct ColDotArgs { const double* Z; const double* R; long len; long stride; double* partials; int grid; };
struct ColDot1Body {          // one pass over the 6 + 6 columns; needs 6 * nthr doubles of shared memory
  static VUS_DEV void run(const ColDotArgs& A, int bid, int tid, int nthr, double* sm) {
    const long chunk = (A.len + A.grid - 1) / A.grid;
    const long i0 = (long)bid * chunk;
    long i1 = i0 + chunk;
    if (i1 > A.len) i1 = A.len;
    double acc[6][6];
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) acc[a][b] = 0.0;
    for (long i = i0 + tid; i < i1; i += nthr) {
      double z[6], r[6];
#pragma unroll
      for (int c = 0; c < 6; ++c) { z[c] = A.Z[(long)c * A.stride + i]; r[c] = A.R[(long)c * A.stride + i]; }
#pragma unroll
      for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = 0; b < 6; ++b) acc[a][b] += z[a] * r[b];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) cta_sum6(acc[a], A.partials + (long)bid * 36 + a * 6, tid, nthr, sm);
  }
};
// Sb = Hbb - F^T Z (from partials of BorderDot over the 6 Z columns), SbInv = Sb^-1
struct BorderSchurArgs { const double* Hbb; const double* partials; int grid; int nv; double* SbInv; int* fail; const double* corr; };
// column sums of a [grid][stride] partial table into shared memory: out[e] = sum_b partials[b*stride + e], e < nout
// (thread per column, coalesced across columns; replaces one thread walking the whole table)
VUS_DEV void sum_partials(double* out, const double* partials, int grid, int stride, int nout, int tid, int nthr) {
  for (int e = tid; e < nout; e += nthr) {
    double s = 0.0;
    for (int b = 0; b < grid; ++b) s += partials[(long)b * stride + e];
    out[e] = s;
  }
}
struct BorderSchurBody {      // one CTA
  static VUS_DEV void run(const BorderSchurArgs& A, int, int tid, int nthr, double* sm) {
    double* ftz = sm;             // [36]  entry v*6+c = (F^T Z)[c][v]
    double* ztr = sm + 36;        // [36]
    sum_partials(ftz, A.partials, A.grid, A.nv * 6, 36, tid, nthr);
    if (A.corr) sum_partials(ztr, A.corr, A.grid, 36, 36, tid, nthr);
    VUS_SYNC();
    if (tid != 0) return;
    double S[36];
    for (int e = 0; e < 36; ++e) S[e] = A.Hbb[e];
    for (int v = 0; v < 6; ++v)
      for (int c = 0; c < 6; ++c) S[c * 6 + v] -= ftz[v * 6 + c];
    // second-order correction for the finite accuracy of Z: with R = F - M Z,  F^T M^-1 F = F^T Z + Z^T R + O(|dZ|^2)
    if (A.corr) {
      for (int e = 0; e < 36; ++e) {
        const int a = e / 6, c = e - a * 6;
        S[a * 6 + c] -= 0.5 * ztr[e];            // symmetrised: Z^T R is symmetric up to rounding
        S[c * 6 + a] -= 0.5 * ztr[e];
      }
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
struct BorderSolveBody {      // one CTA
  static VUS_DEV void run(const BorderSolveArgs& A, int, int tid, int nthr, double* sm) {
    sum_partials(sm, A.partials, A.grid, 6, 6, tid, nthr);
    VUS_SYNC();
    for (int r = tid; r < 6; r += nthr) {
      double s = 0.0;
      for (int c = 0; c < 6; ++c) s += A.SbInv[r * 6 + c] * (A.rb[c] - sm[c]);
      A.xb[r] = s;
    }
  }
};
// yb = F^T x (partials) + Hbb xb
struct BorderRowArgs { const double* Hbb; const double* xb; const double* partials; int grid; double* yb; };
struct BorderRowBody {        // one CTA
  static VUS_DEV void run(const BorderRowArgs& A, int, int tid, int nthr, double* sm) {
    sum_partials(sm, A.partials, A.grid, 6, 6, tid, nthr);
    VUS_SYNC();
    for (int r = tid; r < 6; r += nthr) {
      double s = sm[r];
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
// R_c[i] = F[i][c] - Y_c[i]  (residual of the border columns, for one step of iterative refinement of Z = M^-1 F)
struct BorderResidBody {
  static VUS_DEV void run(const struct BorderColsArgs& A, long w);
};
// Z_c[i] = F[i][c]  (border columns as right-hand sides)
struct BorderColsArgs { const double* F; double* Z; long len; long zstride; double* R; };
struct BorderColsBody {
  static VUS_DEV void run(const BorderColsArgs& A, long w) {
    const long i = w % A.len;
    const int c = (int)(w / A.len);
    A.Z[(long)c * A.zstride + i] = A.F[i * 6 + c];
  }
};
VUS_DEV void BorderResidBody::run(const BorderColsArgs& A, long w) {
  const long i = w % A.len;
  const int c = (int)(w / A.len);
  A.R[(long)c * A.zstride + i] = A.F[i * 6 + c] - A.R[(long)c * A.zstride + i];
}
struct AddVecBody {     // y[i] += x[i]
  static VUS_DEV void run(const VecArgs& A, long i) { A.y[i] += A.x[i]; }
};

// layout conversion between the caller's row-major [n][dim] tables and the component-major device tables
struct TransposeArgs { const double* src; double* dst; long n; int dim; int to_soa; };
struct TransposeBody {
  static VUS_DEV void run(const TransposeArgs& A, long w) {
    const long i = w / A.dim;
    const int c = (int)(w - i * A.dim);
    if (A.to_soa) A.dst[(long)c * A.n + i] = A.src[w];
    else A.dst[w] = A.src[(long)c * A.n + i];
  }
};
// sort keys for the landmark-major observation order: (landmark << 32) | pose, value = row
struct SortKeyArgs { const int* idx; long n; unsigned long long* keys; int* vals; };
struct SortKeyBody {
  static VUS_DEV void run(const SortKeyArgs& A, long f) {
    A.keys[f] = ((unsigned long long)(unsigned)A.idx[A.n + f] << 32) | (unsigned)A.idx[f];
    A.vals[f] = (int)f;
  }
};
// the same re-ordering for the int32 index table [slots][n]; flags a non-identity order
struct GatherIdxArgs { const int* src; int* dst; const int* perm; long n; int slots; int* moved; };
struct GatherIdxBody {
  static VUS_DEV void run(const GatherIdxArgs& A, long w) {
    const long f = w % A.n;
    const long c = w / A.n;
    const int s = A.perm[f];
    A.dst[c * A.n + f] = A.src[c * A.n + s];
    if (c == 0 && s != (int)f) *A.moved = 1;
  }
};
// sort keys for the pose-major observation lists: key = pose of row f, value = row
struct PoseKeyArgs { const int* idx; long n; unsigned* keys; int* vals; };
struct PoseKeyBody {
  static VUS_DEV void run(const PoseKeyArgs& A, long f) { A.keys[f] = (unsigned)A.idx[f]; A.vals[f] = (int)f; }
};
// dst[c][f] = src[c][perm[f]]  (re-ordering of a component-major table, used once per graph by vus_analyze)
struct GatherArgs { const double* src; double* dst; const int* perm; long n; int comps; };
struct GatherBody {
  static VUS_DEV void run(const GatherArgs& A, long w) {
    const long f = w % A.n;
    const long c = w / A.n;
    A.dst[c * A.n + f] = A.src[c * A.n + A.perm[f]];
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
    for (int c = 0; c < 6; ++c) A.bias_out[c * A.nb + w] = A.bias[c * A.nb + w] + A.xb[6 * w + c];
  }
};

}  // namespace vus

#include "bcr.cuh"
#include "batch.cuh"
