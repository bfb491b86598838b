// Kernel 3b -- block cyclic reduction (BCR) of the supernode block-tridiagonal band: the exact band solve used as
// the PCG preconditioner (and, when the graph has no off-band blocks, as the solve itself).
//
// This replaces what gtsam's multifrontal Cholesky does for the chain part of the graph under
// LevenbergMarquardtOptimizer::optimize() (/root/reference/batch.py:337).  Blocks are B x B, B = k*D <= 96.
//
// Level with stride s: nodes j = s*(2m+1) are eliminated, nodes c = 2*m*s survive.
//   elim  (per j):  Dinv_j = D_j^-1,  Gl_j = U_{j-s} Dinv_j,  Gr_j = U_j^T Dinv_j           (U_i = A(i, i+s))
//   update(per c):  D_c   -= Gr_{c-s} U_{c-s} + Gl_{c+s} U_c^T ,   Unext_c = -Gl_{c+s} U_{c+s}
//   solve fwd (c):  b_c   -= Gr_{c-s} b_{c-s} + Gl_{c+s} b_{c+s}
//   solve bwd (j):  x_j    = Dinv_j b_j - Gl_j^T x_{j-s} - Gr_j^T x_{j+s}
//
// sm_100a implementation: every dense product runs on the FP64 tensor path (mma.sync.m8n8k4.f64, SASS DMMA) with
// the operands staged in shared memory and the accumulators in registers; the SPD inverse is a blocked (8 x 8 pivot
// tiles) Gauss-Jordan that keeps the whole block in DMMA accumulator fragments and only moves the pivot panels through
// shared memory.  Measured on B200: DMMA 37.1 TFLOP/s, DFMA 34.1 TFLOP/s peak -- the tensor path is used because it
// needs ~8x fewer issued instructions per flop, which is what bounded the scalar version (profiles/r1_00_*).
// The VUS_EMU build (tests only) runs the same algorithm with plain loops.
#pragma once
#ifdef VUS_EMU
#include <vector>
#endif

namespace vus {

VUS_HD int bcr_kp(int B) { return (B + 3) & ~3; }                                       // K padded to the DMMA k=4
VUS_HD int bcr_ld(int B) { const int kp = bcr_kp(B); return (kp & 7) == 4 ? kp : kp + 4; }   // row stride = 4 mod 8: conflict-free fragments
VUS_HD int bcr_tiles(int B) { return (B + 7) >> 3; }
VUS_HD long bcr_buf_doubles(int B) { return (long)bcr_kp(B) * bcr_ld(B); }
// Gauss-Jordan panel scratch (aliases the operand buffers): 2 x {Rraw, Rnew [8][100], Craw, Cnew [96][12]} + per-warp 8x8
#define VUS_GJ_LDP 100
#define VUS_GJ_LDQ 12
#define VUS_GJ_SET (2 * 8 * VUS_GJ_LDP + 2 * 96 * VUS_GJ_LDQ)
#define VUS_GJ_DOUBLES (2 * VUS_GJ_SET + 8 * 64)
VUS_HD long bcr_smem_doubles(int B) {
  const long a = 2 * bcr_buf_doubles(B);
  return a > VUS_GJ_DOUBLES ? a : VUS_GJ_DOUBLES;
}

struct BcrArgs {
  long Ns; int B; long s;          // level stride
  double* Dw;                      // working diagonal blocks [Ns]
  const double* Ucur; double* Unext;   // couplings at this level / next level, indexed by node id
  double* Dinv; double* Gl; double* Gr;   // per eliminated node
  int* fail;
  // solve
  double* X; long xstride; int nrhs;
};

#ifndef VUS_EMU
// =====================================================================================  sm_100a: DMMA tile engine
VUS_DEV void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// A 256-thread CTA = 8 warps in a 4 x 2 grid; warp (wr, wc) owns the 8x8 tiles ti in [3wr, 3wr+3), tj in [6wc, 6wc+6)
// of a (<= 96 x 96) block.  Inside a tile lane (g = lane/4, t = lane%4) holds (row g, cols 2t, 2t+1).
struct Tiles {
  int B, T, KP, LD, ti0, tj0, na, nb, g, t, warp, lane;
  VUS_DEV Tiles(int B_, int tid) {
    B = B_; T = bcr_tiles(B); KP = bcr_kp(B); LD = bcr_ld(B);
    warp = tid >> 5; lane = tid & 31; g = lane >> 2; t = lane & 3;
    ti0 = 3 * (warp >> 1); tj0 = 6 * (warp & 1);
    na = T - ti0; na = na < 0 ? 0 : (na > 3 ? 3 : na);
    nb = T - tj0; nb = nb < 0 ? 0 : (nb > 6 ? 6 : nb);
  }
};
typedef double Acc[3][6][2];

VUS_DEV void acc_zero(Acc& c) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 6; ++b) c[a][b][0] = c[a][b][1] = 0.0;
}
// acc += op(A) op(B)  (NEG: acc -= op(A) op(B)); sA / sB operand buffers [KP][LD] (zero padded); TA: A^T is stored, TB: B^T is stored
template <bool TA, bool TB, bool NEG>
VUS_DEV void mma_gemm(Acc& c, const double* sA, const double* sB, const Tiles& G) {
  if (G.na == 0 || G.nb == 0) return;
  int ia[3], jb[6];
#pragma unroll
  for (int a = 0; a < 3; ++a) { const int i = (G.ti0 + a) * 8 + G.g; ia[a] = i < G.B ? i : G.B - 1; }
#pragma unroll
  for (int b = 0; b < 6; ++b) { const int j = (G.tj0 + b) * 8 + G.g; jb[b] = j < G.B ? j : G.B - 1; }
  const int LD = G.LD;
#pragma unroll 2
  for (int k0 = 0; k0 < G.KP; k0 += 4) {
    const int kk = k0 + G.t;
    double af[3], bf[6];
#pragma unroll
    for (int a = 0; a < 3; ++a) { const double v = TA ? sA[kk * LD + ia[a]] : sA[ia[a] * LD + kk]; af[a] = NEG ? -v : v; }
#pragma unroll
    for (int b = 0; b < 6; ++b) bf[b] = TB ? sB[jb[b] * LD + kk] : sB[kk * LD + jb[b]];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 6; ++b)
        if (a < G.na && b < G.nb) dmma884(c[a][b][0], c[a][b][1], af[a], bf[b]);
  }
}
// global row-major B x B block -> operand buffer [KP][LD], zero padded.  Warp per row, lanes across columns (coalesced);
// every element is an asynchronous 8-byte global->shared copy (LDGSTS), so all of a thread's copies are in flight at once.
// Follow with stage_wait() and a barrier.
VUS_DEV void stage_block(double* s, const double* g, const Tiles& G) {
  for (int r = G.warp; r < G.KP; r += 8)
    for (int c = G.lane; c < G.LD; c += 32) {
      double* dst = s + r * G.LD + c;
      if (r < G.B && c < G.B) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(g + (long)r * G.B + c) : "memory");
      } else {
        *dst = 0.0;
      }
    }
}
VUS_DEV void stage_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }
// accumulator tiles -> global row-major block; dst = alpha * acc (+ dst if ADD); optional second destination
template <bool ADD>
VUS_DEV void acc_store_global(double* dst, const Acc& c, double alpha, const Tiles& G) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      if (a >= G.na || b >= G.nb) continue;
      const int i = (G.ti0 + a) * 8 + G.g, j = (G.tj0 + b) * 8 + 2 * G.t;
      if (i >= G.B) continue;
      double* p = dst + (long)i * G.B + j;
      if (j < G.B) p[0] = ADD ? p[0] + alpha * c[a][b][0] : alpha * c[a][b][0];
      if (j + 1 < G.B) p[1] = ADD ? p[1] + alpha * c[a][b][1] : alpha * c[a][b][1];
    }
}
// accumulator tiles -> operand buffer [KP][LD] (zero outside the B x B block)
VUS_DEV void acc_store_smem(double* s, const Acc& c, const Tiles& G) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      if (a >= G.na || b >= G.nb) continue;
      const int i = (G.ti0 + a) * 8 + G.g, j = (G.tj0 + b) * 8 + 2 * G.t;
      if (i >= G.KP) continue;
      if (j < G.LD) s[i * G.LD + j] = (i < G.B && j < G.B) ? c[a][b][0] : 0.0;
      if (j + 1 < G.LD) s[i * G.LD + j + 1] = (i < G.B && j + 1 < G.B) ? c[a][b][1] : 0.0;
    }
}
// global row-major SPD block -> accumulator tiles, identity on the padding
VUS_DEV void acc_load_global(Acc& c, const double* src, const Tiles& G) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const int i = (G.ti0 + a) * 8 + G.g, j = (G.tj0 + b) * 8 + 2 * G.t;
      const bool in = a < G.na && b < G.nb && i < G.B;
      c[a][b][0] = (in && j < G.B) ? src[(long)i * G.B + j] : (i == j ? 1.0 : 0.0);
      c[a][b][1] = (in && j + 1 < G.B) ? src[(long)i * G.B + j + 1] : (i == j + 1 ? 1.0 : 0.0);
    }
}

// In-place inverse of the SPD block held in accumulator tiles: blocked Gauss-Jordan over 8 x 8 pivot tiles, no pivoting.
//   step p:  P = M[p][p]^-1 ;  M[i][j] += (-M[i][p] P) M[p][j]  (i, j != p) ;  M[p][j] = P M[p][j] ;  M[i][p] = -M[i][p] P ;  M[p][p] = P
// The rank-8 updates are DMMAs on the resident accumulators; only the pivot row / column panels go through shared
// memory (double buffered: two barriers per step).  `sm` needs VUS_GJ_DOUBLES doubles.
VUS_DEV void mma_gj_inverse(Acc& c, double* sm, const Tiles& G, int* fail) {
  const int LDP = VUS_GJ_LDP, LDQ = VUS_GJ_LDQ;
  double* Pw = sm + 2 * VUS_GJ_SET + G.warp * 64;
  const int g = G.g, t = G.t;
  for (int p = 0; p < G.T; ++p) {
    double* Rraw = sm + (p & 1) * VUS_GJ_SET;
    double* Rnew = Rraw + 8 * LDP;
    double* Craw = Rnew + 8 * LDP;
    double* Cnew = Craw + 96 * LDQ;
    // (a) owners publish the raw pivot row / column panels
#pragma unroll
    for (int a = 0; a < 3; ++a)
      if (G.ti0 + a == p) {
#pragma unroll
        for (int b = 0; b < 6; ++b)
          if (b < G.nb) { const int col = (G.tj0 + b) * 8 + 2 * t; Rraw[g * LDP + col] = c[a][b][0]; Rraw[g * LDP + col + 1] = c[a][b][1]; }
      }
#pragma unroll
    for (int b = 0; b < 6; ++b)
      if (G.tj0 + b == p) {
#pragma unroll
        for (int a = 0; a < 3; ++a)
          if (a < G.na) { const int row = (G.ti0 + a) * 8 + g; Craw[row * LDQ + 2 * t] = c[a][b][0]; Craw[row * LDQ + 2 * t + 1] = c[a][b][1]; }
      }
    __syncthreads();
    // (b) every warp inverts the 8x8 pivot tile redundantly: lane r (mod 8) owns row r, pivot rows travel by shuffle
    {
      const int r = G.lane & 7;
      double row[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) row[q] = Rraw[r * LDP + 8 * p + q];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        double prow[8];
#pragma unroll
        for (int q2 = 0; q2 < 8; ++q2) prow[q2] = __shfl_sync(0xffffffffu, row[q2], q);
        const double piv = prow[q];
        if (G.lane == 0 && !(piv > 0.0)) *fail = 1;
        const double d = 1.0 / piv;
        const double f = row[q] * d;
        if (r == q) {
#pragma unroll
          for (int q2 = 0; q2 < 8; ++q2) row[q2] = (q2 == q) ? d : prow[q2] * d;
        } else {
#pragma unroll
          for (int q2 = 0; q2 < 8; ++q2) row[q2] = (q2 == q) ? -f : row[q2] - f * prow[q2];
        }
      }
      __syncwarp();
      if (G.lane < 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) Pw[r * 8 + q] = row[q];
      }
      __syncwarp();
    }
    // (c) new panels, spread over the warps:  Rnew[q] = P Rraw[q],  Cnew[q] = -Craw[q] P   (q != p)
    for (int item = G.warp; item < 2 * G.T; item += 8) {
      const int q = item < G.T ? item : item - G.T;
      if (q == p) continue;
      double c0 = 0.0, c1 = 0.0;
      if (item < G.T) {
        dmma884(c0, c1, Pw[g * 8 + t], Rraw[t * LDP + 8 * q + g]);
        dmma884(c0, c1, Pw[g * 8 + 4 + t], Rraw[(4 + t) * LDP + 8 * q + g]);
        Rnew[g * LDP + 8 * q + 2 * t] = c0; Rnew[g * LDP + 8 * q + 2 * t + 1] = c1;
      } else {
        dmma884(c0, c1, Craw[(8 * q + g) * LDQ + t], Pw[t * 8 + g]);
        dmma884(c0, c1, Craw[(8 * q + g) * LDQ + 4 + t], Pw[(4 + t) * 8 + g]);
        Cnew[(8 * q + g) * LDQ + 2 * t] = -c0; Cnew[(8 * q + g) * LDQ + 2 * t + 1] = -c1;
      }
    }
    __syncthreads();
    // (d) update the resident tiles
    double af[3][2], bf[6][2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const int row = (G.ti0 + a) * 8 + g;
      const bool use = a < G.na && G.ti0 + a != p;
      af[a][0] = use ? Cnew[row * LDQ + t] : 0.0;
      af[a][1] = use ? Cnew[row * LDQ + 4 + t] : 0.0;
    }
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const int col = (G.tj0 + b) * 8 + g;
      const bool use = b < G.nb && G.tj0 + b != p;
      bf[b][0] = use ? Rraw[t * LDP + col] : 0.0;
      bf[b][1] = use ? Rraw[(4 + t) * LDP + col] : 0.0;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 6; ++b) {
        if (a >= G.na || b >= G.nb) continue;
        const int ti = G.ti0 + a, tj = G.tj0 + b;
        if (ti == p && tj == p) { c[a][b][0] = Pw[g * 8 + 2 * t]; c[a][b][1] = Pw[g * 8 + 2 * t + 1]; }
        else if (ti == p) { c[a][b][0] = Rnew[g * LDP + 8 * tj + 2 * t]; c[a][b][1] = Rnew[g * LDP + 8 * tj + 2 * t + 1]; }
        else if (tj == p) { c[a][b][0] = Cnew[(8 * ti + g) * LDQ + 2 * t]; c[a][b][1] = Cnew[(8 * ti + g) * LDQ + 2 * t + 1]; }
        else { dmma884(c[a][b][0], c[a][b][1], af[a][0], bf[b][0]); dmma884(c[a][b][0], c[a][b][1], af[a][1], bf[b][1]); }
      }
    __syncwarp();
  }
  __syncthreads();
}

// per eliminated node j = s*(2m+1): Dinv_j, Gl_j = U[j-s] Dinv_j, Gr_j = U[j]^T Dinv_j      (256 threads)
struct BcrElimBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int, double* sm) {
    const Tiles G(A.B, tid);
    const long BB = (long)A.B * A.B;
    const long j = A.s * (2L * m + 1);
    double* buf0 = sm;
    double* buf1 = sm + bcr_buf_doubles(A.B);
    Acc c;
    acc_load_global(c, A.Dw + j * BB, G);
    mma_gj_inverse(c, sm, G, A.fail);
    acc_store_global<false>(A.Dinv + j * BB, c, 1.0, G);
    acc_store_smem(buf1, c, G);
    stage_block(buf0, A.Ucur + (j - A.s) * BB, G);
    stage_wait();
    __syncthreads();
    acc_zero(c);
    mma_gemm<false, false, false>(c, buf0, buf1, G);
    if (j + A.s < A.Ns) {
      __syncthreads();
      stage_block(buf0, A.Ucur + j * BB, G);           // in flight while Gl is written out
      acc_store_global<false>(A.Gl + j * BB, c, 1.0, G);
      stage_wait();
      __syncthreads();
      acc_zero(c);
      mma_gemm<true, false, false>(c, buf0, buf1, G);
      acc_store_global<false>(A.Gr + j * BB, c, 1.0, G);
    } else {
      acc_store_global<false>(A.Gl + j * BB, c, 1.0, G);
    }
  }
};
// per surviving node c = 2*m*s: Dw_c -= Gr_{c-s} U_{c-s} + Gl_{c+s} U_c^T ; Unext_c = -Gl_{c+s} U_{c+s}
struct BcrUpdateBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int, double* sm) {
    const Tiles G(A.B, tid);
    const long BB = (long)A.B * A.B;
    const long c = 2L * m * A.s;
    double* buf0 = sm;
    double* buf1 = sm + bcr_buf_doubles(A.B);
    const bool lo = c - A.s >= 0, hi = c + A.s < A.Ns;
    if (lo) {
      stage_block(buf0, A.Gr + (c - A.s) * BB, G);
      stage_block(buf1, A.Ucur + (c - A.s) * BB, G);
    }
    Acc acc;
    acc_load_global(acc, A.Dw + c * BB, G);            // D_c rides in the accumulators; the products are subtracted
    if (lo) {
      stage_wait();
      __syncthreads();
      mma_gemm<false, false, true>(acc, buf0, buf1, G);
      __syncthreads();
    }
    if (hi) {
      const long j = c + A.s;
      stage_block(buf0, A.Gl + j * BB, G);
      stage_block(buf1, A.Ucur + c * BB, G);
      stage_wait();
      __syncthreads();
      mma_gemm<false, true, true>(acc, buf0, buf1, G);
      if (j + A.s < A.Ns) {
        __syncthreads();
        stage_block(buf1, A.Ucur + j * BB, G);          // in flight while D_c is written out
        acc_store_global<false>(A.Dw + c * BB, acc, 1.0, G);
        stage_wait();
        __syncthreads();
        acc_zero(acc);
        mma_gemm<false, false, true>(acc, buf0, buf1, G);
        acc_store_global<false>(A.Unext + c * BB, acc, 1.0, G);
        return;
      }
    }
    acc_store_global<false>(A.Dw + c * BB, acc, 1.0, G);
  }
};
// root: Dinv_0 = inv(Dw_0)
struct BcrRootBody {
  static VUS_DEV void run(const BcrArgs& A, int, int tid, int, double* sm) {
    const Tiles G(A.B, tid);
    Acc c;
    acc_load_global(c, A.Dw, G);
    mma_gj_inverse(c, sm, G, A.fail);
    acc_store_global<false>(A.Dinv, c, 1.0, G);
  }
};


namespace rt {
template <> struct CoopBounds<BcrElimBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 2; };
template <> struct CoopBounds<BcrUpdateBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 2; };
template <> struct CoopBounds<BcrRootBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 1; };
}  // namespace rt

#else
// =====================================================================================  host emulation (tests only)
inline void emu_spd_inverse(const double* M, double* out, int B, int* fail) {
  for (int e = 0; e < B * B; ++e) out[e] = M[e];
  std::vector<double> rowp(B), colp(B);
  for (int p = 0; p < B; ++p) {
    for (int i = 0; i < B; ++i) { colp[i] = out[(long)i * B + p]; rowp[i] = out[(long)p * B + i]; }
    const double piv = rowp[p];
    if (!(piv > 0.0)) *fail = 1;
    const double d = 1.0 / piv;
    for (int i = 0; i < B; ++i) {
      double* row = out + (long)i * B;
      if (i == p) { for (int j = 0; j < B; ++j) row[j] = (j == p) ? d : rowp[j] * d; }
      else { const double ci = colp[i] * d; for (int j = 0; j < B; ++j) row[j] = (j == p) ? -ci : row[j] - ci * rowp[j]; }
    }
  }
}
// C = beta*C + alpha * op(A) op(B), all row-major B x B
inline void emu_gemm(double* C, const double* A, bool ta, const double* Bm, bool tb, int B, double alpha, double beta) {
  std::vector<double> out((size_t)B * B);
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < B; ++j) {
      double s = 0.0;
      for (int k = 0; k < B; ++k) s += (ta ? A[(long)k * B + i] : A[(long)i * B + k]) * (tb ? Bm[(long)j * B + k] : Bm[(long)k * B + j]);
      out[(long)i * B + j] = alpha * s + (beta != 0.0 ? beta * C[(long)i * B + j] : 0.0);
    }
  for (long e = 0; e < (long)B * B; ++e) C[e] = out[e];
}
struct BcrElimBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int, int, double*) {
    const int B = A.B;
    const long BB = (long)B * B;
    const long j = A.s * (2L * m + 1);
    emu_spd_inverse(A.Dw + j * BB, A.Dinv + j * BB, B, A.fail);
    emu_gemm(A.Gl + j * BB, A.Ucur + (j - A.s) * BB, false, A.Dinv + j * BB, false, B, 1.0, 0.0);
    if (j + A.s < A.Ns) emu_gemm(A.Gr + j * BB, A.Ucur + j * BB, true, A.Dinv + j * BB, false, B, 1.0, 0.0);
  }
};
struct BcrUpdateBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int, int, double*) {
    const int B = A.B;
    const long BB = (long)B * B;
    const long c = 2L * m * A.s;
    if (c - A.s >= 0) emu_gemm(A.Dw + c * BB, A.Gr + (c - A.s) * BB, false, A.Ucur + (c - A.s) * BB, false, B, -1.0, 1.0);
    if (c + A.s < A.Ns) {
      const long j = c + A.s;
      emu_gemm(A.Dw + c * BB, A.Gl + j * BB, false, A.Ucur + c * BB, true, B, -1.0, 1.0);
      if (j + A.s < A.Ns) emu_gemm(A.Unext + c * BB, A.Gl + j * BB, false, A.Ucur + j * BB, false, B, -1.0, 0.0);
    }
  }
};
struct BcrRootBody {
  static VUS_DEV void run(const BcrArgs& A, int, int, int, double*) { emu_spd_inverse(A.Dw, A.Dinv, A.B, A.fail); }
};
#endif

// =====================================================================================  triangular-free solve sweeps
// Shared-memory layout of the solve kernels: xs [3][nrhs][B] staged vectors, part [ngrp][nrhs][B] partial sums.
VUS_HD long bcr_solve_smem_doubles(int B, int nrhs, int nthr) {
  int ngrp = nthr / B; if (ngrp < 1) ngrp = 1;
  return (long)(3 + ngrp) * nrhs * B;
}
// out[v][r] = sum_k M[r][k] x[v][k]   (row-major block times staged vectors; one warp per row, shuffle reduction)
VUS_DEV void cta_rowdot_sub(double* X, long xstride, long node, const double* M, const double* xs, int B, int nrhs, int tid, int nthr) {
#ifdef VUS_EMU
  for (int e = tid; e < nrhs * B; e += nthr) {
    const int v = e / B, r = e - v * B;
    double s = 0.0;
    for (int k = 0; k < B; ++k) s += M[(long)r * B + k] * xs[v * B + k];
    X[(long)v * xstride + node * B + r] -= s;
  }
#else
  const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
  for (int r = warp; r < B; r += nw) {
    const double* row = M + (long)r * B;
    const double m0 = lane < B ? row[lane] : 0.0;
    const double m1 = lane + 32 < B ? row[lane + 32] : 0.0;
    const double m2 = lane + 64 < B ? row[lane + 64] : 0.0;
    for (int v = 0; v < nrhs; ++v) {
      const double* x = xs + v * B;
      double s = lane < B ? m0 * x[lane] : 0.0;
      if (lane + 32 < B) s += m1 * x[lane + 32];
      if (lane + 64 < B) s += m2 * x[lane + 64];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) X[(long)v * xstride + node * B + r] -= s;
    }
  }
#endif
}
// forward sweep, per surviving node c: b_c -= Gr_{c-s} b_{c-s} + Gl_{c+s} b_{c+s}
struct BcrFwdBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int nthr, double* sm) {
    const int B = A.B;
    const long BB = (long)B * B;
    const long c = 2L * m * A.s;
    double* xs = sm;   // [2][nrhs][B] neighbour rhs
    const long jl = c - A.s, jh = c + A.s;
    for (int e = tid; e < A.nrhs * B; e += nthr) {
      const int v = e / B, r = e - v * B;
      if (jl >= 0) xs[e] = A.X[(long)v * A.xstride + jl * B + r];
      if (jh < A.Ns) xs[A.nrhs * B + e] = A.X[(long)v * A.xstride + jh * B + r];
    }
    VUS_SYNC();
    if (jl >= 0) cta_rowdot_sub(A.X, A.xstride, c, A.Gr + jl * BB, xs, B, A.nrhs, tid, nthr);
    VUS_SYNC();
    if (jh < A.Ns) cta_rowdot_sub(A.X, A.xstride, c, A.Gl + jh * BB, xs + A.nrhs * B, B, A.nrhs, tid, nthr);
  }
};
// backward sweep, per eliminated node j: x_j = Dinv_j b_j - Gl_j^T x_{j-s} - Gr_j^T x_{j+s}
// threads (grp, r): column r of each block over a k-range (coalesced rows), partial sums combined through shared memory
struct BcrBwdBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int nthr, double* sm) {
    const int B = A.B, nrhs = A.nrhs;
    const long BB = (long)B * B;
    const long j = A.s * (2L * m + 1);
    const long jl = j - A.s, jh = j + A.s;
    const bool hl = jl >= 0 && A.s > 0, hh = jh < A.Ns && A.s > 0;
    int ngrp = nthr / B; if (ngrp < 1) ngrp = 1;
    double* xs = sm;                       // [3][nrhs][B]: b_j, x_{j-s}, x_{j+s}
    double* part = sm + 3 * nrhs * B;      // [ngrp][nrhs][B]
    for (int e = tid; e < nrhs * B; e += nthr) {
      const int v = e / B, r = e - v * B;
      xs[e] = A.X[(long)v * A.xstride + j * B + r];
      xs[nrhs * B + e] = hl ? A.X[(long)v * A.xstride + jl * B + r] : 0.0;
      xs[2 * nrhs * B + e] = hh ? A.X[(long)v * A.xstride + jh * B + r] : 0.0;
    }
    VUS_SYNC();
    const int kchunk = (B + ngrp - 1) / ngrp;
    const double* Di = A.Dinv + j * BB;
    const double* Gl = A.Gl + j * BB;
    const double* Gr = A.Gr + j * BB;
    for (int e = tid; e < ngrp * B; e += nthr) {
      const int grp = e / B, r = e - grp * B;
      const int k0 = grp * kchunk;
      int k1 = k0 + kchunk; if (k1 > B) k1 = B;
      for (int v0 = 0; v0 < nrhs; v0 += 6) {
        double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const int nv = nrhs - v0 < 6 ? nrhs - v0 : 6;
        for (int k = k0; k < k1; ++k) {
          const double d = Di[(long)k * B + r];                  // Dinv symmetric: column r read as row k
          const double gl = hl ? Gl[(long)k * B + r] : 0.0;
          const double gr = hh ? Gr[(long)k * B + r] : 0.0;
#pragma unroll
          for (int v = 0; v < 6; ++v)
            if (v < nv) s[v] += d * xs[(v0 + v) * B + k] - gl * xs[(nrhs + v0 + v) * B + k] - gr * xs[(2 * nrhs + v0 + v) * B + k];
        }
#pragma unroll
        for (int v = 0; v < 6; ++v)
          if (v < nv) part[(grp * nrhs + v0 + v) * B + r] = s[v];
      }
    }
    VUS_SYNC();
    for (int e = tid; e < nrhs * B; e += nthr) {
      const int v = e / B, r = e - v * B;
      double s = 0.0;
      for (int grp = 0; grp < ngrp; ++grp) s += part[(grp * nrhs + v) * B + r];
      A.X[(long)v * A.xstride + j * B + r] = s;
    }
  }
};
// root solve: x_0 = Dinv_0 b_0   (the backward body with no neighbours: s = 0, m such that j = 0)
struct BcrRootSolveBody {
  static VUS_DEV void run(const BcrArgs& A, int, int tid, int nthr, double* sm) {
    BcrArgs R = A;
    R.s = 0;
    BcrBwdBody::run(R, 0, tid, nthr, sm);
  }
};

}  // namespace vus
