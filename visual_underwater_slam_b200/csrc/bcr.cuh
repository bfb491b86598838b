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

// Every array the reduction itself produces (Dw, level couplings, Dinv, Gl, Gr) is stored PADDED, one [KP][LD] tile
// per supernode with zero padding -- exactly the shared-memory operand layout -- so a block is 16-byte aligned and
// moves global -> shared as ONE bulk copy (cp.async.bulk, SASS UBLKCP: the TMA engine), completion on an mbarrier.
// The assembled system (SD, SU) uses the same tile layout, so level 1 and the band operator stream it the same way.
struct BcrArgs {
  long Ns; int B; long s;          // level stride
  int small_g;                     // small-block kernels: supernodes per CTA
  long root_stride;                // the nodes left after the last level are the multiples of root_stride (>= Ns: node 0 only)
  const double* Dsrc; int d_ld; long d_stride;    // diagonal blocks read at this level (level 1: SD, plain; else Dw, padded)
  double* Dw;                      // working diagonal blocks [Ns], padded
  const double* Ucur; int u_ld; long u_stride;    // couplings at this level (level 1: SU, plain; else padded)
  double* Unext;                   // couplings of the next level, padded
  double* Dinv; double* Gl; double* Gr;   // per eliminated node, padded
  int* fail;                       // [ncomp] non-positive pivot seen in a block of this component (single graph: one entry)
  const int* node_comp; int knodes; // batched mode: component of every node, nodes per supernode (null: everything is component 0)
  // solve
  double* X; long xstride; int nrhs;
};
VUS_HD long bcr_bbp(int B) { return bcr_buf_doubles(B); }          // doubles per padded block
// where supernode j reports a failed pivot: the flag of the component its first node belongs to
VUS_HD int* bcr_fail_of(const BcrArgs& A, long j) { return A.node_comp ? A.fail + A.node_comp[j * A.knodes] : A.fail; }

#ifndef VUS_EMU
// =====================================================================================  sm_100a: DMMA tile engine
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
// ---- bulk (TMA) global -> shared copy of one padded block, completion on an mbarrier in shared memory
VUS_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
struct Mbar {
  unsigned long long* bar; unsigned parity;
  VUS_DEV void init(unsigned long long* b, int tid) {        // call once per kernel, followed by a barrier
    bar = b; parity = 0;
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  // thread 0 only, after a barrier that retired every earlier generic access to the destinations
  VUS_DEV void expect(unsigned bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  }
  VUS_DEV void copy(double* dst, const double* src, unsigned bytes) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
  }
  VUS_DEV void wait() {                                       // all threads
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
    parity ^= 1;
  }
};

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
VUS_DEV void acc_store_global(double* dst, int ld, const Acc& c, double alpha, const Tiles& G) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      if (a >= G.na || b >= G.nb) continue;
      const int i = (G.ti0 + a) * 8 + G.g, j = (G.tj0 + b) * 8 + 2 * G.t;
      if (i >= G.B) continue;
      double* p = dst + (long)i * ld + j;
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
VUS_DEV void acc_load_global(Acc& c, const double* src, int ld, const Tiles& G) {
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const int i = (G.ti0 + a) * 8 + G.g, j = (G.tj0 + b) * 8 + 2 * G.t;
      const bool in = a < G.na && b < G.nb && i < G.B;
      c[a][b][0] = (in && j < G.B) ? src[(long)i * ld + j] : (i == j ? 1.0 : 0.0);
      c[a][b][1] = (in && j + 1 < G.B) ? src[(long)i * ld + j + 1] : (i == j + 1 ? 1.0 : 0.0);
    }
}

// In-place inverse of the SPD block held in accumulator tiles: blocked Gauss-Jordan over 8 x 8 pivot tiles, no pivoting.
//   step p:  P = M[p][p]^-1 ;  M[i][j] += (-M[i][p] P) M[p][j]  (i, j != p) ;  M[p][j] = P M[p][j] ;  M[i][p] = -M[i][p] P ;  M[p][p] = P
// The rank-8 updates are DMMAs on the resident accumulators; only the pivot row / column panels go through shared
// memory (double buffered: two barriers per step).  `sm` needs VUS_GJ_DOUBLES doubles.
// A non-positive pivot raises *fail and the block's inverse is returned as ZERO: the node then passes nothing on to its
// neighbours (Gl = Gr = 0), so a bad block cannot leak Inf / NaN into the other components of a batched system.
VUS_DEV void mma_gj_inverse(Acc& c, double* sm, const Tiles& G, int* fail) {
  const int LDP = VUS_GJ_LDP, LDQ = VUS_GJ_LDQ;
  const int g = G.g, t = G.t;
  __shared__ int bad_block_;
  if (G.warp == 0 && G.lane == 0) bad_block_ = 0;      // published by the first barrier of step 0
  for (int p = 0; p < G.T; ++p) {
    double* Rraw = sm + (p & 1) * VUS_GJ_SET;
    double* Rnew = Rraw + 8 * LDP;
    double* Craw = Rnew + 8 * LDP;
    double* Cnew = Craw + 96 * LDQ;
    double* Pw = sm + 2 * VUS_GJ_SET + (p & 1) * 64;   // inverse of the pivot tile, double buffered like the panels
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
    // (b) ONE warp inverts the 8x8 pivot tile (lane r mod 8 owns row r, pivot rows travel by shuffle); the others wait at
    //     the barrier, leaving the issue slots to the co-resident CTA instead of repeating the same serial chain 8 times
    if (G.warp == (p & 7)) {
      const int r = G.lane & 7;
      double row[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) row[q] = Rraw[r * LDP + 8 * p + q];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        double prow[8];
#pragma unroll
        for (int q2 = 0; q2 < 8; ++q2) prow[q2] = __shfl_sync(0xffffffffu, row[q2], q);
        double piv = prow[q];
        if (!(piv > 0.0)) { if (G.lane == 0) { *fail = 1; bad_block_ = 1; } piv = 1.0; }   // keep the sweep finite; the flag decides
        // reciprocal by hardware seed + two Newton steps (full FP64 accuracy for normal pivots): the IEEE division's
        // slow-path checks sit on the serial chain of every sweep
        double d;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(d) : "d"(piv));
        d = fma(fma(-piv, d, 1.0), d, d);
        d = fma(fma(-piv, d, 1.0), d, d);
        const double f = row[q] * d;
        if (r == q) {
#pragma unroll
          for (int q2 = 0; q2 < 8; ++q2) row[q2] = (q2 == q) ? d : prow[q2] * d;
        } else {
#pragma unroll
          for (int q2 = 0; q2 < 8; ++q2) row[q2] = (q2 == q) ? -f : row[q2] - f * prow[q2];
        }
      }
      if (G.lane < 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) Pw[r * 8 + q] = row[q];
      }
    }
    __syncthreads();
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
  if (bad_block_) acc_zero(c);
}

// one operand of a factor kernel: padded blocks arrive by one bulk copy, the plain level-1 inputs by 8-byte cp.async
struct Staged { bool bulk; };
VUS_DEV Staged stage_any(double* buf, const double* src, int ld, const Tiles& G, Mbar& mb, int tid) {
  Staged st;
  st.bulk = ld == G.LD;
  if (st.bulk) {
    if (tid == 0) { const unsigned bytes = (unsigned)(bcr_buf_doubles(G.B) * sizeof(double)); mb.expect(bytes); mb.copy(buf, src, bytes); }
  } else {
    stage_block(buf, src, G);
  }
  return st;
}
VUS_DEV void stage_done(const Staged& st, Mbar& mb) {            // then __syncthreads() before the operand is read
  if (st.bulk) mb.wait(); else stage_wait();
}

// per eliminated node j = s*(2m+1): Dinv_j, Gl_j = U[j-s] Dinv_j, Gr_j = U[j]^T Dinv_j      (256 threads)
struct BcrElimBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int, double* sm) {
    __shared__ unsigned long long bar_;
    const Tiles G(A.B, tid);
    const long BBP = bcr_bbp(A.B);
    const long j = A.s * (2L * m + 1);
    double* buf0 = sm;
    double* buf1 = sm + bcr_buf_doubles(A.B);
    Mbar mb;
    mb.init(&bar_, tid);
    Acc c;
    acc_load_global(c, A.Dsrc + j * A.d_stride, A.d_ld, G);
    mma_gj_inverse(c, sm, G, bcr_fail_of(A, j));         // ends with a barrier (also publishes the mbarrier init)
    acc_store_global<false>(A.Dinv + j * BBP, G.LD, c, 1.0, G);
    acc_store_smem(buf1, c, G);
    Staged st = stage_any(buf0, A.Ucur + (j - A.s) * A.u_stride, A.u_ld, G, mb, tid);
    stage_done(st, mb);
    __syncthreads();
    acc_zero(c);
    mma_gemm<false, false, false>(c, buf0, buf1, G);
    if (j + A.s < A.Ns) {
      __syncthreads();
      st = stage_any(buf0, A.Ucur + j * A.u_stride, A.u_ld, G, mb, tid);   // in flight while Gl is written out
      acc_store_global<false>(A.Gl + j * BBP, G.LD, c, 1.0, G);
      stage_done(st, mb);
      __syncthreads();
      acc_zero(c);
      mma_gemm<true, false, false>(c, buf0, buf1, G);
      acc_store_global<false>(A.Gr + j * BBP, G.LD, c, 1.0, G);
    } else {
      acc_store_global<false>(A.Gl + j * BBP, G.LD, c, 1.0, G);
    }
  }
};
// per surviving node c = 2*m*s: Dw_c = D_c - Gr_{c-s} U_{c-s} - Gl_{c+s} U_c^T ; Unext_c = -Gl_{c+s} U_{c+s}
struct BcrUpdateBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int, double* sm) {
    __shared__ unsigned long long bar_;
    const Tiles G(A.B, tid);
    const long BBP = bcr_bbp(A.B);
    const unsigned blk_bytes = (unsigned)(BBP * sizeof(double));
    const long c = 2L * m * A.s;
    double* buf0 = sm;
    double* buf1 = sm + bcr_buf_doubles(A.B);
    const bool lo = c - A.s >= 0, hi = c + A.s < A.Ns;
    const bool ubulk = A.u_ld == G.LD;
    Mbar mb;
    mb.init(&bar_, tid);
    __syncthreads();
    if (lo) {                                            // Gr is always padded -> bulk; U bulk from level 2 on
      if (tid == 0) {
        mb.expect(ubulk ? 2 * blk_bytes : blk_bytes);
        mb.copy(buf0, A.Gr + (c - A.s) * BBP, blk_bytes);
        if (ubulk) mb.copy(buf1, A.Ucur + (c - A.s) * A.u_stride, blk_bytes);
      }
      if (!ubulk) stage_block(buf1, A.Ucur + (c - A.s) * A.u_stride, G);
    }
    Acc acc;
    acc_load_global(acc, A.Dsrc + c * A.d_stride, A.d_ld, G);    // D_c rides in the accumulators; the products are subtracted
    if (lo) {
      mb.wait();
      if (!ubulk) stage_wait();
      __syncthreads();
      mma_gemm<false, false, true>(acc, buf0, buf1, G);
      __syncthreads();
    }
    if (hi) {
      const long j = c + A.s;
      if (tid == 0) {
        mb.expect(ubulk ? 2 * blk_bytes : blk_bytes);
        mb.copy(buf0, A.Gl + j * BBP, blk_bytes);
        if (ubulk) mb.copy(buf1, A.Ucur + c * A.u_stride, blk_bytes);
      }
      if (!ubulk) stage_block(buf1, A.Ucur + c * A.u_stride, G);
      mb.wait();
      if (!ubulk) stage_wait();
      __syncthreads();
      mma_gemm<false, true, true>(acc, buf0, buf1, G);
      if (j + A.s < A.Ns) {
        __syncthreads();
        Staged st = stage_any(buf1, A.Ucur + j * A.u_stride, A.u_ld, G, mb, tid);   // in flight while D_c is written out
        acc_store_global<false>(A.Dw + c * BBP, G.LD, acc, 1.0, G);
        stage_done(st, mb);
        __syncthreads();
        acc_zero(acc);
        mma_gemm<false, false, true>(acc, buf0, buf1, G);
        acc_store_global<false>(A.Unext + c * BBP, G.LD, acc, 1.0, G);
        return;
      }
    }
    acc_store_global<false>(A.Dw + c * BBP, G.LD, acc, 1.0, G);
  }
};
// root: Dinv_0 = inv(D_0)
struct BcrRootBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int, double* sm) {
    const Tiles G(A.B, tid);
    const long j = (long)m * A.root_stride;
    Acc c;
    acc_load_global(c, A.Dsrc + j * A.d_stride, A.d_ld, G);
    mma_gj_inverse(c, sm, G, bcr_fail_of(A, j));
    acc_store_global<false>(A.Dinv + j * bcr_bbp(A.B), G.LD, c, 1.0, G);
  }
};


namespace rt {
template <> struct CoopBounds<BcrElimBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 2; };
template <> struct CoopBounds<BcrUpdateBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 2; };
template <> struct CoopBounds<BcrRootBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 1; };
}  // namespace rt

#else
// =====================================================================================  host emulation (tests only)
inline void emu_spd_inverse(const double* M, int ldm, double* out, int ldo, int B, int* fail) {
  std::vector<double> w((size_t)B * B), rowp(B), colp(B);
  bool failed = false;                                   // a failed block is returned as zero (see mma_gj_inverse)
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < B; ++j) w[(size_t)i * B + j] = M[(long)i * ldm + j];
  for (int p = 0; p < B; ++p) {
    for (int i = 0; i < B; ++i) { colp[i] = w[(size_t)i * B + p]; rowp[i] = w[(size_t)p * B + i]; }
    double piv = rowp[p];
    if (!(piv > 0.0)) { *fail = 1; failed = true; piv = 1.0; }
    const double d = 1.0 / piv;
    for (int i = 0; i < B; ++i) {
      double* row = w.data() + (size_t)i * B;
      if (i == p) { for (int j = 0; j < B; ++j) row[j] = (j == p) ? d : rowp[j] * d; }
      else { const double ci = colp[i] * d; for (int j = 0; j < B; ++j) row[j] = (j == p) ? -ci : row[j] - ci * rowp[j]; }
    }
  }
  bool bad = false;
  for (int p = 0; p < B && !bad; ++p) bad = !(w[(size_t)p * B + p] == w[(size_t)p * B + p]);
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < B; ++j) out[(long)i * ldo + j] = (failed || bad) ? 0.0 : w[(size_t)i * B + j];
}
// C = Cin + alpha * op(A) op(B)   (Cin may be null); every operand is a B x B block with its own row stride
inline void emu_gemm(double* C, int ldc, const double* Cin, int ldi, const double* A, int lda, bool ta, const double* Bm, int ldb, bool tb,
                     int B, double alpha) {
  std::vector<double> out((size_t)B * B);
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < B; ++j) {
      double s = 0.0;
      for (int k = 0; k < B; ++k) s += (ta ? A[(long)k * lda + i] : A[(long)i * lda + k]) * (tb ? Bm[(long)j * ldb + k] : Bm[(long)k * ldb + j]);
      out[(size_t)i * B + j] = alpha * s + (Cin ? Cin[(long)i * ldi + j] : 0.0);
    }
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < B; ++j) C[(long)i * ldc + j] = out[(size_t)i * B + j];
}
struct BcrElimBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int, int, double*) {
    const int B = A.B, LD = bcr_ld(B);
    const long BBP = bcr_bbp(B);
    const long j = A.s * (2L * m + 1);
    emu_spd_inverse(A.Dsrc + j * A.d_stride, A.d_ld, A.Dinv + j * BBP, LD, B, bcr_fail_of(A, j));
    emu_gemm(A.Gl + j * BBP, LD, nullptr, 0, A.Ucur + (j - A.s) * A.u_stride, A.u_ld, false, A.Dinv + j * BBP, LD, false, B, 1.0);
    if (j + A.s < A.Ns) emu_gemm(A.Gr + j * BBP, LD, nullptr, 0, A.Ucur + j * A.u_stride, A.u_ld, true, A.Dinv + j * BBP, LD, false, B, 1.0);
  }
};
struct BcrUpdateBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int, int, double*) {
    const int B = A.B, LD = bcr_ld(B);
    const long BBP = bcr_bbp(B);
    const long c = 2L * m * A.s;
    const double* Cin = A.Dsrc + c * A.d_stride;
    int ldi = A.d_ld;
    double* D = A.Dw + c * BBP;
    if (c - A.s >= 0) {
      emu_gemm(D, LD, Cin, ldi, A.Gr + (c - A.s) * BBP, LD, false, A.Ucur + (c - A.s) * A.u_stride, A.u_ld, false, B, -1.0);
      Cin = D; ldi = LD;
    }
    if (c + A.s < A.Ns) {
      const long j = c + A.s;
      emu_gemm(D, LD, Cin, ldi, A.Gl + j * BBP, LD, false, A.Ucur + c * A.u_stride, A.u_ld, true, B, -1.0);
      Cin = D; ldi = LD;
      if (j + A.s < A.Ns) emu_gemm(A.Unext + c * BBP, LD, nullptr, 0, A.Gl + j * BBP, LD, false, A.Ucur + j * A.u_stride, A.u_ld, false, B, -1.0);
    }
    if (Cin != D)                                            // isolated node: plain copy
      for (int i = 0; i < B; ++i)
        for (int jj = 0; jj < B; ++jj) D[(long)i * LD + jj] = Cin[(long)i * ldi + jj];
  }
};
struct BcrRootBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int, int, double*) {
    const long j = (long)m * A.root_stride;
    emu_spd_inverse(A.Dsrc + j * A.d_stride, A.d_ld, A.Dinv + j * bcr_bbp(A.B), bcr_ld(A.B), A.B, bcr_fail_of(A, j));
  }
};
#endif

// =====================================================================================  streaming block x panel kernels
// The solve sweeps and the band operator are HBM-bound: every B x B block is read once per application and multiplied
// by a thin panel of nv <= 6 vectors.  sm_100a: the block is copied global -> shared with asynchronous 8-byte copies
// (all of a thread's copies in flight at once, no registers held), the panel is a [KP][12] shared tile, and the product
// runs as m8n8k4 DMMAs (N padded to 8), so the transposed and the non-transposed products read the block the same
// coalesced way and the kernel issues ~10x fewer instructions than a shuffle-reduced dot product per (row, vector).
// One operand buffer + one panel per CTA (64.5 KB) -> three CTAs per SM overlap copy and compute.
// Launch with 256 threads; nv <= 6.
#define VUS_MAXV 6
#define VUS_LDX 12
VUS_HD long blk_slot_doubles(int B) { return bcr_buf_doubles(B) + (long)bcr_kp(B) * VUS_LDX + 8; }   // one block + its panel
VUS_HD long blk_smem_doubles(int B, int nv, int) {
#ifdef VUS_EMU
  return (long)4 * nv * B;
#else
  return bcr_buf_doubles(B) + (long)bcr_kp(B) * VUS_LDX + 8;
#endif
}

#ifndef VUS_EMU
struct Panel {          // accumulators of one warp: row tiles warp, warp + 8; columns = vectors 2t, 2t+1
  double c[2][2];
  VUS_DEV void zero() { c[0][0] = c[0][1] = c[1][0] = c[1][1] = 0.0; }
};
// stage panel x[v][k] (global, vector stride xstride) as sX[k][v], zero padded to [KP][8]
VUS_DEV void stage_panel(double* sX, const double* x, long xstride, bool present, const Tiles& G, int nv, int tid) {
  for (int e = tid; e < G.KP * 8; e += 256) {
    const int k = e >> 3, v = e & 7;
    sX[k * VUS_LDX + v] = (present && k < G.B && v < nv) ? x[(long)v * xstride + k] : 0.0;
  }
}
// acc (+/-)= op(M) Xs ; sM operand buffer [KP][LD], sX [KP][VUS_LDX]
template <bool TA, bool NEG>
VUS_DEV void panel_mma(Panel& P, const double* sM, const double* sX, const Tiles& G) {
  int ia[2];
#pragma unroll
  for (int a = 0; a < 2; ++a) { const int i = (G.warp + 8 * a) * 8 + G.g; ia[a] = i < G.B ? i : G.B - 1; }
  const bool two = G.warp + 8 < G.T;
  if (G.warp >= G.T) return;
#pragma unroll 3
  for (int k0 = 0; k0 < G.KP; k0 += 4) {
    const int kk = k0 + G.t;
    const double b = sX[kk * VUS_LDX + G.g];
    double a0 = TA ? sM[kk * G.LD + ia[0]] : sM[ia[0] * G.LD + kk];
    dmma884(P.c[0][0], P.c[0][1], NEG ? -a0 : a0, b);
    if (two) {
      double a1 = TA ? sM[kk * G.LD + ia[1]] : sM[ia[1] * G.LD + kk];
      dmma884(P.c[1][0], P.c[1][1], NEG ? -a1 : a1, b);
    }
  }
}
// one streamed product: copy block, stage panel, wait, multiply-accumulate.  Ends with a barrier (buffers reusable).
// BULK: the block is padded in global memory and arrives by one bulk (TMA) copy; else 8-byte cp.async of a plain block.
template <bool TA, bool NEG, bool BULK>
VUS_DEV void blk_stream(Panel& P, double* buf, double* sX, const double* M, const double* x, long xstride, const Tiles& G, int nv, int tid, Mbar& mb) {
  if (BULK) {
    if (tid == 0) { const unsigned bytes = (unsigned)(bcr_buf_doubles(G.B) * sizeof(double)); mb.expect(bytes); mb.copy(buf, M, bytes); }
  } else {
    stage_block(buf, M, G);
  }
  stage_panel(sX, x, xstride, true, G, nv, tid);
  if (BULK) mb.wait(); else stage_wait();
  __syncthreads();
  panel_mma<TA, NEG>(P, buf, sX, G);
  __syncthreads();
}
// visit (row r, vector v, value) of a Panel
#define VUS_PANEL_FOREACH(P, G, nv, BODY)                                   \
  _Pragma("unroll") for (int a_ = 0; a_ < 2; ++a_) {                        \
    const int r = (G.warp + 8 * a_) * 8 + G.g;                              \
    if (G.warp + 8 * a_ < G.T && r < G.B) {                                 \
      _Pragma("unroll") for (int h_ = 0; h_ < 2; ++h_) {                    \
        const int v = 2 * G.t + h_;                                         \
        if (v < nv) { const double val = P.c[a_][h_]; BODY }               \
      }                                                                     \
    }                                                                       \
  }
#else
// host emulation helpers: out[v][r] += sign * sum_k op(M)[r][k] x[v][k]
inline void emu_blk_accum(double* out, const double* M, int ld, bool ta, const double* x, long xstride, double sign, int B, int nv) {
  for (int v = 0; v < nv; ++v)
    for (int r = 0; r < B; ++r) {
      double s = 0.0;
      for (int k = 0; k < B; ++k) s += (ta ? M[(long)k * ld + r] : M[(long)r * ld + k]) * x[(long)v * xstride + k];
      out[v * B + r] += sign * s;
    }
}
#endif

// forward sweep, per surviving node c: b_c -= Gr_{c-s} b_{c-s} + Gl_{c+s} b_{c+s}
// DEEP (levels with at most one CTA per SM): one operand buffer + panel per block, all copies in flight at once --
// a lone CTA otherwise pays the DRAM latency once per block.
template <bool DEEP>
struct BcrFwdBodyT {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int, double* sm) {
    const int B = A.B, nv = A.nrhs;
    const long BBP = bcr_bbp(B);
    const long c = 2L * m * A.s;
    const long jl = c - A.s, jh = c + A.s;
#ifdef VUS_EMU
    const int LD = bcr_ld(B);
    double* dl = sm;
    for (int e = 0; e < nv * B; ++e) dl[e] = 0.0;
    if (jl >= 0) emu_blk_accum(dl, A.Gr + jl * BBP, LD, false, A.X + jl * B, A.xstride, 1.0, B, nv);
    if (jh < A.Ns) emu_blk_accum(dl, A.Gl + jh * BBP, LD, false, A.X + jh * B, A.xstride, 1.0, B, nv);
    for (int v = 0; v < nv; ++v)
      for (int r = 0; r < B; ++r) A.X[(long)v * A.xstride + c * B + r] -= dl[v * B + r];
    (void)tid;
#else
    __shared__ unsigned long long bar_;
    const Tiles G(B, tid);
    const long slot = blk_slot_doubles(B), po = bcr_buf_doubles(B);
    const unsigned blk_bytes = (unsigned)(BBP * sizeof(double));
    Mbar mb;
    mb.init(&bar_, tid);
    __syncthreads();
    Panel P;
    P.zero();
    if (DEEP) {
      if (tid == 0) {
        mb.expect(((jl >= 0) + (jh < A.Ns)) * blk_bytes);
        if (jl >= 0) mb.copy(sm, A.Gr + jl * BBP, blk_bytes);
        if (jh < A.Ns) mb.copy(sm + slot, A.Gl + jh * BBP, blk_bytes);
      }
      if (jl >= 0) stage_panel(sm + po, A.X + jl * B, A.xstride, true, G, nv, tid);
      if (jh < A.Ns) stage_panel(sm + slot + po, A.X + jh * B, A.xstride, true, G, nv, tid);
      mb.wait();
      __syncthreads();
      if (jl >= 0) panel_mma<false, false>(P, sm, sm + po, G);
      if (jh < A.Ns) panel_mma<false, false>(P, sm + slot, sm + slot + po, G);
    } else {
      double* buf = sm;
      double* sX = sm + po;
      if (jl >= 0) blk_stream<false, false, true>(P, buf, sX, A.Gr + jl * BBP, A.X + jl * B, A.xstride, G, nv, tid, mb);
      if (jh < A.Ns) blk_stream<false, false, true>(P, buf, sX, A.Gl + jh * BBP, A.X + jh * B, A.xstride, G, nv, tid, mb);
    }
    VUS_PANEL_FOREACH(P, G, nv, { A.X[(long)v * A.xstride + c * B + r] -= val; })
#endif
  }
};
typedef BcrFwdBodyT<false> BcrFwdBody;
typedef BcrFwdBodyT<true> BcrFwdDeepBody;
// backward sweep, per eliminated node j: x_j = Dinv_j b_j - Gl_j^T x_{j-s} - Gr_j^T x_{j+s}
template <bool DEEP>
struct BcrBwdBodyT {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int, double* sm) {
    const int B = A.B, nv = A.nrhs;
    const long BBP = bcr_bbp(B);
    const long j = A.s > 0 ? A.s * (2L * m + 1) : (long)m * A.root_stride;     // s = 0: a root
    const long jl = j - A.s, jh = j + A.s;
    const bool hl = jl >= 0 && A.s > 0, hh = jh < A.Ns && A.s > 0;
#ifdef VUS_EMU
    const int LD = bcr_ld(B);
    double* dl = sm;
    for (int e = 0; e < nv * B; ++e) dl[e] = 0.0;
    emu_blk_accum(dl, A.Dinv + j * BBP, LD, false, A.X + j * B, A.xstride, 1.0, B, nv);
    if (hl) emu_blk_accum(dl, A.Gl + j * BBP, LD, true, A.X + jl * B, A.xstride, -1.0, B, nv);
    if (hh) emu_blk_accum(dl, A.Gr + j * BBP, LD, true, A.X + jh * B, A.xstride, -1.0, B, nv);
    for (int v = 0; v < nv; ++v)
      for (int r = 0; r < B; ++r) A.X[(long)v * A.xstride + j * B + r] = dl[v * B + r];
    (void)tid;
#else
    __shared__ unsigned long long bar_;
    const Tiles G(B, tid);
    const long slot = blk_slot_doubles(B), po = bcr_buf_doubles(B);
    const unsigned blk_bytes = (unsigned)(BBP * sizeof(double));
    Mbar mb;
    mb.init(&bar_, tid);
    __syncthreads();
    Panel P;
    P.zero();
    if (DEEP) {
      if (tid == 0) {
        mb.expect((1 + hl + hh) * blk_bytes);
        mb.copy(sm, A.Dinv + j * BBP, blk_bytes);
        if (hl) mb.copy(sm + slot, A.Gl + j * BBP, blk_bytes);
        if (hh) mb.copy(sm + 2 * slot, A.Gr + j * BBP, blk_bytes);
      }
      stage_panel(sm + po, A.X + j * B, A.xstride, true, G, nv, tid);
      if (hl) stage_panel(sm + slot + po, A.X + jl * B, A.xstride, true, G, nv, tid);
      if (hh) stage_panel(sm + 2 * slot + po, A.X + jh * B, A.xstride, true, G, nv, tid);
      mb.wait();
      __syncthreads();
      panel_mma<false, false>(P, sm, sm + po, G);
      if (hl) panel_mma<true, true>(P, sm + slot, sm + slot + po, G);
      if (hh) panel_mma<true, true>(P, sm + 2 * slot, sm + 2 * slot + po, G);
    } else {
      double* buf = sm;
      double* sX = sm + po;
      blk_stream<false, false, true>(P, buf, sX, A.Dinv + j * BBP, A.X + j * B, A.xstride, G, nv, tid, mb);
      if (hl) blk_stream<true, true, true>(P, buf, sX, A.Gl + j * BBP, A.X + jl * B, A.xstride, G, nv, tid, mb);
      if (hh) blk_stream<true, true, true>(P, buf, sX, A.Gr + j * BBP, A.X + jh * B, A.xstride, G, nv, tid, mb);
    }
    VUS_PANEL_FOREACH(P, G, nv, { A.X[(long)v * A.xstride + j * B + r] = val; })
#endif
  }
};
typedef BcrBwdBodyT<false> BcrBwdBody;
typedef BcrBwdBodyT<true> BcrBwdDeepBody;
// root solve: x_j = Dinv_j b_j for the root(s) j = m * root_stride   (the backward body with no neighbours: s = 0)
struct BcrRootSolveBody {
  static VUS_DEV void run(const BcrArgs& A, int m, int tid, int nthr, double* sm) {
    BcrArgs R = A;
    R.s = 0;
    BcrBwdBody::run(R, m, tid, nthr, sm);
  }
};

// =====================================================================================  Kernel 3a: band operator
// one CTA per supernode I:  y_I = SD_I x_I + SU_I x_{I+1} + SU_{I-1}^T x_{I-1}   for nv vectors
struct BandMatvecBody {
  static VUS_DEV void run(const MatvecArgs& A, int I, int tid, int, double* sm) {
    const int B = A.B, nv = A.nv;
    const long BBP = bcr_bbp(B);
    const bool up = I + 1 < A.Ns, dn = I > 0;
#ifdef VUS_EMU
    const int LD = bcr_ld(B);
    double* dl = sm;
    for (int e = 0; e < nv * B; ++e) dl[e] = 0.0;
    emu_blk_accum(dl, A.SD + I * BBP, LD, false, A.x + (long)I * B, A.xstride, 1.0, B, nv);
    if (up) emu_blk_accum(dl, A.SU + I * BBP, LD, false, A.x + (long)(I + 1) * B, A.xstride, 1.0, B, nv);
    if (dn) emu_blk_accum(dl, A.SU + (I - 1) * BBP, LD, true, A.x + (long)(I - 1) * B, A.xstride, 1.0, B, nv);
    for (int v = 0; v < nv; ++v)
      for (int r = 0; r < B; ++r) A.y[(long)v * A.ystride + (long)I * B + r] = dl[v * B + r];
    (void)tid;
#else
    __shared__ unsigned long long bar_;
    const Tiles G(B, tid);
    double* buf = sm;
    double* sX = sm + bcr_buf_doubles(B);
    Mbar mb;
    mb.init(&bar_, tid);
    __syncthreads();
    Panel P;
    P.zero();
    blk_stream<false, false, true>(P, buf, sX, A.SD + I * BBP, A.x + (long)I * B, A.xstride, G, nv, tid, mb);
    if (up) blk_stream<false, false, true>(P, buf, sX, A.SU + I * BBP, A.x + (long)(I + 1) * B, A.xstride, G, nv, tid, mb);
    if (dn) blk_stream<true, false, true>(P, buf, sX, A.SU + (I - 1) * BBP, A.x + (long)(I - 1) * B, A.xstride, G, nv, tid, mb);
    VUS_PANEL_FOREACH(P, G, nv, { A.y[(long)v * A.ystride + (long)I * B + r] = val; })
#endif
  }
};

// =====================================================================================  small blocks (B <= 16)
// Pure chain graphs (no landmark tracks: k = 1, B = D = 6 or 9 -- BASELINE configs 1, 4, 5) have hundreds of thousands of
// tiny supernodes; a 256-thread CTA with a TMA pipeline per 9 x 9 block spends its time on CTA launch and barrier
// latency.  Here one thread owns one (supernode, row, vector) item and a CTA covers VUS_SMALLB_G supernodes: a tile is
// read once, by neighbouring lanes (rows are contiguous; transposed operands are read along the lanes), straight from
// global memory.  Same recurrences as the streaming bodies above; single source for the device and the emulation.
#define VUS_SMALLB_MAX 16
#define VUS_SMALLB_G 16
#define VUS_SMALLB_MAXV 8
template <int BT>           // BT: compile-time block size (0: A.B) so the k loops unroll and their loads overlap
struct SmallFwdBody {      // per surviving node c = 2 m s:  b_c -= Gr_{c-s} b_{c-s} + Gl_{c+s} b_{c+s}
  static VUS_DEV void run(const BcrArgs& A, int blk, int tid, int nthr, double*) {
    const int B = BT ? BT : A.B, nv = A.nrhs, LD = bcr_ld(B), GN = A.small_g;
    const long BBP = bcr_bbp(B);
    const long nsv = ((A.Ns + A.s - 1) / A.s + 1) / 2;
    for (int e = tid; e < GN * B; e += nthr) {           // item = (node, row): the tile row is read once for all vectors
      const int r = e % B, g = e / B;
      const long m = (long)blk * GN + g;
      if (m >= nsv) continue;
      const long c = 2L * m * A.s, jl = c - A.s, jh = c + A.s;
      double acc[VUS_SMALLB_MAXV];
#pragma unroll
      for (int v = 0; v < VUS_SMALLB_MAXV; ++v) acc[v] = 0.0;
      if (jl >= 0) {
        const double* G = A.Gr + jl * BBP + (long)r * LD;
#pragma unroll
        for (int k = 0; k < B; ++k) {
          const double gk = G[k];
#pragma unroll
          for (int v = 0; v < VUS_SMALLB_MAXV; ++v) if (v < nv) acc[v] += gk * A.X[(long)v * A.xstride + jl * B + k];
        }
      }
      if (jh < A.Ns) {
        const double* G = A.Gl + jh * BBP + (long)r * LD;
#pragma unroll
        for (int k = 0; k < B; ++k) {
          const double gk = G[k];
#pragma unroll
          for (int v = 0; v < VUS_SMALLB_MAXV; ++v) if (v < nv) acc[v] += gk * A.X[(long)v * A.xstride + jh * B + k];
        }
      }
#pragma unroll
      for (int v = 0; v < VUS_SMALLB_MAXV; ++v) if (v < nv) A.X[(long)v * A.xstride + c * B + r] -= acc[v];
    }
  }
};
template <int BT>
struct SmallBwdBody {      // per eliminated node j = s (2 m + 1):  x_j = Dinv_j b_j - Gl_j^T x_{j-s} - Gr_j^T x_{j+s};  s = 0: the root
  static VUS_DEV void run(const BcrArgs& A, int blk, int tid, int nthr, double* sm) {
    const int B = BT ? BT : A.B, nv = A.nrhs, LD = bcr_ld(B), GN = A.small_g;
    const long BBP = bcr_bbp(B);
    const long nel = A.s > 0 ? ((A.Ns + A.s - 1) / A.s) / 2 : (A.Ns + A.root_stride - 1) / A.root_stride;
    for (int e = tid; e < GN * B; e += nthr) {
      const int r = e % B, g = e / B;
      const long m = (long)blk * GN + g;
      if (m >= nel) continue;
      const long j = A.s > 0 ? A.s * (2L * m + 1) : m * A.root_stride, jl = j - A.s, jh = j + A.s;
      const double* Dj = A.Dinv + j * BBP + (long)r * LD;
      double acc[VUS_SMALLB_MAXV];
#pragma unroll
      for (int v = 0; v < VUS_SMALLB_MAXV; ++v) acc[v] = 0.0;
#pragma unroll
      for (int k = 0; k < B; ++k) {
        const double dk = Dj[k];
#pragma unroll
        for (int v = 0; v < VUS_SMALLB_MAXV; ++v) if (v < nv) acc[v] += dk * A.X[(long)v * A.xstride + j * B + k];
      }
      if (A.s > 0 && jl >= 0) {
        const double* G = A.Gl + j * BBP + r;
#pragma unroll
        for (int k = 0; k < B; ++k) {
          const double gk = G[(long)k * LD];
#pragma unroll
          for (int v = 0; v < VUS_SMALLB_MAXV; ++v) if (v < nv) acc[v] -= gk * A.X[(long)v * A.xstride + jl * B + k];
        }
      }
      if (A.s > 0 && jh < A.Ns) {
        const double* G = A.Gr + j * BBP + r;
#pragma unroll
        for (int k = 0; k < B; ++k) {
          const double gk = G[(long)k * LD];
#pragma unroll
          for (int v = 0; v < VUS_SMALLB_MAXV; ++v) if (v < nv) acc[v] -= gk * A.X[(long)v * A.xstride + jh * B + k];
        }
      }
#pragma unroll
      for (int v = 0; v < VUS_SMALLB_MAXV; ++v) if (v < nv) sm[v * GN * B + e] = acc[v];
    }
    VUS_SYNC();                                           // x_j overwrites b_j: every row of the node is computed first
    for (int e = tid; e < GN * B * nv; e += nthr) {
      const int r = e % B, g = (e / B) % GN, v = e / (B * GN);
      const long m = (long)blk * GN + g;
      if (m >= nel) continue;
      A.X[(long)v * A.xstride + (A.s > 0 ? A.s * (2L * m + 1) : m * A.root_stride) * B + r] = sm[e];
    }
  }
};
// factorization, per eliminated node j = s (2 m + 1): Dinv_j = D_j^-1 (unpivoted Gauss-Jordan in shared memory, the nodes of
// a CTA advance pivot by pivot in lock step), Gl_j = U_{j-s} Dinv_j, Gr_j = U_j^T Dinv_j.   smem: G * (B * B + 1) doubles
struct SmallElimBody {
  static VUS_DEV void run(const BcrArgs& A, int blk, int tid, int nthr, double* sm) {
    const int B = A.B, LD = bcr_ld(B), BB = B * B;
    const long BBP = bcr_bbp(B);
    const long nel = ((A.Ns + A.s - 1) / A.s) / 2;
    double* piv = sm + VUS_SMALLB_G * BB;
    VUS_SHARED int bad_node_[VUS_SMALLB_G];              // a node with a non-positive pivot gets a ZERO inverse (see mma_gj_inverse)
    for (int g = tid; g < VUS_SMALLB_G; g += nthr) bad_node_[g] = 0;
    for (int e = tid; e < VUS_SMALLB_G * BB; e += nthr) {
      const int g = e / BB, r = (e % BB) / B, c = e % B;
      const long m = (long)blk * VUS_SMALLB_G + g;
      sm[e] = m < nel ? A.Dsrc[A.s * (2L * m + 1) * A.d_stride + (long)r * A.d_ld + c] : (r == c ? 1.0 : 0.0);
    }
    VUS_SYNC();
    for (int p = 0; p < B; ++p) {
      for (int g = tid; g < VUS_SMALLB_G; g += nthr) {
        double v = sm[g * BB + p * B + p];
        const long mg = (long)blk * VUS_SMALLB_G + g;
        if (!(v > 0.0)) { if (mg < nel) { *bcr_fail_of(A, A.s * (2L * mg + 1)) = 1; bad_node_[g] = 1; } v = 1.0; }
        piv[g] = 1.0 / v;
      }
      VUS_SYNC();
      for (int e = tid; e < VUS_SMALLB_G * B; e += nthr) {       // pivot row
        const int g = e / B, c = e % B;
        double* row = sm + g * BB + p * B;
        row[c] = (c == p) ? piv[g] : row[c] * piv[g];
      }
      VUS_SYNC();
      for (int e = tid; e < VUS_SMALLB_G * B; e += nthr) {       // the other rows
        const int g = e / B, r = e % B;
        if (r == p) continue;
        double* M = sm + g * BB;
        const double f = M[r * B + p];
        for (int c = 0; c < B; ++c) M[r * B + c] = (c == p) ? -f * M[p * B + p] : M[r * B + c] - f * M[p * B + c];
      }
      VUS_SYNC();
    }
    for (int e = tid; e < VUS_SMALLB_G * BB; e += nthr) if (bad_node_[e / BB]) sm[e] = 0.0;
    VUS_SYNC();
    for (int e = tid; e < VUS_SMALLB_G * BB; e += nthr) {
      const int g = e / BB, r = (e % BB) / B, c = e % B;
      const long m = (long)blk * VUS_SMALLB_G + g;
      if (m >= nel) continue;
      const long j = A.s * (2L * m + 1);
      const double* Di = sm + g * BB;
      A.Dinv[j * BBP + (long)r * LD + c] = Di[r * B + c];
      const double* Ul = A.Ucur + (j - A.s) * A.u_stride + (long)r * A.u_ld;
      double gl = 0.0;
      for (int k = 0; k < B; ++k) gl += Ul[k] * Di[k * B + c];
      A.Gl[j * BBP + (long)r * LD + c] = gl;
      if (j + A.s < A.Ns) {
        const double* Uj = A.Ucur + j * A.u_stride + r;
        double gr = 0.0;
        for (int k = 0; k < B; ++k) gr += Uj[(long)k * A.u_ld] * Di[k * B + c];
        A.Gr[j * BBP + (long)r * LD + c] = gr;
      }
    }
  }
};
// per surviving node c = 2 m s, work item (m, r, col):  Dw_c = D_c - Gr_{c-s} U_{c-s} - Gl_{c+s} U_c^T ;  Unext_c = -Gl_{c+s} U_{c+s}
struct SmallUpdateBody {
  static VUS_DEV void run(const BcrArgs& A, long w) {
    const int B = A.B, LD = bcr_ld(B);
    const long BBP = bcr_bbp(B);
    const int col = (int)(w % B), r = (int)((w / B) % B);
    const long c = 2L * (w / ((long)B * B)) * A.s;
    double acc = A.Dsrc[c * A.d_stride + (long)r * A.d_ld + col];
    if (c - A.s >= 0) {
      const double* G = A.Gr + (c - A.s) * BBP + (long)r * LD;
      const double* U = A.Ucur + (c - A.s) * A.u_stride + col;
      for (int k = 0; k < B; ++k) acc -= G[k] * U[(long)k * A.u_ld];
    }
    if (c + A.s < A.Ns) {
      const long j = c + A.s;
      const double* G = A.Gl + j * BBP + (long)r * LD;
      const double* U = A.Ucur + c * A.u_stride + (long)col * A.u_ld;
      for (int k = 0; k < B; ++k) acc -= G[k] * U[k];
      if (j + A.s < A.Ns) {
        const double* U2 = A.Ucur + j * A.u_stride + col;
        double un = 0.0;
        for (int k = 0; k < B; ++k) un -= G[k] * U2[(long)k * A.u_ld];
        A.Unext[c * BBP + (long)r * LD + col] = un;
      }
    }
    A.Dw[c * BBP + (long)r * LD + col] = acc;
  }
};
template <int BT>
struct SmallMatvecBody {   // work item (v, I, r):  y_I[r] = SD_I[r,:] x_I + SU_I[r,:] x_{I+1} + SU_{I-1}[:,r] x_{I-1}
  static VUS_DEV void run(const MatvecArgs& A, long w) {
    const int B = BT ? BT : A.B, LD = bcr_ld(B);
    const long BBP = bcr_bbp(B);
    const int r = (int)(w % B);
    const long I = (w / B) % A.Nrows;
    const int v = (int)(w / ((long)B * A.Nrows));
    const double* x = A.x + (long)v * A.xstride;
    const double* d = A.SD + I * BBP + (long)r * LD;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < B; ++k) acc += d[k] * x[I * B + k];
    if (I + 1 < A.Ns) {
      const double* u = A.SU + I * BBP + (long)r * LD;
#pragma unroll
      for (int k = 0; k < B; ++k) acc += u[k] * x[(I + 1) * B + k];
    }
    if (I > 0) {
      const double* u = A.SU + (I - 1) * BBP + r;
#pragma unroll
      for (int k = 0; k < B; ++k) acc += u[(long)k * LD] * x[(I - 1) * B + k];
    }
    A.y[(long)v * A.ystride + I * B + r] = acc;
  }
};

#ifndef VUS_EMU
namespace rt {
template <> struct CoopBounds<BcrFwdBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 3; };
template <> struct CoopBounds<BcrBwdBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 3; };
template <> struct CoopBounds<BcrFwdDeepBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 1; };
template <> struct CoopBounds<BcrBwdDeepBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 1; };
template <> struct CoopBounds<BcrRootSolveBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 3; };
template <> struct CoopBounds<BandMatvecBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 3; };
}  // namespace rt
#endif

}  // namespace vus
