// Kernel 3b' -- the band factorization at stereo-graph scale: block Cholesky inside chunks, cyclic reduction across separators.
//
// Block cyclic reduction (bcr.cuh) of the WHOLE supernode chain costs ~12 B^3 flops per supernode (explicit inverses, three
// unsymmetric products per eliminated and per surviving node): 71 GFLOP at BASELINE config 3 against the ~7 GFLOP a
// sequential band Cholesky needs -- but a sequential sweep over 11 112 supernodes cannot fill a GPU.  Here the chain is cut
// into P chunks (one per SM) separated by single supernodes:
//
//     [ chunk 0 ] s0 [ chunk 1 ] s1 [ chunk 2 ] ... s(P-2) [ chunk P-1 ]
//
// Every CTA eliminates the interior of its chunk left to right with a block Cholesky (what gtsam's multifrontal elimination
// does along the chain under /root/reference/batch.py:337); the coupling to the separator on its left travels along as a
// "spike" S_i.  With D_i' the running diagonal block, U_i = A(i, i+1) and S_i = A'(sepL, i):
//
//     D_i' = L_i L_i^T ,  Linv_i = L_i^-1                                       (blocked Cholesky + triangular inverse)
//     X_i  = U_i^T Linv_i^T        (= L(i+1, i))          D_{i+1}' = D_{i+1} - X_i X_i^T
//     W_i  = S_i  Linv_i^T        (= L(sepL, i))          S_{i+1}  = -W_i X_i^T ,   E += W_i W_i^T
//
// The last node of a chunk hands X X^T to the separator on its right (SepR), E to the one on its left (SepL) and -W X^T
// becomes the coupling between the two separators (SepU): the separators form a block-tridiagonal system of P-1 supernodes,
// which the cyclic reduction of bcr.cuh factors in log2(P) levels.  Flops per supernode: ~B^3/3 (Cholesky) + B^3/3 (inverse)
// + 5 products of which every one is triangular or symmetric -- and the original couplings U_i are block-lower-triangular in
// D x D node blocks (a factor reaches at most k nodes ahead), which the products skip: ~2.6 B^3 instead of 12 B^3.
//
// Solve:  forward  y_i = Linv_i (b_i - X_{i-1} y_{i-1}),  separator right-hand sides b_s -= X_last y_last + sum_i W_i y_i;
//         separators by cyclic reduction;  backward  x_i = Linv_i^T (y_i - X_i^T x_{i+1} - W_i^T x_sepL).
//
// sm_100a: one 512-thread CTA per chunk, every product on the FP64 tensor path (mma.sync.m8n8k4.f64, SASS DMMA) from three
// shared-memory tile buffers (169 KB), the next coupling tile arriving by bulk (TMA) copy while the Cholesky runs; the
// sweeps stream Linv / X / W tiles through a three-slot TMA ring.  The VUS_EMU build (tests only) runs the same recurrences
// with plain loops.
#pragma once
#include "bcr.cuh"

namespace vus {

struct ChunkGeom { long Ns; int P; long base; int rem; };
VUS_HD ChunkGeom chunk_geom(long Ns, int P) {
  ChunkGeom g; g.Ns = Ns; g.P = P;
  const long interior = Ns - (P - 1);
  g.base = interior / P; g.rem = (int)(interior % P);
  return g;
}
VUS_HD long chunk_len(const ChunkGeom& g, int c) { return g.base + (c < g.rem ? 1 : 0); }
VUS_HD long chunk_first(const ChunkGeom& g, int c) { return (long)c * (g.base + 1) + (c < g.rem ? c : g.rem); }
VUS_HD long chunk_sep(const ChunkGeom& g, int c) { return chunk_first(g, c) + chunk_len(g, c); }   // separator right of chunk c (c < P-1)

struct ChunkArgs {
  ChunkGeom G; int B, D;            // D: node dof -- the couplings SU are block-lower-triangular in D x D blocks
  const double* SD; const double* SU;        // assembled band, padded [KP][LD] tiles
  double* Linv; double* X; double* W;        // per interior node (indexed by supernode), padded tiles
  double* SepL; double* SepR; double* SepU;  // per chunk: contributions to the separator system, padded tiles
  double* Dsep; double* Usep;                // the separator system [P-1] / [P-2], padded tiles
  int* fail;
  // solve
  double* Xv; long xstride; int nrhs;        // right-hand sides / solutions, in place
  double* tL; double* tR;                    // [P][8][B] separator right-hand-side contributions of every chunk
  double* xsep; long sepstride;              // compact separator vectors [nrhs][(P-1) B]
};
#define VUS_CHUNK_MAXV 8

// separator system from the chunk contributions:  Dsep_q = SD[sep_q] - SepR[q] - SepL[q+1] ;  Usep_q = SepU[q+1]
struct SepAssembleBody {
  static VUS_DEV void run(const ChunkArgs& A, long w) {
    const long BBP = bcr_bbp(A.B);
    const long q = w / BBP, e = w % BBP;
    A.Dsep[w] = A.SD[chunk_sep(A.G, (int)q) * BBP + e] - A.SepR[q * BBP + e] - A.SepL[(q + 1) * BBP + e];
    if (q + 2 < A.G.P) A.Usep[w] = A.SepU[(q + 1) * BBP + e];
  }
};
// compact separator right-hand sides:  xsep[v][q B + r] = b[sep_q][r] - tR[q][v][r] - tL[q+1][v][r]
struct SepRhsBody {
  static VUS_DEV void run(const ChunkArgs& A, long w) {
    const int B = A.B;
    const long nsep = A.G.P - 1;
    const int r = (int)(w % B);
    const long q = (w / B) % nsep;
    const int v = (int)(w / (B * nsep));
    A.xsep[(long)v * A.sepstride + q * B + r] = A.Xv[(long)v * A.xstride + chunk_sep(A.G, (int)q) * B + r]
        - A.tR[(q * VUS_CHUNK_MAXV + v) * B + r] - A.tL[((q + 1) * VUS_CHUNK_MAXV + v) * B + r];
  }
};

#ifndef VUS_EMU
// =====================================================================================  sm_100a
// 512 threads = 16 warps (8 measured slower: 105 vs 86 ms of factorization per config-3 solve).  Lower-triangle tiles (Cholesky, D', E)
// are owned cyclically: tile idx = ti (ti + 1) / 2 + tj belongs to warp idx % 16, slot idx / 16 (at most 5 slots for T <= 12).  Full
// products run on a 4 x 4 warp grid of 3 x 3 tile blocks.  Inside a tile lane (g = lane / 4, t = lane % 4) holds (row g, cols 2t, 2t + 1), as in bcr.cuh.
#ifndef VUS_CH_WARPS
#define VUS_CH_WARPS 16
#endif
#define VUS_CH_THREADS (32 * VUS_CH_WARPS)
#define VUS_CH_LSLOTS ((78 + VUS_CH_WARPS - 1) / VUS_CH_WARPS)
#define VUS_CH_FBC (VUS_CH_WARPS == 16 ? 3 : 6)      /* tile columns per warp block (rows: 3) */
#define VUS_CH_FSLOTS (3 * VUS_CH_FBC)
#define VUS_CH_LDQ 12
#define VUS_CH_LDP 100
// scratch of the blocked Cholesky: pivot tile inverse [64] | Craw, Cnew [96][12] | Rraw, Rnew [8][100]
#define VUS_CH_SCRATCH (64 + 2 * 96 * VUS_CH_LDQ + 2 * 8 * VUS_CH_LDP)
VUS_HD size_t chunk_factor_smem(int B) { return (size_t)(3 * bcr_buf_doubles(B) + VUS_CH_SCRATCH) * sizeof(double); }

struct CT {
  int B, T, KP, LD, warp, lane, g, t;
  int lti[VUS_CH_LSLOTS], ltj[VUS_CH_LSLOTS], nl;      // this warp's lower-triangle tiles
  VUS_DEV CT(int B_, int tid) {
    B = B_; T = bcr_tiles(B); KP = bcr_kp(B); LD = bcr_ld(B);
    warp = tid >> 5; lane = tid & 31; g = lane >> 2; t = lane & 3;
    nl = 0;
    const int NT = T * (T + 1) / 2;
#pragma unroll
    for (int s = 0; s < VUS_CH_LSLOTS; ++s) {
      const int idx = warp + VUS_CH_WARPS * s;
      lti[s] = ltj[s] = -1;
      if (idx < NT) {
        int ti = 0;
        while ((ti + 1) * (ti + 2) / 2 <= idx) ++ti;
        lti[s] = ti; ltj[s] = idx - ti * (ti + 1) / 2; nl = s + 1;
      }
    }
  }
};
typedef double LAcc[VUS_CH_LSLOTS][2];
typedef double FAcc[VUS_CH_FSLOTS][2];

// one output tile: c += sign * sum_{k in [kbeg, kend)} opA(ti, k) opB(k, tj).   TA: sA holds A^T;  TB: sB holds B^T.
template <bool TA, bool TB, bool NEG>
VUS_DEV void tile_mma(double& c0, double& c1, const double* sA, const double* sB, int ti, int tj, int kbeg, int kend, const CT& G) {
  int ia = ti * 8 + G.g, jb = tj * 8 + G.g;
  ia = ia < G.B ? ia : G.B - 1;                       // padded rows / columns re-read the last real one; their results are never kept
  jb = jb < G.B ? jb : G.B - 1;
  const int LD = G.LD;
#pragma unroll 4
  for (int k = kbeg + G.t; k < kend; k += 4) {
    const double a = TA ? sA[k * LD + ia] : sA[ia * LD + k];
    const double b = TB ? sB[jb * LD + k] : sB[k * LD + jb];
    dmma884(c0, c1, NEG ? -a : a, b);
  }
}
// Full T x T product on a 4 x 2 warp grid of 3 x 6 tile blocks (8 warps; 4 x 4 of 3 x 3 with 16): per k step a warp loads three A and
// six B fragments for eighteen independent DMMAs.  kb / ke give every TILE its own k range (structural
// zeros of the operands are skipped); a block runs over the union.
struct KRange { int kb[VUS_CH_FSLOTS], ke[VUS_CH_FSLOTS], kmin, kmax; };
VUS_DEV int fblock_r(const CT& G) { return VUS_CH_WARPS == 16 ? (G.warp >> 2) : (G.warp >> 1); }
VUS_DEV int fblock_c(const CT& G) { return VUS_CH_WARPS == 16 ? (((G.warp & 3) - (G.warp >> 2)) & 3) : ((G.warp ^ (G.warp >> 2)) & 1); }
template <bool TA, bool TB, bool NEG>
VUS_DEV void block_mma(FAcc& r, const double* sA, const double* sB, const KRange& K, const CT& G) {
  const int wr = fblock_r(G), wc = fblock_c(G), LD = G.LD;
  int ia[3], jb[VUS_CH_FBC];
#pragma unroll
  for (int a = 0; a < 3; ++a) { const int i = (3 * wr + a) * 8 + G.g; ia[a] = i < G.B ? i : G.B - 1; }   // padded rows / columns re-read the
#pragma unroll
  for (int b = 0; b < VUS_CH_FBC; ++b) { const int j = (VUS_CH_FBC * wc + b) * 8 + G.g; jb[b] = j < G.B ? j : G.B - 1; }   // last real one
#pragma unroll
  for (int s = 0; s < VUS_CH_FSLOTS; ++s) r[s][0] = r[s][1] = 0.0;
#pragma unroll 2
  for (int k0 = K.kmin; k0 < K.kmax; k0 += 4) {
    const int k = k0 + G.t;
    double af[3], bf[VUS_CH_FBC];
#pragma unroll
    for (int a = 0; a < 3; ++a) { const double v = TA ? sA[k * LD + ia[a]] : sA[ia[a] * LD + k]; af[a] = NEG ? -v : v; }
#pragma unroll
    for (int b = 0; b < VUS_CH_FBC; ++b) bf[b] = TB ? sB[jb[b] * LD + k] : sB[k * LD + jb[b]];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < VUS_CH_FBC; ++b)
        if (k0 >= K.kb[VUS_CH_FBC * a + b] && k0 < K.ke[VUS_CH_FBC * a + b]) dmma884(r[VUS_CH_FBC * a + b][0], r[VUS_CH_FBC * a + b][1], af[a], bf[b]);
  }
}
// The warps of one scheduler (warp % 4: each SM sub-partition has its own FP64 tensor unit) get different row blocks AND different
// column blocks, so k ranges that shrink along rows or along columns load the four tensor units evenly.
VUS_DEV int ftile_i(const CT& G, int s) { return 3 * fblock_r(G) + s / VUS_CH_FBC; }
VUS_DEV int ftile_j(const CT& G, int s) { return VUS_CH_FBC * fblock_c(G) + s % VUS_CH_FBC; }
// symmetric product on the lower tiles a warp owns, all slots interleaved (independent accumulators):
//   c[s] += sign * sum_{k >= kb[s]} A[ti][k] A[tj][k]
template <bool NEG>
VUS_DEV void lower_syrk(LAcc& c, const double* sA, const int (&kb)[VUS_CH_LSLOTS], int kmin, const CT& G) {
  const int LD = G.LD;
  int ia[VUS_CH_LSLOTS], jb[VUS_CH_LSLOTS];
#pragma unroll
  for (int s = 0; s < VUS_CH_LSLOTS; ++s) {
    int i = G.lti[s] * 8 + G.g, j = G.ltj[s] * 8 + G.g;
    ia[s] = (s < G.nl && i < G.B) ? i : G.B - 1;
    jb[s] = (s < G.nl && j < G.B) ? j : G.B - 1;
  }
#pragma unroll 2
  for (int k0 = kmin; k0 < G.KP; k0 += 4) {
    const int k = k0 + G.t;
#pragma unroll
    for (int s = 0; s < VUS_CH_LSLOTS; ++s) {
      if (s >= G.nl || k0 < kb[s]) continue;
      const double a = sA[ia[s] * LD + k], b = sA[jb[s] * LD + k];
      dmma884(c[s][0], c[s][1], NEG ? -a : a, b);
    }
  }
}
// accumulator tile -> operand buffer [KP][LD] (zero outside the B x B block), optionally also to a padded global tile
VUS_DEV void tile_store(double* s, double c0, double c1, int ti, int tj, const CT& G) {
  const int i = ti * 8 + G.g, j = tj * 8 + 2 * G.t;
  if (i >= G.KP || j >= G.LD) return;                  // LD and j are even: the pair (j, j + 1) is inside the row and 16-byte aligned
  const bool ri = i < G.B;
  double2 v;
  v.x = (ri && j < G.B) ? c0 : 0.0;
  v.y = (ri && j + 1 < G.B) ? c1 : 0.0;
  *reinterpret_cast<double2*>(s + i * G.LD + j) = v;
}
// finished operand buffer -> padded global tile as ONE asynchronous bulk (TMA) store.  Thread 0, after a barrier that made the
// buffer's generic-proxy writes visible; bulk_store_wait_read() before the buffer is overwritten.
VUS_DEV void bulk_store(double* gl, const double* s, unsigned bytes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gl), "r"(smem_u32(s)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
VUS_DEV void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
VUS_DEV void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// symmetric result held as lower tiles -> full padded global tile (the tile and its mirror image)
VUS_DEV void ltile_store_global_sym(double* gl, double c0, double c1, int ti, int tj, double alpha, const CT& G) {
  const int i = ti * 8 + G.g, j = tj * 8 + 2 * G.t;
  if (i >= G.B) return;
  if (j < G.B) { gl[i * G.LD + j] = alpha * c0; gl[j * G.LD + i] = alpha * c0; }
  if (j + 1 < G.B) { gl[i * G.LD + j + 1] = alpha * c1; gl[(j + 1) * G.LD + i] = alpha * c1; }
}
// lower tiles of an SPD block from a padded global tile, identity on the padding
VUS_DEV void ltile_load(double& c0, double& c1, const double* src, int ti, int tj, const CT& G) {
  const int i = ti * 8 + G.g, j = tj * 8 + 2 * G.t;
  c0 = (i < G.B && j < G.B) ? src[i * G.LD + j] : (i == j ? 1.0 : 0.0);
  c1 = (i < G.B && j + 1 < G.B) ? src[i * G.LD + j + 1] : (i == j + 1 ? 1.0 : 0.0);
}
VUS_DEV void ltile_fix_padding(double& c0, double& c1, int ti, int tj, const CT& G) {
  const int i = ti * 8 + G.g, j = tj * 8 + 2 * G.t;
  if (i >= G.B || j >= G.B) c0 = (i == j ? 1.0 : 0.0);
  if (i >= G.B || j + 1 >= G.B) c1 = (i == j + 1 ? 1.0 : 0.0);
}

// 1 / sqrt(d): hardware seed + two Newton steps (full FP64 accuracy for normal d); the IEEE sqrt / division slow paths would
// sit on the serial chain of every pivot
VUS_DEV double fast_rsqrt(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  e = fma(-d * y, y, 1.0);
  y = fma(0.5 * y, e, y);
  return y;
}
// 8 x 8 SPD tile M (rows at `src`, row stride `ld`) -> P^-1 with M = P P^T (P lower), written row-major to Pinv[64].
// One warp: lane r (mod 8) owns row r of the tile AND row r of the inverse, which is built alongside the right-looking
// Cholesky (forward elimination of [M | I]); only the pivot's reciprocal square root and one broadcast sit on the serial chain.
VUS_DEV void tile_chol_inv(const double* src, int ld, double* Pinv, int lane, int* fail) {
  const int r = lane & 7;
  double row[8], z[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { row[q] = src[r * ld + q]; z[q] = (q == r) ? 1.0 : 0.0; }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    double d = __shfl_sync(0xffffffffu, row[q], q);
    if (!(d > 0.0)) { if (lane == 0) *fail = 1; d = 1.0; }
    const double inv = fast_rsqrt(d);
    const double l = row[q] * inv;                   // L[r][q] (meaningful for r >= q)
#pragma unroll
    for (int j = q + 1; j < 8; ++j) row[j] -= l * __shfl_sync(0xffffffffu, l, j);
    // row q of the inverse is final: Pinv[q][c] = z_q[c] / L[q][q]; the rows below eliminate it
#pragma unroll
    for (int c = 0; c <= q; ++c) {
      const double pq = __shfl_sync(0xffffffffu, z[c], q) * inv;
      if (r == q) z[c] = pq;
      else if (r > q) z[c] -= l * pq;
    }
  }
  if (lane < 8) {
#pragma unroll
    for (int c = 0; c < 8; ++c) Pinv[r * 8 + c] = (c <= r) ? z[c] : 0.0;
  }
}

// In place: the SPD block held as lower tiles -> Linv = L^-1 (lower tiles), D = L L^T.  Blocked right-looking Cholesky over
// 8 x 8 pivot tiles that builds the inverse alongside: after step p, tile (i, j) holds  the trailing matrix entry (j > p),
// or Y_ij = -sum_{q <= p} L_iq Linv_qj (j <= p, i > p), or the final Linv_ij (i <= p).
VUS_DEV void chunk_chol_inv(LAcc& c, double* scr, const CT& G, int* fail) {
  const int LDQ = VUS_CH_LDQ, LDP = VUS_CH_LDP, g = G.g, t = G.t;
  double* Pinv = scr;
  double* Craw = scr + 64;
  double* Cnew = Craw + 96 * LDQ;
  double* Rraw = Cnew + 96 * LDQ;
  double* Rnew = Rraw + 8 * LDP;
  for (int p = 0; p < G.T; ++p) {
    // (1) owners publish the raw pivot column (tiles (i, p), i >= p) and pivot row (tiles (p, j), j < p)
#pragma unroll
    for (int s = 0; s < VUS_CH_LSLOTS; ++s) {
      if (s >= G.nl) continue;
      const int ti = G.lti[s], tj = G.ltj[s];
      if (tj == p && ti >= p) { Craw[(8 * ti + g) * LDQ + 2 * t] = c[s][0]; Craw[(8 * ti + g) * LDQ + 2 * t + 1] = c[s][1]; }
      if (ti == p && tj < p) { Rraw[g * LDP + 8 * tj + 2 * t] = c[s][0]; Rraw[g * LDP + 8 * tj + 2 * t + 1] = c[s][1]; }
    }
    __syncthreads();
    // (2) one warp factors and inverts the pivot tile
    if (G.warp == (p & (VUS_CH_WARPS - 1))) tile_chol_inv(Craw + 8 * p * LDQ, LDQ, Pinv, G.lane, fail);
    __syncthreads();
    // (3) panels, spread over the warps:  Cnew_i = Craw_i Pinv^T (i > p) ;  Rnew_j = Pinv Rraw_j (j < p) ;  Rnew_p = Pinv
    for (int item = G.warp; item < G.T; item += VUS_CH_WARPS) {
      double c0 = 0.0, c1 = 0.0;
      if (item > p) {
        dmma884(c0, c1, Craw[(8 * item + g) * LDQ + t], Pinv[g * 8 + t]);
        dmma884(c0, c1, Craw[(8 * item + g) * LDQ + 4 + t], Pinv[g * 8 + 4 + t]);
        Cnew[(8 * item + g) * LDQ + 2 * t] = c0; Cnew[(8 * item + g) * LDQ + 2 * t + 1] = c1;
      } else if (item < p) {
        dmma884(c0, c1, Pinv[g * 8 + t], Rraw[t * LDP + 8 * item + g]);
        dmma884(c0, c1, Pinv[g * 8 + 4 + t], Rraw[(4 + t) * LDP + 8 * item + g]);
        Rnew[g * LDP + 8 * item + 2 * t] = c0; Rnew[g * LDP + 8 * item + 2 * t + 1] = c1;
      } else {
        Rnew[g * LDP + 8 * p + 2 * t] = Pinv[g * 8 + 2 * t]; Rnew[g * LDP + 8 * p + 2 * t + 1] = Pinv[g * 8 + 2 * t + 1];
      }
    }
    __syncthreads();
    // (4) every warp updates the tiles it owns
#pragma unroll
    for (int s = 0; s < VUS_CH_LSLOTS; ++s) {
      if (s >= G.nl) continue;
      const int ti = G.lti[s], tj = G.ltj[s];
      if (ti < p) continue;
      if (ti == p) { c[s][0] = Rnew[g * LDP + 8 * tj + 2 * t]; c[s][1] = Rnew[g * LDP + 8 * tj + 2 * t + 1]; continue; }
      if (tj == p) c[s][0] = c[s][1] = 0.0;
      const double a0 = -Cnew[(8 * ti + g) * LDQ + t], a1 = -Cnew[(8 * ti + g) * LDQ + 4 + t];
      if (tj > p) {
        dmma884(c[s][0], c[s][1], a0, Cnew[(8 * tj + g) * LDQ + t]);
        dmma884(c[s][0], c[s][1], a1, Cnew[(8 * tj + g) * LDQ + 4 + t]);
      } else {
        dmma884(c[s][0], c[s][1], a0, Rnew[t * LDP + 8 * tj + g]);
        dmma884(c[s][0], c[s][1], a1, Rnew[(4 + t) * LDP + 8 * tj + g]);
      }
    }
    // the next step's (1) writes Craw / Rraw, which (3) of THIS step has finished reading (barrier above); its (3) rewrites
    // Cnew / Rnew only after two more barriers
  }
}

struct ChunkFactorBody {
  static VUS_DEV void run(const ChunkArgs& A, int c, int tid, int, double* sm) {
    __shared__ unsigned long long bar_;
    const CT G(A.B, tid);
    const int B = A.B, T = G.T, KP = G.KP, D = A.D;
    const long BBP = bcr_bbp(B);
    const unsigned blk_bytes = (unsigned)(BBP * sizeof(double));
    double* bufL = sm;
    double* bufU = sm + BBP;
    double* bufS = sm + 2 * BBP;
    double* scr = sm + 3 * BBP;
    const long first = chunk_first(A.G, c), len = chunk_len(A.G, c);
    const bool has_left = c > 0, has_right = c + 1 < A.G.P;
    Mbar mb;
    mb.init(&bar_, tid);
    for (long e = tid; e < BBP; e += VUS_CH_THREADS) bufL[e] = 0.0;       // the upper triangle of Linv stays zero throughout
    __syncthreads();
    // prologue: the first coupling tile and the left separator's coupling in flight, the first diagonal block into registers
    const bool first_has_u = first + 1 < A.G.Ns;
    if (tid == 0) {
      mb.expect((first_has_u ? blk_bytes : 0) + (has_left ? blk_bytes : 0));
      if (first_has_u) mb.copy(bufU, A.SU + first * BBP, blk_bytes);
      if (has_left) mb.copy(bufS, A.SU + (first - 1) * BBP, blk_bytes);
    }
    LAcc acc, accE;
#pragma unroll
    for (int s = 0; s < VUS_CH_LSLOTS; ++s) {
      accE[s][0] = accE[s][1] = 0.0;
      acc[s][0] = acc[s][1] = 0.0;
      if (s < G.nl) ltile_load(acc[s][0], acc[s][1], A.SD + first * BBP, G.lti[s], G.ltj[s], G);
    }
    bool pending_copy = first_has_u || has_left;
    for (long i = first; i < first + len; ++i) {
      const bool last = i + 1 == first + len;
      const bool has_u = i + 1 < A.G.Ns;                     // only the very last supernode of the chain has no coupling
      // ---- Cholesky + inverse of the running diagonal block
      chunk_chol_inv(acc, scr, G, A.fail);
#pragma unroll
      for (int s = 0; s < VUS_CH_LSLOTS; ++s)
        if (s < G.nl) tile_store(bufL, acc[s][0], acc[s][1], G.lti[s], G.ltj[s], G);
      // the accumulators are free until the D' update: the next diagonal block's global loads fly during the X product
      if (has_u && !last) {
#pragma unroll
        for (int s = 0; s < VUS_CH_LSLOTS; ++s)
          if (s < G.nl) ltile_load(acc[s][0], acc[s][1], A.SD + (i + 1) * BBP, G.lti[s], G.ltj[s], G);
      }
      if (pending_copy) { mb.wait(); pending_copy = false; }
      __syncthreads();
      if (tid == 0) bulk_store(A.Linv + i * BBP, bufL, blk_bytes);
      FAcc r;
      KRange K;
      if (has_u) {
        // ---- X = U^T Linv^T : X[a][c] = sum_k U[k][a] Linv[c][k],  D blk(a) <= k <= c
        K.kmin = KP; K.kmax = 0;
#pragma unroll
        for (int s = 0; s < VUS_CH_FSLOTS; ++s) {
          const int ti = ftile_i(G, s), tj = ftile_j(G, s);
          K.kb[s] = (D * ((8 * ti) / D)) & ~3; K.ke[s] = (8 * tj + 8) < KP ? (8 * tj + 8) : KP;
          if (ti >= T || tj >= T || K.kb[s] >= K.ke[s]) { K.kb[s] = KP; K.ke[s] = 0; continue; }
          K.kmin = K.kb[s] < K.kmin ? K.kb[s] : K.kmin; K.kmax = K.ke[s] > K.kmax ? K.ke[s] : K.kmax;
        }
        block_mma<true, true, false>(r, bufU, bufL, K, G);
        __syncthreads();
#pragma unroll
        for (int s = 0; s < VUS_CH_FSLOTS; ++s)
          if (ftile_i(G, s) < T && ftile_j(G, s) < T) tile_store(bufU, r[s][0], r[s][1], ftile_i(G, s), ftile_j(G, s), G);
        __syncthreads();
        if (tid == 0) bulk_store(A.X + i * BBP, bufU, blk_bytes);
        // ---- next diagonal block (or the right separator's contribution):  D' = D_{i+1} - X X^T, lower tiles
        int kbl[VUS_CH_LSLOTS], kminl = KP;
#pragma unroll
        for (int s = 0; s < VUS_CH_LSLOTS; ++s) {
          kbl[s] = KP;
          if (s >= G.nl) continue;
          if (last) acc[s][0] = acc[s][1] = 0.0;         // else: D_{i+1}, loaded right after the Cholesky
          kbl[s] = (D * ((8 * G.lti[s]) / D)) & ~3;
          kminl = kbl[s] < kminl ? kbl[s] : kminl;
        }
        lower_syrk<true>(acc, bufU, kbl, kminl, G);
#pragma unroll
        for (int s = 0; s < VUS_CH_LSLOTS; ++s) {
          if (s >= G.nl) continue;
          if (last) ltile_store_global_sym(A.SepR + (long)c * BBP, acc[s][0], acc[s][1], G.lti[s], G.ltj[s], -1.0, G);
          else ltile_fix_padding(acc[s][0], acc[s][1], G.lti[s], G.ltj[s], G);
        }
      }
      if (has_left) {
        // ---- W = S Linv^T : W[r][c] = sum_{k <= c} S[r][k] Linv[c][k]
        K.kmin = 0; K.kmax = 0;
#pragma unroll
        for (int s = 0; s < VUS_CH_FSLOTS; ++s) {
          const int ti = ftile_i(G, s), tj = ftile_j(G, s);
          K.kb[s] = 0; K.ke[s] = (8 * tj + 8) < KP ? (8 * tj + 8) : KP;
          if (ti >= T || tj >= T) { K.kb[s] = KP; K.ke[s] = 0; continue; }
          K.kmax = K.ke[s] > K.kmax ? K.ke[s] : K.kmax;
        }
        block_mma<false, true, false>(r, bufS, bufL, K, G);
        __syncthreads();
#pragma unroll
        for (int s = 0; s < VUS_CH_FSLOTS; ++s)
          if (ftile_i(G, s) < T && ftile_j(G, s) < T) tile_store(bufS, r[s][0], r[s][1], ftile_i(G, s), ftile_j(G, s), G);
        __syncthreads();
        if (tid == 0) bulk_store(A.W + i * BBP, bufS, blk_bytes);
        // ---- E += W W^T (lower tiles, kept in registers over the whole chunk)
        {
          int kb0[VUS_CH_LSLOTS];
#pragma unroll
          for (int s = 0; s < VUS_CH_LSLOTS; ++s) kb0[s] = 0;
          lower_syrk<false>(accE, bufS, kb0, 0, G);
        }
        if (has_u) {
          // ---- S' = -W X^T : S'[r][b] = -sum_{c >= D blk(b)} W[r][c] X[b][c]   (the separator coupling SepU at the last node)
          K.kmin = KP; K.kmax = KP;
#pragma unroll
          for (int s = 0; s < VUS_CH_FSLOTS; ++s) {
            const int ti = ftile_i(G, s), tj = ftile_j(G, s);
            K.kb[s] = (D * ((8 * tj) / D)) & ~3; K.ke[s] = KP;
            if (ti >= T || tj >= T) { K.kb[s] = KP; K.ke[s] = 0; continue; }
            K.kmin = K.kb[s] < K.kmin ? K.kb[s] : K.kmin;
          }
          block_mma<false, true, true>(r, bufS, bufU, K, G);
          if (tid == 0) bulk_store_wait_read();               // the bulk store of W has read bufS
          __syncthreads();
#pragma unroll
          for (int s = 0; s < VUS_CH_FSLOTS; ++s)
            if (ftile_i(G, s) < T && ftile_j(G, s) < T) tile_store(bufS, r[s][0], r[s][1], ftile_i(G, s), ftile_j(G, s), G);
          if (last && has_right) {
            __syncthreads();
            if (tid == 0) bulk_store(A.SepU + (long)c * BBP, bufS, blk_bytes);
          }
        }
      }
      if (tid == 0) bulk_store_wait_read();                   // every bulk store of this node has read its buffer (bufL, bufU, bufS)
      __syncthreads();                                        // every read of bufU / bufS of this node is done
      if (!last && i + 2 < A.G.Ns + 0 && tid == 0) {           // next coupling tile in flight while the next Cholesky runs
        mb.expect(blk_bytes);
        mb.copy(bufU, A.SU + (i + 1) * BBP, blk_bytes);
      }
      if (!last && i + 2 < A.G.Ns + 0) pending_copy = true;
    }
    if (has_left) {
#pragma unroll
      for (int s = 0; s < VUS_CH_LSLOTS; ++s)
        if (s < G.nl) ltile_store_global_sym(A.SepL + (long)c * BBP, accE[s][0], accE[s][1], G.lti[s], G.ltj[s], 1.0, G);
    }
    if (tid == 0) bulk_store_wait_all();
  }
};

// ---- sweeps: a three-slot ring of TMA-fed tile buffers, panels of up to 8 right-hand sides (bcr.cuh: Panel, panel_mma)
#define VUS_CH_RING 3
VUS_HD size_t chunk_sweep_smem(int B) { return (size_t)(VUS_CH_RING * bcr_buf_doubles(B) + 3 * (long)bcr_kp(B) * VUS_LDX + 16) * sizeof(double); }
struct TileRing {
  double* buf[VUS_CH_RING];
  unsigned long long* bar;            // [VUS_CH_RING] in shared memory
  unsigned parity[VUS_CH_RING];
  unsigned bytes;
  VUS_DEV void init(double* base, long stride, unsigned long long* bars, unsigned nbytes, int tid) {
    bar = bars; bytes = nbytes;
#pragma unroll
    for (int s = 0; s < VUS_CH_RING; ++s) { buf[s] = base + s * stride; parity[s] = 0; }
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < VUS_CH_RING; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + s)));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  VUS_DEV void issue(int slot, const double* src) {          // thread 0, after a barrier that retired every read of the slot
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + slot)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(buf[slot])), "l"(src), "r"(bytes), "r"(smem_u32(bar + slot)) : "memory");
  }
  VUS_DEV void wait(int slot) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 ::"r"(smem_u32(bar + slot)), "r"(parity[slot]) : "memory");
    parity[slot] ^= 1;
  }
};
// tile n of a chunk's stream: per node [first tile kinds...]; forward order Linv, (W), X -- backward order X, (W), Linv
VUS_DEV const double* chunk_stream_tile(const ChunkArgs& A, long first, long len, bool has_left, bool backward, long n, long* node, int* kind) {
  const int tpn = has_left ? 3 : 2;
  const long step = n / tpn;
  const int k = (int)(n % tpn);
  const long i = backward ? first + len - 1 - step : first + step;
  int kd;                                              // 0 Linv, 1 W, 2 X
  if (!backward) kd = k == 0 ? 0 : (has_left && k == 1 ? 1 : 2);
  else kd = k == 0 ? 2 : (has_left && k == 1 ? 1 : 0);
  *node = i; *kind = kd;
  const long BBP = bcr_bbp(A.B);
  return (kd == 0 ? A.Linv : kd == 1 ? A.W : A.X) + i * BBP;
}

// forward sweep of one chunk:  y_i = Linv_i (b_i - X_{i-1} y_{i-1}) ;  tL = sum_i W_i y_i ;  tR = X_last y_last
struct ChunkFwdBody {
  static VUS_DEV void run(const ChunkArgs& A, int c, int tid, int, double* sm) {
    __shared__ unsigned long long bars_[VUS_CH_RING];
    const int B = A.B, nv = A.nrhs;
    const Tiles G(B, tid);
    const long BBP = bcr_bbp(B);
    const long first = chunk_first(A.G, c), len = chunk_len(A.G, c);
    const bool has_left = c > 0;
    const int tpn = has_left ? 3 : 2;
    const long ntiles = len * tpn;
    double* sZ = sm + VUS_CH_RING * BBP;                 // panel of b_i - X_{i-1} y_{i-1}
    double* sY = sZ + (long)G.KP * VUS_LDX;              // panel of y_i
    TileRing ring;
    ring.init(sm, BBP, bars_, (unsigned)(BBP * sizeof(double)), tid);
    for (int e = tid; e < 2 * G.KP * VUS_LDX; e += 256) sZ[e] = 0.0;
    __syncthreads();
    long issued = 0;
    if (tid == 0)
      for (; issued < VUS_CH_RING && issued < ntiles; ++issued) {
        long nd; int kd;
        ring.issue((int)(issued % VUS_CH_RING), chunk_stream_tile(A, first, len, has_left, false, issued, &nd, &kd));
      }
    issued = ntiles < VUS_CH_RING ? ntiles : VUS_CH_RING;
    Panel pend, PtL, P;
    pend.zero(); PtL.zero();
    long n = 0;
    for (long i = first; i < first + len; ++i) {
      // z = b_i - pend
      VUS_PANEL_FOREACH(pend, G, nv, { sZ[r * VUS_LDX + v] = A.Xv[(long)v * A.xstride + i * B + r] - val; })
      __syncthreads();
      // y_i = Linv_i z
      ring.wait((int)(n % VUS_CH_RING));
      P.zero();
      panel_mma<false, false>(P, ring.buf[n % VUS_CH_RING], sZ, G);
      VUS_PANEL_FOREACH(P, G, nv, { A.Xv[(long)v * A.xstride + i * B + r] = val; sY[r * VUS_LDX + v] = val; })
      __syncthreads();
      if (tid == 0 && issued < ntiles) { long nd; int kd; ring.issue((int)(n % VUS_CH_RING), chunk_stream_tile(A, first, len, has_left, false, issued, &nd, &kd)); }
      if (issued < ntiles) ++issued;
      ++n;
      if (has_left) {                                    // tL += W_i y_i
        ring.wait((int)(n % VUS_CH_RING));
        panel_mma<false, false>(PtL, ring.buf[n % VUS_CH_RING], sY, G);
        __syncthreads();
        if (tid == 0 && issued < ntiles) { long nd; int kd; ring.issue((int)(n % VUS_CH_RING), chunk_stream_tile(A, first, len, has_left, false, issued, &nd, &kd)); }
        if (issued < ntiles) ++issued;
        ++n;
      }
      // pend = X_i y_i  (contribution to node i + 1, or to the right separator)
      ring.wait((int)(n % VUS_CH_RING));
      pend.zero();
      if (i + 1 < A.G.Ns) panel_mma<false, false>(pend, ring.buf[n % VUS_CH_RING], sY, G);
      __syncthreads();
      if (tid == 0 && issued < ntiles) { long nd; int kd; ring.issue((int)(n % VUS_CH_RING), chunk_stream_tile(A, first, len, has_left, false, issued, &nd, &kd)); }
      if (issued < ntiles) ++issued;
      ++n;
    }
    VUS_PANEL_FOREACH(pend, G, nv, { A.tR[((long)c * VUS_CHUNK_MAXV + v) * B + r] = val; })
    VUS_PANEL_FOREACH(PtL, G, nv, { A.tL[((long)c * VUS_CHUNK_MAXV + v) * B + r] = val; })
  }
};
// backward sweep of one chunk:  x_i = Linv_i^T (y_i - X_i^T x_{i+1} - W_i^T x_sepL);  also writes the right separator's solution
struct ChunkBwdBody {
  static VUS_DEV void run(const ChunkArgs& A, int c, int tid, int, double* sm) {
    __shared__ unsigned long long bars_[VUS_CH_RING];
    const int B = A.B, nv = A.nrhs;
    const Tiles G(B, tid);
    const long BBP = bcr_bbp(B);
    const long first = chunk_first(A.G, c), len = chunk_len(A.G, c);
    const bool has_left = c > 0, has_right = c + 1 < A.G.P;
    const int tpn = has_left ? 3 : 2;
    const long ntiles = len * tpn;
    double* sXn = sm + VUS_CH_RING * BBP;                // panel of x_{i+1}
    double* sXs = sXn + (long)G.KP * VUS_LDX;            // panel of x_sepL
    double* sZ = sXs + (long)G.KP * VUS_LDX;             // panel of y_i - ...
    TileRing ring;
    ring.init(sm, BBP, bars_, (unsigned)(BBP * sizeof(double)), tid);
    for (int e = tid; e < 3 * G.KP * VUS_LDX; e += 256) sXn[e] = 0.0;
    __syncthreads();
    long issued = 0;
    if (tid == 0)
      for (; issued < VUS_CH_RING && issued < ntiles; ++issued) {
        long nd; int kd;
        ring.issue((int)(issued % VUS_CH_RING), chunk_stream_tile(A, first, len, has_left, true, issued, &nd, &kd));
      }
    issued = ntiles < VUS_CH_RING ? ntiles : VUS_CH_RING;
    // separator solutions: right one into sXn (and out to the full vector), left one into sXs
    for (int e = tid; e < B * nv; e += 256) {
      const int r = e % B, v = e / B;
      if (has_right) {
        const double xs = A.xsep[(long)v * A.sepstride + (long)c * B + r];
        sXn[r * VUS_LDX + v] = xs;
        A.Xv[(long)v * A.xstride + chunk_sep(A.G, c) * B + r] = xs;
      }
      if (has_left) sXs[r * VUS_LDX + v] = A.xsep[(long)v * A.sepstride + (long)(c - 1) * B + r];
    }
    __syncthreads();
    Panel P;
    long n = 0;
    for (long i = first + len - 1; i >= first; --i) {
      P.zero();
      ring.wait((int)(n % VUS_CH_RING));                 // X_i^T x_{i+1}
      if (i + 1 < A.G.Ns) panel_mma<true, false>(P, ring.buf[n % VUS_CH_RING], sXn, G);
      __syncthreads();
      if (tid == 0 && issued < ntiles) { long nd; int kd; ring.issue((int)(n % VUS_CH_RING), chunk_stream_tile(A, first, len, has_left, true, issued, &nd, &kd)); }
      if (issued < ntiles) ++issued;
      ++n;
      if (has_left) {                                    // + W_i^T x_sepL
        ring.wait((int)(n % VUS_CH_RING));
        panel_mma<true, false>(P, ring.buf[n % VUS_CH_RING], sXs, G);
        __syncthreads();
        if (tid == 0 && issued < ntiles) { long nd; int kd; ring.issue((int)(n % VUS_CH_RING), chunk_stream_tile(A, first, len, has_left, true, issued, &nd, &kd)); }
        if (issued < ntiles) ++issued;
        ++n;
      }
      VUS_PANEL_FOREACH(P, G, nv, { sZ[r * VUS_LDX + v] = A.Xv[(long)v * A.xstride + i * B + r] - val; })
      __syncthreads();
      ring.wait((int)(n % VUS_CH_RING));                 // x_i = Linv_i^T z
      P.zero();
      panel_mma<true, false>(P, ring.buf[n % VUS_CH_RING], sZ, G);
      VUS_PANEL_FOREACH(P, G, nv, { A.Xv[(long)v * A.xstride + i * B + r] = val; sXn[r * VUS_LDX + v] = val; })
      __syncthreads();
      if (tid == 0 && issued < ntiles) { long nd; int kd; ring.issue((int)(n % VUS_CH_RING), chunk_stream_tile(A, first, len, has_left, true, issued, &nd, &kd)); }
      if (issued < ntiles) ++issued;
      ++n;
    }
  }
};

namespace rt {
template <> struct CoopBounds<ChunkFactorBody> { static constexpr int kMaxThreads = VUS_CH_THREADS, kMinBlocks = 1; };
template <> struct CoopBounds<ChunkFwdBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 1; };
template <> struct CoopBounds<ChunkBwdBody> { static constexpr int kMaxThreads = 256, kMinBlocks = 1; };
}  // namespace rt

#else
// =====================================================================================  host emulation (tests only)
#define VUS_CH_THREADS 1
inline size_t chunk_factor_smem(int) { return 64; }
inline size_t chunk_sweep_smem(int) { return 64; }
namespace chunk_emu {
typedef std::vector<double> Mat;   // dense B x B, row-major
inline Mat load(const double* tile, int B, int LD) { Mat m((size_t)B * B); for (int i = 0; i < B; ++i) for (int j = 0; j < B; ++j) m[(size_t)i * B + j] = tile[(long)i * LD + j]; return m; }
inline void store(double* tile, const Mat& m, int B, int LD) { for (int i = 0; i < B; ++i) for (int j = 0; j < B; ++j) tile[(long)i * LD + j] = m[(size_t)i * B + j]; }
// C = alpha * op(A) op(B) + beta * C
inline void gemm(Mat& C, const Mat& A, bool ta, const Mat& Bm, bool tb, int B, double alpha, double beta) {
  Mat out((size_t)B * B);
  for (int i = 0; i < B; ++i)
    for (int j = 0; j < B; ++j) {
      double s = 0.0;
      for (int k = 0; k < B; ++k) s += (ta ? A[(size_t)k * B + i] : A[(size_t)i * B + k]) * (tb ? Bm[(size_t)j * B + k] : Bm[(size_t)k * B + j]);
      out[(size_t)i * B + j] = alpha * s + beta * C[(size_t)i * B + j];
    }
  C.swap(out);
}
// D = L L^T -> L^-1 (lower); a non-positive pivot raises the flag (and is replaced by 1, as on the device)
inline Mat chol_inv(const Mat& Dm, int B, int* fail) {
  Mat L((size_t)B * B, 0.0), Li((size_t)B * B, 0.0);
  for (int j = 0; j < B; ++j) {
    double d = Dm[(size_t)j * B + j];
    for (int k = 0; k < j; ++k) d -= L[(size_t)j * B + k] * L[(size_t)j * B + k];
    if (!(d > 0.0)) { *fail = 1; d = 1.0; }
    const double dj = std::sqrt(d);
    L[(size_t)j * B + j] = dj;
    for (int i = j + 1; i < B; ++i) {
      double v = Dm[(size_t)i * B + j];
      for (int k = 0; k < j; ++k) v -= L[(size_t)i * B + k] * L[(size_t)j * B + k];
      L[(size_t)i * B + j] = v / dj;
    }
  }
  for (int cidx = 0; cidx < B; ++cidx)
    for (int i = cidx; i < B; ++i) {
      double s = (i == cidx) ? 1.0 : 0.0;
      for (int k = cidx; k < i; ++k) s -= L[(size_t)i * B + k] * Li[(size_t)k * B + cidx];
      Li[(size_t)i * B + cidx] = s / L[(size_t)i * B + i];
    }
  return Li;
}
}  // namespace chunk_emu

struct ChunkFactorBody {
  static VUS_DEV void run(const ChunkArgs& A, int c, int, int, double*) {
    using namespace chunk_emu;
    const int B = A.B, LD = bcr_ld(B);
    const long BBP = bcr_bbp(B);
    const long first = chunk_first(A.G, c), len = chunk_len(A.G, c);
    const bool has_left = c > 0, has_right = c + 1 < A.G.P;
    Mat Dc = load(A.SD + first * BBP, B, LD), S, E((size_t)B * B, 0.0), Z((size_t)B * B, 0.0);
    if (has_left) S = load(A.SU + (first - 1) * BBP, B, LD);
    for (long i = first; i < first + len; ++i) {
      const bool last = i + 1 == first + len, has_u = i + 1 < A.G.Ns;
      Mat Li = chol_inv(Dc, B, A.fail);
      store(A.Linv + i * BBP, Li, B, LD);
      Mat Xm;
      if (has_u) {
        Mat U = load(A.SU + i * BBP, B, LD);
        Xm = Z; gemm(Xm, U, true, Li, true, B, 1.0, 0.0);            // X = U^T Linv^T
        store(A.X + i * BBP, Xm, B, LD);
        if (!last) { Dc = load(A.SD + (i + 1) * BBP, B, LD); gemm(Dc, Xm, false, Xm, true, B, -1.0, 1.0); }
        else { Mat R = Z; gemm(R, Xm, false, Xm, true, B, 1.0, 0.0); store(A.SepR + (long)c * BBP, R, B, LD); }
      }
      if (has_left) {
        Mat Wm = Z; gemm(Wm, S, false, Li, true, B, 1.0, 0.0);       // W = S Linv^T
        store(A.W + i * BBP, Wm, B, LD);
        gemm(E, Wm, false, Wm, true, B, 1.0, 1.0);
        if (has_u) {
          S = Z; gemm(S, Wm, false, Xm, true, B, -1.0, 0.0);         // S' = -W X^T
          if (last && has_right) store(A.SepU + (long)c * BBP, S, B, LD);
        }
      }
    }
    if (has_left) store(A.SepL + (long)c * BBP, E, B, LD);
  }
};
struct ChunkFwdBody {
  static VUS_DEV void run(const ChunkArgs& A, int c, int, int, double*) {
    const int B = A.B, LD = bcr_ld(B), nv = A.nrhs;
    const long BBP = bcr_bbp(B);
    const long first = chunk_first(A.G, c), len = chunk_len(A.G, c);
    const bool has_left = c > 0;
    std::vector<double> pend((size_t)nv * B, 0.0), tl((size_t)nv * B, 0.0), y((size_t)nv * B), z((size_t)nv * B);
    for (long i = first; i < first + len; ++i) {
      for (int v = 0; v < nv; ++v) for (int r = 0; r < B; ++r) z[(size_t)v * B + r] = A.Xv[(long)v * A.xstride + i * B + r] - pend[(size_t)v * B + r];
      std::fill(y.begin(), y.end(), 0.0);
      emu_blk_accum(y.data(), A.Linv + i * BBP, LD, false, z.data(), B, 1.0, B, nv);
      for (int v = 0; v < nv; ++v) for (int r = 0; r < B; ++r) A.Xv[(long)v * A.xstride + i * B + r] = y[(size_t)v * B + r];
      if (has_left) emu_blk_accum(tl.data(), A.W + i * BBP, LD, false, y.data(), B, 1.0, B, nv);
      std::fill(pend.begin(), pend.end(), 0.0);
      if (i + 1 < A.G.Ns) emu_blk_accum(pend.data(), A.X + i * BBP, LD, false, y.data(), B, 1.0, B, nv);
    }
    for (int v = 0; v < nv; ++v)
      for (int r = 0; r < B; ++r) {
        A.tR[((long)c * VUS_CHUNK_MAXV + v) * B + r] = pend[(size_t)v * B + r];
        A.tL[((long)c * VUS_CHUNK_MAXV + v) * B + r] = tl[(size_t)v * B + r];
      }
  }
};
struct ChunkBwdBody {
  static VUS_DEV void run(const ChunkArgs& A, int c, int, int, double*) {
    const int B = A.B, LD = bcr_ld(B), nv = A.nrhs;
    const long BBP = bcr_bbp(B);
    const long first = chunk_first(A.G, c), len = chunk_len(A.G, c);
    const bool has_left = c > 0, has_right = c + 1 < A.G.P;
    std::vector<double> xn((size_t)nv * B, 0.0), xs((size_t)nv * B, 0.0), acc((size_t)nv * B), z((size_t)nv * B), x((size_t)nv * B);
    for (int v = 0; v < nv; ++v)
      for (int r = 0; r < B; ++r) {
        if (has_right) {
          xn[(size_t)v * B + r] = A.xsep[(long)v * A.sepstride + (long)c * B + r];
          A.Xv[(long)v * A.xstride + chunk_sep(A.G, c) * B + r] = xn[(size_t)v * B + r];
        }
        if (has_left) xs[(size_t)v * B + r] = A.xsep[(long)v * A.sepstride + (long)(c - 1) * B + r];
      }
    for (long i = first + len - 1; i >= first; --i) {
      std::fill(acc.begin(), acc.end(), 0.0);
      if (i + 1 < A.G.Ns) emu_blk_accum(acc.data(), A.X + i * BBP, LD, true, xn.data(), B, 1.0, B, nv);
      if (has_left) emu_blk_accum(acc.data(), A.W + i * BBP, LD, true, xs.data(), B, 1.0, B, nv);
      for (int v = 0; v < nv; ++v) for (int r = 0; r < B; ++r) z[(size_t)v * B + r] = A.Xv[(long)v * A.xstride + i * B + r] - acc[(size_t)v * B + r];
      std::fill(x.begin(), x.end(), 0.0);
      emu_blk_accum(x.data(), A.Linv + i * BBP, LD, true, z.data(), B, 1.0, B, nv);
      for (int v = 0; v < nv; ++v) for (int r = 0; r < B; ++r) A.Xv[(long)v * A.xstride + i * B + r] = x[(size_t)v * B + r];
      xn = x;
    }
  }
};
#endif

}  // namespace vus
