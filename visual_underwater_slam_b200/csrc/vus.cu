// libvus: handle, symbolic analysis, the LM driver and the C-ABI of include/vus.h.
//
// Replaces gtsam's LevenbergMarquardtOptimizer::optimize() for the graph of /root/reference/batch.py
// (call site batch.py:337).  Host code here only sequences kernels and takes the accept/reject
// decisions from device-computed scalars; all arithmetic on graph data is in kernels.cuh.
#include "../../include/vus.h"
#include "rt.h"
#include "kernels.cuh"
#include "chunk.cuh"
#include "spike.cuh"
#include "frontend.cuh"
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <map>
#include <numeric>
#include <tuple>
#ifndef VUS_EMU
#include <cub/device/device_radix_sort.cuh>
#include <dlfcn.h>
#endif

using namespace vus;
using rt::DBuf;

namespace {

// launch counter / class / profiler are per host thread: one host thread drives one handle (include/vus.h), several
// handles may run concurrently from several threads (independent trajectories on one GPU)
thread_local long g_launches = 0;

// kernel classes for the per-class device timing (vus_lm_result.ms_class / launches_class)
enum { KC_LINEARIZE = 0, KC_ERROR, KC_LINERR, KC_ASSEMBLE, KC_STEREO_ASM, KC_SCHUR, KC_BCR_FACTOR, KC_BCR_SOLVE,
       KC_MATVEC, KC_BORDER, KC_VECTOR, KC_RETRACT, KC_COUNT };
thread_local int g_class = KC_VECTOR;
struct ClassGuard { int prev; explicit ClassGuard(int c) : prev(g_class) { g_class = c; } ~ClassGuard() { g_class = prev; } };

struct Profiler {   // one instance per host thread (g_prof below)
  bool on = false;
  double ms[16] = {0};
  long count[16] = {0};
#ifndef VUS_EMU
  std::vector<cudaEvent_t> pool;
  std::vector<int> cls;
  size_t used = 0;
  void begin(rt::stream_t st) {
    if (used + 2 > pool.size()) { pool.resize(pool.size() + 2); cudaEventCreate(&pool[pool.size() - 2]); cudaEventCreate(&pool[pool.size() - 1]); }
    cudaEventRecord(pool[used], st);
  }
  void end(rt::stream_t st) { cudaEventRecord(pool[used + 1], st); cls.push_back(g_class); used += 2; }
  void collect() {
    if (!used) return;
    cudaEventSynchronize(pool[used - 1]);
    for (size_t i = 0; i < used; i += 2) { float t = 0; cudaEventElapsedTime(&t, pool[i], pool[i + 1]); ms[cls[i / 2]] += t; count[cls[i / 2]]++; }
    used = 0; cls.clear();
  }
#else
  void begin(rt::stream_t) {}
  void end(rt::stream_t) { count[g_class]++; }
  void collect() {}
#endif
  void reset() { collect(); for (int i = 0; i < 16; ++i) { ms[i] = 0; count[i] = 0; } }
};
thread_local Profiler g_prof;

template <class Body, class Args>
void L_elem(long n, rt::stream_t st, const Args& a) {
  if (n <= 0) return;
  ++g_launches;
  if (g_prof.on) g_prof.begin(st);
  rt::launch_elem<Body, Args>(n, st, a);
  if (g_prof.on) g_prof.end(st);
}
template <class Body, class Args>
void L_coop(int grid, int block, size_t smem, rt::stream_t st, const Args& a) {
  if (grid <= 0) return;
  ++g_launches;
  if (g_prof.on) g_prof.begin(st);
  rt::launch_coop<Body, Args>(grid, block, smem, st, a);
  if (g_prof.on) g_prof.end(st);
}

// A fixed sequence of launches (one band factorization, one band solve of a given right-hand side) replayed as a CUDA
// graph: one host call instead of ~30, so the solve does not depend on how fast the host can enqueue kernels.
struct GraphCache {
  bool warm = false;
  long launches = 0;
#ifndef VUS_EMU
  cudaGraphExec_t exec = nullptr;
  void reset() { if (exec) cudaGraphExecDestroy(exec); exec = nullptr; warm = false; launches = 0; }
#else
  void reset() { warm = false; launches = 0; }
#endif
};

thread_local bool g_capturing = false;       // a fixed sequence is being captured: sequences nested in it are recorded inline
template <class F>
void run_graphed(GraphCache& gc, rt::stream_t st, F&& body) {
#ifndef VUS_EMU
  if (!g_prof.on && st != 0 && !g_capturing) {
    if (!gc.exec) {
      if (!gc.warm) { body(); gc.warm = true; return; }       // first use runs eagerly (sets kernel attributes)
      const long l0 = g_launches;
      cudaGraph_t graph = nullptr;
      rt::check(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal), "graph capture begin");
      g_capturing = true;
      try { body(); } catch (...) { g_capturing = false; cudaStreamEndCapture(st, &graph); if (graph) cudaGraphDestroy(graph); throw; }
      g_capturing = false;
      rt::check(cudaStreamEndCapture(st, &graph), "graph capture end");
      gc.launches = g_launches - l0;
      g_launches = l0;
      rt::check(cudaGraphInstantiate(&gc.exec, graph, 0), "graph instantiate");
      cudaGraphDestroy(graph);
    }
    rt::check(cudaGraphLaunch(gc.exec, st), "graph launch");
    g_launches += gc.launches;
    return;
  }
#endif
  body();
}

struct FactorTable {
  long n = 0;
  std::vector<int> h_idx;        // [slots][n]
  std::vector<int64_t> orig;
  std::vector<int> perm;         // stereo only: table row f holds the caller's row perm[f] (rows are kept landmark-major)
  DBuf<int> idx;
  DBuf<double> meas, sinfo, r, J;
  DBuf<int> d_order;             // stereo only: pending row re-ordering whose orig / perm bookkeeping has not been derived yet
  long e_off = 0;                // offset into the concatenated per-factor error buffer
};

const int kVarDim[4] = {12, 3, 6, 3};

// Caller-order bookkeeping of a re-ordered table: row f now holds what row order[f] held.  Only the calls that hand
// per-factor results back in the caller's order need it, so vus_analyze leaves the order on the device.
void finish_order(FactorTable& T, rt::stream_t st) {
  if (!T.d_order.p) return;
  std::vector<int> order(T.n);
  rt::d2h(order.data(), T.d_order.p, T.n * sizeof(int), st);
  rt::sync(st);
  std::vector<int64_t> norig(T.n);
  std::vector<int> nperm(T.n);
  for (long o = 0; o < T.n; ++o) {
    norig[o] = T.orig[order[o]];
    nperm[o] = T.perm.empty() ? order[o] : T.perm[order[o]];
  }
  T.orig.swap(norig); T.perm.swap(nperm);
  T.d_order.release();
}


#ifndef VUS_EMU
// NCCL is loaded at run time (dlopen of the libnccl.so.2 the process already uses -- torch's -- or the system one): the
// library has no link-time dependency on a communication stack and loads on a box without one.  Only the handful of entry
// points the pose-range partition needs are bound; the types are NCCL's ABI (nccl.h: 128-byte unique id, ncclDouble = 8,
// ncclSum = 0).
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
  std::string why;
};
NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { api.why = std::string("libnccl.so.2 not found: ") + dlerror(); return api; }
  auto bind = [&](const char* name) { void* p = dlsym(lib, name); if (!p) api.why = std::string("missing NCCL symbol ") + name; return p; };
  api.GetUniqueId = (int (*)(NcclApi::UniqueId*))bind("ncclGetUniqueId");
  api.CommInitRank = (int (*)(void**, int, NcclApi::UniqueId, int))bind("ncclCommInitRank");
  api.CommDestroy = (int (*)(void*))bind("ncclCommDestroy");
  api.AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))bind("ncclAllReduce");
  api.Send = (int (*)(const void*, size_t, int, int, void*, cudaStream_t))bind("ncclSend");
  api.Recv = (int (*)(void*, size_t, int, int, void*, cudaStream_t))bind("ncclRecv");
  api.GroupStart = (int (*)())bind("ncclGroupStart");
  api.GroupEnd = (int (*)())bind("ncclGroupEnd");
  api.GetErrorString = (const char* (*)(int))bind("ncclGetErrorString");
  api.ok = api.why.empty();
  return api;
}
void nccl_check(int rc, const char* what) {
  if (rc != 0) {
    NcclApi& n = nccl_api();
    throw std::runtime_error(std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"));
  }
}
#endif

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct vus_handle {
  int device = 0;
  std::string err;
  vus_lm_params prm;
  double K[6] = {1, 1, 0, 0, 0, 1};
  double grav[3] = {0, 0, -9.81};
  int opts = VUS_OPT_TANGENT;    // gtsam build switches (vus_set_gtsam_build): tangent preintegration on, slow BetweenFactor off
  // variables
  long nvar[4] = {0, 0, 0, 0};
  std::vector<uint64_t> keys[4];
  DBuf<double> val[2][4];
  DBuf<double> saved[4];
  bool has_saved = false;
  int cur = 0;
  // factors
  FactorTable ft[VUS_F_NTYPES];
  long nfactors = 0;
  bool analyzed = false;
  // layout
  int D = 6, k = 1, B = 6, has_bias = 0;
  long N = 0, Ns = 0, Npad = 0, Lc = 0, L = 0;   // Lc = Npad*D camera dofs, L = Lc + 6*has_bias
  long Ns_band = 0;              // supernodes the band preconditioner / operator rows cover: all, or the owned prefix of a partition
  long nrem = 0, sd_off = 0, su_off = 0, rem_off = 0, hlen = 0;
  DBuf<double> H0, H;            // SD | SU | REM   (undamped base / damped + Schur)
  bool H_synced = false;         // H holds a full copy of H0's structure (form_system then refreshes only the blocks that change)
  DBuf<int> na_ptr, na_code, na_fac;      // node -> incident chain factors (NodeAsmBody)
  DBuf<int> pg_ptr, pg_code, pg_fac;      // node pair -> its two-node factors (PairAsmBody)
  DBuf<PairDst> pg_dst; long npairs = 0;
  DBuf<int> rem_ptr, rem_col;
  DBuf<double> g0, gs, F, Hbb0, Hbb, gb;     // gs = [reduced camera gradient ; gb] (length L)
  // stereo
  long nobs = 0, nposes_obs = 0;
  DBuf<int> pose_ptr, pose_obs, pose_ids, lm_ptr, lm_long, long_ids;
  long nlong = 0;
  // long tracks in the preconditioner: segments that fit in the band (SchurArgs), and the preconditioner's own band matrix
  DBuf<int> obs_seg, seg_ptr; long nseg = 0;
  DBuf<double> CinvSeg, Hp;
  DBuf<double> ulong;
  double cur_lambda = 0.0;
  int schur_ndj = 1;
  DBuf<int> partner;
  double last_rel_res = 0.0;     // true relative residual ||rhs - A x|| / ||rhs|| at the end of the last pcg()
  bool z0_valid = false;         // column 6 of Z holds M^-1 gs (band part of the first preconditioner application)
  DBuf<double> C, gl, Cinv, E, Pp, Pl;
  // BCR
  DBuf<double> Dw, U1, U2, Dinv, Gl, Gr, Z, Zr, SbInv;
  // chunked band factorization (chunk.cuh): P chunks + the separator system, which is the one the cyclic reduction then factors.
  // Dinv / Gl / Gr hold Linv / X / W of the chunk interiors.
  int chunk_P = 0;               // 0: plain cyclic reduction of the whole chain
  ChunkGeom cgeom;
  DBuf<double> SepL, SepR, SepU, Dsep, Usep, sDw, sU1, sU2, sDinv, sGl, sGr, tL, tR, xsep;
  // PCG
  DBuf<double> x, r, z, p, Ap, d, xl, scal, partials, bpart, bpart2, e_all, le_all;
  DBuf<int> fail;
  int red_grid = 0;
  // pose-range partition (one rank of a graph split across GPUs): local nodes = [owned | halo]
  long n_owned = -1;             // -1: not partitioned
  int64_t nf_owned[VUS_F_NTYPES] = {0, 0, 0, 0, 0, 0};
  long Lr = 0;                   // length of the vector prefix that enters dot products (owned dofs)
  DBuf<double> e_mask;           // 1 for factors this rank owns, 0 for duplicates of a neighbour's factor
  vus_comm_fn comm = nullptr;
  void* comm_ctx = nullptr;
  // collectives issued by the library itself (vus_comm_init): NCCL on the stream the solve runs on, inside the captured sequences
  void* nccl = nullptr;          // ncclComm_t
  int nccl_rank = 0, nccl_nranks = 1;
  std::vector<int> halo_peers;                   // ranks this one exchanges halo nodes with
  std::vector<long> halo_send_ptr;               // [npeers + 1] into halo_send_idx: the owned nodes each peer needs
  std::vector<long> halo_recv_off, halo_recv_cnt;   // [npeers] where each peer's nodes sit in the halo (contiguous per owner)
  DBuf<int> halo_send_idx;
  DBuf<double> halo_sendbuf;
  long halo_nsend = 0;
  bool comm_stream_ordered = false;   // the callback enqueues its collective on the stream the library runs on: no host sync around it
  // exact band across ranks (spike.cuh): chain position of this rank and the halo nodes that continue the chain
  int sp_rank = 0, sp_nranks = 1;
  std::vector<long> sp_prev, sp_next; // local index of the j-th pose before the first / after the last owned pose (-1: none)
  int spike_state = 0;                // 0 undecided, 1 on (every rank can), -1 off (block-Jacobi across ranks)
  DBuf<PairDst> spPairs;              // [2][k][k] coupling blocks to the next / previous rank
  DBuf<double> spZ, spG, spR, spRinv, spY, spCoef;
  // batched mode: independent components (trajectories) in one block-diagonal system (batch.cuh)
  int ncomp = 1;
  long bcr_stop = 1L << 62;      // cyclic reduction needs no stride beyond the longest component (in supernodes)
  std::vector<long> comp_start;  // [ncomp + 1] node ranges, vus_set_components
  DBuf<int> node_comp, cf_ptr, cf_list, ci_ptr, ci_list;
  DBuf<long> seg;
  DBuf<double> scal_b, lam_c, acc_c, cdots, cdots2, err_c;
  DBuf<int> fail_mask;           // [ncomp] components whose factorization failed in the current round
  std::vector<vus_component_result> comp_res;
  // loop closures of a batched graph by capacitance (batch.cuh): R = 12 * (max closures per component) columns, 0 = off
  int wb_R = 0, wb_lmax = 0;
  DBuf<int> wb_node;
  DBuf<long> wb_blk;
  DBuf<double> wbY, wbCapInv, wbW;
  // CUDA graphs of the fixed launch sequences + the stream used when the caller passes NULL (capture needs a real stream)
  GraphCache g_factor, g_pcg_iter;
  std::map<std::tuple<const double*, long, int>, GraphCache> g_solve;
  rt::stream_t own_stream = 0;
  void drop_graphs() { g_factor.reset(); g_pcg_iter.reset(); for (auto& kv : g_solve) kv.second.reset(); g_solve.clear(); }
  ~vus_handle() {      // vus_destroy holds the DeviceGuard; members (device buffers) are freed after this body, on own_stream's order
    drop_graphs();
  }
  // stats
  vus_lm_result res;
  std::vector<vus_lm_try> trace;   // one entry per lambda try of the last single-graph optimize()
};

namespace {

int fail(vus_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

ValuesView view_of(vus_handle* h, int which) {
  ValuesView V;
  V.pose = h->val[which][0].p; V.nx = h->nvar[0];
  V.vel = h->val[which][1].p; V.nv = h->nvar[1];
  V.bias = h->val[which][2].p; V.nb = h->nvar[2];
  V.lm = h->val[which][3].p; V.nl = h->nvar[3];
  return V;
}

template <bool WJ>
void launch_lin_type(vus_handle* h, int t, const LinArgs& a, rt::stream_t st) {
  const long n = h->ft[t].n;
  switch (t) {
    case VUS_F_PRIOR_POSE: L_elem<LinBody<VUS_F_PRIOR_POSE, WJ>>(n, st, a); break;
    case VUS_F_PRIOR_VEL: L_elem<LinBody<VUS_F_PRIOR_VEL, WJ>>(n, st, a); break;
    case VUS_F_BETWEEN: L_elem<LinBody<VUS_F_BETWEEN, WJ>>(n, st, a); break;
    case VUS_F_DVL: L_elem<LinBody<VUS_F_DVL, WJ>>(n, st, a); break;
    case VUS_F_STEREO: L_elem<LinBody<VUS_F_STEREO, WJ>>(n, st, a); break;
    default: L_elem<LinBody<VUS_F_IMU, WJ>>(n, st, a); break;
  }
}

// kernel 1 over every factor type; with_J: also write r and J (linearize), else residual norms only
void run_factors(vus_handle* h, int which, bool with_J, rt::stream_t st) {
  ClassGuard kc_guard(with_J ? KC_LINEARIZE : KC_ERROR);
  for (int t = 0; t < VUS_F_NTYPES; ++t) {
    FactorTable& T = h->ft[t];
    if (!T.n) continue;
    LinArgs a;
    a.V = view_of(h, which);
    a.F.n = T.n; a.F.idx = T.idx.p; a.F.meas = T.meas.p; a.F.sinfo = T.sinfo.p;
    a.O.r = with_J ? T.r.p : nullptr;
    a.O.J = with_J ? T.J.p : nullptr;
    a.O.e2 = h->e_all.p + T.e_off;
    const bool fused = with_J && t == VUS_F_STEREO && which == h->cur;
    a.O.sE = fused ? h->E.p : nullptr; a.O.sPp = fused ? h->Pp.p : nullptr; a.O.sPl = fused ? h->Pl.p : nullptr;
    for (int i = 0; i < 6; ++i) a.K[i] = h->K[i];
    for (int i = 0; i < 3; ++i) a.g[i] = h->grav[i];
    a.type = t;
    a.opts = h->opts;
    if (fused) L_coop<LinStereoTileBody>((int)((T.n + VUS_LIN_TILE - 1) / VUS_LIN_TILE), VUS_LIN_TILE, (size_t)VUS_LIN_TILE * VUS_LIN_ROW * sizeof(double), st, a);
    else if (with_J) launch_lin_type<true>(h, t, a, st);
    else launch_lin_type<false>(h, t, a, st);
  }
}

// deterministic sum / dot -> scal[slot] with post-op
struct HaloPackArgs { const double* vec; double* out; const int* idx; long n; int D; };
struct HaloPackBody {       // out[e][c] = vec[idx[e]][c]: the owned nodes other ranks hold as halo, packed per peer
  static VUS_DEV void run(const HaloPackArgs& A, long w) { A.out[w] = A.vec[(long)A.idx[w / A.D] * A.D + w % A.D]; }
};
bool has_comm(const vus_handle* h) { return h->comm != nullptr || h->nccl != nullptr; }
void comm_call(vus_handle* h, int op, void* buf, long count, rt::stream_t st) {
#ifndef VUS_EMU
  if (h->nccl) {                                       // stream-ordered, capturable
    NcclApi& n = nccl_api();
    if (op == VUS_COMM_ALLREDUCE_SUM) {
      nccl_check(n.AllReduce(buf, buf, (size_t)count, 8 /* ncclDouble */, 0 /* ncclSum */, h->nccl, st), "ncclAllReduce");
      return;
    }
    const int D = (int)count;
    double* vec = (double*)buf;
    if (h->halo_nsend) {
      HaloPackArgs a; a.vec = vec; a.out = h->halo_sendbuf.p; a.idx = h->halo_send_idx.p; a.n = h->halo_nsend; a.D = D;
      L_elem<HaloPackBody>(h->halo_nsend * D, st, a);
    }
    nccl_check(n.GroupStart(), "ncclGroupStart");
    for (size_t q = 0; q < h->halo_peers.size(); ++q) {
      const long ns = h->halo_send_ptr[q + 1] - h->halo_send_ptr[q];
      if (ns) nccl_check(n.Send(h->halo_sendbuf.p + h->halo_send_ptr[q] * D, (size_t)ns * D, 8, h->halo_peers[q], h->nccl, st), "ncclSend");
      if (h->halo_recv_cnt[q])
        nccl_check(n.Recv(vec + (h->n_owned + h->halo_recv_off[q]) * D, (size_t)h->halo_recv_cnt[q] * D, 8, h->halo_peers[q], h->nccl, st), "ncclRecv");
    }
    nccl_check(n.GroupEnd(), "ncclGroupEnd");
    return;
  }
#endif
  if (!h->comm_stream_ordered) rt::sync(st);           // host-synchronous callbacks: the collective runs on the caller's own stream
  if (h->comm(h->comm_ctx, op, buf, (int64_t)count) != 0) throw std::runtime_error("communication callback failed");
}
void reduce(vus_handle* h, const double* a, const double* b, long n, int slot, int op, rt::stream_t st) {
  RedArgs r1; r1.a = a; r1.b = b; r1.n = n; r1.partials = h->partials.p; r1.grid = h->red_grid;
  L_coop<Red1Body>(h->red_grid, 256, 256 * sizeof(double), st, r1);
  Red2Args r2; r2.partials = h->partials.p; r2.grid = h->red_grid; r2.scal = h->scal.p; r2.slot = slot; r2.op = op;
  if (!has_comm(h)) {
    L_coop<Red2Body>(1, 256, 256 * sizeof(double), st, r2);
    return;
  }
  Red2Args rl = r2; rl.slot = S_COMM; rl.op = RED_STORE;             // local partial -> all-reduce -> post-op
  L_coop<Red2Body>(1, 256, 256 * sizeof(double), st, rl);
  comm_call(h, VUS_COMM_ALLREDUCE_SUM, h->scal.p + S_COMM, 1, st);
  L_elem<RedPostBody>(1, st, r2);
}
// fill the halo entries of a node vector (D doubles per node) from their owners
void halo(vus_handle* h, double* vec, rt::stream_t st) {
  if (has_comm(h) && h->n_owned >= 0 && h->n_owned < h->N) comm_call(h, VUS_COMM_HALO, vec, h->D, st);
}
void zero_halo(vus_handle* h, double* vec, rt::stream_t st) {
  if (h->n_owned >= 0 && h->Lc > h->Lr) rt::dzero(vec + h->Lr, (size_t)(h->Lc - h->Lr) * sizeof(double), st);
}

double read_scalar(vus_handle* h, int slot, rt::stream_t st) {
  double v = 0;
  rt::d2h(&v, h->scal.p + slot, sizeof(double), st);
  rt::sync(st);
  return v;
}

double graph_error(vus_handle* h, int which, rt::stream_t st) {
  run_factors(h, which, false, st);
  reduce(h, h->e_all.p, h->n_owned >= 0 ? h->e_mask.p : nullptr, h->nfactors, S_TMP, RED_STORE, st);
  return read_scalar(h, S_TMP, st);
}

// copy a caller table (host/device, component-major or row-major) into a component-major device table [dim][n]
void import_table(double* dst, const double* src, long n, int dim, int mem, rt::stream_t st) {
  const size_t bytes = (size_t)dim * n * sizeof(double);
  if (!bytes) return;
  if (mem == VUS_MEM_HOST) { rt::h2d(dst, src, bytes, st); return; }
  if (mem == VUS_MEM_DEVICE) { rt::d2d(dst, src, bytes, st); return; }
  DBuf<double> tmp;
  const double* rows = src;
  if (mem == VUS_MEM_HOST_ROWS) { tmp.alloc((size_t)dim * n); rt::h2d(tmp.p, src, bytes, st); rows = tmp.p; }
  TransposeArgs t; t.src = rows; t.dst = dst; t.n = n; t.dim = dim; t.to_soa = 1;
  L_elem<TransposeBody>((long)dim * n, st, t);
  rt::sync(st);                                        // tmp is released on return
}
void export_table(double* dst, const double* src, long n, int dim, int mem, rt::stream_t st) {
  const size_t bytes = (size_t)dim * n * sizeof(double);
  if (!bytes) return;
  if (mem == VUS_MEM_HOST) { rt::d2h(dst, src, bytes, st); return; }
  if (mem == VUS_MEM_DEVICE) { rt::d2d(dst, src, bytes, st); return; }
  DBuf<double> tmp;
  double* rows = dst;
  if (mem == VUS_MEM_HOST_ROWS) { tmp.alloc((size_t)dim * n); rows = tmp.p; }
  TransposeArgs t; t.src = src; t.dst = rows; t.n = n; t.dim = dim; t.to_soa = 0;
  L_elem<TransposeBody>((long)dim * n, st, t);
  if (mem == VUS_MEM_HOST_ROWS) rt::d2h(dst, tmp.p, bytes, st);
  rt::sync(st);
}

// ------------------------------------------------------------------ symbolic analysis
// Off-band node pairs (both orientations) -> index of their remainder block, blocks in (p, q) lexicographic order.  A sorted
// key vector with binary search: a std::map of the 12 M pairs of config 5 cost ~1.2 us per insert and per look-up.
struct RemIndex {
  std::vector<uint64_t> keys;                            // (p << 32) | q, sorted and unique after finalize()
  static uint64_t key(long p, long q) { return ((uint64_t)p << 32) | (uint64_t)(uint32_t)q; }
  void add(long p, long q) { keys.push_back(key(p, q)); }
  void finalize() { std::sort(keys.begin(), keys.end()); keys.erase(std::unique(keys.begin(), keys.end()), keys.end()); }
  long find(long p, long q) const {
    const uint64_t k = key(p, q);
    auto it = std::lower_bound(keys.begin(), keys.end(), k);
    return (it != keys.end() && *it == k) ? (long)(it - keys.begin()) : -1;
  }
  long at(long p, long q) const { const long i = find(p, q); if (i < 0) throw std::out_of_range("no remainder block for this node pair"); return i; }
  bool has(long p, long q) const { return find(p, q) >= 0; }
  bool empty() const { return keys.empty(); }
  long size() const { return (long)keys.size(); }
  long p_of(long i) const { return (long)(keys[i] >> 32); }
  long q_of(long i) const { return (long)(keys[i] & 0xffffffffu); }
};

PairDst band_dst(const vus_handle* h, long p, long q, const RemIndex& rem_index) {
  // destination of block (p,q), p != q allowed to be in any order
  PairDst d; d.pad = 0;
  const int D = h->D, k = h->k, B = bcr_ld(h->B);          // B: row stride of the padded supernode tiles
  const long I = p / k, J = q / k;
  const int rp = (int)(p % k), rq = (int)(q % k);
  const long BB = bcr_bbp(h->B);
  if (I == J) {
    d.off = h->sd_off + I * BB + (long)(rp * D) * B + rq * D; d.ld = B; d.transposed = 0;
    d.moff = h->sd_off + I * BB + (long)(rq * D) * B + rp * D; d.mld = B;
  } else if (J == I + 1) {
    d.off = h->su_off + I * BB + (long)(rp * D) * B + rq * D; d.ld = B; d.transposed = 0; d.moff = -1; d.mld = 0;
  } else if (J == I - 1) {
    d.off = h->su_off + J * BB + (long)(rq * D) * B + rp * D; d.ld = B; d.transposed = 1; d.moff = -1; d.mld = 0;
  } else {
    const long b1 = rem_index.at(p, q), b2 = rem_index.at(q, p);
    d.off = h->rem_off + b1 * D * D; d.ld = D; d.transposed = 0;
    d.moff = h->rem_off + b2 * D * D; d.mld = D;
  }
  return d;
}

SchurArgs schur_args(vus_handle* h, double lambda);

// batched mode: component of every node, dof segments, per-component factor lists (errors) and IMU lists (bias blocks)
int analyze_components(vus_handle* h, rt::stream_t st, const RemIndex& rem_index) {
  const int nc = h->ncomp;
  const long NX = h->nvar[0];
  std::vector<int> node_comp(h->Npad, nc - 1);
  std::vector<long> seg(nc + 1);
  for (int c = 0; c < nc; ++c) {
    if (h->comp_start[c + 1] <= h->comp_start[c]) return fail(h, VUS_ERR_INVALID, "vus_set_components: empty component");
    for (long i = h->comp_start[c]; i < h->comp_start[c + 1]; ++i) node_comp[i] = c;
    seg[c] = h->comp_start[c] * h->D;
  }
  long longest = 1;
  for (int c = 0; c < nc; ++c) longest = std::max(longest, h->comp_start[c + 1] - h->comp_start[c]);
  h->bcr_stop = (longest + h->k - 1) / h->k + 1;          // supernodes a component can touch
  seg[nc] = NX * h->D;                                   // padding nodes carry identity rows and zero vectors
  const int slot_kind[VUS_F_NTYPES][5] = {{0}, {1}, {0, 0}, {1, 0}, {0, 3}, {0, 1, 0, 1, 2}};
  std::vector<int> fcomp(h->nfactors), cf_ptr(nc + 1, 0), ci_ptr(nc + 1, 0);
  for (int t = 0; t < VUS_F_NTYPES; ++t) {
    FactorTable& T = h->ft[t];
    for (long f = 0; f < T.n; ++f) {
      const int c = node_comp[T.h_idx[f]];
      for (int sl = 1; sl < kFactorSlots[t]; ++sl) {
        const int v = T.h_idx[sl * T.n + f];
        const int cv = slot_kind[t][sl] == 2 ? v : node_comp[v];
        if (cv != c) return fail(h, VUS_ERR_INVALID, "batched graph: a factor connects two components");
      }
      fcomp[T.e_off + f] = c;
      cf_ptr[c + 1]++;
      if (t == VUS_F_IMU) ci_ptr[c + 1]++;
    }
  }
  for (int c = 0; c < nc; ++c) { cf_ptr[c + 1] += cf_ptr[c]; ci_ptr[c + 1] += ci_ptr[c]; }
  std::vector<int> cf_list(h->nfactors), ci_list(h->ft[VUS_F_IMU].n);
  {
    std::vector<int> fill(cf_ptr.begin(), cf_ptr.end() - 1);
    for (long e = 0; e < h->nfactors; ++e) cf_list[fill[fcomp[e]]++] = (int)e;
    std::vector<int> filli(ci_ptr.begin(), ci_ptr.end() - 1);
    const FactorTable& I = h->ft[VUS_F_IMU];
    for (long f = 0; f < I.n; ++f) ci_list[filli[fcomp[I.e_off + f]]++] = (int)f;
  }
  h->node_comp.upload(node_comp, st); h->seg.upload(seg, st);
  h->cf_ptr.upload(cf_ptr, st); h->cf_list.upload(cf_list, st);
  h->ci_ptr.upload(ci_ptr, st); h->ci_list.upload(ci_list, st);
  // closures per component (unique off-band node pairs p < q); capacitance solve when every component has at most 8
  h->wb_R = 0; h->wb_lmax = 0;
  if (!rem_index.empty()) {
    std::vector<int> cnt(nc, 0);
    for (long b = 0; b < rem_index.size(); ++b)
      if (rem_index.p_of(b) < rem_index.q_of(b)) cnt[node_comp[rem_index.p_of(b)]]++;
    const int lmax = *std::max_element(cnt.begin(), cnt.end());
    if (lmax >= 1 && lmax <= 8) {
      std::vector<int> wnode((size_t)nc * lmax * 2, -1);
      std::vector<long> wblk((size_t)nc * lmax, 0);
      std::fill(cnt.begin(), cnt.end(), 0);
      for (long b = 0; b < rem_index.size(); ++b) {
        const long p = rem_index.p_of(b), q = rem_index.q_of(b);
        if (p >= q) continue;
        const int c = node_comp[p], l = cnt[c]++;
        wnode[((size_t)c * lmax + l) * 2] = (int)p; wnode[((size_t)c * lmax + l) * 2 + 1] = (int)q;
        wblk[(size_t)c * lmax + l] = h->rem_off + b * h->D * h->D;
      }
      h->wb_lmax = lmax; h->wb_R = 12 * lmax;
      h->wb_node.upload(wnode, st); h->wb_blk.upload(wblk, st);
      h->wbY.alloc((size_t)h->wb_R * h->L); h->wbCapInv.alloc((size_t)nc * h->wb_R * h->wb_R); h->wbW.alloc((size_t)nc * h->wb_R);
    }
  }
  h->scal_b.alloc((size_t)nc * SB_STRIDE); h->scal_b.zero(st);
  h->fail_mask.alloc(nc); h->lam_c.alloc(nc); h->acc_c.alloc(nc); h->cdots.alloc((size_t)nc * 36); h->cdots2.alloc((size_t)nc * 36); h->err_c.alloc(nc);
  rt::sync(st);
  return VUS_OK;
}

int analyze(vus_handle* h, rt::stream_t st) {
  const double t_an0 = now_ms();
  h->drop_graphs();
  double t_tick = t_an0;
  auto tick = [&](const char* what) {
    if (h->prm.verbose > 1) { const double t = now_ms(); std::fprintf(stderr, "  analyze %-28s %.1f ms\n", what, t - t_tick); t_tick = t; }
  };
  const long NX = h->nvar[0], NV = h->nvar[1], NB = h->nvar[2], NL = h->nvar[3];
  if (NX == 0) return fail(h, VUS_ERR_INVALID, "no Pose3 variables");
  if (h->ncomp <= 1 && NB > 1) return fail(h, VUS_ERR_UNSUPPORTED, "more than one imuBias variable: only the shared B(0) of batch.py:238/:274 is supported (or one per component, vus_set_components)");
  if (h->ncomp > 1) {
    if (NB != 0 && NB != h->ncomp) return fail(h, VUS_ERR_INVALID, "batched graph: one imuBias per component (or none)");
    if (h->ft[VUS_F_STEREO].n) return fail(h, VUS_ERR_UNSUPPORTED, "batched graph: stereo factors are not supported");
    if (h->n_owned >= 0) return fail(h, VUS_ERR_UNSUPPORTED, "batched graph: cannot be combined with a pose-range partition");
    if (h->comp_start.back() != NX) return fail(h, VUS_ERR_INVALID, "vus_set_components: node ranges must cover every pose");
  }
  h->D = NV > 0 ? 9 : 6;
  if (NV > 0) {
    if (NV != NX) return fail(h, VUS_ERR_UNSUPPORTED, "every pose X(i) needs a velocity V(i) (batch.py:283-288)");
    const uint64_t mask = (uint64_t(1) << 56) - 1;
    for (long i = 0; i < NX; ++i)
      if ((h->keys[0][i] & mask) != (h->keys[1][i] & mask))
        return fail(h, VUS_ERR_UNSUPPORTED, "velocity and pose symbol indices must match (V(i) with X(i))");
  }
  h->has_bias = NB ? 1 : 0;
  const int D = h->D;
  h->N = NX;
  // ---- node pairs from two-node factors and landmark tracks
  FactorTable& FB = h->ft[VUS_F_BETWEEN];
  FactorTable& FI = h->ft[VUS_F_IMU];
  FactorTable& FD = h->ft[VUS_F_DVL];
  FactorTable& FS = h->ft[VUS_F_STEREO];
  if (D == 6 && (FI.n || FD.n || h->ft[VUS_F_PRIOR_VEL].n))
    return fail(h, VUS_ERR_INVALID, "velocity factors without velocity variables");
  for (long f = 0; f < FI.n; ++f)
    if (FI.h_idx[f] != FI.h_idx[FI.n + f] || FI.h_idx[2 * FI.n + f] != FI.h_idx[3 * FI.n + f])
      return fail(h, VUS_ERR_UNSUPPORTED, "ImuFactor must connect (X(i),V(i)) to (X(j),V(j)) (batch.py:238)");
  for (long f = 0; f < FD.n; ++f)
    if (FD.h_idx[f] != FD.h_idx[FD.n + f])
      return fail(h, VUS_ERR_UNSUPPORTED, "DVL factor must connect V(i) and X(i) of the same keyframe (batch.py:247)");
  // stereo: keep the observation table landmark-major, pose-sorted inside a landmark (stable in factor order), so a
  // landmark's observations -- and the per-observation products the Schur kernels stream -- are contiguous
  h->nobs = FS.n;
  std::vector<int> lm_ptr(NL + 1, 0), pose_cnt(NX, 0);
  bool pose_obs_on_device = false;
#ifndef VUS_EMU
  if (FS.n) {
    // Stable device radix sort of (landmark, pose) keys gives the landmark-major row order; the index / measurement /
    // noise tables are re-ordered on the device and only the re-ordered indices come back to the host.  The caller-order
    // bookkeeping (orig, perm) is not needed by optimize(): it is derived from the order on first use (finish_order).
    finish_order(FS, st);
    DBuf<unsigned long long> k_in, k_out;
    DBuf<int> v_in, v_out, nidx, moved;
    k_in.alloc(FS.n); k_out.alloc(FS.n); v_in.alloc(FS.n); v_out.alloc(FS.n); nidx.alloc(2 * FS.n); moved.alloc(1); moved.zero(st);
    SortKeyArgs ka; ka.idx = FS.idx.p; ka.n = FS.n; ka.keys = k_in.p; ka.vals = v_in.p;
    L_elem<SortKeyBody>(FS.n, st, ka);
    int end_bit = 33;
    while (end_bit < 64 && (NL >> (end_bit - 32)) != 0) ++end_bit;
    size_t tmp_bytes = 0, tmp_bytes2 = 0;
    rt::check(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, (int)FS.n, 0, end_bit, st), "radix sort size");
    int pose_bits = 1;
    while (pose_bits < 32 && (NX >> pose_bits) != 0) ++pose_bits;
    DBuf<unsigned> pk_in, pk_out;
    DBuf<int> pv_in, pobs;
    pk_in.alloc(FS.n); pk_out.alloc(FS.n); pv_in.alloc(FS.n); pobs.alloc(FS.n);
    rt::check(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes2, pk_in.p, pk_out.p, pv_in.p, pobs.p, (int)FS.n, 0, pose_bits, st), "radix sort size");
    DBuf<unsigned char> tmp; tmp.alloc(std::max(tmp_bytes, tmp_bytes2));
    rt::check(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, k_in.p, k_out.p, v_in.p, v_out.p, (int)FS.n, 0, end_bit, st), "radix sort");
    GatherIdxArgs gi; gi.src = FS.idx.p; gi.dst = nidx.p; gi.perm = v_out.p; gi.n = FS.n; gi.slots = 2; gi.moved = moved.p;
    L_elem<GatherIdxBody>(2 * FS.n, st, gi);
    // pose-major observation lists of the re-ordered rows (stable: rows of one pose stay in row order)
    PoseKeyArgs pa; pa.idx = nidx.p; pa.n = FS.n; pa.keys = pk_in.p; pa.vals = pv_in.p;
    L_elem<PoseKeyBody>(FS.n, st, pa);
    rt::check(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes2, pk_in.p, pk_out.p, pv_in.p, pobs.p, (int)FS.n, 0, pose_bits, st), "radix sort");
    int h_moved = 0;
    rt::d2h(&h_moved, moved.p, sizeof(int), st);
    rt::d2h(FS.h_idx.data(), nidx.p, 2 * FS.n * sizeof(int), st);
    rt::sync(st);
    if (h_moved) {
      FS.idx.swap(nidx);
      DBuf<double> tmpd; tmpd.alloc((size_t)3 * FS.n);
      GatherArgs ga; ga.perm = v_out.p; ga.n = FS.n; ga.comps = 3;
      ga.src = FS.meas.p; ga.dst = tmpd.p;
      L_elem<GatherBody>(3 * FS.n, st, ga);
      FS.meas.swap(tmpd);
      ga.src = FS.sinfo.p; ga.dst = tmpd.p;
      L_elem<GatherBody>(3 * FS.n, st, ga);
      FS.sinfo.swap(tmpd);
      FS.d_order.swap(v_out);                           // orig / perm follow lazily
      rt::sync(st);
    } else if (FS.perm.empty()) {
      FS.perm.resize(FS.n);
      std::iota(FS.perm.begin(), FS.perm.end(), 0);
    }
    h->pose_obs.swap(pobs);
    pose_obs_on_device = true;
  }
  for (long o = 0; o < FS.n; ++o) lm_ptr[FS.h_idx[FS.n + o] + 1]++;
  for (long l = 0; l < NL; ++l) lm_ptr[l + 1] += lm_ptr[l];
#else
  for (long o = 0; o < FS.n; ++o) lm_ptr[FS.h_idx[FS.n + o] + 1]++;
  for (long l = 0; l < NL; ++l) lm_ptr[l + 1] += lm_ptr[l];
  if (FS.n) {
    std::vector<int> order(FS.n);
    {
      std::vector<int> fill(lm_ptr.begin(), lm_ptr.end() - 1);
      for (long o = 0; o < FS.n; ++o) order[fill[FS.h_idx[FS.n + o]]++] = (int)o;
    }
    for (long l = 0; l < NL; ++l) {                     // insertion sort by pose: tracks arrive (nearly) sorted
      int* q = order.data() + lm_ptr[l];
      const int len = lm_ptr[l + 1] - lm_ptr[l];
      for (int a = 1; a < len; ++a) {
        const int v = q[a];
        const int pv = FS.h_idx[v];
        int b = a - 1;
        while (b >= 0 && FS.h_idx[q[b]] > pv) { q[b + 1] = q[b]; --b; }
        q[b + 1] = v;
      }
    }
    bool identity = true;
    for (long o = 0; o < FS.n && identity; ++o) identity = order[o] == o;
    if (!identity) {
      std::vector<int> nidx(2 * FS.n);
      std::vector<int64_t> norig(FS.n);
      std::vector<int> nperm(FS.n);
      for (long o = 0; o < FS.n; ++o) {
        nidx[o] = FS.h_idx[order[o]]; nidx[FS.n + o] = FS.h_idx[FS.n + order[o]];
        norig[o] = FS.orig[order[o]];
        nperm[o] = FS.perm.empty() ? order[o] : FS.perm[order[o]];
      }
      FS.h_idx.swap(nidx); FS.orig.swap(norig); FS.perm.swap(nperm);
      FS.idx.upload(FS.h_idx, st);
      DBuf<int> dperm; dperm.upload(order, st);
      DBuf<double> tmp; tmp.alloc((size_t)3 * FS.n);
      GatherArgs ga; ga.perm = dperm.p; ga.n = FS.n; ga.comps = 3;
      ga.src = FS.meas.p; ga.dst = tmp.p;
      L_elem<GatherBody>(3 * FS.n, st, ga);
      rt::d2d(FS.meas.p, tmp.p, (size_t)3 * FS.n * sizeof(double), st);
      ga.src = FS.sinfo.p;
      L_elem<GatherBody>(3 * FS.n, st, ga);
      rt::d2d(FS.sinfo.p, tmp.p, (size_t)3 * FS.n * sizeof(double), st);
      rt::sync(st);
    } else if (FS.perm.empty()) {
      FS.perm.resize(FS.n);
      std::iota(FS.perm.begin(), FS.perm.end(), 0);
    }
  }
#endif
  tick("stereo landmark-major sort");
  for (long o = 0; o < FS.n; ++o) pose_cnt[FS.h_idx[o]]++;
  std::vector<int> pose_ids, pose_ptr(1, 0), pose_obs;
  {
    std::vector<int> slot(NX, -1);
    for (long i = 0; i < NX; ++i)
      if (pose_cnt[i]) { slot[i] = (int)pose_ids.size(); pose_ids.push_back((int)i); pose_ptr.push_back(pose_ptr.back() + pose_cnt[i]); }
    if (!pose_obs_on_device) {
      pose_obs.resize(FS.n);
      std::vector<int> fill(pose_ptr.begin(), pose_ptr.end() - 1);
      for (long o = 0; o < FS.n; ++o) pose_obs[fill[slot[FS.h_idx[o]]]++] = (int)o;
    }
  }
  h->nposes_obs = (long)pose_ids.size();
  for (long l = 0; l < NL; ++l)
    if (lm_ptr[l + 1] == lm_ptr[l]) return fail(h, VUS_ERR_INVALID, "a landmark has no stereo factor (indeterminate system)");
  tick("pose CSR");
  // ---- band width k
  const int kcap = h->prm.max_supernode > 0 ? h->prm.max_supernode : 96 / D;
  long span = 1;
  // spans up to kcap widen the band; longer links (loop closures, very long tracks) stay in the off-band remainder.
  // NOTE: the band is only guaranteed positive definite when every landmark track fits in it (the Schur complement
  // subtracts from the blocks it touches); tracks longer than kcap poses degrade the preconditioner.
  // (a rank of a partitioned graph: owned poses only -- the local index distance to a halo pose means nothing, and the couplings
  // to the neighbouring ranks' poses are tied in by spike.cuh)
  auto consider = [&](long p, long q) {
    if (h->n_owned >= 0 && (p >= h->n_owned || q >= h->n_owned)) return;
    long s = p > q ? p - q : q - p; if (s <= kcap && s > span) span = s;
  };
  for (long f = 0; f < FB.n; ++f) consider(FB.h_idx[f], FB.h_idx[FB.n + f]);
  for (long f = 0; f < FI.n; ++f) consider(FI.h_idx[f], FI.h_idx[2 * FI.n + f]);
  for (long l = 0; l < NL; ++l) {
    const int a = lm_ptr[l], b = lm_ptr[l + 1] - 1;
    // a track longer than the cap still wants a wide band: most such tracks then fit in two neighbouring supernodes after
    // all, and the segments of the others (below) are long.  Nine poses per supernode, not the cap of ten: 81 x 81 is the
    // block size every BASELINE stereo configuration runs (and the GPU test-suite covers).
    const long pa = FS.h_idx[a], pb = FS.h_idx[b], kwide = std::min<long>(kcap, 9);
    consider(pa, pb - pa > kcap ? pa + kwide : pb);
  }
  h->k = (int)std::min<long>(span, kcap);
  if (h->k < 1) h->k = 1;
  const int k = h->k;
  h->B = k * D;
  h->Ns = (NX + k - 1) / k;
  // a rank of a partitioned graph factors / multiplies only the supernodes of the nodes it owns: the halo nodes' rows belong
  // to their owners (with random loop closures the halo is as large as the owned range -- measured: twice the work otherwise)
  h->Ns_band = h->n_owned >= 0 ? std::max<long>(1, std::min<long>(h->Ns, (h->n_owned + k - 1) / k)) : h->Ns;
  h->Npad = h->Ns * k;
  h->Lc = h->Npad * D;
  h->L = h->Lc + 6 * NB;
  const long BB = bcr_bbp(h->B);                     // SD / SU are padded [KP][LD] tiles (bulk-copy layout, bcr.cuh)
  auto inband = [&](long p, long q) { long I = p / k, J = q / k; return (I - J <= 1) && (J - I <= 1); };
  // ---- off-band remainder blocks
  RemIndex rem_index;
  // (Measured and dropped: keeping sparse loop closures out of the band entirely -- diagonal blocks too -- makes the
  // operator a rank-6 instead of rank-12 update per closure and saves ~30 % of the PCG iterations, but the preconditioned
  // operator then has very large eigenvalues, the attainable true residual degrades from 1e-12 to ~1e-9 at small lambda
  // and the LM path leaves the oracle's.  The diagonal blocks of every factor stay in the band.)
  auto add_rem = [&](long p, long q) { if (p != q && !inband(p, q)) { rem_index.add(p, q); rem_index.add(q, p); } };
  for (long f = 0; f < FB.n; ++f) add_rem(FB.h_idx[f], FB.h_idx[FB.n + f]);
  for (long f = 0; f < FI.n; ++f) add_rem(FI.h_idx[f], FI.h_idx[2 * FI.n + f]);
  rem_index.finalize();
  // landmarks whose track does not fit inside the band keep their Schur term implicit (LongSchur*Body)
  std::vector<int> lm_long(NL, 0), long_ids;
  long track_span = 0;                                  // longest pose span of a track that is folded into the band
  for (long l = 0; l < NL; ++l) {
    const long pf = FS.h_idx[lm_ptr[l]], pl = FS.h_idx[lm_ptr[l + 1] - 1];   // observations are pose-sorted
    bool twice = false;                                 // seen twice from one pose: SchurBlockBody expects one partner per pose
    for (int q = lm_ptr[l] + 1; q < lm_ptr[l + 1] && !twice; ++q) twice = FS.h_idx[q] == FS.h_idx[q - 1];
    if (pl / k - pf / k <= 1 && !twice) { track_span = std::max(track_span, pl - pf); continue; }   // whole track inside the band
    long_ids.push_back((int)l);
    lm_long[l] = (int)long_ids.size();                  // 1 + position in long_ids
  }
  // The exact Schur term of a long track stays implicit in the OPERATOR.  For the PRECONDITIONER the track is cut into segments
  // that each fit in the band (two neighbouring supernodes, every pose at most once) and every segment acts as a landmark of
  // its own: the exact Schur complement of the problem in which the copies of the landmark are not tied together -- symmetric
  // positive definite, below the true reduced matrix in the SPD order, and far closer to it than a band without the track.
  h->nseg = 0;
  if (!long_ids.empty()) {
    std::vector<int> obs_seg(FS.n, -1), seg_ptr;          // seg_ptr: [begin, end) observation rows of every segment
    for (int l : long_ids) {
      int a = lm_ptr[l];
      while (a < lm_ptr[l + 1]) {
        int c = a + 1;
        while (c < lm_ptr[l + 1] && FS.h_idx[c] != FS.h_idx[c - 1] && FS.h_idx[c] / k - FS.h_idx[a] / k <= 1) ++c;
        for (int o = a; o < c; ++o) obs_seg[o] = (int)(seg_ptr.size() / 2);
        track_span = std::max<long>(track_span, FS.h_idx[c - 1] - FS.h_idx[a]);
        seg_ptr.push_back(a); seg_ptr.push_back(c);
        a = c;
      }
    }
    h->nseg = (long)(seg_ptr.size() / 2);
    h->obs_seg.upload(obs_seg, st); h->seg_ptr.upload(seg_ptr, st);
    h->CinvSeg.alloc((size_t)9 * h->nseg);
  }
  h->schur_ndj = (int)track_span + 1;
  h->nlong = (long)long_ids.size();
  h->lm_long.upload(lm_long, st);
  if (h->nlong) { h->long_ids.upload(long_ids, st); h->ulong.alloc((size_t)3 * h->nlong); }
  std::vector<int> rem_ptr(NX + 1, 0), rem_col;
  {
    rem_col.resize(rem_index.size());
    for (long b = 0; b < rem_index.size(); ++b) { rem_ptr[rem_index.p_of(b) + 1]++; rem_col[b] = (int)rem_index.q_of(b); }
    for (long i = 0; i < NX; ++i) rem_ptr[i + 1] += rem_ptr[i];
    h->nrem = rem_index.size();
  }
  h->sd_off = 0;
  h->su_off = h->Ns * BB;
  h->rem_off = h->su_off + (h->Ns > 1 ? (h->Ns - 1) * BB : 0);
  h->hlen = h->rem_off + h->nrem * D * D;
  tick("band width + remainder");
  // ---- gather lists of the chain-factor assembly (NodeAsmBody / PairAsmBody): node -> incident factors, node pair -> factors
  {
    FactorTable* tabs[VUS_F_NTYPES] = {&h->ft[VUS_F_PRIOR_POSE], &h->ft[VUS_F_PRIOR_VEL], &FB, &h->ft[VUS_F_DVL], nullptr, &FI};
    const int slot_p[VUS_F_NTYPES] = {0, 0, 0, 1, 0, 0}, slot_q[VUS_F_NTYPES] = {-1, -1, 1, -1, -1, 2};
    std::vector<int> nptr(NX + 1, 0);
    for (int t = 0; t < VUS_F_NTYPES; ++t) {
      if (!tabs[t]) continue;
      const FactorTable& T = *tabs[t];
      for (long f = 0; f < T.n; ++f) {
        const long p = T.h_idx[slot_p[t] * T.n + f];
        nptr[p + 1]++;
        if (slot_q[t] >= 0) {
          const long q = T.h_idx[slot_q[t] * T.n + f];
          if (p == q) return fail(h, VUS_ERR_INVALID, "a two-pose factor connects a pose with itself");
          nptr[q + 1]++;
        }
      }
    }
    for (long i = 0; i < NX; ++i) nptr[i + 1] += nptr[i];
    std::vector<int> ncode(nptr[NX]), nfac(nptr[NX]);
    // pair groups in order of first appearance (factor order: neighbouring groups read neighbouring table rows); an in-band
    // pair is found through a direct table over (lo, hi - lo), an off-band one through its remainder block
    std::vector<int> inband_grp((size_t)NX * 2 * k, -1), rem_grp(h->nrem, -1), gcount, fgroup[VUS_F_NTYPES];
    std::vector<PairDst> gdst;
    {
      std::vector<int> fill(nptr.begin(), nptr.end() - 1);
      for (int t = 0; t < VUS_F_NTYPES; ++t) {          // types in enum order, rows in insertion order, p before q: the summation order
        if (!tabs[t]) continue;
        const FactorTable& T = *tabs[t];
        if (slot_q[t] >= 0) fgroup[t].resize(T.n);
        for (long f = 0; f < T.n; ++f) {
          const long p = T.h_idx[slot_p[t] * T.n + f];
          ncode[fill[p]] = 2 * t; nfac[fill[p]++] = (int)f;
          if (slot_q[t] < 0) continue;
          const long q = T.h_idx[slot_q[t] * T.n + f];
          ncode[fill[q]] = 2 * t + 1; nfac[fill[q]++] = (int)f;
          const long lo = std::min(p, q), hi = std::max(p, q);
          int* slot = inband(lo, hi) ? &inband_grp[(size_t)lo * 2 * k + (hi - lo)] : &rem_grp[rem_index.at(lo, hi)];
          if (*slot < 0) { *slot = (int)gdst.size(); gdst.push_back(band_dst(h, lo, hi, rem_index)); gcount.push_back(0); }
          fgroup[t][f] = *slot;
          gcount[*slot]++;
        }
      }
    }
    h->npairs = (long)gdst.size();
    std::vector<int> gptr(h->npairs + 1, 0);
    for (long gi = 0; gi < h->npairs; ++gi) gptr[gi + 1] = gptr[gi] + gcount[gi];
    std::vector<int> gcode(gptr[h->npairs]), gfac(gptr[h->npairs]);
    {
      std::vector<int> fill(gptr.begin(), gptr.end() - 1);
      for (int t = 0; t < VUS_F_NTYPES; ++t) {
        if (!tabs[t] || slot_q[t] < 0) continue;
        const FactorTable& T = *tabs[t];
        for (long f = 0; f < T.n; ++f) {
          const int gi = fgroup[t][f];
          const bool flip = T.h_idx[slot_p[t] * T.n + f] > T.h_idx[slot_q[t] * T.n + f];
          gcode[fill[gi]] = 2 * t + (flip ? 1 : 0); gfac[fill[gi]++] = (int)f;
        }
      }
    }
    h->na_ptr.upload(nptr, st); h->na_code.upload(ncode, st); h->na_fac.upload(nfac, st);
    h->pg_ptr.upload(gptr, st); h->pg_code.upload(gcode, st); h->pg_fac.upload(gfac, st); h->pg_dst.upload(gdst, st);
  }
  tick("chain-factor gather lists");
  // ---- pose-range partition: the blocks that couple the owned chain to the neighbouring ranks (spike.cuh)
  h->spike_state = 0;
  if (h->n_owned >= 2 * k && h->sp_nranks > 1 && (long)h->sp_prev.size() >= k && (long)h->sp_next.size() >= k) {
    std::vector<PairDst> pairs((size_t)2 * k * k);
    for (auto& d : pairs) { d.off = -1; d.moff = -1; d.ld = d.mld = d.transposed = d.pad = 0; }
    auto coupled = [&](long p, long q) { return q >= 0 && q < NX && (inband(p, q) || rem_index.has(p, q)); };
    for (int p = 0; p < k; ++p)
      for (int j = 0; j < k; ++j) {
        const long last_p = h->n_owned - k + p, nxt = h->sp_next[j];           // next pose j: global b + j
        if (coupled(last_p, nxt)) pairs[(size_t)p * k + j] = band_dst(h, last_p, nxt, rem_index);
        const long prv = h->sp_prev[k - 1 - j];                                 // previous pose j: global a - k + j
        if (coupled(p, prv)) pairs[(size_t)(k + p) * k + j] = band_dst(h, p, prv, rem_index);
      }
    h->spPairs.upload(pairs, st);
  } else if (h->sp_nranks > 1) {
    h->spike_state = -2;                                  // cannot here: says so in the first set-up's all-reduce
  }
  // ---- uploads / allocations
  h->rem_ptr.upload(rem_ptr, st); h->rem_col.upload(rem_col, st);
  h->pose_ptr.upload(pose_ptr, st); if (!pose_obs_on_device) h->pose_obs.upload(pose_obs, st); h->pose_ids.upload(pose_ids, st);
  h->lm_ptr.upload(lm_ptr, st);
  h->H0.alloc(h->hlen); h->H.alloc(h->hlen); h->H_synced = false;
  if (h->nseg) h->Hp.alloc(h->hlen);
  h->g0.alloc(h->Lc); h->gs.alloc(h->L);
  h->F.alloc(h->Lc * 6);
  // zero-filled once: the assembly kernels store every entry a factor can touch, nothing else ever writes the rest
  h->H0.zero(st); h->g0.zero(st); h->F.zero(st);
  h->Hbb0.alloc(36 * std::max<long>(NB, 1)); h->Hbb.alloc(36 * std::max<long>(NB, 1)); h->gb.alloc(6 * std::max<long>(NB, 1));
  h->C.alloc(9 * NL); h->gl.alloc(3 * NL); h->Cinv.alloc(9 * NL); h->E.alloc(18 * FS.n); h->Pp.alloc(28 * FS.n); h->Pl.alloc(12 * FS.n);
  {   // the reduction's own blocks are padded [KP][LD] tiles; the padding must be (and stays) zero
    // Stereo-scale supernodes on one long chain: block Cholesky inside P chunks (one per SM), cyclic reduction across the
    // P - 1 separators only (chunk.cuh).  Tiny supernodes (chain graphs), batched and partitioned graphs keep the plain reduction.
    h->chunk_P = 0;
    if (h->ncomp <= 1 && h->n_owned < 0 && h->B > VUS_SMALLB_MAX && h->prm.band_chunks >= 0) {
      long P = h->prm.band_chunks > 0 ? h->prm.band_chunks : std::min<long>(rt::sm_count(), h->Ns / 6);
      P = std::min<long>(P, (h->Ns + 1) / 2);            // every chunk needs at least one interior supernode
      if (P >= 2) { h->chunk_P = (int)P; h->cgeom = chunk_geom(h->Ns, (int)P); }
    }
    const size_t nb = (size_t)h->Ns * bcr_bbp(h->B);
    for (DBuf<double>* b : {&h->Dinv, &h->Gl, &h->Gr}) { b->alloc(nb); b->zero(st); }
    if (!h->chunk_P) {
      for (DBuf<double>* b : {&h->Dw, &h->U1, &h->U2}) { b->alloc(nb); b->zero(st); }
    } else {
      const size_t np = (size_t)h->chunk_P * bcr_bbp(h->B);
      for (DBuf<double>* b : {&h->SepL, &h->SepR, &h->SepU, &h->Dsep, &h->Usep, &h->sDw, &h->sU1, &h->sU2, &h->sDinv, &h->sGl, &h->sGr}) { b->alloc(np); b->zero(st); }
      h->tL.alloc((size_t)h->chunk_P * VUS_CHUNK_MAXV * h->B); h->tL.zero(st);
      h->tR.alloc((size_t)h->chunk_P * VUS_CHUNK_MAXV * h->B); h->tR.zero(st);
      h->xsep.alloc((size_t)VUS_CHUNK_MAXV * h->chunk_P * h->B); h->xsep.zero(st);
    }
  }
  h->Z.alloc(7 * h->Lc); h->Zr.alloc(6 * h->Lc); h->SbInv.alloc(36 * std::max<long>(NB, 1));
  if (h->n_owned >= 0) {
    if (h->has_bias || FS.n || NV) return fail(h, VUS_ERR_UNSUPPORTED, "pose-range partition supports pose graphs (PriorFactorPose3 / BetweenFactorPose3) only");
    if (h->n_owned > NX) return fail(h, VUS_ERR_INVALID, "vus_set_partition: more owned nodes than nodes");
    h->Lr = h->n_owned * D;
    std::vector<double> mask((size_t)h->nfactors, 0.0);
    long off = 0;
    for (int t = 0; t < VUS_F_NTYPES; ++t) {
      if (h->nf_owned[t] > h->ft[t].n) return fail(h, VUS_ERR_INVALID, "vus_set_partition: more owned factors than factors");
      for (long f = 0; f < h->nf_owned[t]; ++f) mask[off + f] = 1.0;
      off += h->ft[t].n;
    }
    h->e_mask.upload(mask, st);
  } else {
    h->Lr = h->L;
  }
  h->x.alloc(h->L); h->d.alloc(h->L); h->r.alloc(h->L); h->z.alloc(h->L); h->p.alloc(h->L); h->Ap.alloc(h->L);
  h->xl.alloc(3 * NL);
  h->scal.alloc(S_COUNT); h->scal.zero(st);
  h->red_grid = std::min<long>(std::max<long>(1, (std::max(h->L, h->nfactors) + 4095) / 4096), 2L * rt::sm_count());
  h->partials.alloc(h->red_grid);
  h->bpart.alloc((size_t)h->red_grid * 42);
  h->bpart2.alloc((size_t)h->red_grid * 36);
  h->fail.alloc(std::max(h->ncomp, 1)); h->fail.zero(st);     // one flag per component (batched mode), else one
  long eoff = 0;
  for (int t = 0; t < VUS_F_NTYPES; ++t) {
    FactorTable& T = h->ft[t];
    T.e_off = eoff; eoff += T.n;
    T.r.alloc((size_t)kFactorM[t] * T.n);
    T.J.alloc((size_t)kFactorM[t] * kFactorCols[t] * T.n);
  }
  h->e_all.alloc(eoff); h->le_all.alloc(eoff);
  if (h->ncomp > 1) {
    const int rc = analyze_components(h, st, rem_index);
    if (rc != VUS_OK) return rc;
  }
  for (int kind = 0; kind < 4; ++kind) h->val[1 - h->cur][kind].alloc((size_t)kVarDim[kind] * h->nvar[kind]);
  if (FS.n) {                                           // partner table of the per-lambda Schur kernel
    h->partner.alloc((size_t)FS.n * h->schur_ndj);
    SchurArgs pa = schur_args(h, 0.0);
    L_elem<SchurPartnerBody>(FS.n * h->schur_ndj, st, pa);
  }
  rt::sync(st);
  tick("uploads + allocations");
  h->analyzed = true;
  if (h->prm.verbose) std::fprintf(stderr, "vus_analyze: %.1f ms (D=%d k=%d Ns=%ld nrem=%ld)\n", now_ms() - t_an0, h->D, h->k, h->Ns, h->nrem);
  return VUS_OK;
}

// ------------------------------------------------------------------ kernel 2 driver: base (undamped) system
ChainTables chain_tables(vus_handle* h) {
  ChainTables T;
  for (int t = 0; t < VUS_F_NTYPES; ++t) { T.n[t] = h->ft[t].n; T.J[t] = h->ft[t].J.p; T.r[t] = h->ft[t].r.p; }
  return T;
}
void assemble_base(vus_handle* h, rt::stream_t st) {
  ClassGuard kc_guard(KC_ASSEMBLE);
  h->Hbb0.zero(st); h->gb.zero(st);
  {
    NodeAsmArgs a;
    a.T = chain_tables(h); a.nnodes = h->N; a.D = h->D; a.k = h->k; a.ld = bcr_ld(h->B); a.bs = bcr_bbp(h->B);
    a.has_bias = h->has_bias ? 1 : 0;
    a.ptr = h->na_ptr.p; a.code = h->na_code.p; a.fac = h->na_fac.p;
    a.SD = h->H0.p + h->sd_off; a.g = h->g0.p; a.F = h->F.p;
    L_coop<NodeAsmBody>((int)((h->N + 31) / 32), 32 * h->D, 0, st, a);
  }
  if (h->npairs) {
    PairAsmArgs a;
    a.T = chain_tables(h); a.ngroups = h->npairs; a.D = h->D;
    a.ptr = h->pg_ptr.p; a.code = h->pg_code.p; a.fac = h->pg_fac.p; a.dst = h->pg_dst.p; a.Hval = h->H0.p;
    L_coop<PairAsmBody>((int)((h->npairs + 31) / 32), 32 * h->D, 0, st, a);
  }
  if (h->ft[VUS_F_IMU].n) {
    FactorTable& I = h->ft[VUS_F_IMU];
    if (h->ncomp > 1) {
      if (h->has_bias) {
        BImuBiasArgs bb; bb.n = I.n; bb.J = I.J.p; bb.r = I.r.p; bb.ptr = h->ci_ptr.p; bb.list = h->ci_list.p; bb.Hbb = h->Hbb0.p; bb.gb = h->gb.p;
        L_coop<BImuBiasBody>(h->ncomp, 128, 128 * sizeof(double), st, bb);
      }
    } else {
    ImuBiasArgs b; b.n = I.n; b.J = I.J.p; b.r = I.r.p; b.partials = h->bpart.p; b.grid = h->red_grid; b.Hbb = h->Hbb0.p; b.gb = h->gb.p;
    L_coop<ImuBias1Body>(h->red_grid, 256, 256 * sizeof(double), st, b);
    L_elem<ImuBias2Body>(42, st, b);
    }
  }
  FactorTable& S = h->ft[VUS_F_STEREO];
  if (S.n) {
    ClassGuard kc_stereo(KC_STEREO_ASM);
    StereoAsmArgs a;
    a.n = S.n; a.idx = S.idx.p; a.J = S.J.p; a.r = S.r.p; a.D = h->D; a.k = h->k; a.B = h->B; a.ld = bcr_ld(h->B); a.bs = bcr_bbp(h->B);
    a.SD = h->H0.p + h->sd_off; a.g = h->g0.p; a.C = h->C.p; a.gl = h->gl.p; a.E = h->E.p; a.nl = h->nvar[3];
    a.pose_ptr = h->pose_ptr.p; a.pose_obs = h->pose_obs.p; a.pose_ids = h->pose_ids.p; a.nposes_obs = h->nposes_obs;
    a.lm_ptr = h->lm_ptr.p;
    a.Pp = h->Pp.p; a.Pl = h->Pl.p;
    L_elem<StereoPoseBody>(h->nposes_obs * 28, st, a);
    L_elem<StereoLmBody>(h->nvar[3] * 12, st, a);
  }
}

SchurArgs schur_args(vus_handle* h, double lambda) {
  FactorTable& S = h->ft[VUS_F_STEREO];
  SchurArgs a;
  a.n = S.n; a.nl = h->nvar[3]; a.idx = S.idx.p; a.C = h->C.p; a.gl = h->gl.p; a.Cinv = h->Cinv.p; a.E = h->E.p;
  a.lambda = lambda; a.D = h->D; a.k = h->k; a.B = h->B; a.ld = bcr_ld(h->B); a.bs = bcr_bbp(h->B);
  a.SD = h->H.p + h->sd_off; a.SU = h->H.p + h->su_off; a.gs = h->gs.p;
  a.SD0 = h->H0.p + h->sd_off; a.SU0 = h->H0.p + h->su_off;
  a.pose_ptr = h->pose_ptr.p; a.pose_obs = h->pose_obs.p; a.pose_ids = h->pose_ids.p; a.nposes_obs = h->nposes_obs;
  a.lm_ptr = h->lm_ptr.p; a.lm_long = h->lm_long.p;
  a.fail = h->fail.p; a.xc = h->x.p; a.xl = h->xl.p;
  a.nlong = h->nlong; a.long_ids = h->long_ids.p; a.ulong = h->ulong.p; a.xin = nullptr; a.yout = nullptr;
  a.partner = h->partner.p; a.ndj = h->schur_ndj;
  a.obs_seg = h->nseg ? h->obs_seg.p : nullptr; a.seg_ptr = h->seg_ptr.p; a.CinvSeg = h->CinvSeg.p; a.nseg = h->nseg; a.Pl = h->Pl.p; a.long_pass = 0; a.has_long = h->nlong > 0;
  return a;
}

// damped + Schur-reduced system for this lambda
void form_system(vus_handle* h, double lambda, rt::stream_t st) {
  ClassGuard kc_guard(KC_SCHUR);
  h->cur_lambda = lambda;
  // The first system of a graph is a full copy of the base system (2.3 GB at config 3); after that only the blocks that can
  // change are refreshed (CopyBlocksBody: ~0.2 GB) -- the rest of SD | SU | REM is structurally zero or assigned by the Schur kernel.
  const bool full = !h->H_synced;
  if (full) {
    rt::d2d(h->H.p, h->H0.p, h->hlen * sizeof(double), st);
    h->H_synced = true;
  } else {
    CopyBlocksArgs c; c.H0 = h->H0.p; c.H = h->H.p; c.nnodes = h->N; c.ngroups = h->npairs; c.D = h->D; c.k = h->k; c.ld = bcr_ld(h->B);
    c.bs = bcr_bbp(h->B); c.sd_off = h->sd_off; c.dst = h->pg_dst.p;
    L_elem<CopyBlocksBody>((h->N + h->npairs) * (long)(h->D * h->D), st, c);
  }
  rt::d2d(h->Hbb.p, h->Hbb0.p, 36 * sizeof(double), st);
  rt::d2d(h->gs.p, h->g0.p, h->Lc * sizeof(double), st);
  if (h->has_bias) rt::d2d(h->gs.p + h->Lc, h->gb.p, 6 * sizeof(double), st);
  DampArgs d; d.SD = h->H.p + h->sd_off; d.Hbb = h->Hbb.p; d.ndof = h->Lc; d.nreal = h->N * h->D; d.B = h->B; d.lambda = lambda; d.ld = bcr_ld(h->B); d.bs = bcr_bbp(h->B);
  d.pad_add = full ? 1.0 : 0.0;
  L_elem<DampBody>(h->Lc + (h->has_bias ? 6 : 0), st, d);
  if (h->nobs) {
    SchurArgs a = schur_args(h, lambda);
    L_elem<LmInvertBody>(a.nl, st, a);
    L_elem<SchurBlockBody>(6 * h->nposes_obs * h->schur_ndj, st, a);
    if (h->nseg) {                                       // the preconditioner's band: the finished system minus the segment terms
      L_elem<SegInvertBody>(h->nseg, st, a);
      rt::d2d(h->Hp.p, h->H.p, h->hlen * sizeof(double), st);
      SchurArgs b = a;
      b.long_pass = 1; b.SD = h->Hp.p + h->sd_off; b.SU = h->Hp.p + h->su_off;
      L_elem<SchurBlockBody>(6 * h->nposes_obs * h->schur_ndj, st, b);
    }
  }
}
// the band matrix the preconditioner factors (and that the border elimination refers to): the damped system itself, or -- with
// long tracks -- the copy that also carries their segment terms
double* band_matrix(vus_handle* h) { return h->nseg ? h->Hp.p : h->H.p; }

// ------------------------------------------------------------------ kernel 3 drivers
// The reduction stops at the first stride that no coupling can span: the whole chain for one graph; the longest component
// for a batch of independent trajectories (their couplings across component boundaries are zero), where the nodes left
// -- the multiples of that stride, one or none per component -- are all roots and are inverted / solved in one launch.
// A block-tridiagonal system the cyclic reduction works on: the whole supernode chain, or -- in chunk mode -- the separators.
struct BandSys {
  long Ns;
  const double* D; const double* U;                  // input blocks (padded tiles)
  double* Dw; double* U1; double* U2;                // working blocks
  double* Dinv; double* Gl; double* Gr;              // factors
};
BandSys band_sys(vus_handle* h) {
  BandSys y;
  if (h->chunk_P) {
    y.Ns = h->chunk_P - 1; y.D = h->Dsep.p; y.U = h->Usep.p;
    y.Dw = h->sDw.p; y.U1 = h->sU1.p; y.U2 = h->sU2.p; y.Dinv = h->sDinv.p; y.Gl = h->sGl.p; y.Gr = h->sGr.p;
  } else {
    y.Ns = h->Ns_band; y.D = band_matrix(h) + h->sd_off; y.U = band_matrix(h) + h->su_off;
    y.Dw = h->Dw.p; y.U1 = h->U1.p; y.U2 = h->U2.p; y.Dinv = h->Dinv.p; y.Gl = h->Gl.p; y.Gr = h->Gr.p;
  }
  return y;
}
long bcr_root_stride(const vus_handle* h, long Ns) {
  long s = 1;
  while (s < Ns && s < h->bcr_stop) s <<= 1;
  return s;
}
size_t bcr_smem(int B) { return (size_t)bcr_smem_doubles(B) * sizeof(double); }

BcrArgs bcr_args(vus_handle* h, const BandSys& y) {
  BcrArgs a;
  a.Ns = y.Ns; a.B = h->B; a.s = 1; a.root_stride = bcr_root_stride(h, y.Ns); a.small_g = VUS_SMALLB_G;
  a.Dsrc = nullptr; a.d_ld = 0; a.d_stride = 0; a.Dw = y.Dw;
  a.Ucur = nullptr; a.u_ld = 0; a.u_stride = 0; a.Unext = nullptr;
  a.Dinv = y.Dinv; a.Gl = y.Gl; a.Gr = y.Gr; a.fail = h->fail.p;
  a.node_comp = h->ncomp > 1 ? h->node_comp.p : nullptr; a.knodes = h->k;
  a.X = nullptr; a.xstride = 0; a.nrhs = 1;
  return a;
}
ChunkArgs chunk_args(vus_handle* h) {
  ChunkArgs c;
  c.G = h->cgeom; c.B = h->B; c.D = h->D;
  c.SD = band_matrix(h) + h->sd_off; c.SU = band_matrix(h) + h->su_off;
  c.Linv = h->Dinv.p; c.X = h->Gl.p; c.W = h->Gr.p;
  c.SepL = h->SepL.p; c.SepR = h->SepR.p; c.SepU = h->SepU.p; c.Dsep = h->Dsep.p; c.Usep = h->Usep.p;
  c.fail = h->fail.p;
  c.Xv = nullptr; c.xstride = 0; c.nrhs = 1;
  c.tL = h->tL.p; c.tR = h->tR.p; c.xsep = h->xsep.p; c.sepstride = (long)(h->chunk_P - 1) * h->B;
  return c;
}

void bcr_factor_launches(vus_handle* h, rt::stream_t st);
void bcr_factor(vus_handle* h, rt::stream_t st) {
  run_graphed(h->g_factor, st, [&] { bcr_factor_launches(h, st); });
}
void bcr_factor_launches(vus_handle* h, rt::stream_t st) {
  ClassGuard kc_guard(KC_BCR_FACTOR);
  const long BBP = bcr_bbp(h->B);
  const int LD = bcr_ld(h->B);
  if (h->chunk_P) {                                    // chunk interiors first: they produce the separator system
    ChunkArgs c = chunk_args(h);
    L_coop<ChunkFactorBody>(h->chunk_P, VUS_CH_THREADS, chunk_factor_smem(h->B), st, c);
    L_elem<SepAssembleBody>((long)(h->chunk_P - 1) * BBP, st, c);
  }
  const BandSys y = band_sys(h);
  BcrArgs a = bcr_args(h, y);
  // level 1 reads the assembled system in place; deeper levels read the working arrays (all padded tiles)
  a.Dsrc = y.D; a.d_ld = LD; a.d_stride = BBP;
  a.Ucur = y.U; a.u_ld = LD; a.u_stride = BBP;
  double* bufs[2] = {y.U1, y.U2};
  int w = 0;
  const long s_root = bcr_root_stride(h, y.Ns);
  for (long s = 1; s < s_root; s <<= 1) {
    const long nact = (y.Ns + s - 1) / s;
    const int nel = (int)(nact / 2), nsv = (int)((nact + 1) / 2);
    a.s = s; a.Unext = bufs[w];
    if (h->B <= VUS_SMALLB_MAX) {
      L_coop<SmallElimBody>((nel + VUS_SMALLB_G - 1) / VUS_SMALLB_G, 256, (size_t)VUS_SMALLB_G * (h->B * h->B + 1) * sizeof(double), st, a);
      L_elem<SmallUpdateBody>((long)nsv * h->B * h->B, st, a);
    } else {
      L_coop<BcrElimBody>(nel, 256, bcr_smem(h->B), st, a);
      L_coop<BcrUpdateBody>(nsv, 256, bcr_smem(h->B), st, a);
    }
    a.Dsrc = y.Dw; a.d_ld = LD; a.d_stride = BBP;
    a.Ucur = bufs[w]; a.u_ld = LD; a.u_stride = BBP;
    w ^= 1;
  }
  L_coop<BcrRootBody>((int)((y.Ns + s_root - 1) / s_root), 256, bcr_smem(h->B), st, a);
}

// in-place solve of the band system for nrhs vectors X[v*xstride + ...]
void bcr_solve_launches(vus_handle* h, double* X, long xstride, int nrhs, rt::stream_t st);
void bcr_solve(vus_handle* h, double* X, long xstride, int nrhs, rt::stream_t st, bool cacheable = true) {
  if (!cacheable) { bcr_solve_launches(h, X, xstride, nrhs, st); return; }
  GraphCache& gc = h->g_solve[std::make_tuple((const double*)X, xstride, nrhs)];
  run_graphed(gc, st, [&] { bcr_solve_launches(h, X, xstride, nrhs, st); });
}
void bcr_system_solve_launches(vus_handle* h, const BandSys& y, double* X, long xstride, int nrhs, rt::stream_t st);
void bcr_solve_launches(vus_handle* h, double* X, long xstride, int nrhs, rt::stream_t st) {
  ClassGuard kc_guard(KC_BCR_SOLVE);
  const BandSys y = band_sys(h);
  if (!h->chunk_P) { bcr_system_solve_launches(h, y, X, xstride, nrhs, st); return; }
  if (nrhs > VUS_CHUNK_MAXV) throw std::runtime_error("band solve: at most 8 right-hand sides at a time");
  ChunkArgs c = chunk_args(h);
  c.Xv = X; c.xstride = xstride; c.nrhs = nrhs;
  const int P = h->chunk_P;
  L_coop<ChunkFwdBody>(P, 256, chunk_sweep_smem(h->B), st, c);          // interiors forward, separator right-hand sides
  L_elem<SepRhsBody>((long)nrhs * (P - 1) * h->B, st, c);
  bcr_system_solve_launches(h, y, h->xsep.p, c.sepstride, nrhs, st);    // separators by cyclic reduction
  L_coop<ChunkBwdBody>(P, 256, chunk_sweep_smem(h->B), st, c);          // interiors backward
}
void bcr_system_solve_launches(vus_handle* h, const BandSys& y, double* X, long xstride, int nrhs, rt::stream_t st) {
  BcrArgs a = bcr_args(h, y);
  a.X = X; a.xstride = xstride; a.nrhs = nrhs;
  const int nthr = 256;
  std::vector<long> levels;
  const long s_root = bcr_root_stride(h, y.Ns);
  const int nroot = (int)((y.Ns + s_root - 1) / s_root);
  for (long s = 1; s < s_root; s <<= 1) levels.push_back(s);
  if (h->B <= VUS_SMALLB_MAX) {                          // tiny supernodes: thread per (node, row), VUS_SMALLB_G nodes per CTA
    const int GN = std::max(1, std::min(64, nthr / h->B));     // one (node, row) item per thread, all vectors of the item in registers
    a.small_g = GN;
    const size_t sm_small = (size_t)GN * h->B * nrhs * sizeof(double);
    auto grid_of = [GN](long n) { return (int)((n + GN - 1) / GN); };
    auto fwd = [&](int grid) {
      if (h->B == 9) L_coop<SmallFwdBody<9>>(grid, nthr, 0, st, a);
      else if (h->B == 12) L_coop<SmallFwdBody<12>>(grid, nthr, 0, st, a);
      else if (h->B == 6) L_coop<SmallFwdBody<6>>(grid, nthr, 0, st, a);
      else L_coop<SmallFwdBody<0>>(grid, nthr, 0, st, a);
    };
    auto bwd = [&](int grid) {
      if (h->B == 9) L_coop<SmallBwdBody<9>>(grid, nthr, sm_small, st, a);
      else if (h->B == 12) L_coop<SmallBwdBody<12>>(grid, nthr, sm_small, st, a);
      else if (h->B == 6) L_coop<SmallBwdBody<6>>(grid, nthr, sm_small, st, a);
      else L_coop<SmallBwdBody<0>>(grid, nthr, sm_small, st, a);
    };
    for (long s : levels) {
      a.s = s;
      fwd(grid_of(((y.Ns + s - 1) / s + 1) / 2));
    }
    a.s = 0;
    bwd(grid_of(nroot));
    for (auto it = levels.rbegin(); it != levels.rend(); ++it) {
      a.s = *it;
      bwd(grid_of(((y.Ns + a.s - 1) / a.s) / 2));
    }
    return;
  }
  const size_t smem = (size_t)blk_smem_doubles(h->B, nrhs, nthr) * sizeof(double);
  // levels with at most one CTA per SM use the DEEP bodies (every block of a node in flight at once)
  const size_t smem_deep = (size_t)3 * blk_slot_doubles(h->B) * sizeof(double);
  const int deep_max = smem_deep <= 220 * 1024 ? rt::sm_count() : 0;
  for (long s : levels) {
    const long nact = (y.Ns + s - 1) / s;
    a.s = s;
    const int grid = (int)((nact + 1) / 2);
    if (grid <= deep_max) L_coop<BcrFwdDeepBody>(grid, nthr, smem_deep, st, a);
    else L_coop<BcrFwdBody>(grid, nthr, smem, st, a);
  }
  L_coop<BcrRootSolveBody>(nroot, nthr, smem, st, a);
  for (auto it = levels.rbegin(); it != levels.rend(); ++it) {
    const long s = *it;
    const long nact = (y.Ns + s - 1) / s;
    a.s = s;
    const int grid = (int)(nact / 2);
    if (grid <= deep_max) L_coop<BcrBwdDeepBody>(grid, nthr, smem_deep, st, a);
    else L_coop<BcrBwdBody>(grid, nthr, smem, st, a);
  }
}

void border_dot(vus_handle* h, const double* Y, long ystride, int nv, rt::stream_t st) {
  ClassGuard kc_guard(KC_BORDER);
  BorderDotArgs a; a.F = h->F.p; a.Y = Y; a.len = h->Lc; a.ystride = ystride; a.nv = nv; a.partials = h->bpart.p; a.grid = h->red_grid;
  L_coop<BorderDot1Body>(h->red_grid, 256, 6 * 256 * sizeof(double), st, a);
}

void small_matvec(vus_handle* h, long items, const MatvecArgs& a, rt::stream_t st) {
  if (h->B == 9) L_elem<SmallMatvecBody<9>>(items, st, a);
  else if (h->B == 12) L_elem<SmallMatvecBody<12>>(items, st, a);     // pose graphs with skip factors (config 5): k = 2, D = 6
  else if (h->B == 6) L_elem<SmallMatvecBody<6>>(items, st, a);
  else L_elem<SmallMatvecBody<0>>(items, st, a);
}
// Y = band(SD, SU) X for nv vectors (no remainder, no border)
void apply_band(vus_handle* h, double* Y, long ystride, const double* X, long xstride, int nv, rt::stream_t st) {
  ClassGuard kc_guard(KC_MATVEC);
  MatvecArgs a;
  a.SD = band_matrix(h) + h->sd_off; a.SU = band_matrix(h) + h->su_off; a.Ns = h->Ns; a.B = h->B; a.x = X; a.y = Y;
  a.nv = nv; a.xstride = xstride; a.ystride = ystride; a.Nrows = h->Ns_band;
  a.rem_ptr = nullptr; a.rem_col = nullptr; a.rem_val = nullptr; a.nnodes = h->N; a.D = h->D;
  a.F = nullptr; a.Hbb = nullptr; a.xb = nullptr; a.yb = nullptr; a.has_bias = 0;
  if (h->B <= VUS_SMALLB_MAX) { small_matvec(h, h->Ns_band * h->B * nv, a, st); return; }
  const size_t smem = (size_t)blk_smem_doubles(h->B, nv, 256) * sizeof(double);
  L_coop<BandMatvecBody>((int)h->Ns_band, 256, smem, st, a);
}

bool spike_enabled(vus_handle* h, rt::stream_t st);
void spike_setup(vus_handle* h, rt::stream_t st);
void spike_apply(vus_handle* h, double* z, rt::stream_t st);
// preconditioner set-up for the current damped system: BCR of the band, Z = M^-1 F, Sb^-1
// with_rhs: the band solve of the right-hand side gs rides along as a seventh vector (the DMMA panel is 8 wide, so it
// costs nothing): the first preconditioner application of the PCG that follows takes it from there.
void precond_setup(vus_handle* h, rt::stream_t st, bool with_rhs = false) {
  bcr_factor(h, st);
  h->z0_valid = false;
  if (spike_enabled(h, st)) spike_setup(h, st);
  if (h->has_bias) {
    BorderColsArgs c; c.F = h->F.p; c.Z = h->Z.p; c.len = h->Lc; c.zstride = h->Lc; c.R = h->Zr.p;
    L_elem<BorderColsBody>(h->Lc * 6, st, c);
    if (with_rhs) {
      rt::d2d(h->Z.p + 6 * h->Lc, h->gs.p, h->Lc * sizeof(double), st);
      bcr_solve(h, h->Z.p, h->Lc, 7, st);
      h->z0_valid = true;
    } else {
      bcr_solve(h, h->Z.p, h->Lc, 6, st);
    }
    // The bias Schur complement Hbb - F^T M^-1 F cancels to ~1e-8 of its terms on short trajectories (the bias is barely
    // observable), far below the raw ~1e-7 accuracy of the band solve.  With the residual R = F - M Z of the computed
    // Z,  F^T M^-1 F = F^T Z + Z^T R + O(|dZ|^2): one 6-vector band product and a 6x6 dot product restore the
    // complement to second order without a second band solve (Z itself only feeds the preconditioner).
    apply_band(h, h->Zr.p, h->Lc, h->Z.p, h->Lc, 6, st);
    L_elem<BorderResidBody>(h->Lc * 6, st, c);
    border_dot(h, h->Z.p, h->Lc, 6, st);
    {
      ClassGuard kc_b(KC_BORDER);
      ColDotArgs d; d.Z = h->Z.p; d.R = h->Zr.p; d.len = h->Lc; d.stride = h->Lc; d.partials = h->bpart2.p; d.grid = h->red_grid;
      L_coop<ColDot1Body>(h->red_grid, 256, 6 * 256 * sizeof(double), st, d);
    }
    BorderSchurArgs s; s.Hbb = h->Hbb.p; s.partials = h->bpart.p; s.grid = h->red_grid; s.nv = 6; s.SbInv = h->SbInv.p; s.fail = h->fail.p;
    s.corr = h->bpart2.p;
    L_coop<BorderSchurBody>(1, 128, 72 * sizeof(double), st, s);
  }
}

// ---- exact band across the ranks of a pose-range partition (spike.cuh)
SpikeArgs spike_args(vus_handle* h) {
  SpikeArgs a;
  a.P = h->sp_nranks; a.rank = h->sp_rank; a.k = h->k; a.K = 6 * h->k; a.Ls = h->Ns_band * h->B; a.n_owned = h->n_owned;
  a.pairs = h->spPairs.p; a.Hval = h->H.p;
  a.Z = h->spZ.p; a.G = h->spG.p; a.R = h->spR.p; a.Rinv = h->spRinv.p; a.Y = h->spY.p; a.coef = h->spCoef.p;
  a.z = nullptr; a.fail = h->fail.p;
  return a;
}
// Every rank must take part in the collectives or none, with the same supernode width: the first set-up all-reduces
// (cannot, k, k^2) and the scheme is on only if nobody said "cannot" and every rank has the same k.
bool spike_enabled(vus_handle* h, rt::stream_t st) {
  if (h->n_owned < 0 || h->sp_nranks <= 1 || !has_comm(h)) return false;
  if (h->spike_state == 0 || h->spike_state == -2) {
    const int P = h->sp_nranks, K = 6 * h->k;
    // a rank whose owned range is not a whole number of supernodes shares its last supernode with a halo pose
    const bool can = h->spike_state == 0 && h->D == 6 && !h->has_bias && h->B <= VUS_SMALLB_MAX && h->ncomp <= 1 && 2 * K * P <= 384 &&
                     (h->n_owned % h->k == 0 || h->sp_rank == P - 1);
    double v[3] = {can ? 0.0 : 1.0, (double)h->k, (double)h->k * h->k};
    h->spCoef.alloc(std::max(3, 2 * K));
    rt::h2d(h->spCoef.p, v, sizeof v, st);
    comm_call(h, VUS_COMM_ALLREDUCE_SUM, h->spCoef.p, 3, st);
    rt::d2h(v, h->spCoef.p, sizeof v, st);
    rt::sync(st);
    const bool same_k = std::fabs(P * v[2] - v[1] * v[1]) < 0.5;
    h->spike_state = (v[0] > 0.0 || !same_k) ? -1 : 1;
    if (h->spike_state == 1) {
      const size_t n = (size_t)2 * K * P;
      h->spZ.alloc((size_t)2 * K * h->Ns_band * h->B); h->spG.alloc((size_t)P * 4 * K * K); h->spR.alloc(n * n); h->spRinv.alloc(n * n);
      h->spY.alloc(n);
    }
    if (h->prm.verbose) std::fprintf(stderr, "pose-range partition: band tied across %d ranks exactly: %s\n", P, h->spike_state == 1 ? "yes" : "no (block-Jacobi)");
  }
  return h->spike_state == 1;
}
void spike_setup(vus_handle* h, rt::stream_t st) {
  ClassGuard kc_guard(KC_BORDER);
  SpikeArgs a = spike_args(h);
  const int K = a.K;
  h->spZ.zero(st);
  L_elem<SpikeRhsBody>((long)2 * K * K, st, a);
  for (int c = 0; c < 2 * K; c += VUS_SMALLB_MAXV)
    bcr_solve(h, h->spZ.p + (long)c * a.Ls, a.Ls, std::min(VUS_SMALLB_MAXV, 2 * K - c), st);
  h->spG.zero(st);
  L_elem<SpikeGatherBody>((long)4 * K * K, st, a);
  comm_call(h, VUS_COMM_ALLREDUCE_SUM, h->spG.p, (long)a.P * 4 * K * K, st);
  L_coop<SpikeInvertBody>(1, 256, (size_t)(2 * K * a.P + 2) * sizeof(double), st, a);
}
void spike_apply(vus_handle* h, double* z, rt::stream_t st) {
  ClassGuard kc_guard(KC_BORDER);
  SpikeArgs a = spike_args(h);
  a.z = z;
  h->spY.zero(st);
  L_elem<SpikeYBody>(2 * a.K, st, a);
  comm_call(h, VUS_COMM_ALLREDUCE_SUM, h->spY.p, (long)2 * a.K * a.P, st);
  L_elem<SpikeCoefBody>(2 * a.K, st, a);
  L_elem<SpikeCorrBody>(h->n_owned * 6, st, a);
}

// z = P^-1 r  (z and r are full vectors of length L)
void precond_apply(vus_handle* h, double* z, const double* r, rt::stream_t st) {
  if (h->z0_valid) {                                   // r is gs: its band solve was done with the border columns
    h->z0_valid = false;
    rt::d2d(z, h->Z.p + 6 * h->Lc, h->Lc * sizeof(double), st);
    rt::d2d(z + h->Lc, r + h->Lc, (h->L - h->Lc) * sizeof(double), st);
  } else {
    rt::d2d(z, r, h->L * sizeof(double), st);
    bcr_solve(h, z, h->Lc, 1, st);
    if (h->spike_state == 1) spike_apply(h, z, st);
  }
  if (h->has_bias) {
    border_dot(h, z, h->Lc, 1, st);
    BorderSolveArgs b; b.SbInv = h->SbInv.p; b.rb = r + h->Lc; b.partials = h->bpart.p; b.grid = h->red_grid; b.xb = z + h->Lc;
    L_coop<BorderSolveBody>(1, 32, 8 * sizeof(double), st, b);
    VecArgs v; v.y = z; v.x = z; v.z = nullptr; v.scal = nullptr; v.slot = 0; v.n = h->Lc; v.Z = h->Z.p; v.xb = z + h->Lc; v.zstride = h->Lc;
    L_elem<SubZxbBody>(h->Lc, st, v);
  }
}

// y = A x
void apply_A(vus_handle* h, double* y, const double* x, rt::stream_t st) {
  ClassGuard kc_guard(KC_MATVEC);
  MatvecArgs a;
  a.SD = h->H.p + h->sd_off; a.SU = h->H.p + h->su_off; a.Ns = h->Ns; a.B = h->B; a.x = x; a.y = y;
  a.nv = 1; a.xstride = 0; a.ystride = 0; a.Nrows = h->Ns_band;
  a.rem_ptr = h->nrem ? h->rem_ptr.p : nullptr; a.rem_col = h->rem_col.p; a.rem_val = h->H.p + h->rem_off; a.nnodes = h->N; a.D = h->D;
  a.F = h->F.p; a.Hbb = h->Hbb.p; a.xb = x + h->Lc; a.yb = y + h->Lc; a.has_bias = h->has_bias;
  if (h->B <= VUS_SMALLB_MAX) {
    small_matvec(h, h->Ns_band * h->B, a, st);
  } else {
    const size_t smem = (size_t)blk_smem_doubles(h->B, 1, 256) * sizeof(double);
    L_coop<BandMatvecBody>((int)h->Ns_band, 256, smem, st, a);
  }
  // off-band blocks and border: only the rows of the nodes this rank owns (the others are zeroed by zero_halo right after)
  const long nrow_nodes = h->n_owned >= 0 ? std::min<long>(h->N, h->Ns_band * h->k) : h->N;
  if (h->nrem || h->has_bias) { ClassGuard kc_b(KC_BORDER); L_elem<RemBorderMatvecBody>(nrow_nodes * h->D, st, a); }
  if (h->nlong) {                                      // implicit Schur term of the long-track landmarks
    ClassGuard kc_s(KC_SCHUR);
    SchurArgs sa = schur_args(h, h->cur_lambda);
    sa.xin = x; sa.yout = y;
    L_elem<LongSchur1Body>(h->nlong, st, sa);
    L_elem<LongSchur2Body>(h->nposes_obs * 6, st, sa);
  }
  if (h->has_bias) {
    border_dot(h, x, h->Lc, 1, st);
    BorderRowArgs b; b.Hbb = h->Hbb.p; b.xb = x + h->Lc; b.partials = h->bpart.p; b.grid = h->red_grid; b.yb = y + h->Lc;
    { ClassGuard kc_b(KC_BORDER); L_coop<BorderRowBody>(1, 32, 8 * sizeof(double), st, b); }
  }
}

void axpy(vus_handle* h, double* y, const double* x, int slot, rt::stream_t st) {
  VecArgs v; v.y = y; v.x = x; v.z = nullptr; v.scal = h->scal.p; v.slot = slot; v.n = h->L; v.Z = nullptr; v.xb = nullptr; v.zstride = 0;
  L_elem<AxpyBody>(h->L, st, v);
}
void xpby(vus_handle* h, double* y, const double* x, int slot, rt::stream_t st) {
  VecArgs v; v.y = y; v.x = x; v.z = nullptr; v.scal = h->scal.p; v.slot = slot; v.n = h->L; v.Z = nullptr; v.xb = nullptr; v.zstride = 0;
  L_elem<XpbyBody>(h->L, st, v);
}

const int kPcgBlock = 8;     // most PCG iterations replayed per look of the host at the residual norm
// a PCG run that stopped short of pcg_rel_tol still counts as a solve when its TRUE relative residual is below this
const double kPcgAcceptRelRes = 1e-6;
struct SetScalarArgs { double* dst; double v; };
struct SetScalarBody {
  static VUS_DEV void run(const SetScalarArgs& A, long) { *A.dst = A.v; }
};
// y = x - y   (used for the true residual r = rhs - A x)
struct RsubBody {
  static VUS_DEV void run(const VecArgs& A, long i) { A.y[i] = A.x[i] - A.y[i]; }
};
struct AddBody {
  static VUS_DEV void run(const VecArgs& A, long i) { A.y[i] += A.x[i]; }
};

// PCG on the reduced system with residual replacement: the inner recursion solves A d = r for a correction,
// the outer loop recomputes the TRUE residual r = rhs - A x (the recursive residual drifts on these badly scaled
// systems: IMU information ~1e10 next to lambda ~1e-5).  Returns inner iterations; solution in h->x.
// one PCG iteration as a fixed launch sequence (every scalar it needs -- alpha, beta, the freeze once the residual met the
// tolerance -- lives on the device): replayed as a CUDA graph, several per look of the host at the residual norm
void pcg_iteration_launches(vus_handle* h, rt::stream_t st) {
  precond_apply(h, h->z.p, h->r.p, st);
  reduce(h, h->r.p, h->z.p, h->Lr, S_RZ, RED_RZ, st);   // beta = rz / rz_old (0 on the first iteration: rz_old = 0)
  xpby(h, h->p.p, h->z.p, S_BETA, st);                  // p = z + beta p
  halo(h, h->p.p, st);
  apply_A(h, h->Ap.p, h->p.p, st);
  zero_halo(h, h->Ap.p, st);
  reduce(h, h->p.p, h->Ap.p, h->Lr, S_PAP, RED_PAP, st);
  axpy(h, h->d.p, h->p.p, S_ALPHA, st);
  axpy(h, h->r.p, h->Ap.p, S_NEG_ALPHA, st);
  reduce(h, h->r.p, h->r.p, h->Lr, S_RR, RED_RR, st);
}
void set_scalar(vus_handle* h, int slot, double v, rt::stream_t st) {
  SetScalarArgs q; q.dst = h->scal.p + slot; q.v = v;
  L_elem<SetScalarBody>(1, st, q);
}

// PCG on the reduced system with residual replacement: the inner recursion solves A d = r for a correction,
// the outer loop recomputes the TRUE residual r = rhs - A x (the recursive residual drifts on these badly scaled
// systems: IMU information ~1e10 next to lambda ~1e-5).  Returns inner iterations; solution in h->x.
// The recursion is device resident: the host replays kPcgBlock iterations at a time and reads the residual norm once per
// block; a recursion that met its tolerance inside a block freezes itself (alpha = beta = 0), so the extra iterations of the
// block leave the iterate untouched.
int pcg(vus_handle* h, rt::stream_t st, bool* converged, bool* bad_out = nullptr, double accept_rel = kPcgAcceptRelRes) {
  const long L = h->L;
  VecArgs v; v.z = nullptr; v.scal = h->scal.p; v.slot = 0; v.n = L; v.Z = nullptr; v.xb = nullptr; v.zstride = 0;
  h->x.zero(st);
  zero_halo(h, h->gs.p, st);                           // halo rows belong to another rank's system
  rt::d2d(h->r.p, h->gs.p, L * sizeof(double), st);
  set_scalar(h, S_TOL2, -1.0, st);
  set_scalar(h, S_NAN, 0.0, st);
  set_scalar(h, S_ITS, 0.0, st);
  reduce(h, h->r.p, h->r.p, h->Lr, S_RR, RED_STORE, st);
  const double rr0 = read_scalar(h, S_RR, st);
  *converged = true;
  if (bad_out) *bad_out = false;
  h->last_rel_res = 0.0;
  if (!(rr0 == rr0)) { if (bad_out) *bad_out = true; *converged = false; h->last_rel_res = INFINITY; h->z0_valid = false; return 0; }
  if (!(rr0 > 0.0)) { h->z0_valid = false; return 0; }
  h->last_rel_res = INFINITY;
  const double tol2 = h->prm.pcg_rel_tol * h->prm.pcg_rel_tol * rr0;
  *converged = false;
  // host-side callbacks (partitioned graphs over gloo) cannot be captured: one iteration per look there
  const int block = (h->comm && !h->nccl && !h->comm_stream_ordered) ? 1 : kPcgBlock;
  int it = 0, live_its = 0;                            // launched / live (not frozen) iterations
  double rr_outer = rr0;
  for (int outer = 0; outer < 6 && !*converged; ++outer) {
    // ---- inner PCG: A d = r, d = 0
    h->d.zero(st);
    h->p.zero(st);
    set_scalar(h, S_RZ, 0.0, st);                        // first iteration: beta = 0, p = z
    set_scalar(h, S_RR, rr_outer, st);
    set_scalar(h, S_TOL2, tol2, st);
    int since_best = 0;
    double best = rr_outer;
    bool bad = false, stalled = true;                   // stalled: the recursion ended on a stall, not on its tolerance
    int looks = 0;
    const int live_before = live_its;
    while (it < h->prm.pcg_max_iterations) {
      // an exactly factored band converges in one to three iterations: look after every one of the first three, then let the
      // block grow (closure-dominated graphs run a hundred iterations per solve)
      const int want = looks < 3 ? 1 : std::min(block, 1 << std::min(looks - 2, 3));
      const int nb = std::min(want, h->prm.pcg_max_iterations - it);
      ++looks;
      for (int b = 0; b < nb; ++b) {
        // collectives inside, or the first preconditioner application that rode along the border solve: not the captured sequence
        if ((h->comm && !h->nccl) || h->z0_valid) pcg_iteration_launches(h, st);
        else run_graphed(h->g_pcg_iter, st, [&] { pcg_iteration_launches(h, st); });
      }
      it += nb;
      double sc[S_COUNT];
      rt::d2h(sc, h->scal.p, sizeof(sc), st);
      rt::sync(st);
      const double rr = sc[S_RR];
      const double two[2] = {rr, sc[S_NAN]};
      live_its = (int)sc[S_ITS];
      if (h->prm.verbose > 1) std::fprintf(stderr, "    pcg %d.%d rel_res %.3e\n", outer, it, std::sqrt(rr / rr0));
      if (!(rr == rr) || two[1] > 0.0) { bad = true; break; }
      if (rr <= tol2) { stalled = false; break; }
      if (rr < best) { best = rr; since_best = 0; }
      else if ((since_best += nb) >= 40) break;          // CG residuals are not monotone: only a long stall ends the recursion
    }
    set_scalar(h, S_TOL2, -1.0, st);
    if (bad) { if (bad_out) *bad_out = true; break; }
    // ---- x += d ; true residual
    v.y = h->x.p; v.x = h->d.p;
    L_elem<AddBody>(L, st, v);
    halo(h, h->x.p, st);
    apply_A(h, h->r.p, h->x.p, st);
    v.y = h->r.p; v.x = h->gs.p;
    L_elem<RsubBody>(L, st, v);
    zero_halo(h, h->r.p, st);
    reduce(h, h->r.p, h->r.p, h->Lr, S_RR, RED_STORE, st);
    const double rr_true = read_scalar(h, S_RR, st);
    if (h->prm.verbose > 1) std::fprintf(stderr, "    pcg outer %d true rel_res %.3e\n", outer, std::sqrt(rr_true / rr0));
    h->last_rel_res = std::sqrt(rr_true / rr0);
    if (rr_true <= tol2) { *converged = true; break; }
    if (!(rr_true < 0.25 * rr_outer) || it >= h->prm.pcg_max_iterations) break;   // no further progress possible
    // A recursion that met its tolerance within three iterations (exactly factored band: configs 2 and 3) has not drifted:
    // its true residual already sits at the attainable floor eps |A| |x| / |b| (1e-9 .. 1e-8 on these systems, IMU
    // information 1e10 next to lambda 1e-5), and another pass would spend a band solve and two operator products to
    // find the same floor again (measured: the third iteration never lowered the true residual at config 3).
    if (!stalled && live_its - live_before <= 3 && rr_true <= accept_rel * accept_rel * rr0) break;
    rr_outer = rr_true;
  }
  return live_its;
}

// per-component flags of the last factorization (batched mode); clears them.  Returns whether any was set.
bool read_fail_vec(vus_handle* h, rt::stream_t st, std::vector<int>& out) {
  out.assign(std::max(h->ncomp, 1), 0);
  rt::d2h(out.data(), h->fail.p, out.size() * sizeof(int), st);
  rt::sync(st);
  bool any = false;
  for (int v : out) any |= v != 0;
  if (any) h->fail.zero(st);
  return any;
}
int read_fail(vus_handle* h, rt::stream_t st) {
  int f = 0;
  if (h->ncomp > 1) {
    std::vector<int> v;
    f = read_fail_vec(h, st, v) ? 1 : 0;
  } else {
    rt::d2h(&f, h->fail.p, sizeof(int), st);
    rt::sync(st);
    if (f) h->fail.zero(st);
  }
  if (has_comm(h)) {                                   // every rank must take the same decision
    double v = f ? 1.0 : 0.0;
    rt::h2d(h->scal.p + S_COMM, &v, sizeof(double), st);
    comm_call(h, VUS_COMM_ALLREDUCE_SUM, h->scal.p + S_COMM, 1, st);
    rt::d2h(&v, h->scal.p + S_COMM, sizeof(double), st);
    rt::sync(st);
    f = v > 0.0;
  }
  return f;
}

// solve the damped system at the current linearization; delta in h->x (camera+bias) and h->xl
bool solve_damped(vus_handle* h, double lambda, rt::stream_t st, int* iters, bool timing) {
  double t0 = timing ? now_ms() : 0;
  form_system(h, lambda, st);
  if (timing) { rt::sync(st); double t1 = now_ms(); h->res.ms_schur += t1 - t0; t0 = t1; }
  precond_setup(h, st, h->n_owned < 0);
  if (timing) { rt::sync(st); double t1 = now_ms(); h->res.ms_factor += t1 - t0; t0 = t1; }
  if (read_fail(h, st)) { *iters = 0; return false; }
  bool conv = false, bad = false;
  *iters = pcg(h, st, &conv, &bad);
  h->res.worst_pcg_rel_residual = std::max(h->res.worst_pcg_rel_residual, bad ? (double)INFINITY : h->last_rel_res);
  // gtsam's elimination either solves the damped system or throws (-> "increase lambda"); an iterative solve that ended on a
  // NaN or stalled far from the solution is the same event, not a step: the try fails and lambda is raised.
  if (bad || (!conv && !(h->last_rel_res <= kPcgAcceptRelRes))) {
    h->res.pcg_not_converged++;
    if (timing) { rt::sync(st); double t1 = now_ms(); h->res.ms_pcg += t1 - t0; t0 = t1; }
    return false;
  }
  halo(h, h->x.p, st);                                 // the step of the halo poses comes from their owners
  if (h->nobs) {
    SchurArgs a = schur_args(h, lambda);
    L_elem<LmBacksubBody>(a.nl, st, a);
  }
  if (timing) { rt::sync(st); double t1 = now_ms(); h->res.ms_pcg += t1 - t0; t0 = t1; }
  return true;
}

template <int T>
void launch_linerr(vus_handle* h, rt::stream_t st) {
  FactorTable& F = h->ft[T];
  if (!F.n) return;
  LinErrArgs a;
  a.F.n = F.n; a.F.idx = F.idx.p; a.F.meas = nullptr; a.F.sinfo = nullptr;
  a.r = F.r.p; a.J = F.J.p;
  a.X.xc = h->x.p; a.X.xb = h->x.p + h->Lc; a.X.xl = h->xl.p; a.X.nl = h->nvar[3]; a.X.D = h->D;
  a.out = h->le_all.p + F.e_off; a.type = T;
  L_elem<LinErrBody<T>>(F.n, st, a);
}
void linear_error_launches(vus_handle* h, rt::stream_t st) {
  ClassGuard kc_guard(KC_LINERR);
  launch_linerr<VUS_F_PRIOR_POSE>(h, st); launch_linerr<VUS_F_PRIOR_VEL>(h, st); launch_linerr<VUS_F_BETWEEN>(h, st);
  launch_linerr<VUS_F_DVL>(h, st); launch_linerr<VUS_F_STEREO>(h, st); launch_linerr<VUS_F_IMU>(h, st);
}
double linear_error(vus_handle* h, rt::stream_t st) {
  linear_error_launches(h, st);
  ClassGuard kc_guard(KC_LINERR);
  reduce(h, h->le_all.p, h->n_owned >= 0 ? h->e_mask.p : nullptr, h->nfactors, S_TMP, RED_STORE, st);
  return read_scalar(h, S_TMP, st);
}

void retract(vus_handle* h, rt::stream_t st) {
  ClassGuard kc_guard(KC_RETRACT);
  RetractArgs a;
  const int c = h->cur, t = 1 - h->cur;
  a.pose = h->val[c][0].p; a.pose_out = h->val[t][0].p; a.nx = h->nvar[0];
  a.vel = h->val[c][1].p; a.vel_out = h->val[t][1].p; a.nv = h->nvar[1];
  a.bias = h->val[c][2].p; a.bias_out = h->val[t][2].p; a.nb = h->nvar[2];
  a.lm = h->val[c][3].p; a.lm_out = h->val[t][3].p; a.nl = h->nvar[3];
  a.xc = h->x.p; a.xb = h->x.p + h->Lc; a.xl = h->xl.p; a.D = h->D;
  L_elem<RetractBody>(a.nx + a.nv + a.nl + a.nb, st, a);
}

// ------------------------------------------------------------------ kernel 4 driver: the LM loop
int optimize(vus_handle* h, rt::stream_t st) {
  vus_lm_result& R = h->res;
  R = vus_lm_result();
  h->trace.clear();
  const vus_lm_params& P = h->prm;
  const long launches0 = g_launches;
  g_prof.reset();
  g_prof.on = P.profile_kernels != 0;
  const double t_begin = now_ms();
  double lambda = P.lambda_initial;
  double err = graph_error(h, h->cur, st);
  R.initial_error = err;
  int iterations = 0;
  if (!(err <= P.error_tol) && P.max_iterations > 0 && std::isfinite(err)) {
    while (true) {
      const double cur_err = err;
      double t0 = now_ms();
      run_factors(h, h->cur, true, st);
      rt::sync(st);
      double t1 = now_ms();
      R.ms_linearize += t1 - t0;
      assemble_base(h, st);
      rt::sync(st);
      R.ms_assemble += now_ms() - t1;
      R.linearizations++;
      R.factors_linearized += h->nfactors;
      while (true) {                                   // tryLambda
        int its = 0;
        const bool solved = solve_damped(h, lambda, st, &its, true);
        R.pcg_iterations += its;
        R.inner_iterations++;
        if (!solved) R.solve_failures++;
        bool success = false, stop = false;
        double new_err = INFINITY;
        double t2 = now_ms();
        if (solved) {
          const double old_lin = err;
          const double new_lin = linear_error(h, st);
          const double lin_change = old_lin - new_lin;
          if (lin_change >= 0) {
            retract(h, st);
            new_err = graph_error(h, 1 - h->cur, st);
            const double cost_change = err - new_err;
            if (lin_change > VUS_EPS * old_lin) {
              const double fidelity = cost_change / lin_change;
              success = fidelity > P.min_model_fidelity;
            }
            if (std::fabs(cost_change) < P.relative_error_tol * err) stop = true;
          }
        }
        R.ms_update += now_ms() - t2;
        if (P.verbose) std::fprintf(stderr, "  try lam=%.3e solved=%d pcg=%d new_err=%.9e success=%d\n", lambda, (int)solved, its, new_err, (int)success);
        { vus_lm_try tr; tr.lambda = lambda; tr.new_error = new_err; tr.solved = solved; tr.accepted = success; tr.pcg_iterations = its; tr.reserved = 0; h->trace.push_back(tr); }
        if (success) {
          h->cur = 1 - h->cur;
          err = new_err;
          lambda = std::max(P.lambda_lower_bound, lambda / P.lambda_factor);
          iterations++;
          break;
        } else if (!stop) {
          lambda *= P.lambda_factor;
          if (lambda >= P.lambda_upper_bound) break;
        } else {
          break;
        }
      }
      if (P.verbose) std::fprintf(stderr, "iter %d err=%.12e lam=%.3e\n", iterations, err, lambda);
      if (iterations >= P.max_iterations || !std::isfinite(cur_err)) break;
      if (err <= P.error_tol) break;
      const double absdec = cur_err - err, reldec = absdec / cur_err;
      if ((P.relative_error_tol != 0.0 && reldec <= P.relative_error_tol) || absdec <= P.absolute_error_tol) break;
    }
  }
  rt::sync(st);
  R.iterations = iterations;
  R.final_error = err;
  R.final_lambda = lambda;
  R.ms_total = now_ms() - t_begin;
  R.kernel_launches = g_launches - launches0;
  g_prof.collect();
  g_prof.on = false;
  for (int i = 0; i < 16; ++i) { R.ms_class[i] = g_prof.ms[i]; R.launches_class[i] = g_prof.count[i]; }
  return VUS_OK;
}

// ------------------------------------------------------------------ batched mode (batch.cuh): many independent trajectories
BCtx bctx(vus_handle* h) {
  BCtx c; c.node_comp = h->node_comp.p; c.seg = h->seg.p; c.ncomp = h->ncomp; c.D = h->D; c.has_bias = h->has_bias; c.Lc = h->Lc;
  return c;
}
void form_system_b(vus_handle* h, rt::stream_t st) {       // damping from h->lam_c
  ClassGuard kc_guard(KC_SCHUR);
  const long nb = h->nvar[2];
  rt::d2d(h->H.p, h->H0.p, h->hlen * sizeof(double), st);
  rt::d2d(h->gs.p, h->g0.p, h->Lc * sizeof(double), st);
  if (nb) {
    rt::d2d(h->Hbb.p, h->Hbb0.p, 36 * nb * sizeof(double), st);
    rt::d2d(h->gs.p + h->Lc, h->gb.p, 6 * nb * sizeof(double), st);
  }
  BDampArgs d; d.C = bctx(h); d.SD = h->H.p + h->sd_off; d.Hbb = h->Hbb.p; d.nreal = h->N * h->D; d.B = h->B; d.lam = h->lam_c.p;
  d.ld = bcr_ld(h->B); d.bs = bcr_bbp(h->B);
  L_elem<BDampBody>(h->Lc + 6 * nb, st, d);
}
void bdot(vus_handle* h, const double* a, const double* b, int slot, int op, rt::stream_t st) {
  BDotArgs r; r.C = bctx(h); r.a = a; r.b = b; r.scal = h->scal_b.p; r.slot = slot; r.op = op; r.tol = h->prm.pcg_rel_tol;
  L_coop<BDotBody>(h->ncomp, 256, 256 * sizeof(double), st, r);
}
void bborder_dot(vus_handle* h, const double* Y, long ystride, int nv, double* out, rt::stream_t st) {
  ClassGuard kc_guard(KC_BORDER);
  BBorderDotArgs a; a.C = bctx(h); a.F = h->F.p; a.Y = Y; a.ystride = ystride; a.nv = nv; a.out = out;
  L_coop<BBorderDotBody>(h->ncomp, 256, 6 * 256 * sizeof(double), st, a);
}
void precond_setup_b(vus_handle* h, rt::stream_t st) {
  bcr_factor(h, st);
  h->z0_valid = false;
  if (!h->has_bias) return;
  BorderColsArgs c; c.F = h->F.p; c.Z = h->Z.p; c.len = h->Lc; c.zstride = h->Lc; c.R = h->Zr.p;
  L_elem<BorderColsBody>(h->Lc * 6, st, c);
  rt::d2d(h->Z.p + 6 * h->Lc, h->gs.p, h->Lc * sizeof(double), st);      // the right-hand side rides along (precond_setup)
  bcr_solve(h, h->Z.p, h->Lc, 7, st);
  h->z0_valid = true;
  apply_band(h, h->Zr.p, h->Lc, h->Z.p, h->Lc, 6, st);     // second-order correction of every component's complement
  L_elem<BorderResidBody>(h->Lc * 6, st, c);
  bborder_dot(h, h->Z.p, h->Lc, 6, h->cdots.p, st);
  ClassGuard kc_b(KC_BORDER);
  BColDotArgs d; d.C = bctx(h); d.Z = h->Z.p; d.R = h->Zr.p; d.stride = h->Lc; d.out = h->cdots2.p;
  L_coop<BColDotBody>(h->ncomp, 128, 6 * 128 * sizeof(double), st, d);
  BBorderSchurArgs sa; sa.ncomp = h->ncomp; sa.Hbb = h->Hbb.p; sa.ftz = h->cdots.p; sa.ztr = h->cdots2.p; sa.SbInv = h->SbInv.p; sa.fail = h->fail.p;
  L_elem<BBorderSchurBody>(h->ncomp, st, sa);
}
void precond_apply_b(vus_handle* h, double* z, const double* r, rt::stream_t st) {
  if (h->z0_valid) {
    h->z0_valid = false;
    rt::d2d(z, h->Z.p + 6 * h->Lc, h->Lc * sizeof(double), st);
    rt::d2d(z + h->Lc, r + h->Lc, (h->L - h->Lc) * sizeof(double), st);
  } else {
    rt::d2d(z, r, h->L * sizeof(double), st);
    bcr_solve(h, z, h->Lc, 1, st);
  }
  if (!h->has_bias) return;
  bborder_dot(h, z, h->Lc, 1, h->cdots.p, st);
  ClassGuard kc_b(KC_BORDER);
  BBorderSmallArgs b; b.M = h->SbInv.p; b.rb = r + h->Lc; b.dots = h->cdots.p; b.out = z + h->Lc;
  L_elem<BBorderSolveBody>(6L * h->ncomp, st, b);
  BVecArgs v; v.C = bctx(h); v.y = z; v.x = z; v.scal = nullptr; v.slot = 0; v.Z = h->Z.p; v.xb = z + h->Lc; v.zstride = h->Lc;
  L_elem<BSubZxbBody>(h->Lc, st, v);
}
void apply_A_b(vus_handle* h, double* y, const double* x, rt::stream_t st) {
  apply_band(h, y, 0, x, 0, 1, st);
  ClassGuard kc_b(KC_BORDER);
  if (h->nrem || h->has_bias) {
    BRemBorderArgs a; a.C = bctx(h); a.rem_ptr = h->nrem ? h->rem_ptr.p : nullptr; a.rem_col = h->rem_col.p; a.rem_val = h->H.p + h->rem_off;
    a.F = h->F.p; a.x = x; a.xb = x + h->Lc; a.y = y;
    L_elem<BRemBorderBody>(h->N * h->D, st, a);
  }
  if (h->has_bias) {
    bborder_dot(h, x, h->Lc, 1, h->cdots.p, st);
    BBorderSmallArgs b; b.M = h->Hbb.p; b.rb = x + h->Lc; b.dots = h->cdots.p; b.out = y + h->Lc;
    L_elem<BBorderRowBody>(6L * h->ncomp, st, b);
  }
}
void baxpy(vus_handle* h, double* y, const double* x, int slot, rt::stream_t st) {
  BVecArgs v; v.C = bctx(h); v.y = y; v.x = x; v.scal = h->scal_b.p; v.slot = slot; v.Z = nullptr; v.xb = nullptr; v.zstride = 0;
  L_elem<BAxpyBody>(h->L, st, v);
}
void bxpby(vus_handle* h, double* y, const double* x, int slot, rt::stream_t st) {
  BVecArgs v; v.C = bctx(h); v.y = y; v.x = x; v.scal = h->scal_b.p; v.slot = slot; v.Z = nullptr; v.xb = nullptr; v.zstride = 0;
  L_elem<BXpbyBody>(h->L, st, v);
}
BWbCtx wbctx(vus_handle* h) {
  BWbCtx w; w.node = h->wb_node.p; w.blk = h->wb_blk.p; w.lmax = h->wb_lmax; w.R = h->wb_R;
  return w;
}
// capacitance set-up for the current damped system (after precond_setup_b): Y = P^-1 U six columns at a time, CapInv per component
void wb_setup(vus_handle* h, rt::stream_t st) {
  const int R = h->wb_R;
  const long L = h->L;
  h->wbY.zero(st);
  for (int j0 = 0; j0 < R; j0 += 6) {
    double* Yb = h->wbY.p + (long)j0 * L;
    BWbUnitArgs u; u.C = bctx(h); u.W = wbctx(h); u.Y = Yb; u.L = L; u.j0 = j0;
    L_elem<BWbUnitBody>(6L * h->ncomp, st, u);
    bcr_solve(h, Yb, L, 6, st);
    if (h->has_bias) {
      bborder_dot(h, Yb, L, 6, h->cdots.p, st);
      ClassGuard kc_b(KC_BORDER);
      BBorderMultiArgs b; b.C = bctx(h); b.SbInv = h->SbInv.p; b.dots = h->cdots.p; b.Y = Yb; b.L = L; b.nv = 6; b.Z = h->Z.p; b.zstride = h->Lc;
      L_elem<BBorderSolveMultiBody>(36L * h->ncomp, st, b);
      L_elem<BSubZxbMultiBody>(6 * h->Lc, st, b);
    }
  }
  ClassGuard kc_b(KC_BORDER);
  BWbCapArgs c; c.C = bctx(h); c.W = wbctx(h); c.Y = h->wbY.p; c.L = L; c.Hval = h->H.p; c.CapInv = h->wbCapInv.p; c.fail = h->fail.p;
  L_coop<BWbCapBody>(h->ncomp, 256, ((size_t)R * 2 * R + R + 2) * sizeof(double), st, c);
}
// z -= Y (I + S K)^-1 S U^T z   on top of z = P^-1 r
void wb_apply(vus_handle* h, double* z, rt::stream_t st) {
  ClassGuard kc_b(KC_BORDER);
  BWbSmallArgs a; a.C = bctx(h); a.W = wbctx(h); a.z = z; a.Hval = h->H.p; a.CapInv = h->wbCapInv.p; a.wout = h->wbW.p;
  L_coop<BWbSmallBody>(h->ncomp, 128, (size_t)2 * h->wb_R * sizeof(double), st, a);
  BWbSubArgs b; b.C = bctx(h); b.Y = h->wbY.p; b.L = h->L; b.w = h->wbW.p; b.R = h->wb_R; b.z = z;
  L_elem<BWbSubBody>(h->L, st, b);
}
// the full preconditioner of batched mode: band + border exactly, loop closures by capacitance when set up
void precond_full_b(vus_handle* h, double* z, const double* r, rt::stream_t st) {
  precond_apply_b(h, z, r, st);
  if (h->wb_R) wb_apply(h, z, st);
}
// worst ratio rr_c / tol2_c over the components (<= 1: every component met its tolerance); NaN-safe
double worst_ratio(vus_handle* h, std::vector<double>& sb, rt::stream_t st, bool* bad) {
  rt::d2h(sb.data(), h->scal_b.p, sb.size() * sizeof(double), st);
  rt::sync(st);
  double w = 0.0;
  for (int c = 0; c < h->ncomp; ++c) {
    const double rr = sb[(size_t)c * SB_STRIDE + S_RR], t2 = sb[(size_t)c * SB_STRIDE + SB_TOL2];
    if (!(rr == rr)) { *bad = true; return INFINITY; }
    if (rr <= t2) continue;
    w = std::max(w, t2 > 0.0 ? rr / t2 : INFINITY);
  }
  return w;
}
// PCG with per-component scalars: the block-diagonal system is solved as ncomp independent CG recursions advanced in
// lock step (one set of kernels for all of them); a component whose residual met the tolerance stops moving.
int pcg_b(vus_handle* h, rt::stream_t st, bool* converged, bool* bad_out = nullptr) {
  const long L = h->L;
  VecArgs v; v.z = nullptr; v.scal = nullptr; v.slot = 0; v.n = L; v.Z = nullptr; v.xb = nullptr; v.zstride = 0;
  std::vector<double> sb((size_t)h->ncomp * SB_STRIDE);
  h->x.zero(st);
  rt::d2d(h->r.p, h->gs.p, L * sizeof(double), st);
  bdot(h, h->r.p, h->r.p, S_RR, BOP_RR0, st);
  bool bad = false;
  double worst = worst_ratio(h, sb, st, &bad);
  *converged = !bad && worst <= 1.0;
  if (bad_out) *bad_out = bad;
  if (*converged || bad) { h->z0_valid = false; return 0; }
  int it = 0;
  double worst_outer = worst;
  for (int outer = 0; outer < 6 && !*converged; ++outer) {
    h->d.zero(st);
    precond_full_b(h, h->z.p, h->r.p, st);
    rt::d2d(h->p.p, h->z.p, L * sizeof(double), st);
    bdot(h, h->r.p, h->z.p, S_RZ, BOP_RZ0, st);
    int since_best = 0;
    double best = worst_outer;
    const int it_before = it;
    bool stalled = true;                                 // the recursion ended on a stall, not on its tolerance
    while (it < h->prm.pcg_max_iterations) {
      apply_A_b(h, h->Ap.p, h->p.p, st);
      bdot(h, h->p.p, h->Ap.p, S_PAP, BOP_PAP, st);
      baxpy(h, h->d.p, h->p.p, S_ALPHA, st);
      baxpy(h, h->r.p, h->Ap.p, S_NEG_ALPHA, st);
      ++it;
      bdot(h, h->r.p, h->r.p, S_RR, BOP_STORE, st);
      worst = worst_ratio(h, sb, st, &bad);
      if (h->prm.verbose > 1) std::fprintf(stderr, "    pcg_b %d.%d worst rr/tol2 %.3e\n", outer, it, worst);
      if (bad || worst <= 1.0) { stalled = false; break; }
      if (worst < best) { best = worst; since_best = 0; }
      else if (++since_best >= 40) break;
      precond_full_b(h, h->z.p, h->r.p, st);
      bdot(h, h->r.p, h->z.p, S_RZ, BOP_RZ, st);
      bxpby(h, h->p.p, h->z.p, S_BETA, st);
    }
    if (bad) { if (bad_out) *bad_out = true; break; }
    v.y = h->x.p; v.x = h->d.p;
    L_elem<AddBody>(L, st, v);
    apply_A_b(h, h->r.p, h->x.p, st);
    v.y = h->r.p; v.x = h->gs.p;
    L_elem<RsubBody>(L, st, v);
    bdot(h, h->r.p, h->r.p, S_RR, BOP_STORE, st);
    const double worst_true = worst_ratio(h, sb, st, &bad);
    if (h->prm.verbose > 1) std::fprintf(stderr, "    pcg_b outer %d true worst rr/tol2 %.3e\n", outer, worst_true);
    if (bad) { if (bad_out) *bad_out = true; break; }
    if (worst_true <= 1.0) { *converged = true; break; }
    if (!(worst_true < 0.25 * worst_outer) || it >= h->prm.pcg_max_iterations) break;
    {                                                    // as in pcg(): a recursion of at most three iterations sits at its floor
      const double a = kPcgAcceptRelRes / h->prm.pcg_rel_tol;
      if (!stalled && it - it_before <= 3 && worst_true <= a * a) break;
    }
    worst_outer = worst_true;
  }
  return it;
}
void comp_sums(vus_handle* h, const double* e, std::vector<double>& out, rt::stream_t st) {
  BErrSumArgs a; a.e = e; a.ptr = h->cf_ptr.p; a.list = h->cf_list.p; a.out = h->err_c.p;
  L_coop<BErrSumBody>(h->ncomp, 128, 128 * sizeof(double), st, a);
  out.resize(h->ncomp);
  rt::d2h(out.data(), h->err_c.p, h->ncomp * sizeof(double), st);
  rt::sync(st);
}
// gtsam's LM loop run for every component at once: each round linearizes at the current values, every
// active component tries its own lambda, and the accept / reject / stop decisions of optimize() are taken per component.
// A component whose try fails is re-solved next round from the same values (hence the same linearization) with 10 x lambda.
int optimize_batched(vus_handle* h, rt::stream_t st) {
  vus_lm_result& R = h->res;
  R = vus_lm_result();
  const vus_lm_params& P = h->prm;
  const int nc = h->ncomp;
  const long launches0 = g_launches;
  g_prof.reset();
  g_prof.on = P.profile_kernels != 0;
  const double t_begin = now_ms();
  std::vector<double> lam(nc, P.lambda_initial), err, cur_err, new_err, new_lin, accept(nc, 0.0);
  std::vector<char> active(nc, 0);
  h->comp_res.assign(nc, vus_component_result());
  run_factors(h, h->cur, false, st);
  comp_sums(h, h->e_all.p, err, st);
  int n_active = 0;
  for (int c = 0; c < nc; ++c) {
    h->comp_res[c].initial_error = err[c];
    R.initial_error += err[c];
    active[c] = !(err[c] <= P.error_tol) && P.max_iterations > 0 && std::isfinite(err[c]);
    n_active += active[c];
  }
  cur_err = err;
  while (n_active > 0) {
    run_factors(h, h->cur, true, st);
    assemble_base(h, st);
    R.linearizations++;
    R.factors_linearized += h->nfactors;
    rt::h2d(h->lam_c.p, lam.data(), nc * sizeof(double), st);
    form_system_b(h, st);
    precond_setup_b(h, st);
    if (h->wb_R) wb_setup(h, st);
    // a non-positive pivot fails the try of THE COMPONENT it belongs to (as one optimizer per trajectory would see it): its
    // right-hand side is zeroed so its PCG recursion never starts, every other component is solved as usual
    std::vector<int> failc;
    const bool any_fail = read_fail_vec(h, st, failc);
    if (any_fail) {
      rt::h2d(h->fail_mask.p, failc.data(), nc * sizeof(int), st);
      BZeroCompArgs z; z.C = bctx(h); z.mask = h->fail_mask.p; z.v = h->gs.p;
      L_elem<BZeroCompBody>(h->L, st, z);
      z.v = h->Z.p + 6 * h->Lc;                          // the right-hand side that rode along the border solve
      if (h->z0_valid) L_elem<BZeroCompBody>(h->Lc, st, z);
      for (int c = 0; c < nc; ++c) if (failc[c] && active[c]) R.solve_failures++;
    }
    bool pcg_bad = false;
    {
      bool conv = false;
      R.pcg_iterations += pcg_b(h, st, &conv, &pcg_bad);
      if (pcg_bad) { R.solve_failures++; R.pcg_not_converged++; }
    }
    if (!pcg_bad) {
      linear_error_launches(h, st);
      comp_sums(h, h->le_all.p, new_lin, st);
      retract(h, st);
      run_factors(h, 1 - h->cur, false, st);
      comp_sums(h, h->e_all.p, new_err, st);
    }
    R.inner_iterations++;
    for (int c = 0; c < nc; ++c) {
      accept[c] = 0.0;
      if (!active[c]) continue;
      vus_component_result& cr = h->comp_res[c];
      cr.inner_iterations++;
      bool success = false, stop = false;
      double ne = INFINITY;
      const bool solved = !pcg_bad && !failc[c];
      if (solved) {
        const double old_lin = err[c];
        const double lin_change = old_lin - new_lin[c];
        if (lin_change >= 0) {
          ne = new_err[c];
          const double cost_change = err[c] - ne;
          if (lin_change > VUS_EPS * old_lin) success = cost_change / lin_change > P.min_model_fidelity;
          if (std::fabs(cost_change) < P.relative_error_tol * err[c]) stop = true;
        }
      }
      if (P.verbose > 1) std::fprintf(stderr, "  comp %d try lam=%.3e new_err=%.9e success=%d\n", c, lam[c], ne, (int)success);
      bool outer_done = false;                         // the tryLambda loop of this component ended
      if (success) {
        accept[c] = 1.0;
        err[c] = ne;
        lam[c] = std::max(P.lambda_lower_bound, lam[c] / P.lambda_factor);
        cr.iterations++;
        outer_done = true;
      } else if (!stop) {
        lam[c] *= P.lambda_factor;
        if (lam[c] >= P.lambda_upper_bound) outer_done = true;
      } else {
        outer_done = true;
      }
      if (outer_done) {                                // NonlinearOptimizer::defaultOptimize's convergence test
        bool done = cr.iterations >= P.max_iterations || !std::isfinite(cur_err[c]) || err[c] <= P.error_tol;
        const double absdec = cur_err[c] - err[c], reldec = absdec / cur_err[c];
        if ((P.relative_error_tol != 0.0 && reldec <= P.relative_error_tol) || absdec <= P.absolute_error_tol) done = true;
        cur_err[c] = err[c];
        if (done) { active[c] = 0; --n_active; }
      }
    }
    if (!pcg_bad) {
      rt::h2d(h->acc_c.p, accept.data(), nc * sizeof(double), st);
      BCommitArgs a; a.node_comp = h->node_comp.p; a.accept = h->acc_c.p; a.nx = h->nvar[0]; a.nv = h->nvar[1]; a.nb = h->nvar[2];
      const int cu = h->cur, tr = 1 - h->cur;
      a.pose = h->val[cu][0].p; a.pose_t = h->val[tr][0].p; a.vel = h->val[cu][1].p; a.vel_t = h->val[tr][1].p;
      a.bias = h->val[cu][2].p; a.bias_t = h->val[tr][2].p;
      ClassGuard kc_guard(KC_RETRACT);
      L_elem<BCommitBody>(a.nx + a.nv + a.nb, st, a);
      rt::sync(st);                                    // accept[] is reused by the next round
    }
    if (P.verbose) std::fprintf(stderr, "round %d: %d components active\n", R.inner_iterations, n_active);
  }
  rt::sync(st);
  int max_it = 0;
  for (int c = 0; c < nc; ++c) {
    h->comp_res[c].final_error = err[c];
    h->comp_res[c].final_lambda = lam[c];
    R.final_error += err[c];
    max_it = std::max(max_it, (int)h->comp_res[c].iterations);
  }
  R.iterations = max_it;
  R.final_lambda = lam[0];
  R.ms_total = now_ms() - t_begin;
  R.kernel_launches = g_launches - launches0;
  g_prof.collect();
  g_prof.on = false;
  for (int i = 0; i < 16; ++i) { R.ms_class[i] = g_prof.ms[i]; R.launches_class[i] = g_prof.count[i]; }
  return VUS_OK;
}

// ------------------------------------------------------------------ gtsam::Marginals
// right-hand side = one unit vector of the FULL system (camera dof `pos`, or column `col` of landmark l, whose reduced
// right-hand side is -E_o C_l^-1 e_col on the poses of its track).  One thread: a track is a handful of observations.
struct UnitRhsArgs { double* gs; double* gl; const double* E; const double* Cinv; const int* idx; const int* lm_ptr; long nl, l, pos; int col, D; };
struct UnitRhsBody {
  static VUS_DEV void run(const UnitRhsArgs& A, long) {
    if (A.pos >= 0) { A.gs[A.pos] = 1.0; return; }
    A.gl[A.col * A.nl + A.l] = 1.0;
    for (int q = A.lm_ptr[A.l]; q < A.lm_ptr[A.l + 1]; ++q) {
      const long node = A.idx[q];
      for (int a = 0; a < 6; ++a) {
        double s = 0.0;
        for (int c = 0; c < 3; ++c) s += A.E[(long)q * 18 + a * 3 + c] * A.Cinv[(c * 3 + A.col) * A.nl + A.l];
        A.gs[node * A.D + a] -= s;
      }
    }
  }
};

int marginal_covariance(vus_handle* h, rt::stream_t st, long nq, const int32_t* kinds, const int32_t* idx, double* out) {
  const int kTan[4] = {6, 3, 6, 3};
  std::vector<long> off(nq + 1, 0);
  bool any_lm = false;
  for (long q = 0; q < nq; ++q) {
    const int kd = kinds[q];
    if (kd < 0 || kd > 3 || idx[q] < 0 || idx[q] >= h->nvar[kd]) return fail(h, VUS_ERR_INVALID, "vus_marginal_covariance: query out of range");
    if (kd == VUS_VAR_BIAS && !h->has_bias) return fail(h, VUS_ERR_INVALID, "vus_marginal_covariance: the graph has no factor on the bias");
    if (kd == VUS_VAR_VEL && h->D < 9) return fail(h, VUS_ERR_INVALID, "vus_marginal_covariance: the graph has no factor on velocities");
    any_lm |= kd == VUS_VAR_LM;
    off[q + 1] = off[q] + kTan[kd];
  }
  const long M = off[nq];
  const long nl = h->nvar[3];
  // undamped system at the current values, factored once
  run_factors(h, h->cur, true, st);
  assemble_base(h, st);
  form_system(h, 0.0, st);
  precond_setup(h, st);
  if (read_fail(h, st)) return fail(h, VUS_ERR_STATE, "vus_marginal_covariance: the undamped system is not positive definite (indeterminate linear system)");
  std::vector<double> cov((size_t)M * M, 0.0);
  double buf[6];
  for (long q = 0; q < nq; ++q) {
    const int kd = kinds[q];
    for (int j = 0; j < kTan[kd]; ++j) {
      h->gs.zero(st);
      if (nl) h->gl.zero(st);
      UnitRhsArgs u;
      u.gs = h->gs.p; u.gl = h->gl.p; u.E = h->E.p; u.Cinv = h->Cinv.p; u.idx = h->ft[VUS_F_STEREO].idx.p; u.lm_ptr = h->lm_ptr.p;
      u.nl = nl; u.l = idx[q]; u.col = j; u.D = h->D; u.pos = -1;
      if (kd == VUS_VAR_POSE) u.pos = (long)idx[q] * h->D + j;
      else if (kd == VUS_VAR_VEL) u.pos = (long)idx[q] * h->D + 6 + j;
      else if (kd == VUS_VAR_BIAS) u.pos = h->Lc + j;
      L_elem<UnitRhsBody>(1, st, u);
      bool conv = false;
      pcg(h, st, &conv, nullptr, 1e-7);
      // The undamped system can be far harder than the damped ones of the LM loop (no lambda, unit right-hand sides that excite
      // the weakest directions; graphs whose tracks exceed the band leave most of the stereo information to PCG): refuse to
      // hand back a covariance column whose solve did not reach a small true residual.
      if (!conv && !(h->last_rel_res <= 1e-7)) {
        char msg[200];
        std::snprintf(msg, sizeof msg, "vus_marginal_covariance: the undamped solve of column %d of query %ld did not converge "
                      "(true relative residual %.2e)", j, q, h->last_rel_res);
        return fail(h, VUS_ERR_STATE, msg);
      }
      if (any_lm) {
        SchurArgs a = schur_args(h, 0.0);
        L_elem<LmBacksubBody>(a.nl, st, a);
      }
      const long colj = off[q] + j;
      for (long p = 0; p < nq; ++p) {
        const int kp = kinds[p];
        if (kp == VUS_VAR_LM) {
          for (int c = 0; c < 3; ++c) rt::d2h(buf + c, h->xl.p + c * nl + idx[p], sizeof(double), st);
        } else {
          const long pos = kp == VUS_VAR_BIAS ? h->Lc : (long)idx[p] * h->D + (kp == VUS_VAR_VEL ? 6 : 0);
          rt::d2h(buf, h->x.p + pos, kTan[kp] * sizeof(double), st);
        }
        rt::sync(st);
        for (int c = 0; c < kTan[kp]; ++c) cov[(size_t)(off[p] + c) * M + colj] = buf[c];
      }
    }
  }
  for (long a = 0; a < M; ++a)
    for (long b = 0; b < M; ++b) out[a * M + b] = 0.5 * (cov[(size_t)a * M + b] + cov[(size_t)b * M + a]);
  return VUS_OK;
}

}  // namespace

// =====================================================================================
// C-ABI
// =====================================================================================
// Every entry point that touches the device runs with the handle's device current on the CALLING thread (a handle may be
// created on one thread and driven from another, and one process may hold handles on several GPUs) and restores the
// caller's device on the way out.
struct DeviceGuard {
  int prev = -1, dev;
  rt::stream_t prev_stream;
  int prev_tl;
  explicit DeviceGuard(const vus_handle* h) : dev(h ? h->device : 0), prev_stream(rt::tl_stream()), prev_tl(rt::tl_device()) {
#ifndef VUS_EMU
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
#endif
    rt::tl_device() = dev;
    rt::tl_stream() = h ? h->own_stream : 0;
  }
  ~DeviceGuard() {
#ifndef VUS_EMU
    if (prev >= 0 && prev != dev) cudaSetDevice(prev);
#endif
    rt::tl_device() = prev_tl;
    rt::tl_stream() = prev_stream;
  }
};
// the stream an entry point works on (the caller's, or the handle's own); temporaries are freed in its order
inline rt::stream_t use_stream(vus_handle* h, void* stream) {
  rt::stream_t st = stream ? (rt::stream_t)stream : h->own_stream;
  rt::tl_stream() = st;
  return st;
}
#define VUS_TRY(h) try { DeviceGuard vus_device_guard_(h);
#define VUS_CATCH(h)                                                         \
  } catch (const std::exception& e) { return fail(h, VUS_ERR_CUDA, e.what()); } \
  catch (...) { return fail(h, VUS_ERR_CUDA, "unknown error"); }

extern "C" {

void vus_default_lm_params(vus_lm_params* p) {
  p->max_iterations = 100; p->relative_error_tol = 1e-5; p->absolute_error_tol = 1e-5; p->error_tol = 0.0;
  p->lambda_initial = 1e-5; p->lambda_factor = 10.0; p->lambda_upper_bound = 1e5; p->lambda_lower_bound = 0.0;
  p->min_model_fidelity = 1e-3; p->pcg_max_iterations = 500; p->pcg_rel_tol = 1e-12; p->max_supernode = 0; p->verbose = 0; p->profile_kernels = 0;
  p->band_chunks = 0;
}

int vus_create(int device, vus_handle** out) {
  if (!out) return VUS_ERR_INVALID;
  *out = nullptr;
#ifndef VUS_EMU
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return VUS_ERR_CUDA;
  int prev_device = -1;
  cudaGetDevice(&prev_device);
  if (cudaSetDevice(device) != cudaSuccess) return VUS_ERR_CUDA;
#endif
  rt::pool_setup(device);
  vus_handle* h = new vus_handle();
  h->device = device;
#ifndef VUS_EMU
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) h->own_stream = 0;
  if (prev_device >= 0 && prev_device != device) cudaSetDevice(prev_device);      // the caller's current device is left as it was
#endif
  vus_default_lm_params(&h->prm);
  *out = h;
  return VUS_OK;
}

void vus_destroy(vus_handle* h) {
  if (!h) return;
  DeviceGuard guard(h);
  rt::stream_t st = h->own_stream;
  try { rt::sync(st); } catch (...) {}
#ifndef VUS_EMU
  if (h->nccl) { h->drop_graphs(); nccl_api().CommDestroy(h->nccl); h->nccl = nullptr; }
#endif
  delete h;                                  // device buffers are released in the order of `st` (rt::dfree)
#ifndef VUS_EMU
  if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
#endif
}

const char* vus_last_error(const vus_handle* h) { return h ? h->err.c_str() : "null handle"; }

int vus_set_variables(vus_handle* h, int kind, int64_t n, const uint64_t* keys, const double* data, int mem) {
  if (!h || kind < 0 || kind >= 4 || n < 0 || mem < 0 || mem > 3) return fail(h, VUS_ERR_INVALID, "vus_set_variables: bad arguments");
  VUS_TRY(h)
  for (int64_t i = 1; i < n; ++i)
    if (keys[i] <= keys[i - 1]) return fail(h, VUS_ERR_INVALID, "vus_set_variables: keys must be strictly ascending");
  h->keys[kind].assign(keys, keys + n);
  h->nvar[kind] = n;
  DBuf<double>& b = h->val[h->cur][kind];
  b.alloc((size_t)kVarDim[kind] * n);
  import_table(b.p, data, n, kVarDim[kind], mem, h->own_stream);
  rt::sync(h->own_stream);
  h->analyzed = false;
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_get_variables(vus_handle* h, int kind, double* out, int mem) {
  if (!h || kind < 0 || kind >= 4) return fail(h, VUS_ERR_INVALID, "vus_get_variables: bad arguments");
  VUS_TRY(h)
  export_table(out, h->val[h->cur][kind].p, h->nvar[kind], kVarDim[kind], mem, h->own_stream);
  rt::sync(h->own_stream);
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_save_values(vus_handle* h) {
  if (!h) return VUS_ERR_INVALID;
  VUS_TRY(h)
  for (int kind = 0; kind < 4; ++kind) {
    const size_t n = (size_t)kVarDim[kind] * h->nvar[kind];
    h->saved[kind].alloc(n);
    rt::d2d(h->saved[kind].p, h->val[h->cur][kind].p, n * sizeof(double), h->own_stream);
  }
  rt::sync(h->own_stream);
  h->has_saved = true;
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_restore_values(vus_handle* h) {
  if (!h) return VUS_ERR_INVALID;
  if (!h->has_saved) return fail(h, VUS_ERR_STATE, "vus_restore_values: nothing saved");
  VUS_TRY(h)
  for (int kind = 0; kind < 4; ++kind)
    rt::d2d(h->val[h->cur][kind].p, h->saved[kind].p, (size_t)kVarDim[kind] * h->nvar[kind] * sizeof(double), h->own_stream);
  rt::sync(h->own_stream);
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_add_factors(vus_handle* h, int type, int64_t n, const int32_t* var_idx, const double* meas, const double* sqrt_info,
                    const int64_t* orig_index, int mem) {
  if (!h || type < 0 || type >= VUS_F_NTYPES || n < 0) return fail(h, VUS_ERR_INVALID, "vus_add_factors: bad arguments");
  VUS_TRY(h)
  FactorTable& T = h->ft[type];
  if (T.n) return fail(h, VUS_ERR_STATE, "vus_add_factors: this factor type was already added (one call per type)");
  T.n = n;
  T.h_idx.assign(var_idx, var_idx + (size_t)kFactorSlots[type] * n);
  T.orig.assign(orig_index, orig_index + n);
  T.idx.upload(T.h_idx, h->own_stream);
  T.meas.alloc((size_t)kFactorMeas[type] * n);
  T.sinfo.alloc((size_t)kFactorInfo[type] * n);
  import_table(T.meas.p, meas, n, kFactorMeas[type], mem, h->own_stream);
  import_table(T.sinfo.p, sqrt_info, n, kFactorInfo[type], mem, h->own_stream);
  rt::sync(h->own_stream);
  h->nfactors += n;
  h->analyzed = false;
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_set_calibration(vus_handle* h, const double K[6]) {
  if (!h || !K) return VUS_ERR_INVALID;
  for (int i = 0; i < 6; ++i) h->K[i] = K[i];
  return VUS_OK;
}
int vus_set_gravity(vus_handle* h, const double g[3]) {
  if (!h || !g) return VUS_ERR_INVALID;
  for (int i = 0; i < 3; ++i) h->grav[i] = g[i];
  return VUS_OK;
}
int vus_set_gtsam_build(vus_handle* h, int tangent_preintegration, int slow_but_correct_betweenfactor) {
  if (!h) return VUS_ERR_INVALID;
  h->opts = (tangent_preintegration ? VUS_OPT_TANGENT : 0) | (slow_but_correct_betweenfactor ? VUS_OPT_SLOW_BETWEEN : 0);
  return VUS_OK;
}
int vus_set_lm_params(vus_handle* h, const vus_lm_params* p) {
  if (!h || !p) return VUS_ERR_INVALID;
  h->prm = *p;
  if (const char* v = std::getenv("VUS_VERBOSE")) h->prm.verbose = std::max(h->prm.verbose, std::atoi(v));   // diagnostics without touching the caller
  return VUS_OK;
}

int vus_set_partition(vus_handle* h, int64_t n_owned_nodes, const int64_t n_owned_factors[6]) {
  if (!h || n_owned_nodes < 0 || !n_owned_factors) return fail(h, VUS_ERR_INVALID, "vus_set_partition: bad arguments");
  h->n_owned = n_owned_nodes;
  for (int t = 0; t < VUS_F_NTYPES; ++t) h->nf_owned[t] = n_owned_factors[t];
  h->analyzed = false;
  return VUS_OK;
}

int vus_set_partition_chain(vus_handle* h, int32_t nside, const int64_t* prev_local, const int64_t* next_local, int32_t rank, int32_t nranks) {
  if (!h || nside < 0 || (nside > 0 && (!prev_local || !next_local)) || rank < 0 || nranks < 1 || rank >= nranks)
    return fail(h, VUS_ERR_INVALID, "vus_set_partition_chain: bad arguments");
  h->sp_prev.assign(prev_local, prev_local + nside); h->sp_next.assign(next_local, next_local + nside);
  h->sp_rank = rank; h->sp_nranks = nranks;
  h->analyzed = false;
  return VUS_OK;
}

int vus_set_comm(vus_handle* h, vus_comm_fn fn, void* ctx) {
  if (!h) return VUS_ERR_INVALID;
  h->comm = fn; h->comm_ctx = ctx;
  return VUS_OK;
}

int vus_nccl_unique_id(void* out128) {
  if (!out128) return VUS_ERR_INVALID;
#ifndef VUS_EMU
  NcclApi& n = nccl_api();
  if (!n.ok) return VUS_ERR_UNSUPPORTED;
  NcclApi::UniqueId id;
  if (n.GetUniqueId(&id) != 0) return VUS_ERR_CUDA;
  std::memcpy(out128, id.internal, 128);
  return VUS_OK;
#else
  std::memset(out128, 0, 128);
  return VUS_ERR_UNSUPPORTED;
#endif
}

int vus_comm_init(vus_handle* h, const void* unique_id128, int rank, int nranks) {
  if (!h || !unique_id128 || rank < 0 || nranks < 1 || rank >= nranks) return fail(h, VUS_ERR_INVALID, "vus_comm_init: bad arguments");
#ifndef VUS_EMU
  VUS_TRY(h)
  NcclApi& n = nccl_api();
  if (!n.ok) return fail(h, VUS_ERR_UNSUPPORTED, "vus_comm_init: " + n.why);
  if (h->nccl) { n.CommDestroy(h->nccl); h->nccl = nullptr; }
  NcclApi::UniqueId id;
  std::memcpy(id.internal, unique_id128, 128);
  nccl_check(n.CommInitRank(&h->nccl, nranks, id, rank), "ncclCommInitRank");
  h->nccl_rank = rank; h->nccl_nranks = nranks;
  h->comm_stream_ordered = true;
  h->drop_graphs();
  return VUS_OK;
  VUS_CATCH(h)
#else
  return fail(h, VUS_ERR_UNSUPPORTED, "vus_comm_init: the host emulation has no NCCL (use vus_set_comm with caller-supplied collectives)");
#endif
}

int vus_set_halo(vus_handle* h, int32_t npeers, const int32_t* peers, const int64_t* send_ptr, const int32_t* send_idx,
                 const int64_t* recv_off, const int64_t* recv_cnt) {
  if (!h || npeers < 0 || (npeers > 0 && (!peers || !send_ptr || !recv_off || !recv_cnt))) return fail(h, VUS_ERR_INVALID, "vus_set_halo: bad arguments");
  VUS_TRY(h)
  h->halo_peers.assign(peers, peers + npeers);
  h->halo_send_ptr.assign(1, 0);
  if (npeers) h->halo_send_ptr.assign(send_ptr, send_ptr + npeers + 1);
  h->halo_recv_off.assign(recv_off, recv_off + npeers);
  h->halo_recv_cnt.assign(recv_cnt, recv_cnt + npeers);
  h->halo_nsend = h->halo_send_ptr.back();
  if (h->halo_nsend) {
    if (!send_idx) return fail(h, VUS_ERR_INVALID, "vus_set_halo: send_idx missing");
    h->halo_send_idx.upload(send_idx, (size_t)h->halo_nsend, h->own_stream);
    h->halo_sendbuf.alloc((size_t)h->halo_nsend * 12);
    rt::sync(h->own_stream);
  }
  h->drop_graphs();
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_set_components(vus_handle* h, int64_t ncomp, const int64_t* node_start) {
  if (!h || ncomp < 1 || !node_start) return VUS_ERR_INVALID;
  if (node_start[0] != 0) return fail(h, VUS_ERR_INVALID, "vus_set_components: node_start[0] must be 0");
  for (int64_t c = 0; c < ncomp; ++c)
    if (node_start[c + 1] <= node_start[c]) return fail(h, VUS_ERR_INVALID, "vus_set_components: node ranges must be ascending and non-empty");
  h->ncomp = (int)ncomp;
  h->comp_start.assign(node_start, node_start + ncomp + 1);
  h->analyzed = false;
  return VUS_OK;
}

int vus_get_component_results(vus_handle* h, vus_component_result* out) {
  if (!h || !out) return VUS_ERR_INVALID;
  if (h->ncomp <= 1 || (int)h->comp_res.size() != h->ncomp) return fail(h, VUS_ERR_STATE, "vus_get_component_results: no batched optimize() has run");
  std::copy(h->comp_res.begin(), h->comp_res.end(), out);
  return VUS_OK;
}

int vus_set_comm_mode(vus_handle* h, int stream_ordered) {
  if (!h) return VUS_ERR_INVALID;
  h->comm_stream_ordered = stream_ordered != 0;
  return VUS_OK;
}

int vus_analyze(vus_handle* h) {
  if (!h) return VUS_ERR_INVALID;
  VUS_TRY(h)
  // validate indices
  const int slot_kind[VUS_F_NTYPES][5] = {{0}, {1}, {0, 0}, {1, 0}, {0, 3}, {0, 1, 0, 1, 2}};
  for (int t = 0; t < VUS_F_NTYPES; ++t) {
    FactorTable& T = h->ft[t];
    for (int s = 0; s < kFactorSlots[t]; ++s)
      for (long f = 0; f < T.n; ++f) {
        const int v = T.h_idx[s * T.n + f];
        if (v < 0 || v >= h->nvar[slot_kind[t][s]]) return fail(h, VUS_ERR_INVALID, "vus_analyze: factor references a variable index out of range");
      }
  }
  return analyze(h, h->own_stream);
  VUS_CATCH(h)
}

int vus_get_layout(vus_handle* h, int64_t out[12]) {
  if (!h || !h->analyzed) return fail(h, VUS_ERR_STATE, "vus_get_layout: call vus_analyze first");
  out[0] = h->D; out[1] = h->k; out[2] = h->Ns; out[3] = h->nrem; out[4] = h->nobs; out[5] = h->B; out[6] = h->L; out[7] = h->nfactors;
  out[8] = h->chunk_P; out[9] = h->Ns_band; out[10] = h->nvar[3]; out[11] = 0;
  return VUS_OK;
}

int vus_optimize(vus_handle* h, void* stream, vus_lm_result* result) {
  if (!h) return VUS_ERR_INVALID;
  if (!h->analyzed) return fail(h, VUS_ERR_STATE, "vus_optimize: call vus_analyze first");
  VUS_TRY(h)
  rt::stream_t st_ = use_stream(h, stream);
  const int rc = h->ncomp > 1 ? optimize_batched(h, st_) : optimize(h, st_);
  if (result) *result = h->res;
  return rc;
  VUS_CATCH(h)
}

int vus_get_trace(vus_handle* h, int32_t capacity, vus_lm_try* out, int32_t* count) {
  if (!h || !count) return VUS_ERR_INVALID;
  *count = (int32_t)h->trace.size();
  if (out) std::copy(h->trace.begin(), h->trace.begin() + std::min<size_t>(h->trace.size(), capacity > 0 ? capacity : 0), out);
  return VUS_OK;
}

int vus_error(vus_handle* h, void* stream, double* out) {
  if (!h || !out) return VUS_ERR_INVALID;
  if (!h->analyzed) return fail(h, VUS_ERR_STATE, "vus_error: call vus_analyze first");
  VUS_TRY(h)
  *out = graph_error(h, h->cur, use_stream(h, stream));
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_factor_errors(vus_handle* h, void* stream, double* out) {
  if (!h || !out) return VUS_ERR_INVALID;
  if (!h->analyzed) return fail(h, VUS_ERR_STATE, "vus_factor_errors: call vus_analyze first");
  VUS_TRY(h)
  rt::stream_t st = use_stream(h, stream);
  run_factors(h, h->cur, false, st);
  finish_order(h->ft[VUS_F_STEREO], st);
  std::vector<double> tmp(h->nfactors);
  rt::d2h(tmp.data(), h->e_all.p, h->nfactors * sizeof(double), st);
  rt::sync(st);
  for (int t = 0; t < VUS_F_NTYPES; ++t) {
    FactorTable& T = h->ft[t];
    for (long f = 0; f < T.n; ++f) {
      if (T.orig[f] < 0 || T.orig[f] >= h->nfactors) return fail(h, VUS_ERR_INVALID, "orig_index out of range");
      out[T.orig[f]] = tmp[T.e_off + f];
    }
  }
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_linearize(vus_handle* h, void* stream, int type, double* r_out, double* J_out) {
  if (!h || type < 0 || type >= VUS_F_NTYPES) return VUS_ERR_INVALID;
  if (!h->analyzed) return fail(h, VUS_ERR_STATE, "vus_linearize: call vus_analyze first");
  VUS_TRY(h)
  rt::stream_t st = use_stream(h, stream);
  run_factors(h, h->cur, true, st);
  FactorTable& T = h->ft[type];
  finish_order(T, st);
  if (r_out) rt::d2h(r_out, T.r.p, (size_t)kFactorM[type] * T.n * sizeof(double), st);
  if (J_out) rt::d2h(J_out, T.J.p, (size_t)kFactorM[type] * kFactorCols[type] * T.n * sizeof(double), st);
  rt::sync(st);
  if (!T.perm.empty()) {                               // rows are stored landmark-major: hand them back in the caller's order
    std::vector<double> tmp(T.n);
    auto unpermute = [&](double* out, int comps) {
      for (int c = 0; c < comps; ++c) {
        double* row = out + (size_t)c * T.n;
        for (long f = 0; f < T.n; ++f) tmp[T.perm[f]] = row[f];
        std::copy(tmp.begin(), tmp.end(), row);
      }
    };
    if (r_out) unpermute(r_out, kFactorM[type]);
    if (J_out) unpermute(J_out, kFactorM[type] * kFactorCols[type]);
  }
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_solve_step(vus_handle* h, void* stream, double lambda, double* d_pose, double* d_vel, double* d_bias, double* d_lm,
                   int32_t* pcg_iterations) {
  if (!h) return VUS_ERR_INVALID;
  if (!h->analyzed) return fail(h, VUS_ERR_STATE, "vus_solve_step: call vus_analyze first");
  if (h->ncomp > 1) return fail(h, VUS_ERR_UNSUPPORTED, "vus_solve_step: not available on a batched graph (vus_set_components)");
  VUS_TRY(h)
  rt::stream_t st = use_stream(h, stream);
  run_factors(h, h->cur, true, st);
  assemble_base(h, st);
  int its = 0;
  const bool ok = solve_damped(h, lambda, st, &its, false);
  if (pcg_iterations) *pcg_iterations = its;
  if (!ok) return fail(h, VUS_ERR_STATE, "vus_solve_step: damped system is not positive definite");
  std::vector<double> x(h->L), xl(3 * h->nvar[3]);
  rt::d2h(x.data(), h->x.p, h->L * sizeof(double), st);
  rt::d2h(xl.data(), h->xl.p, xl.size() * sizeof(double), st);
  rt::sync(st);
  const int D = h->D;
  for (long i = 0; i < h->nvar[0]; ++i)
    for (int c = 0; c < 6; ++c) if (d_pose) d_pose[i * 6 + c] = x[i * D + c];
  for (long i = 0; i < h->nvar[1]; ++i)
    for (int c = 0; c < 3; ++c) if (d_vel) d_vel[i * 3 + c] = x[i * D + 6 + c];
  if (h->has_bias && d_bias) for (int c = 0; c < 6; ++c) d_bias[c] = x[h->Lc + c];
  for (long l = 0; l < h->nvar[3]; ++l)
    for (int c = 0; c < 3; ++c) if (d_lm) d_lm[l * 3 + c] = xl[c * h->nvar[3] + l];
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_marginal_covariance(vus_handle* h, void* stream, int64_t nq, const int32_t* kinds, const int32_t* idx, double* cov_out) {
  if (!h || nq <= 0 || !kinds || !idx || !cov_out) return VUS_ERR_INVALID;
  if (!h->analyzed) return fail(h, VUS_ERR_STATE, "vus_marginal_covariance: call vus_analyze first");
  if (h->ncomp > 1) return fail(h, VUS_ERR_UNSUPPORTED, "vus_marginal_covariance: not available on a batched graph (vus_set_components)");
  if (h->n_owned >= 0) return fail(h, VUS_ERR_UNSUPPORTED, "vus_marginal_covariance: not available on a partitioned graph");
  VUS_TRY(h)
  return marginal_covariance(h, use_stream(h, stream), nq, kinds, idx, cov_out);
  VUS_CATCH(h)
}

int vus_debug_band_solve(vus_handle* h, void* stream, double lambda, double* SD_out, double* SU_out, double* x_inout, int nrhs) {
  if (!h) return VUS_ERR_INVALID;
  if (!h->analyzed) return fail(h, VUS_ERR_STATE, "vus_debug_band_solve: call vus_analyze first");
  if (h->ncomp > 1) return fail(h, VUS_ERR_UNSUPPORTED, "vus_debug_band_solve: not available on a batched graph (vus_set_components)");
  VUS_TRY(h)
  rt::stream_t st = use_stream(h, stream);
  const long BB = (long)h->B * h->B, BBP = bcr_bbp(h->B);
  const int LD = bcr_ld(h->B);
  run_factors(h, h->cur, true, st);
  assemble_base(h, st);
  form_system(h, lambda, st);
  auto unpad = [&](double* out, const double* dev, long nblk) {     // padded device tiles -> dense B x B blocks
    std::vector<double> tmp((size_t)nblk * BBP);
    rt::d2h(tmp.data(), dev, tmp.size() * sizeof(double), st);
    rt::sync(st);
    for (long b = 0; b < nblk; ++b)
      for (int i = 0; i < h->B; ++i)
        for (int j = 0; j < h->B; ++j) out[b * BB + (long)i * h->B + j] = tmp[b * BBP + (long)i * LD + j];
  };
  if (SD_out) unpad(SD_out, band_matrix(h) + h->sd_off, h->Ns);
  if (SU_out && h->Ns > 1) unpad(SU_out, band_matrix(h) + h->su_off, h->Ns - 1);
  bcr_factor(h, st);
  if (x_inout && nrhs > 0) {
    DBuf<double> X;
    X.alloc((size_t)nrhs * h->Lc);
    rt::h2d(X.p, x_inout, (size_t)nrhs * h->Lc * sizeof(double), st);
    bcr_solve(h, X.p, h->Lc, nrhs, st, false);
    rt::d2h(x_inout, X.p, (size_t)nrhs * h->Lc * sizeof(double), st);
    rt::sync(st);
  }
  rt::sync(st);
  return read_fail(h, st) ? 1 : VUS_OK;
  VUS_CATCH(h)
}

int vus_preintegrate_imu(vus_handle* h, void* stream, int64_t n, int32_t k, const double* acc, const double* gyro, double dt,
                         const double bias_hat[6], const double acc_cov[9], const double gyro_cov[9], const double int_cov[9],
                         double* pim_out, double* sqrt_info_out, int mem) {
  if (!h || n < 0 || k <= 0 || !(dt > 0.0) || !acc || !gyro || !pim_out || !sqrt_info_out || (mem != VUS_MEM_HOST_ROWS && mem != VUS_MEM_DEVICE_ROWS))
    return fail(h, VUS_ERR_INVALID, "vus_preintegrate_imu: bad arguments (tables are row-major: mem = VUS_MEM_HOST_ROWS or VUS_MEM_DEVICE_ROWS)");
  VUS_TRY(h)
  rt::stream_t st = use_stream(h, stream);
  if (n == 0) return VUS_OK;
  const size_t in_bytes = (size_t)n * k * 3 * sizeof(double);
  DBuf<double> dacc, dgyr, pim, sinfo;
  PreintArgs P;
  P.n = n; P.k = k; P.dt = dt;
  if (mem == VUS_MEM_HOST_ROWS) {
    dacc.alloc((size_t)n * k * 3); dgyr.alloc((size_t)n * k * 3);
    rt::h2d(dacc.p, acc, in_bytes, st); rt::h2d(dgyr.p, gyro, in_bytes, st);
    P.acc = dacc.p; P.gyro = dgyr.p;
  } else { P.acc = acc; P.gyro = gyro; }
  for (int i = 0; i < 6; ++i) P.bhat[i] = bias_hat ? bias_hat[i] : 0.0;
  for (int i = 0; i < 9; ++i) { P.aC[i] = acc_cov[i]; P.wC[i] = gyro_cov[i]; P.iC[i] = int_cov[i]; }
  pim.alloc((size_t)67 * n); sinfo.alloc((size_t)45 * n);
  P.pim = pim.p; P.sinfo = sinfo.p;
  P.tangent = (h->opts & VUS_OPT_TANGENT) != 0;
  if (!h->fail.p) { h->fail.alloc(1); h->fail.zero(st); }
  P.fail = h->fail.p;
  L_elem<PreintBody>(n, st, P);
  export_table(pim_out, pim.p, n, 67, mem, st);
  export_table(sqrt_info_out, sinfo.p, n, 45, mem, st);
  rt::sync(st);
  if (read_fail(h, st)) return fail(h, VUS_ERR_INVALID, "vus_preintegrate_imu: a preintegrated covariance is not positive definite");
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_backproject_stereo(vus_handle* h, void* stream, int64_t n, const int32_t* pose_idx, const double* meas, double* points_out, int mem) {
  if (!h || n < 0 || !pose_idx || !meas || !points_out || (mem != VUS_MEM_HOST_ROWS && mem != VUS_MEM_DEVICE_ROWS))
    return fail(h, VUS_ERR_INVALID, "vus_backproject_stereo: bad arguments (tables are row-major)");
  if (!h->nvar[0]) return fail(h, VUS_ERR_STATE, "vus_backproject_stereo: set the pose variables first");
  VUS_TRY(h)
  rt::stream_t st = use_stream(h, stream);
  if (n == 0) return VUS_OK;
  for (int64_t o = 0; o < n; ++o)
    if (pose_idx[o] < 0 || pose_idx[o] >= h->nvar[0]) return fail(h, VUS_ERR_INVALID, "vus_backproject_stereo: pose index out of range");
  DBuf<int> didx; didx.upload(pose_idx, (size_t)n, st);
  DBuf<double> dmeas, out;
  dmeas.alloc((size_t)3 * n); out.alloc((size_t)3 * n);
  import_table(dmeas.p, meas, n, 3, mem, st);
  BackprojArgs A;
  A.n = n; A.pose_idx = didx.p; A.pose = h->val[h->cur][0].p; A.nx = h->nvar[0]; A.meas = dmeas.p; A.out = out.p;
  for (int i = 0; i < 6; ++i) A.K[i] = h->K[i];
  if (!h->fail.p) { h->fail.alloc(1); h->fail.zero(st); }
  A.fail = h->fail.p;
  L_elem<BackprojBody>(n, st, A);
  export_table(points_out, out.p, n, 3, mem, st);
  rt::sync(st);
  if (read_fail(h, st)) return fail(h, VUS_ERR_INVALID, "vus_backproject_stereo: non-positive disparity (uL - uR <= 0)");
  return VUS_OK;
  VUS_CATCH(h)
}

int vus_time_linearize(vus_handle* h, void* stream, int reps, double* ms_per_rep) {
  if (!h || !ms_per_rep || reps <= 0) return VUS_ERR_INVALID;
  if (!h->analyzed) return fail(h, VUS_ERR_STATE, "vus_time_linearize: call vus_analyze first");
  VUS_TRY(h)
  rt::stream_t st = use_stream(h, stream);
  run_factors(h, h->cur, true, st);
  rt::sync(st);
#ifndef VUS_EMU
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps; ++i) run_factors(h, h->cur, true, st);
  cudaEventRecord(e1, st);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *ms_per_rep = ms / reps;
#else
  const double t0 = now_ms();
  for (int i = 0; i < reps; ++i) run_factors(h, h->cur, true, st);
  *ms_per_rep = (now_ms() - t0) / reps;
#endif
  return VUS_OK;
  VUS_CATCH(h)
}

}  // extern "C"
