// Minimal runtime layer: device memory, copies and the two launch shapes (see vus_common.h).
// CUDA build: cudaMalloc / cudaMemcpyAsync / <<<>>> on the caller's stream.
// VUS_EMU build (tests only): malloc / memcpy / sequential loops.
#pragma once
#include "vus_common.h"
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>
#include <vector>

namespace vus { namespace rt {

#ifdef VUS_EMU
typedef void* stream_t;
inline void check_last(const char*) {}
inline void* dalloc(size_t bytes) { void* p = std::calloc(bytes ? bytes : 1, 1); if (!p) throw std::runtime_error("emu alloc failed"); return p; }
inline void dfree(void* p) { std::free(p); }
inline void pool_setup(int) {}
inline void h2d(void* d, const void* s, size_t n, stream_t) { std::memcpy(d, s, n); }
inline void d2h(void* d, const void* s, size_t n, stream_t) { std::memcpy(d, s, n); }
inline void d2d(void* d, const void* s, size_t n, stream_t) { std::memmove(d, s, n); }
inline void dzero(void* d, size_t n, stream_t) { std::memset(d, 0, n); }
inline void sync(stream_t) {}
inline int sm_count() { return 4; }
inline int& tl_device() { static thread_local int d = 0; return d; }
inline stream_t& tl_stream() { static thread_local stream_t s = nullptr; return s; }

template <class Body, class Args>
inline void launch_elem(long n, stream_t, const Args& a) {
  for (long i = 0; i < n; ++i) Body::run(a, i);
}
template <class Body, class Args>
inline void launch_coop(int grid, int /*block*/, size_t smem_bytes, stream_t, const Args& a) {
  std::vector<double> smem(smem_bytes / sizeof(double) + 8);
  for (int b = 0; b < grid; ++b) Body::run(a, b, 0, 1, smem.data());
}
#else
typedef cudaStream_t stream_t;
inline void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
inline void check_last(const char* what) { check(cudaGetLastError(), what); }
// Device memory comes from the default stream-ordered pool with an unlimited release threshold: a handle that is
// destroyed and re-created (one optimize() per incoming graph) reuses the pool instead of paying cudaMalloc/cudaFree
// (un)mapping of several GB every time.
inline void pool_setup(int device) {
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    unsigned long long thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
}
// The C-ABI entry points run under a DeviceGuard (vus.cu) that makes the handle's device current on the calling thread and
// records it -- and the stream the handle works on -- here: per-device caches below are keyed by it, and memory is allocated
// and freed in the order of the handle's stream, never on the legacy default stream (which a non-blocking compute stream
// does not wait for).
inline int& tl_device() { static thread_local int d = 0; return d; }
inline stream_t& tl_stream() { static thread_local stream_t s = nullptr; return s; }
inline void* dalloc(size_t bytes) {
  void* p = nullptr;
  check(cudaMallocAsync(&p, bytes ? bytes : 8, tl_stream()), "cudaMallocAsync");
  check(cudaStreamSynchronize(tl_stream()), "alloc sync");       // usable from any stream once the call returns
  return p;
}
inline void dfree(void* p) { if (p) cudaFreeAsync(p, tl_stream()); }
inline void h2d(void* d, const void* s, size_t n, stream_t st) { if (n) check(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st), "h2d"); }
inline void d2h(void* d, const void* s, size_t n, stream_t st) { if (n) check(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st), "d2h"); }
inline void d2d(void* d, const void* s, size_t n, stream_t st) { if (n) check(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st), "d2d"); }
inline void dzero(void* d, size_t n, stream_t st) { if (n) check(cudaMemsetAsync(d, 0, n, st), "memset"); }
inline void sync(stream_t st) { check(cudaStreamSynchronize(st), "stream sync"); }
constexpr int kMaxDevices = 64;
inline int sm_count() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = tl_device() & (kMaxDevices - 1);
  int v = n[dev].load();
  if (!v) { cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, tl_device()); if (v <= 0) v = 148; n[dev].store(v); }
  return v;
}

template <class Body, class Args>
__global__ void __launch_bounds__(256) k_elem(Args a, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) Body::run(a, i);
}
// per-body launch bounds (specialise for register-heavy cooperative bodies that must keep 2 CTAs per SM)
template <class Body> struct CoopBounds { static constexpr int kMaxThreads = 1024, kMinBlocks = 1; };
template <class Body, class Args>
__global__ void __launch_bounds__(CoopBounds<Body>::kMaxThreads, CoopBounds<Body>::kMinBlocks) k_coop(Args a) {
  extern __shared__ double vus_smem[];
  Body::run(a, (int)blockIdx.x, (int)threadIdx.x, (int)blockDim.x, vus_smem);
}
// grid sized in multiples of the SM count (148 on B200), capped by the work available
template <class Body, class Args>
inline void launch_elem(long n, stream_t st, const Args& a) {
  if (n <= 0) return;
  const int block = 256;
  long need = (n + block - 1) / block;
  long cap = (long)sm_count() * 8;
  int grid = (int)(need < cap ? need : cap);
  k_elem<Body, Args><<<grid, block, 0, st>>>(a, n);
  check_last("launch_elem");
}
template <class Body, class Args>
inline void launch_coop(int grid, int block, size_t smem_bytes, stream_t st, const Args& a) {
  if (grid <= 0) return;
  if (smem_bytes > 48 * 1024) {
    static std::atomic<size_t> configured[kMaxDevices];   // per (Body, Args) instantiation AND per device (the attribute is per device)
    std::atomic<size_t>& c = configured[tl_device() & (kMaxDevices - 1)];
    if (smem_bytes > c.load()) {
      check(cudaFuncSetAttribute(k_coop<Body, Args>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes), "smem attr");
      c.store(smem_bytes);
    }
  }
  k_coop<Body, Args><<<grid, block, smem_bytes, st>>>(a);
  check_last("launch_coop");
}
#endif

// typed device buffer
template <class T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() {}
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { dfree(p); }
  void alloc(size_t count) { if (count > n || !p) { dfree(p); p = (T*)dalloc(count * sizeof(T)); n = count; } }
  void upload(const T* h, size_t count, stream_t st) { alloc(count); h2d(p, h, count * sizeof(T), st); }
  void upload(const std::vector<T>& h, stream_t st) { upload(h.data(), h.size(), st); }
  void zero(stream_t st) { dzero(p, n * sizeof(T), st); }
  void swap(DBuf& o) { std::swap(p, o.p); std::swap(n, o.n); }
  void release() { dfree(p); p = nullptr; n = 0; }
};

}}  // namespace vus::rt
