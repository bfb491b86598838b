// streaming benchmark: bulk (TMA) global->shared copies of padded 84x84 FP64 blocks vs 8-byte cp.async of 81x81 blocks
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int KP = 84, LD = 84, BBP = KP * LD;
// each CTA streams `per` consecutive padded blocks through one buffer and accumulates a checksum
__global__ void __launch_bounds__(256, 3) k_tma(const double* blocks, double* out, int per) {
  extern __shared__ __align__(128) double sm[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  double acc = 0.0;
  unsigned parity = 0;
  for (int b = 0; b < per; ++b) {
    const double* src = blocks + ((long)blockIdx.x * per + b) * BBP;
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bar, BBP * 8);
      bulk_g2s(sm, src, BBP * 8, &bar);
    }
    mbar_wait(&bar, parity);
    parity ^= 1;
    for (int e = threadIdx.x; e < BBP; e += 256) acc += sm[e];
    __syncthreads();
  }
  out[(long)blockIdx.x * 256 + threadIdx.x] = acc;
}
// same with 8-byte cp.async of unpadded 81x81 blocks into the padded buffer
__global__ void __launch_bounds__(256, 3) k_ldgsts(const double* blocks, double* out, int per) {
  extern __shared__ __align__(128) double sm[];
  const int B = 81;
  double acc = 0.0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = 0; b < per; ++b) {
    const double* src = blocks + ((long)blockIdx.x * per + b) * B * B;
    for (int r = warp; r < KP; r += 8)
      for (int c = lane; c < LD; c += 32) {
        double* dst = sm + r * LD + c;
        if (r < B && c < B) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src + (long)r * B + c) : "memory");
        } else *dst = 0.0;
      }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    for (int e = threadIdx.x; e < BBP; e += 256) acc += sm[e];
    __syncthreads();
  }
  out[(long)blockIdx.x * 256 + threadIdx.x] = acc;
}

int main() {
  const int per = 3, grid = 5556;
  const long nblk = (long)grid * per;
  std::vector<double> h((size_t)nblk * BBP);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (double)(i % 97) * 0.01;
  double *d, *out;
  CK(cudaMalloc(&d, h.size() * 8)); CK(cudaMalloc(&out, (size_t)grid * 256 * 8));
  CK(cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  const size_t smem = BBP * 8;
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int which = 0; which < 2; ++which) {
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) k_tma<<<grid, 256, smem>>>(d, out, per); else k_ldgsts<<<grid, 256, smem>>>(d, out, per);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = which == 0 ? (double)nblk * BBP * 8 : (double)nblk * 81 * 81 * 8;
      printf("%s: %.3f ms  %.1f GB/s\n", which == 0 ? "tma bulk 84x84" : "ldgsts 81x81  ", ms, bytes / ms / 1e6);
    }
    std::vector<double> ho((size_t)grid * 256);
    CK(cudaMemcpy(ho.data(), out, ho.size() * 8, cudaMemcpyDeviceToHost));
    double s = 0; for (double v : ho) s += v;
    printf("  checksum %.6e\n", s);
  }
  double ref = 0; for (size_t i = 0; i < h.size(); ++i) ref += h[i];
  printf("  ref tma checksum %.6e\n", ref);
  return 0;
}
