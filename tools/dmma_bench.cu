// microbenchmark: FP64 DMMA (mma.sync m8n8k4) vs DFMA register-tile GEMM for batched BxB blocks
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// pure issue throughput
__global__ void k_peak884(double* out, int iters) {
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_peak1688(double* out, int iters) {
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = 0.0;
  double a[4] = {threadIdx.x * 1e-3, 1.0, 2.0, 0.5}, b[2] = {1.0 + threadIdx.x * 1e-4, 0.25};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) dmma1688(c + 4 * i, a, b);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_peakfma(double* out, int iters) {
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = i;
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
  }
  double s = 0; for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- CTA GEMM with DMMA: C[B x B] = A[B x B] * Bm[B x B], operands staged in smem with stride LD
// A row-major [i][k], Bm row-major [k][j].  8 warps in a 4 x 2 grid over the 8x8 tile grid.
template <int B>
struct Cfg {
  static constexpr int KP = (B + 3) & ~3;
  static constexpr int LD = (KP % 8 == 4) ? KP : KP + 4;
  static constexpr int T = (B + 7) / 8;          // tiles per dim
  static constexpr int ROWS = T * 8;             // padded rows
  static constexpr int SM_DOUBLES = ROWS * LD + 8;
};
template <int B, bool TRANS_A, bool TRANS_B>
__device__ __forceinline__ void cta_gemm_mma(double* C, const double* sA, const double* sB, double alpha, int warp, int lane) {
  using F = Cfg<B>;
  constexpr int T = F::T, LD = F::LD;
  constexpr int WR = 4, WC = 2;
  constexpr int TR = (T + WR - 1) / WR, TC = (T + WC - 1) / WC;
  const int wr = warp / WC, wc = warp % WC;
  const int ti0 = wr * TR, tj0 = wc * TC;
  const int g = lane >> 2, t = lane & 3;
  double acc[TR][TC][2];
#pragma unroll
  for (int a = 0; a < TR; ++a)
#pragma unroll
    for (int b = 0; b < TC; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  for (int k0 = 0; k0 < F::KP; k0 += 4) {
    double af[TR], bf[TC];
#pragma unroll
    for (int a = 0; a < TR; ++a) {
      const int i = (ti0 + a) * 8 + g;
      af[a] = (ti0 + a < T) ? (TRANS_A ? sA[(k0 + t) * LD + i] : sA[i * LD + k0 + t]) : 0.0;
    }
#pragma unroll
    for (int b = 0; b < TC; ++b) {
      const int j = (tj0 + b) * 8 + g;
      bf[b] = (tj0 + b < T) ? (TRANS_B ? sB[j * LD + k0 + t] : sB[(k0 + t) * LD + j]) : 0.0;
    }
#pragma unroll
    for (int a = 0; a < TR; ++a)
#pragma unroll
      for (int b = 0; b < TC; ++b)
        if (ti0 + a < T && tj0 + b < T) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
  }
#pragma unroll
  for (int a = 0; a < TR; ++a)
#pragma unroll
    for (int b = 0; b < TC; ++b) {
      const int i = (ti0 + a) * 8 + g, j = (tj0 + b) * 8 + 2 * t;
      if (ti0 + a < T && tj0 + b < T && i < B) {
        if (j < B) C[i * B + j] = alpha * acc[a][b][0];
        if (j + 1 < B) C[i * B + j + 1] = alpha * acc[a][b][1];
      }
    }
}
template <int B>
__device__ __forceinline__ void stage(double* s, const double* g, int tid, int nthr) {
  using F = Cfg<B>;
  for (int e = tid; e < F::ROWS * F::LD + 8; e += nthr) {
    const int i = e / F::LD, j = e - i * F::LD;
    s[e] = (i < B && j < B) ? g[i * B + j] : 0.0;
  }
}
template <int B, bool TA, bool TB>
__global__ void __launch_bounds__(256) k_gemm_mma(const double* A, const double* Bm, double* C, int reps) {
  extern __shared__ double sm[];
  using F = Cfg<B>;
  double* sA = sm;
  double* sB = sm + F::SM_DOUBLES;
  const long o = (long)blockIdx.x * B * B;
  stage<B>(sA, A + o, threadIdx.x, blockDim.x);
  stage<B>(sB, Bm + o, threadIdx.x, blockDim.x);
  __syncthreads();
  for (int r = 0; r < reps; ++r) cta_gemm_mma<B, TA, TB>(C + o, sA, sB, 1.0, threadIdx.x >> 5, threadIdx.x & 31);
}

// ---- the DFMA 8x4 register-tile version currently in kernels.cuh
__device__ __forceinline__ int bcr_ldb(int B) { return (B + 3) & ~3; }
__device__ void cta_gemm_ss(double* C, const double* sA, const double* sB, int B, double alpha, int tid, int nthr) {
  const int ldb = bcr_ldb(B);
  const int TI = (B + 7) / 8, TJ = (B + 3) / 4;
  for (int tile = tid; tile < TI * TJ; tile += nthr) {
    const int i0 = (tile / TJ) * 8, j0 = (tile % TJ) * 4;
    double acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const double* pa = sA + (long)i0 * B;
    const double* pb = sB + j0;
#pragma unroll 3
    for (int k = 0; k < B; ++k) {
      double bv[4];
      const double2 b01 = *reinterpret_cast<const double2*>(pb + (long)k * ldb);
      const double2 b23 = *reinterpret_cast<const double2*>(pb + (long)k * ldb + 2);
      bv[0] = b01.x; bv[1] = b01.y; bv[2] = b23.x; bv[3] = b23.y;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const double av = pa[(long)a * B + k];
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] += av * bv[b];
      }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = i0 + a, j = j0 + b;
        if (i < B && j < B) C[(long)i * B + j] = alpha * acc[a][b];
      }
  }
}
__global__ void __launch_bounds__(256) k_gemm_fma(const double* A, const double* Bm, double* C, int B, int reps) {
  extern __shared__ double sm[];
  const int ldb = bcr_ldb(B);
  double* sA = sm;
  double* sB = sm + ((B + 7) & ~7) * B;
  const long o = (long)blockIdx.x * B * B;
  for (int e = threadIdx.x; e < B * B; e += blockDim.x) { sA[e] = A[o + e]; sB[(e / B) * ldb + e % B] = Bm[o + e]; }
  __syncthreads();
  for (int r = 0; r < reps; ++r) cta_gemm_ss(C + o, sA, sB, B, 1.0, threadIdx.x, blockDim.x);
}

template <class F> float time_ms(F f, int n = 5) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); for (int i = 0; i < n; ++i) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / n;
}

int main() {
  constexpr int B = 81;
  const int nb = 148 * 16;
  double* out; CK(cudaMalloc(&out, 148 * 8 * 256 * 8));
  {
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
      float ms = time_ms([&] { k_peak884<<<148, warps * 32>>>(out, iters); });
      printf("peak m8n8k4   %2d warps/SM: %.2f TFLOP/s\n", warps, 148.0 * warps * iters * 8 * 512 / ms / 1e9);
      ms = time_ms([&] { k_peak1688<<<148, warps * 32>>>(out, iters); });
      printf("peak m16n8k8  %2d warps/SM: %.2f TFLOP/s\n", warps, 148.0 * warps * iters * 4 * 2048 / ms / 1e9);
      ms = time_ms([&] { k_peakfma<<<148, warps * 32>>>(out, iters); });
      printf("peak dfma     %2d warps/SM: %.2f TFLOP/s\n", warps, 148.0 * warps * 32 * iters * 16 * 2 / ms / 1e9);
    }
  }
  std::vector<double> hA((size_t)nb * B * B), hB(hA.size()), hC(hA.size());
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = (rand() % 2001 - 1000) * 1e-3; hB[i] = (rand() % 2001 - 1000) * 1e-3; }
  double *dA, *dB, *dC;
  CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dB, hA.size() * 8)); CK(cudaMalloc(&dC, hA.size() * 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hA.size() * 8, cudaMemcpyHostToDevice));
  auto check = [&](const char* name, bool ta, bool tb) {
    CK(cudaMemcpy(hC.data(), dC, hA.size() * 8, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int blk : {0, nb - 1}) {
      const double* a = &hA[(size_t)blk * B * B]; const double* b = &hB[(size_t)blk * B * B]; const double* c = &hC[(size_t)blk * B * B];
      for (int i = 0; i < B; ++i) for (int j = 0; j < B; ++j) {
        double s = 0; for (int k = 0; k < B; ++k) s += (ta ? a[k * B + i] : a[i * B + k]) * (tb ? b[j * B + k] : b[k * B + j]);
        maxerr = fmax(maxerr, fabs(s - c[i * B + j]));
      }
    }
    printf("%s maxerr %.3e\n", name, maxerr);
  };
  const double flop = 2.0 * B * B * B * nb;
  {
    using F = Cfg<B>;
    const size_t smem = 2 * F::SM_DOUBLES * 8;
    printf("mma smem %zu LD %d\n", smem, F::LD);
    CK(cudaFuncSetAttribute(k_gemm_mma<B, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_gemm_mma<B, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_gemm_mma<B, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_mma<B, false, false><<<nb, 256, smem>>>(dA, dB, dC, 1); CK(cudaDeviceSynchronize()); check("mma NN", false, false);
    k_gemm_mma<B, true, false><<<nb, 256, smem>>>(dA, dB, dC, 1); CK(cudaDeviceSynchronize()); check("mma TN", true, false);
    k_gemm_mma<B, false, true><<<nb, 256, smem>>>(dA, dB, dC, 1); CK(cudaDeviceSynchronize()); check("mma NT", false, true);
    for (int reps : {1, 8}) {
      float ms = time_ms([&] { k_gemm_mma<B, false, false><<<nb, 256, smem>>>(dA, dB, dC, reps); });
      printf("mma NN reps %d: %.3f ms  %.2f TFLOP/s useful\n", reps, ms, flop * reps / ms / 1e9);
      ms = time_ms([&] { k_gemm_mma<B, true, false><<<nb, 256, smem>>>(dA, dB, dC, reps); });
      printf("mma TN reps %d: %.3f ms  %.2f TFLOP/s useful\n", reps, ms, flop * reps / ms / 1e9);
      ms = time_ms([&] { k_gemm_mma<B, false, true><<<nb, 256, smem>>>(dA, dB, dC, reps); });
      printf("mma NT reps %d: %.3f ms  %.2f TFLOP/s useful\n", reps, ms, flop * reps / ms / 1e9);
    }
  }
  {
    const size_t smem = (((B + 7) & ~7) * B + B * ((B + 3) & ~3)) * 8;
    CK(cudaFuncSetAttribute(k_gemm_fma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_gemm_fma<<<nb, 256, smem>>>(dA, dB, dC, B, 1); CK(cudaDeviceSynchronize()); check("fma NN", false, false);
    for (int reps : {1, 8}) {
      float ms = time_ms([&] { k_gemm_fma<<<nb, 256, smem>>>(dA, dB, dC, B, reps); });
      printf("fma NN reps %d: %.3f ms  %.2f TFLOP/s useful\n", reps, ms, flop * reps / ms / 1e9);
    }
  }
  return 0;
}
