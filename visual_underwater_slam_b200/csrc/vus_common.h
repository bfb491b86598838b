// Common macros for the single-source kernels.
//
// Every kernel body in csrc/ is written once and compiled two ways:
//   * nvcc -gencode arch=compute_100a,code=sm_100a  -> libvus.so, THE product (B200 only);
//   * g++ -x c++ -DVUS_EMU                          -> tests/emu/libvus_emu.so, a sequential
//     host emulation of the same kernel bodies used ONLY by `-m "not gpu"` tests to check
//     indexing / host logic where no GPU exists.  The product loader never opens it.
//
// Two kernel shapes:
//   elementwise  body(args, i)                      one work item per thread, grid-stride;
//   cooperative  body(args, bid, tid, nthr, smem)   one CTA per task, written as
//                `for (i = tid; i < n; i += nthr)` loops separated by VUS_SYNC(), with no
//                thread-private state carried across a sync -- so tid=0,nthr=1 is a valid
//                sequential execution of the same code.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cmath>

#ifdef VUS_EMU
  #define VUS_HD inline
  #define VUS_DEV inline
  #define VUS_SYNC() ((void)0)
  #define VUS_RESTRICT
  #define VUS_SHARED
#else
  #include <cuda_runtime.h>
  #define VUS_HD __host__ __device__ __forceinline__
  #define VUS_DEV __device__ __forceinline__
  #define VUS_SYNC() __syncthreads()
  #define VUS_RESTRICT __restrict__
  #define VUS_SHARED __shared__
#endif

#define VUS_EPS 2.220446049250313e-16

// factor types (also the C-ABI enum, include/vus.h)
enum { VUS_F_PRIOR_POSE = 0, VUS_F_PRIOR_VEL = 1, VUS_F_BETWEEN = 2, VUS_F_DVL = 3, VUS_F_STEREO = 4, VUS_F_IMU = 5, VUS_F_NTYPES = 6 };
// variable kinds
enum { VUS_V_POSE = 0, VUS_V_VEL = 1, VUS_V_BIAS = 2, VUS_V_LM = 3, VUS_V_NKINDS = 4 };

#ifndef VUS_EMU
// FP64 tensor-core tile product D(8x8) += A(8x4) B(4x8)  (SASS: DMMA).  Lane (g = lane/4, t = lane%4) supplies
// A[g][t], B[t][g] and holds D[g][2t], D[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif
