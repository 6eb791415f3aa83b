// Hand-written sm_100a kernels of the fractional-step time loop (product code).
// Everything here is HBM/L2-bandwidth-bound FP64 + int32 work: no tensor cores (nothing on this
// path is a dense contraction).  Design rules applied: 128-bit coalesced loads of the matrix
// streams (Sell layout, device.cuh), one pass per fused phase, device-side reductions, one
// persistent cooperative kernel per Krylov solve.
#pragma once
#include <cooperative_groups.h>

#include "device.cuh"

namespace cfdb {
namespace cg = cooperative_groups;

// ---- streaming loads: read-only, do not allocate in L1 (matrix values / columns are used once) ----
__device__ __forceinline__ double2 ldnc_d2(const double* p) {
   double2 r;
   asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
   return r;
}
__device__ __forceinline__ int2 ldnc_i2(const int* p) {
   int2 r;
   asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
   return r;
}

// ---- reductions ----------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
   return v;
}

// Sum N values over the block; result valid in every thread.  sh needs N*32 doubles.
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* sh) {
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
   __syncthreads();
#pragma unroll
   for (int i = 0; i < N; i++) {
      double s = warp_sum(v[i]);
      if (lane == 0) sh[i * 32 + w] = s;
   }
   __syncthreads();
#pragma unroll
   for (int i = 0; i < N; i++) {
      double s = (lane < nw) ? sh[i * 32 + lane] : 0.0;
      v[i] = warp_sum(s);
   }
}

// Deterministic grid-wide sum of N values inside a cooperative kernel: block partials -> global
// scratch -> grid.sync -> every block re-reduces all partials in the same order (bit-identical
// result in every block, so all blocks take the same branches).  `buf` must alternate between
// consecutive calls (scratch holds 2 buffers of N_MAX*gridDim doubles).
template <int N>
__device__ __forceinline__ void grid_sum(cg::grid_group& grid, double (&v)[N], double* buf, double* sh) {
   block_sum<N>(v, sh);
   if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < N; i++) buf[(size_t)i * gridDim.x + blockIdx.x] = v[i];
   }
   grid.sync();
#pragma unroll
   for (int i = 0; i < N; i++) {
      double s = 0.0;
      for (int b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += __ldcg(buf + (size_t)i * gridDim.x + b);
      v[i] = s;
   }
   block_sum<N>(v, sh);
}

__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v) {
   // non-negative doubles order like their bit patterns
   atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

}   // namespace cfdb
