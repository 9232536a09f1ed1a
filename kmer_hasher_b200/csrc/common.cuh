// common.cuh -- shared device/host helpers of libkmergpu (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef __CUDACC__
#error "libkmergpu is CUDA-only: there is no CPU path"
#endif

namespace kmg {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;   // bins of one sort pass with 8-bit digits (the sorted build's; the grouped build may use wider ones)
constexpr int MAX_NB = 1024;             // most bins any pass uses (10-bit digits)
constexpr int MAX_PASSES = 8;            // ceil(64 / RADIX_BITS)
constexpr unsigned FULL = 0xffffffffu;
constexpr unsigned FULL_MASK_ = 0xffffffffu;   // for scopes where a template parameter is named FULL

__host__ __device__ inline uint64_t key_mask(int k) {
  return k < 32 ? ((uint64_t(1) << (2 * k)) - 1) : ~uint64_t(0);
}
__host__ __device__ inline int num_passes(int k) { return (2 * k + RADIX_BITS - 1) / RADIX_BITS; }
template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// streaming loads/stores: every array on this path is touched once per kernel
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// one 32-byte sector in ONE request (sm_100 256-bit load); p must be 32-byte aligned
__device__ __forceinline__ void ld_stream_sector(const uint4 *p, uint4 &a, uint4 &b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t *p) {
  uint64_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p) {
  uint64_t r;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(r) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// A bijection of 64-bit words with good avalanche (the murmur3 finaliser) and its inverse.  The grouped
// build sorts records by mix64(key): equal words <=> equal k-mers, and the low 40 bits of the mix tell almost
// all distinct k-mers of a genome apart, whereas all 64 bits of the key itself are needed to do so.
__host__ __device__ inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
__host__ __device__ inline uint64_t unmix64(uint64_t x) {
  x ^= x >> 33; x *= 0x9cb4b2f8129337dbULL;      // inverse of 0xc4ceb9fe1a85ec53 mod 2^64
  x ^= x >> 33; x *= 0x4f74430c22a54005ULL;      // inverse of 0xff51afd7ed558ccd mod 2^64
  x ^= x >> 33;
  return x;
}

// warp-level inclusive scans
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(FULL, v, d);
    if (lane_id() >= (unsigned)d) v += t;
  }
  return v;
}
__device__ __forceinline__ uint64_t warp_incl_scan64(uint64_t v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint64_t t = __shfl_up_sync(FULL, v, d);
    if (lane_id() >= (unsigned)d) v += t;
  }
  return v;
}
__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
  return v;
}

}  // namespace kmg
