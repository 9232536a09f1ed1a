// lookback.cuh -- decoupled look-back tile prefixes (single-pass chained scan).
//
// Tiles take their id from an atomic ticket, so every tile a waiter depends on is already
// resident: the wait is deadlock-free without cooperative launch.
#pragma once
#include "common.cuh"

namespace kmg {

// ---- status word: [63:62] flag, [61:56] epoch, [55:0] value -------------------------------------
// The epoch lets one status buffer serve consecutive launches without being cleared: a word only
// counts as published when its epoch matches the launch's.
constexpr uint64_t ST_AGG = 1, ST_INCL = 2;
__device__ __forceinline__ uint64_t st_pack(uint64_t flag, uint32_t epoch, uint64_t v) {
  return (flag << 62) | (uint64_t(epoch & 63u) << 56) | (v & ((uint64_t(1) << 56) - 1));
}
__device__ __forceinline__ uint64_t st_value(uint64_t w) { return w & ((uint64_t(1) << 56) - 1); }
__device__ __forceinline__ uint64_t st_flag(uint64_t w, uint32_t epoch) {
  return (((w >> 56) & 63u) == (epoch & 63u)) ? (w >> 62) : 0;
}

// (the sort pass does its per-bin look-back inline, one bin per thread: sort.cuh)

// ---- two-value scalar look-back used by the compaction kernels -------------------------------------
// One 16-byte status per tile: (a, b) each carrying the flag in bits [63:62] (values < 2^62).
// The buffer must be zeroed before the launch.  Call with all 32 lanes of one warp.
struct alignas(16) Pair64 { uint64_t a, b; };

__device__ __forceinline__ void st_pair(Pair64 *p, uint64_t flag, uint64_t a, uint64_t b) {
  uint64_t wa = (flag << 62) | a, wb = (flag << 62) | b;
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(wa), "l"(wb) : "memory");
}
__device__ __forceinline__ uint64_t ld_pair(const Pair64 *p, uint64_t &a, uint64_t &b) {
  uint64_t wa, wb;
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(wa), "=l"(wb) : "l"(p) : "memory");
  uint64_t fa = wa >> 62, fb = wb >> 62;
  a = wa & ((uint64_t(1) << 62) - 1);
  b = wb & ((uint64_t(1) << 62) - 1);
  return fa == fb ? fa : 0;   // a torn read counts as "not yet published"
}

__device__ __forceinline__ void pair_lookback(Pair64 *status, uint32_t tile, uint64_t agg_a,
                                              uint64_t agg_b, uint64_t &excl_a, uint64_t &excl_b) {
  const unsigned lane = lane_id();
  if (tile == 0) {
    if (lane == 0) st_pair(status, ST_INCL, agg_a, agg_b);
    excl_a = excl_b = 0;
    return;
  }
  if (lane == 0) st_pair(status + tile, ST_AGG, agg_a, agg_b);
  uint64_t ea = 0, eb = 0;
  int64_t base = (int64_t)tile - 1;          // lane l inspects tile base-l
  while (true) {
    int64_t t = base - (int64_t)lane;
    uint64_t a = 0, b = 0, f = ST_INCL;      // tiles before 0 behave as an inclusive 0
    if (t >= 0) {
      do { f = ld_pair(status + t, a, b); } while (f == 0);
    }
    unsigned incl = __ballot_sync(FULL, f == ST_INCL);
    unsigned upto = incl ? (unsigned)__ffs(incl) - 1 : 31u;   // nearest inclusive lane
    if (lane > upto) a = b = 0;
    ea += warp_sum64(a);
    eb += warp_sum64(b);
    if (incl) break;
    base -= 32;
  }
  if (lane == 0) st_pair(status + tile, ST_INCL, ea + agg_a, eb + agg_b);
  excl_a = ea;
  excl_b = eb;
}

}  // namespace kmg
