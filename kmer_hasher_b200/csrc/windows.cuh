// windows.cuh -- the 2-bit rolling encoder, as a tile loader.
//
// Replaces init_kmer/skip_n (src/kmer_util.c:4-32) + the rolling loop of seq_to_hash
// (src/kmer_pos.c:78-97): instead of one sequential scan, every window start q is an independent
// predicate + funnel-shift over a 2-bit packed copy of the tile in shared memory.
//
// Window rule (verified against the reference engine): the window that starts at
// global byte p is emitted iff
//   (i)   p + k <= L,
//   (ii)  none of seq[p .. p+k) is a breaker, i.e. (c | 0x20) == 'n'   (src/kmer_util.h:10),
//   (iii) NOT (p + k == L and (p == 0 or seq[p-1] is a breaker))        (src/kmer_pos.c:81-83:
//         a freshly primed window that ends on the terminator is dropped).
// key = sum code(seq[p+j]) << 2(k-1-j), code(c) = (c >> 1) & 3         (src/kmer_util.h:8).
#pragma once
#include "common.cuh"

namespace kmg {

// A rank's view of the (global) sequence.  base[0] is the byte of window start s0 and is 16-byte
// aligned; base[-1] is readable when s0 > 0; bytes base[0 .. avail) exist (the buffer itself is
// padded to a multiple of 16 beyond avail).
struct SeqView {
  const uint8_t *base;
  int64_t nstarts;   // window starts q in [0, nstarts) are handled here (p = s0 + q)
  int64_t avail;     // valid bytes from base: min(L, s1 + k - 1) - s0
  int64_t s0;        // global index of base[0]
  int64_t L;         // global sequence length
  int k;
};

// 16 ASCII bytes -> 32 bits of codes (first base in the top two bits) + 16 breaker bits (first
// base in bit 0).
__device__ __forceinline__ uint32_t pack4_codes(uint32_t x) {
  // per byte (c>>1)&3, then gather c0<<6|c1<<4|c2<<2|c3 into the top byte by one multiply
  return (((x >> 1) & 0x03030303u) * 0x40100401u) >> 24;
}
__device__ __forceinline__ uint32_t pack4_breakers(uint32_t x) {
  uint32_t eq = __vcmpeq4(x | 0x20202020u, 0x6E6E6E6Eu) & 0x01010101u;   // 1 per breaker byte
  return ((eq * 0x01020408u) >> 24) & 0xFu;                               // b0 | b1<<1 | b2<<2 | b3<<3
}
__device__ __forceinline__ void pack16(uint4 v, uint32_t &codes, uint32_t &brk) {
  codes = (pack4_codes(v.x) << 24) | (pack4_codes(v.y) << 16) | (pack4_codes(v.z) << 8) | pack4_codes(v.w);
  brk = pack4_breakers(v.x) | (pack4_breakers(v.y) << 4) | (pack4_breakers(v.z) << 8) | (pack4_breakers(v.w) << 12);
}

// Shared-memory image of one tile: TILE window starts need TILE + k - 1 bytes; groups of 16.
template <int TILE>
struct TileCodes {
  static constexpr int GROUPS = TILE / 16 + 3;
  uint32_t codes[GROUPS];
  uint16_t brk[GROUPS];
  int any_breaker;
};

// Stage 1: all threads of the block pack the tile that starts at window q0 (multiple of 16).
// Returns (block-uniform) whether the tile contains a breaker or touches the end of the data.
template <int TILE, int THREADS>
__device__ __forceinline__ bool tile_pack(const SeqView &sv, int64_t q0, TileCodes<TILE> &tc) {
  constexpr int GROUPS = TileCodes<TILE>::GROUPS;
  int local_any = 0;
  for (int g = threadIdx.x; g < GROUPS; g += THREADS) {
    int64_t off = q0 + (int64_t)g * 16;
    uint32_t c = 0, b = 0xFFFFu;
    if (off < sv.avail) {                                   // buffer is padded to 16
      uint4 v = ld_stream_u4(reinterpret_cast<const uint4 *>(sv.base + off));
      pack16(v, c, b);
      int64_t left = sv.avail - off;                        // bytes of this group that exist
      if (left < 16) b |= (0xFFFFu << left) & 0xFFFFu;      // beyond the data = breaker
    }
    tc.codes[g] = c;
    tc.brk[g] = (uint16_t)b;
    local_any |= (b != 0);
  }
  return __syncthreads_or(local_any) != 0;
}

// Stage 2: the window that starts at tile-local index t (global start q0 + t).
template <int TILE>
__device__ __forceinline__ uint64_t tile_key(const TileCodes<TILE> &tc, int t, int k) {
  const int g = t >> 4, sh = (t & 15) * 2;
  uint64_t a = (uint64_t(tc.codes[g]) << 32) | tc.codes[g + 1];
  uint64_t full = (a << sh) | (uint64_t(tc.codes[g + 2]) >> (32 - sh));
  return full >> (64 - 2 * k);
}
// code of the four bases that start at tile-local index t (t < TILE + 32: reads two groups only)
template <int TILE>
__device__ __forceinline__ uint32_t tile_c4(const TileCodes<TILE> &tc, int t) {
  const int g = t >> 4, sh = (t & 15) * 2;
  const uint64_t a = (uint64_t(tc.codes[g]) << 32) | tc.codes[g + 1];
  return (uint32_t)((a << sh) >> 56);
}
template <int TILE>
__device__ __forceinline__ bool tile_window_clean(const TileCodes<TILE> &tc, int t, int k) {
  const int g = t >> 4, o = t & 15;
  uint64_t m = uint64_t(tc.brk[g]) | (uint64_t(tc.brk[g + 1]) << 16) | (uint64_t(tc.brk[g + 2]) << 32);
  return ((m >> o) & ((uint64_t(1) << k) - 1)) == 0;
}
// Rule (iii), only ever true for the single window with p + k == L.
__device__ __forceinline__ bool dropped_last_window(const SeqView &sv, int64_t q) {
  int64_t p = sv.s0 + q;
  if (p + sv.k != sv.L) return false;
  if (p == 0) return true;
  uint8_t c = sv.base[q - 1];
  return (c | 0x20) == 'n';
}
// Full validity of window q = q0 + t (any_special = result of tile_pack).
template <int TILE>
__device__ __forceinline__ bool tile_valid(const SeqView &sv, const TileCodes<TILE> &tc, int64_t q0,
                                           int t, bool any_special) {
  int64_t q = q0 + t;
  if (q >= sv.nstarts) return false;
  if (any_special && !tile_window_clean(tc, t, sv.k)) return false;
  if (sv.s0 + q + sv.k == sv.L && dropped_last_window(sv, q)) return false;
  return true;
}

}  // namespace kmg
