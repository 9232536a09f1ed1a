// csr.cuh -- sorted (key,pos) records -> CSR index {ukeys[U], ustart[U+1], pos[N]} and the kernels
// that read it back out in the layouts of kmer_positions (src/kmer_hash.c:1054-1147).
#pragma once
#include "common.cuh"
#include "lookback.cuh"

namespace kmg {

// Device-resident facts about an index; the host reads them once at the end of the build.
struct IndexStats {
  uint64_t n;        // records (= rows of `pos`)
  uint64_t U;        // distinct k-mers
  uint64_t P;        // sum n(n-1)/2 (= rows of `pair.pos`)
  uint64_t multi;    // k-mers with more than one position
  uint32_t maxc;     // longest position list
  uint32_t unstable; // != 0: some k-mer's positions are not ascending (the sort was not stable): the caller rebuilds
};

// ---- run-length pass: one head per distinct key, in order (single pass, chained scan) ----------------
// The same sweep reads the positions and checks that they ascend inside every run of equal keys: that is what the
// stable sort promises (and the reference's insertion order gives), and the one-atomic rank variant of the sort pass
// only delivers it where the hardware applies colliding shared-memory atomics in lane order.  A violation sets
// st->unstable; the host then rebuilds with the order-independent variant (api.cu, build_from_view).
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS)
rle_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pos, IndexStats *st, uint64_t *__restrict__ ukeys,
           uint32_t *__restrict__ ustart, Pair64 *status, uint32_t *ticket, const bool hashed) {
  constexpr int TILE = THREADS * ITEMS, WARPS = THREADS / 32;
  __shared__ uint32_t s_tile, s_wsum[WARPS];
  __shared__ uint64_t s_base;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t n = st->n;
  const uint64_t q0 = (uint64_t)tile * TILE;
  if (q0 >= n) {
    if (n == 0 && tile == 0 && tid == 0) { ustart[0] = 0; st->U = 0; }
    return;
  }
  const uint64_t w0 = q0 + (uint64_t)warp * (32 * ITEMS);
  uint64_t key[ITEMS];
  uint32_t p[ITEMS];
  uint32_t heads = 0, running = 0;
  uint64_t carry = 0;                                  // key just before this warp item (lane 0's predecessor)
  uint32_t pcarry = 0;                                 // and its position
  bool bad = false;
  // all loads first (independent, all in flight), then the shuffle chain
  if (lane == 0 && w0 > 0 && w0 < n) { carry = ld_stream_u64(keys + w0 - 1); pcarry = ld_stream_u32(pos + w0 - 1); }
  if (q0 + TILE <= n) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) key[i] = ld_stream_u64(keys + w0 + i * 32 + lane);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) p[i] = ld_stream_u32(pos + w0 + i * 32 + lane);
  } else {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const uint64_t idx = w0 + i * 32 + lane;
      key[i] = idx < n ? ld_stream_u64(keys + idx) : 0;
      p[i] = idx < n ? ld_stream_u32(pos + idx) : 0;
    }
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint64_t idx = w0 + i * 32 + lane;
    const bool ok = idx < n;
    uint64_t prev = __shfl_up_sync(FULL, key[i], 1);
    uint32_t pprev = __shfl_up_sync(FULL, p[i], 1);
    if (lane == 0) { prev = carry; pprev = pcarry; }
    const bool head = ok && (idx == 0 || key[i] != prev);
    bad |= ok && !head && p[i] <= pprev;               // same k-mer as the record before: its position must be larger
    running += __popc(__ballot_sync(FULL, head));
    if (head) heads |= 1u << i;
    carry = __shfl_sync(FULL, key[i], 31);             // lane 0 uses it next round
    pcarry = __shfl_sync(FULL, p[i], 31);
  }
  if (__any_sync(FULL, bad) && lane == 0) st->unstable = 1;
  if (lane == 0) s_wsum[warp] = running;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    uint32_t c = s_wsum[w];
    if (w < (int)warp) wbase += c;
    total += c;
  }
  if (warp == 0) {
    uint64_t ea, eb;
    pair_lookback(status, tile, total, 0, ea, eb);
    if (lane == 0) s_base = ea;
  }
  __syncthreads();
  uint64_t base = s_base + wbase;                      // heads before this warp's item i (ballots are cheaper than 16 live registers)
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const bool head = (heads >> i) & 1u;
    const unsigned bal = __ballot_sync(FULL, head);
    if (head) {
      const uint64_t u = base + __popc(bal & lanemask_lt());
      ukeys[u] = hashed ? unmix64(key[i]) : key[i];        // grouped build: records carry mix64(key)
      ustart[u] = (uint32_t)(w0 + i * 32 + lane);
    }
    base += __popc(bal);
  }
  if (q0 + TILE >= n && tid == 0) {                    // the last tile closes the CSR
    const uint64_t U = s_base + total;
    ustart[U] = (uint32_t)n;
    st->U = U;
  }
}

// ---- grouped build: make equal k-mers contiguous after a sort on the low `bits` bits of mix64(key) -------
// Records are ordered by the low bits of h = mix64(key), stably.  A "group" is a run of equal low bits; almost
// every group holds one k-mer.  Where it holds several (a collision of the low bits: ~N^2 / 2^(bits+1) pairs)
// they interleave and the group has to be partitioned by h, stably.  group_detect_kernel finds those groups
// (read-only; groups whose k-mers are already contiguous are left alone), small_fix_kernel sorts the short ones in
// place, big_fix_kernel partitions the long ones.
constexpr int SMALL_GROUP = 64;
constexpr uint32_t CLAIM_SLOTS = 1u << 16;

struct FixLists {
  uint2 *small_tasks;      // (start, end) of dirty groups of at most SMALL_GROUP records
  uint2 *big_tasks;        // (start, end) of longer ones
  uint32_t *counters;      // [0] small tasks, [1] big tasks, [2] overflow flag
  uint32_t *claim;         // CLAIM_SLOTS words, zero: one task per big group
  uint32_t small_cap, big_cap;
};

__device__ __forceinline__ uint64_t first_at_least(const uint64_t *h, uint64_t n, uint64_t lowmask, uint64_t v) {
  uint64_t lo = 0, hi = n;                               // first j with (h[j] & lowmask) >= v
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if ((h[mid] & lowmask) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__device__ __forceinline__ void file_group(const uint64_t *__restrict__ h, uint64_t n, uint64_t lowmask, uint64_t i, const FixLists &fl) {
  // h[i-1] and h[i] share the low bits and differ: a boundary between two k-mers inside one group.
  // The group's extent: almost every group is a handful of records, so walk outwards a few steps (the lines were just
  // streamed: L2 hits) and fall back to binary searches only for long groups (a repeat k-mer that happens to collide).
  const uint64_t low = h[i] & lowmask;
  constexpr int WALK = 24;
  uint64_t s = i, e = i + 1;
  int w = 0;
  while (s > 0 && w < WALK && (h[s - 1] & lowmask) == low) { --s; ++w; }
  if (s > 0 && (h[s - 1] & lowmask) == low) s = first_at_least(h, n, lowmask, low);
  w = 0;
  while (e < n && w < WALK && (h[e] & lowmask) == low) { ++e; ++w; }
  if (e < n && (h[e] & lowmask) == low) e = low == lowmask ? n : first_at_least(h, n, lowmask, low + 1);
  if (e - s <= SMALL_GROUP) {
    bool first = true;                                   // the group's first boundary decides for the group
    const uint64_t h0 = h[s];
    for (uint64_t j = s + 1; j < i; ++j) first &= h[j] == h0;
    if (first) {
      // Only a group in which some k-mer's records are SEPARATED by another's needs work: k-mers that merely share the low
      // bits but already sit one after the other (the usual case: two k-mers with one position each) are left as they
      // are -- the order of the k-mers inside a group means nothing (the key table's bucket depends on the low bits only).
      bool dirty = false;
      for (uint64_t j = s + 1; j < e && !dirty; ++j) {
        const uint64_t x = h[j];
        if (x != h[j - 1])
          for (uint64_t m = s; m + 1 < j; ++m)
            if (h[m] == x) { dirty = true; break; }
      }
      if (dirty) {
        const uint32_t t = atomicAdd(fl.counters + 0, 1u);
        if (t < fl.small_cap) fl.small_tasks[t] = make_uint2((uint32_t)s, (uint32_t)e); else fl.counters[2] = 1;
      }
    }
  } else {
    uint32_t slot = (uint32_t)(mix64(s) & (CLAIM_SLOTS - 1));
    bool mine = false, seen = false;
    for (uint32_t probe = 0; probe < CLAIM_SLOTS && !mine && !seen; ++probe) {
      const uint32_t old = atomicCAS(fl.claim + slot, 0u, (uint32_t)s + 1u);
      mine = old == 0u;
      seen = old == (uint32_t)s + 1u;                    // another boundary of the same group got there first
      slot = (slot + 1) & (CLAIM_SLOTS - 1);
    }
    if (!mine && !seen) fl.counters[2] = 1;              // claim table full: the caller rebuilds sorted by key
    if (mine) {
      const uint32_t t = atomicAdd(fl.counters + 1, 1u);
      if (t < fl.big_cap) fl.big_tasks[t] = make_uint2((uint32_t)s, (uint32_t)e); else fl.counters[2] = 1;
    }
  }
}

// file_group by a whole warp (every lane calls with the same i): the 32 records around the boundary arrive as ONE coalesced
// load instead of a serial walk; the group's extent comes from a ballot, "is this the group's first boundary" from another,
// and "do some k-mer's records lie apart" from __match_any_sync.  Groups that reach the window's edge take the serial path.
__device__ __forceinline__ void file_group_warp(const uint64_t *__restrict__ h, uint64_t n, uint64_t lowmask, uint64_t i, const FixLists &fl) {
  const unsigned lane = lane_id();
  const uint64_t w0 = i >= 16 ? i - 16 : 0;
  const unsigned p = (unsigned)(i - w0);                   // the boundary's second record sits in lane p (>= 1)
  const bool valid = w0 + lane < n;
  const uint64_t x = valid ? h[w0 + lane] : 0;
  const uint64_t low = __shfl_sync(FULL, x, p) & lowmask;
  const unsigned m = __ballot_sync(FULL, valid && (x & lowmask) == low);
  const unsigned below = ~m & ((1u << p) - 1u);            // lanes before p outside the group
  const unsigned s = below ? 32u - __clz(below) : 0u;
  const unsigned above = ~m >> p;                          // bit t: lane p + t outside the group (lane p itself is inside)
  const unsigned e = above ? p + (unsigned)__ffs(above) - 1u : 32u;
  if ((s == 0 && w0 > 0) || (e == 32 && w0 + 32 < n)) {    // may extend beyond the window: the general path
    if (lane == 0) file_group(h, n, lowmask, i, fl);
    return;
  }
  const unsigned run = (e >= 32 ? ~0u : ((1u << e) - 1u)) & ~((1u << s) - 1u);
  const bool in = (run >> lane) & 1u;
  const uint64_t xs = __shfl_sync(FULL, x, s);
  if (__ballot_sync(FULL, in && lane < p && x != xs)) return;          // an earlier boundary of this group decides for it
  const unsigned mm = __match_any_sync(FULL, in ? x : ~uint64_t(lane)) & run;   // lanes of the group with this lane's k-mer
  const unsigned t = mm ? mm >> (__ffs(mm) - 1) : 0u;
  const bool apart = in && (t & (t + 1u)) != 0;            // not one contiguous run
  if (__ballot_sync(FULL, apart) && lane == 0) {
    const uint32_t task = atomicAdd(fl.counters + 0, 1u);
    if (task < fl.small_cap) fl.small_tasks[task] = make_uint2((uint32_t)(w0 + s), (uint32_t)(w0 + e)); else fl.counters[2] = 1;
  }
}

// Four records per thread; boundaries are rare, the stream is the cost -- and the instruction count: the test runs on
// every record, so it is kept branch-free on 32-bit halves.  LOW32: the sorted bits are exactly the low word (the plan of
// every build up to 400 M records); otherwise the two words are masked first (40 bits beyond 400 M records, fewer in tests).
//   E[i]: records i and i+1 (of the seven the thread sees: 4q-2 .. 4q+4) share the sorted bits; D[i]: they differ as k-mers.
//   boundary before record 4q+j  <=>  E[j+1] && D[j+1];  nothing to do if the group is just those two records, i.e. neither
//   E[j] nor E[j+2] -- by far the commonest collision (two k-mers with one position each).
template <bool LOW32>
__global__ void group_detect_kernel(const uint64_t *__restrict__ h, const IndexStats *st, int bits, FixLists fl) {
  const uint64_t n = st->n;
  const uint64_t lowmask = bits >= 64 ? ~uint64_t(0) : ((uint64_t(1) << bits) - 1);
  const uint32_t mlo = (uint32_t)lowmask, mhi = (uint32_t)(lowmask >> 32);
  const uint64_t quads = n / 4;
  const unsigned lane = lane_id();
  // whole warps iterate together: groups that need a closer look are handled by the warp, one after the other
  for (uint64_t q0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); q0 < quads; q0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t q = q0 + lane;
    unsigned need = 0;                                     // bit j: the boundary before record 4q + j needs a closer look
    if (q < quads) {
      // all seven records requested at once (the neighbours' lines are in L1 or L2: other threads stream them)
      const uint4 *h4 = reinterpret_cast<const uint4 *>(h);
      const uint4 a = __ldg(h4 + 2 * q), b = __ldg(h4 + 2 * q + 1);
      const uint4 pv = q ? __ldg(h4 + 2 * q - 1) : make_uint4(0, 0, 0, 0);
      const uint64_t nx = 4 * q + 4 < n ? __ldg(h + 4 * q + 4) : 0;
      const uint32_t lo[7] = {pv.x, pv.z, a.x, a.z, b.x, b.z, (uint32_t)nx};
      const uint32_t hi[7] = {pv.y, pv.w, a.y, a.w, b.y, b.w, (uint32_t)(nx >> 32)};
      unsigned E = 0, D = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        bool e, d;
        if constexpr (LOW32) { e = lo[i] == lo[i + 1]; d = hi[i] != hi[i + 1]; }
        else {
          const uint32_t xl = lo[i] ^ lo[i + 1], xh = hi[i] ^ hi[i + 1];
          e = ((xl & mlo) | (xh & mhi)) == 0;
          d = (xl | xh) != 0;
        }
        E |= (unsigned)e << i;
        D |= (unsigned)d << i;
      }
      // records that do not exist: 4q-2, 4q-1 for q = 0 (bits 0, 1 of E; the boundary before record 0), 4q+4 at the end (bit 5)
      if (q == 0) { E &= ~3u; D &= ~2u; }
      if (4 * q + 4 >= n) E &= ~(1u << 5);
      const unsigned bnd = (E & D) >> 1;                   // bit j: a boundary before record 4q + j (E[j+1] && D[j+1])
      need = bnd & ((E | (E >> 2)) & 0xFu);                // ... whose group is more than the two records
    }
    if (!__any_sync(FULL, need != 0)) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unsigned bal = __ballot_sync(FULL, (need >> j) & 1u);
      while (bal) {
        const int src = __ffs(bal) - 1;
        bal &= bal - 1;
        file_group_warp(h, n, lowmask, 4 * (q0 + (uint64_t)src) + (uint64_t)j, fl);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < n - 4 * quads) {
    const uint64_t i = 4 * quads + threadIdx.x;
    if (i > 0) {
      const uint64_t x = h[i - 1], y = h[i];
      if (((x ^ y) & lowmask) == 0 && x != y) file_group(h, n, lowmask, i, fl);
    }
  }
}

// one thread per short dirty group: stable insertion sort of its records by h, in place
__global__ void small_fix_kernel(uint64_t *__restrict__ h, uint32_t *__restrict__ pos, FixLists fl) {
  const uint32_t nt = min(fl.counters[0], fl.small_cap);
  for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    const uint2 se = fl.small_tasks[t];
    const int c = (int)(se.y - se.x);
    if (c <= 8) {                                          // the common case (2-3 records): sort in place, no staging arrays
      for (int j = 1; j < c; ++j) {
        const uint64_t x = h[se.x + j];
        const uint32_t y = pos[se.x + j];
        int m = j - 1;
        while (m >= 0 && h[se.x + m] > x) { h[se.x + m + 1] = h[se.x + m]; pos[se.x + m + 1] = pos[se.x + m]; --m; }
        if (m + 1 != j) { h[se.x + m + 1] = x; pos[se.x + m + 1] = y; }
      }
      continue;
    }
    uint64_t hh[SMALL_GROUP];
    uint32_t pp[SMALL_GROUP];
    for (int j = 0; j < c; ++j) { hh[j] = h[se.x + j]; pp[j] = pos[se.x + j]; }
    for (int j = 1; j < c; ++j) {
      const uint64_t x = hh[j];
      const uint32_t y = pp[j];
      int m = j - 1;
      while (m >= 0 && hh[m] > x) { hh[m + 1] = hh[m]; pp[m + 1] = pp[m]; --m; }
      hh[m + 1] = x; pp[m + 1] = y;
    }
    for (int j = 0; j < c; ++j) { h[se.x + j] = hh[j]; pos[se.x + j] = pp[j]; }
  }
}

// one CTA per long dirty group: its distinct h values in ascending order, each compacted (stably) into scratch
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
big_fix_kernel(uint64_t *__restrict__ h, uint32_t *__restrict__ pos, uint64_t *__restrict__ sh, uint32_t *__restrict__ sp, FixLists fl) {
  __shared__ uint64_t s_min[THREADS / 32];
  __shared__ uint32_t s_cnt[THREADS / 32];
  __shared__ uint64_t s_cur;
  __shared__ uint32_t s_placed;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t nt = min(fl.counters[1], fl.big_cap);
  for (uint32_t t = blockIdx.x; t < nt; t += gridDim.x) {
    const uint2 se = fl.big_tasks[t];
    const uint32_t s = se.x, e = se.y;
    if (e <= s) continue;
    uint32_t placed = 0;
    bool have_last = false;
    uint64_t last = 0;
    while (placed < e - s) {
      uint64_t m = ~uint64_t(0);                         // smallest h not placed yet (h > last)
      bool any = false;
      for (uint32_t j = s + tid; j < e; j += THREADS) {
        const uint64_t v = h[j];
        if (!have_last || v > last) { if (!any || v < m) m = v; any = true; }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) { const uint64_t o = __shfl_xor_sync(FULL, m, d); m = o < m ? o : m; }
      if (lane == 0) s_min[warp] = m;
      __syncthreads();
      if (tid == 0) {
        uint64_t mm = s_min[0];
        for (int w = 1; w < THREADS / 32; ++w) mm = s_min[w] < mm ? s_min[w] : mm;
        s_cur = mm;                                      // exists: placed < e - s
        s_placed = placed;
      }
      __syncthreads();
      const uint64_t cur = s_cur;
      if (have_last && cur <= last) break;               // nothing left above `last` (only on inconsistent input): never spin
      for (uint32_t base = s; base < e; base += THREADS) {
        const uint32_t j = base + tid;
        const bool hit = j < e && h[j] == cur;
        const unsigned bal = __ballot_sync(FULL, hit);
        if (lane == 0) s_cnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) { if (w < (int)warp) before += s_cnt[w]; total += s_cnt[w]; }
        const uint32_t at = s + s_placed + before + __popc(bal & lanemask_lt());
        if (hit) { sh[at] = cur; sp[at] = pos[j]; }
        __syncthreads();
        if (tid == 0) s_placed += total;
        __syncthreads();
      }
      placed = s_placed;
      last = cur;
      have_last = true;
      __syncthreads();
    }
    for (uint32_t j = s + tid; j < e; j += THREADS) { h[j] = sh[j]; pos[j] = sp[j]; }
    __syncthreads();
  }
}

// ---- per-k-mer facts: P, multi, max count -------------------------------------------------------------
// Not part of the build: only pair.pos (its row count P, the list of k-mers with pairs) and kmg_index_stats need them, so
// they are taken on first demand (api.cu, ensure_stats) -- as the reference only walks the lists for pairs under opt.flag 4
// (src/kmer_hash.c:1113).
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
stats_kernel(const uint32_t *__restrict__ ustart, const uint64_t U, IndexStats *st) {
  uint64_t P = 0, multi = 0;
  uint32_t maxc = 0;
  // four list starts per 16-byte load; the count of the last needs the next group's first start
  const uint64_t groups = U / 4;
  for (uint64_t g = (uint64_t)blockIdx.x * THREADS + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * THREADS) {
    const uint4 v = ld_stream_u4(reinterpret_cast<const uint4 *>(ustart) + g);
    const uint32_t nx = __ldg(ustart + 4 * g + 4);
    const uint32_t c[4] = {v.y - v.x, v.z - v.y, v.w - v.z, nx - v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      P += (uint64_t)c[j] * (c[j] - 1) / 2;
      multi += c[j] > 1;
      maxc = max(maxc, c[j]);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < U - 4 * groups) {
    const uint64_t u = 4 * groups + threadIdx.x;
    const uint64_t c = ustart[u + 1] - ustart[u];
    P += c * (c - 1) / 2;
    multi += c > 1;
    maxc = max(maxc, (uint32_t)c);
  }
  P = warp_sum64(P);
  multi = warp_sum64(multi);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) maxc = max(maxc, __shfl_xor_sync(FULL, maxc, d));
  if (lane_id() == 0) {
    if (P) atomicAdd((unsigned long long *)&st->P, (unsigned long long)P);
    if (multi) atomicAdd((unsigned long long *)&st->multi, (unsigned long long)multi);
    if (maxc) atomicMax(&st->maxc, maxc);
  }
}

// ---- shared helper: which segment does each of a block's rows fall in? -----------------------------------
// off[0..cnt) strictly increasing, off[0] <= r0.  For rows r0 .. r0+T the block computes
// seg[s] = (largest m with off[m] <= r0+s) - m_first, and returns m_first.
// For every block of T consecutive rows starting at first + b*T: the largest m with off[m] <= that row
// (off[0..cnt) strictly increasing, off[0] <= first).  One thread per block; a few microseconds per chunk.
template <typename OffT>
__global__ void segment_starts_kernel(const OffT *__restrict__ off, uint64_t cnt, uint64_t first, uint64_t T, uint64_t nblocks,
                                      uint64_t *__restrict__ out) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const uint64_t r0 = first + b * T;
  uint64_t lo = 0, hi = cnt;                             // invariant: off[lo] <= r0 < off[hi] (off[cnt] = +inf)
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if ((uint64_t)off[mid] <= r0) lo = mid; else hi = mid;
  }
  out[b] = lo;
}

template <int THREADS, int PER, typename OffT>
__device__ __forceinline__ uint64_t block_segments(const OffT *__restrict__ off, uint64_t cnt, uint64_t r0, uint64_t m_first_in,
                                                   uint8_t *s_flag /* THREADS*PER */, uint32_t *s_seg /* THREADS*PER */,
                                                   uint32_t *s_warp /* THREADS/32 */, uint64_t *s_first) {
  static_assert(PER == 8, "flags are read as one 64-bit word per thread");
  constexpr int T = THREADS * PER, WARPS = THREADS / 32;
  const unsigned tid = threadIdx.x;
  if (tid == 0) *s_first = m_first_in;                   // from segment_starts_kernel: largest m with off[m] <= r0
  reinterpret_cast<uint64_t *>(s_flag)[tid] = 0;
  __syncthreads();
  const uint64_t m_first = *s_first;
  for (uint64_t m = m_first + 1 + tid; m < cnt; m += THREADS) {
    const uint64_t o = (uint64_t)off[m];
    if (o >= r0 + T) break;
    s_flag[o - r0] = 1;
  }
  __syncthreads();
  const uint64_t f = reinterpret_cast<const uint64_t *>(s_flag)[tid];
  uint32_t mine = __popcll(f);
  uint32_t incl = warp_incl_scan(mine);
  if (lane_id() == 31) s_warp[tid >> 5] = incl;
  __syncthreads();
  uint32_t run = incl - mine;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) if (w < (int)(tid >> 5)) run += s_warp[w];
  uint32_t out[PER];
#pragma unroll
  for (int j = 0; j < PER; ++j) { run += (uint32_t)((f >> (8 * j)) & 1u); out[j] = run; }
  uint4 *dst = reinterpret_cast<uint4 *>(s_seg + tid * PER);
  dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
  dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
  __syncthreads();
  return m_first;
}

// ---- flag 8: counts ---------------------------------------------------------------------------------------
__global__ void counts_kernel(const uint32_t *__restrict__ ustart, uint64_t U, int32_t *__restrict__ out) {
  // ustart may be offset by a chunk start (any 4-byte alignment): peel to 16-byte alignment of both arrays if possible
  const bool vec = ((reinterpret_cast<uintptr_t>(ustart) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  const uint64_t groups = vec ? U / 4 : 0;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x) {
    const uint4 v = ld_stream_u4(reinterpret_cast<const uint4 *>(ustart) + g);
    const uint32_t nx = __ldg(ustart + 4 * g + 4);
    reinterpret_cast<int4 *>(out)[g] = make_int4((int)(v.y - v.x), (int)(v.z - v.y), (int)(v.w - v.z), (int)(nx - v.w));
  }
  for (uint64_t u = 4 * groups + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < U; u += (uint64_t)gridDim.x * blockDim.x)
    out[u] = (int32_t)(ustart[u + 1] - ustart[u]);
}

// ---- flag 2: interleaved (i,pos), i = 1-based rank of the k-mer ------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
positions_kernel(const uint32_t *__restrict__ ustart, uint64_t U, const uint32_t *__restrict__ pos, uint64_t first,
                 uint64_t nrows, const uint64_t *__restrict__ blk_first, int2 *__restrict__ out, const uint32_t i_base) {
  constexpr int PER = 8, T = THREADS * PER;
  __shared__ __align__(16) uint8_t s_flag[T];
  __shared__ __align__(16) uint32_t s_seg[T];
  __shared__ uint32_t s_warp[THREADS / 32];
  __shared__ uint64_t s_first;
  const uint64_t b0 = (uint64_t)blockIdx.x * T;
  if (b0 >= nrows) return;
  const uint64_t r0 = first + b0;
  const uint64_t u0 = block_segments<THREADS, PER, uint32_t>(ustart, U, r0, blk_first[blockIdx.x], s_flag, s_seg, s_warp, &s_first);
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const uint32_t s = j * THREADS + threadIdx.x;
    if (b0 + s < nrows) out[b0 + s] = make_int2((int)(i_base + u0 + s_seg[s] + 1), (int)ld_stream_u32(pos + r0 + s));
  }
}

// ---- flag 4: pairs ----------------------------------------------------------------------------------------
// Step 1 (once per index): list the k-mers that have pairs and prefix-sum their pair counts.
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS)
pair_index_kernel(const uint32_t *__restrict__ ustart, uint64_t U, uint32_t *__restrict__ multi_u,
                  uint64_t *__restrict__ pair_off, Pair64 *status, uint32_t *ticket) {
  constexpr int TILE = THREADS * ITEMS, WARPS = THREADS / 32;
  __shared__ uint32_t s_tile, s_wcnt[WARPS];
  __shared__ uint64_t s_wsum[WARPS], s_base_cnt, s_base_sum;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint64_t q0 = (uint64_t)tile * TILE;
  if (q0 >= U) return;
  // blocked arrangement: thread owns ITEMS consecutive k-mers
  const uint64_t u0 = q0 + (uint64_t)tid * ITEMS;
  uint32_t c[ITEMS], cnt = 0;
  uint64_t sum = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const uint64_t u = u0 + i;
    c[i] = u < U ? ustart[u + 1] - ustart[u] : 0;
    if (c[i] > 1) { cnt++; sum += (uint64_t)c[i] * (c[i] - 1) / 2; }
  }
  const uint32_t icnt = warp_incl_scan(cnt);
  const uint64_t isum = warp_incl_scan64(sum);
  if (lane == 31) { s_wcnt[warp] = icnt; s_wsum[warp] = isum; }
  __syncthreads();
  uint64_t bcnt = 0, bsum = 0, tcnt = 0, tsum = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    if (w < (int)warp) { bcnt += s_wcnt[w]; bsum += s_wsum[w]; }
    tcnt += s_wcnt[w]; tsum += s_wsum[w];
  }
  if (warp == 0) {
    uint64_t ea, eb;
    pair_lookback(status, tile, tcnt, tsum, ea, eb);
    if (lane == 0) { s_base_cnt = ea; s_base_sum = eb; }
  }
  __syncthreads();
  uint64_t m = s_base_cnt + bcnt + (icnt - cnt);
  uint64_t o = s_base_sum + bsum + (isum - sum);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    if (c[i] > 1) {
      multi_u[m] = (uint32_t)(u0 + i);
      pair_off[m] = o;
      ++m;
      o += (uint64_t)c[i] * (c[i] - 1) / 2;
    }
  }
}

// pair t of a list of n positions, enumerated (0,1),(0,2)..(0,n-1),(1,2).. as the reference's nested
// loops do (src/kmer_hash.c:1108,1114): rows before first index j: S(j) = j(2n-j-1)/2.
__device__ __forceinline__ void unrank_pair(uint64_t t, uint64_t n, uint32_t &a, uint32_t &b) {
  const uint64_t m = 2 * n - 1;
  const uint64_t D = m * m - 8 * t;                 // exact in 64 bits for n <= 2^31
  uint64_t r = (uint64_t)sqrt((double)D);
  while (r * r > D) --r;
  while ((r + 1) * (r + 1) <= D) ++r;
  uint64_t j = (m - r) / 2;                         // candidate, then settle exactly
  if (j > n - 2) j = n - 2;
  while (j > 0 && j * (2 * n - j - 1) / 2 > t) --j;
  while (j + 1 <= n - 2 && (j + 1) * (2 * n - j - 2) / 2 <= t) ++j;
  a = (uint32_t)j;
  b = (uint32_t)(j + 1 + (t - j * (2 * n - j - 1) / 2));
}

// Step 2: rows [first, first+nrows) of the pair matrix.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
pairs_kernel(const uint32_t *__restrict__ ustart, const uint32_t *__restrict__ pos,
             const uint32_t *__restrict__ multi_u, const uint64_t *__restrict__ pair_off, uint64_t n_multi,
             uint64_t first, uint64_t nrows, const uint64_t *__restrict__ blk_first, int32_t *__restrict__ out, const uint32_t i_base) {
  constexpr int PER = 8, T = THREADS * PER;
  __shared__ __align__(16) uint8_t s_flag[T];
  __shared__ __align__(16) uint32_t s_seg[T];
  __shared__ uint32_t s_warp[THREADS / 32];
  __shared__ uint64_t s_first;
  const uint64_t b0 = (uint64_t)blockIdx.x * T;
  if (b0 >= nrows) return;
  const uint64_t r0 = first + b0;
  const uint64_t m0 = block_segments<THREADS, PER, uint64_t>(pair_off, n_multi, r0, blk_first[blockIdx.x], s_flag, s_seg, s_warp, &s_first);
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const uint32_t s = j * THREADS + threadIdx.x;
    if (b0 + s >= nrows) continue;
    const uint64_t m = m0 + s_seg[s];
    const uint32_t u = multi_u[m];
    const uint32_t a0 = ustart[u], cnt = ustart[u + 1] - a0;
    uint32_t a, b;
    unrank_pair(r0 + s - pair_off[m], cnt, a, b);
    int32_t *row = out + 3 * (b0 + s);
    row[0] = (int32_t)(i_base + u + 1);
    row[1] = (int32_t)pos[a0 + a];
    row[2] = (int32_t)pos[a0 + b];
  }
}

// ---- flag 1: k-mer strings (kmer_seq, src/kmer_hash.c:123-133; alphabet A,C,T,G, :21) --------------------
__global__ void kmers_ascii_kernel(const uint64_t *__restrict__ ukeys, uint64_t U, int k, char *__restrict__ out) {
  const uint64_t total = U * (uint64_t)(k + 1);
  for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < total; b += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t u = b / (uint64_t)(k + 1);
    const int j = (int)(b - u * (uint64_t)(k + 1));
    char c = 0;
    if (j < k) c = "ACTG"[(ukeys[u] >> (2 * (k - 1 - j))) & 3u];
    out[b] = c;
  }
}

}  // namespace kmg
