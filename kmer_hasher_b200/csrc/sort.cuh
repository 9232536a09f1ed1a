// sort.cuh -- the index build's core: a stable single-sweep ("onesweep") binning pass.
//
// Replaces kh_put + kv_push (src/kmer_pos.c:36-50, src/khash.h:307-348, src/kvec.h:74-80): instead
// of one random DRAM probe per k-mer, (key,pos) records are ordered by a stable LSD radix sort,
// 8 bits per pass.  Stability keeps each k-mer's positions ascending (what insertion order gives
// the reference, README "positions are sorted").
//
// One kernel does a whole pass:
//   - records come either from HBM arrays or straight from the ASCII sequence (the 2-bit encoder
//     of windows.cuh is fused into the first pass: keys are never written unsorted, and windows
//     that contain an N are dropped by simply not being ranked);
//   - ranks inside the tile come from warp ballots (match on the 8 bin bits), so the pass is
//     insensitive to skew (homopolymers, microsatellites);
//   - the tile is regrouped by bin in shared memory and written out in runs;
//   - tile offsets chain through a decoupled look-back, one status word per (tile, bin);
//   - while the regrouped keys stream out, the histogram of the NEXT pass's digit is taken, so
//     keys are read once per pass and there is no separate multi-digit histogram kernel.
// The same kernel with OwnerBin (key-range owner instead of digit) is the multi-GPU partitioner.
#pragma once
#include "common.cuh"
#include "lookback.cuh"
#include "windows.cuh"

namespace kmg {

struct DigitBin {
  int shift;
  __device__ __forceinline__ uint32_t operator()(uint64_t key) const { return (uint32_t)(key >> shift) & (RADIX - 1); }
};
// owner r holds keys in [spl[r-1], spl[r]): bin = number of splitters <= key
struct OwnerBin {
  const uint64_t *spl;
  int nparts;
  __device__ __forceinline__ uint32_t operator()(uint64_t key) const {
    int lo = 0, hi = nparts - 1;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (__ldg(spl + mid) <= key) lo = mid + 1; else hi = mid;
    }
    return (uint32_t)lo;
  }
};
struct NoBin {
  __device__ __forceinline__ uint32_t operator()(uint64_t) const { return 0; }
};

template <class BinFn, class NextFn>
struct PassParams {
  SeqView sv;                 // FROM_SEQ source
  const uint64_t *keys_in;    // record source
  const uint32_t *pos_in;
  uint64_t *keys_out;
  uint32_t *pos_out;
  const uint32_t *hist_cur;   // [RADIX] global histogram of this pass's bins (complete)
  uint32_t *hist_next;        // [RADIX] accumulates the next pass's histogram, or nullptr
  uint64_t *status;           // [tiles][RADIX] look-back words
  uint32_t *ticket;           // tile id dispenser (zero before launch)
  const uint64_t *n_records;  // record source: exact number of records (device scalar)
  uint32_t epoch;
  unsigned long long *trace;  // tuning runs only: 8 clock64 stamps per tile, or nullptr
  uint32_t dbg;               // tuning runs only (wrong results): 1 no look-back wait, 2 no global stores, 4 no next-histogram
  BinFn bin;
  NextFn next;
};

// block-wide exclusive scan of one value per bin (thread b < RADIX holds bin b). scratch: 8 words.
template <int THREADS>
__device__ __forceinline__ uint32_t bins_excl_scan(uint32_t v, uint32_t *scratch, uint32_t &total) {
  static_assert(THREADS >= RADIX, "one thread per bin");
  const unsigned w = threadIdx.x >> 5;
  uint32_t incl = warp_incl_scan(v);
  if (threadIdx.x < RADIX && lane_id() == 31) scratch[w] = incl;
  __syncthreads();
  uint32_t add = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < RADIX / 32; ++i) {
    uint32_t s = scratch[i];
    if (i < (int)w) add += s;
    tot += s;
  }
  __syncthreads();
  total = tot;
  return incl - v + add;
}

// Tuning knobs of one pass (chosen per launch by the host, see api.cu).
//   POS_ASYNC (record source only): cp.async the tile's 32-bit positions into shared memory at tile
//   start, so they cost no registers and their latency hides behind the ranking (needs a 16-byte
//   aligned pos array); otherwise they are loaded after the ranking.
// Measured and dropped on B200 (profiles/r01_sort_pass_tuning.md): ballot match on the 8 bin bits
// (4x the instructions), __match_any_sync (20 % slower than ballots), several items per match round trip.
//   CHAINS: a warp's ITEMS are ranked as CHAINS independent sub-chunks (own bitmap + count rows),
//   interleaved in the instruction stream, so the atomicOr -> read -> clear round trips of different
//   sub-chunks overlap (the kernel is latency-bound on that chain, profiles/r01_sort_pass_tuning.md).
template <int THREADS_, int ITEMS_, int MINBLOCKS_, bool POS_ASYNC_, int CHAINS_ = 1>
struct PassCfg {
  static constexpr int THREADS = THREADS_, ITEMS = ITEMS_, MINBLOCKS = MINBLOCKS_, CHAINS = CHAINS_;
  static_assert(ITEMS_ % CHAINS_ == 0, "items split evenly over chains");
  static constexpr bool POS_ASYNC = POS_ASYNC_;
  static constexpr int TILE = THREADS * ITEMS;
};

template <class Cfg, bool FROM_SEQ>
struct PassSmem {
  static constexpr int TILE = Cfg::TILE;
  static constexpr int WARPS = Cfg::THREADS / 32;
  static constexpr int VW = WARPS * Cfg::CHAINS;   // ranked sub-chunks ("virtual warps"), in memory order
  using PosT = typename std::conditional<FROM_SEQ, uint16_t, uint32_t>::type;
  uint64_t keys[TILE];
  PosT pos[TILE];
  uint32_t whist[VW][RADIX];      // per-sub-chunk bin counts, later its offset inside the bin
  uint32_t match[VW][RADIX];      // lane bitmaps of the item being matched (self-clearing)
  int32_t goff[RADIX];            // global index of the bin's first record of this tile - its tile slot
  uint32_t start[RADIX];
  uint32_t next[RADIX];
  uint32_t scratch[8];
  uint32_t tile;
  TileCodes<FROM_SEQ ? TILE : 16> tc;
};

// One tile.  FULL: every slot of the tile holds a valid record (no predicates on the hot path).
template <class Cfg, bool FROM_SEQ, bool FULL, class BinFn, class NextFn, bool HAS_NEXT>
__device__ __forceinline__ void pass_tile(const PassParams<BinFn, NextFn> &P, PassSmem<Cfg, FROM_SEQ> &sm,
                                          const uint32_t tile, const int64_t q0, const int64_t n_in,
                                          const uint32_t gbase, const bool special) {
  using S = PassSmem<Cfg, FROM_SEQ>;
  constexpr int TILE = Cfg::TILE, THREADS = Cfg::THREADS, ITEMS = Cfg::ITEMS;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const int t0 = warp * (32 * ITEMS) + lane;
#define KMG_STAMP(slot) do { if (P.trace && tid == 0) P.trace[(size_t)tile * 8 + (slot)] = clock64(); } while (0)
  KMG_STAMP(1);

  // ---- load ITEMS records per thread, warp-striped (item i of lane l = element i*32+l of the
  //      warp's chunk), which is memory order => the ranks below are stable
  uint64_t key[ITEMS];
  uint32_t val[ITEMS];
  uint32_t valid = FULL ? 0xFFFFFFFFu : 0u;
  if constexpr (FROM_SEQ) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int t = t0 + i * 32;
      key[i] = tile_key<TILE>(sm.tc, t, P.sv.k);
      if constexpr (!FULL)
        if (tile_valid<TILE>(P.sv, sm.tc, q0, t, special)) valid |= 1u << i;
    }
  } else {
    if constexpr (Cfg::POS_ASYNC) {
      // stream the tile's positions into shared memory (tile order) while the keys are ranked
      const uint32_t *src = P.pos_in + q0;
      const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(&sm.pos[0]);
#pragma unroll
      for (int c = 0; c < TILE / 4 / THREADS; ++c) {
        const int e = (c * THREADS + tid) * 4;
        if constexpr (FULL) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + e * 4), "l"(src + e) : "memory");
        } else {
          const int64_t left = n_in - q0 - e;
          if (left > 0) {
            const uint32_t bytes = (uint32_t)min((int64_t)16, left * 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst0 + e * 4), "l"(src + e), "r"(bytes) : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const uint64_t *ksrc = P.keys_in + q0 + t0;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if constexpr (FULL) key[i] = ld_stream_u64(ksrc + i * 32);
      else {
        key[i] = 0;
        if (q0 + t0 + i * 32 < n_in) { key[i] = ld_stream_u64(ksrc + i * 32); valid |= 1u << i; }
      }
    }
  }

  // ---- rank inside the warp: match equal bins through the warp's bitmap table (atomicOr the lane
  //      bit, read the bitmap back, lowest peer clears it and bumps the warp's running bin count).
  //      rk = bin << 16 | rank among the warp's earlier records of that bin.
  uint32_t rk[ITEMS];
  const unsigned lt = lanemask_lt();
  const uint32_t lanebit = 1u << lane;
  if (P.trace && tid == 0) { P.trace[(size_t)tile * 8 + 2] = (unsigned long long)(key[0] & 1) + clock64(); }   // first key has arrived
  constexpr int CH = Cfg::CHAINS, PER = ITEMS / CH;    // item i belongs to sub-chunk i / PER
#pragma unroll
  for (int st = 0; st < PER; ++st) {
    uint32_t d[CH], peers[CH], old[CH];
    __syncwarp();                                        // previous step's clears are visible
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = c * PER + st;
      d[c] = P.bin(key[i]);
      if (FULL || ((valid >> i) & 1u)) atomicOr(&sm.match[warp * CH + c][d[c]], lanebit);
    }
    __syncwarp();
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = c * PER + st;
      peers[c] = (FULL || ((valid >> i) & 1u)) ? sm.match[warp * CH + c][d[c]] : 0u;
    }
    __syncwarp();                                        // everyone has read before the clears
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = c * PER + st;
      old[c] = 0;
      // a warp's shared-memory atomics are applied in issue order, so step st sees steps < st
      if ((FULL || ((valid >> i) & 1u)) && (peers[c] & lt) == 0) {
        sm.match[warp * CH + c][d[c]] = 0;
        old[c] = atomicAdd(&sm.whist[warp * CH + c][d[c]], (uint32_t)__popc(peers[c]));
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int i = c * PER + st;
      const uint32_t base = __shfl_sync(FULL_MASK_, old[c], FULL ? (__ffs(peers[c]) - 1) : (peers[c] ? __ffs(peers[c]) - 1 : 0));
      rk[i] = (d[c] << 16) | (base + __popc(peers[c] & lt));
    }
  }
  KMG_STAMP(3);                                          // this warp's ranking done
  if constexpr (!FROM_SEQ && Cfg::POS_ASYNC) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  KMG_STAMP(4);
  if constexpr (!FROM_SEQ && Cfg::POS_ASYNC) {           // every thread's copies have landed
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) val[i] = sm.pos[t0 + i * 32];
  }

  // ---- per-bin totals of the tile; publish them, then place the tile's bins in shared memory
  uint32_t cnt = 0;
  if (tid < RADIX) {
#pragma unroll
    for (int w = 0; w < S::VW; ++w) {
      uint32_t c = sm.whist[w][tid];
      sm.whist[w][tid] = cnt;                 // becomes the sub-chunk's offset inside the bin
      cnt += c;
    }
    st_relaxed_u64(P.status + (size_t)tile * RADIX + tid, st_pack(tile == 0 ? ST_INCL : ST_AGG, P.epoch, cnt));
  }
  uint32_t tile_count;
  const uint32_t lstart = bins_excl_scan<THREADS>(cnt, sm.scratch, tile_count);
  if (tid < RADIX) sm.start[tid] = lstart;
  __syncthreads();

#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {                     // rk becomes the slot in the regrouped tile
    const uint32_t d = rk[i] >> 16;
    rk[i] = sm.start[d] + sm.whist[warp * CH + i / PER][d] + (rk[i] & 0xFFFFu);
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    if (FULL || ((valid >> i) & 1u)) sm.keys[rk[i]] = key[i];
  if constexpr (FROM_SEQ) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
      if (FULL || ((valid >> i) & 1u)) sm.pos[rk[i]] = (uint16_t)(t0 + i * 32);
  } else {
    if constexpr (!Cfg::POS_ASYNC) {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i)
        val[i] = (FULL || ((valid >> i) & 1u)) ? ld_stream_u32(P.pos_in + q0 + t0 + i * 32) : 0;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
      if (FULL || ((valid >> i) & 1u)) sm.pos[rk[i]] = val[i];
  }

  KMG_STAMP(5);                                          // regrouped in shared memory
  // ---- look back over earlier tiles (one bin per thread, four status words per round trip);
  //      predecessors have had the whole regrouping above to publish
  if (tid < RADIX) {
    uint64_t excl = 0;
    if (tile > 0 && !(P.dbg & 1u)) {
      int64_t t = (int64_t)tile - 1;
      bool done = false;
      while (!done) {
        uint64_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          w[j] = t - j >= 0 ? ld_relaxed_u64(P.status + (size_t)(t - j) * RADIX + tid) : st_pack(ST_INCL, P.epoch, 0);
        if (st_flag(w[0], P.epoch) == 0) { __nanosleep(40); continue; }   // not published yet: back off, poll again
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (done) break;
          const uint64_t f = st_flag(w[j], P.epoch);
          if (f == 0) break;
          excl += st_value(w[j]);
          --t;
          if (f == ST_INCL) done = true;
        }
      }
      st_relaxed_u64(P.status + (size_t)tile * RADIX + tid, st_pack(ST_INCL, P.epoch, excl + cnt));
    }
    sm.goff[tid] = (int32_t)((int64_t)gbase + (int64_t)excl - (int64_t)lstart);
  }
  __syncthreads();
  KMG_STAMP(6);                                          // look-back finished for all bins

  // ---- stream the regrouped tile out: consecutive threads -> consecutive slots -> runs per bin
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int s = j * THREADS + tid;
    if (FULL || s < (int)tile_count) {
      const uint64_t kk = sm.keys[s];
      const int dst = sm.goff[P.bin(kk)] + s;
      if (!(P.dbg & 2u)) {
        P.keys_out[dst] = kk;
        if constexpr (FROM_SEQ) P.pos_out[dst] = (uint32_t)(P.sv.s0 + q0 + 1) + sm.pos[s];   // 1-based start
        else P.pos_out[dst] = sm.pos[s];
      }
      if (HAS_NEXT && !(P.dbg & 4u)) atomicAdd(&sm.next[P.next(kk)], 1u);
    }
  }
  KMG_STAMP(7);                                          // stores issued
#undef KMG_STAMP
}

template <class Cfg, bool FROM_SEQ, class BinFn, class NextFn, bool HAS_NEXT>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINBLOCKS)
scatter_pass_kernel(const PassParams<BinFn, NextFn> P) {
  using S = PassSmem<Cfg, FROM_SEQ>;
  constexpr int TILE = Cfg::TILE, THREADS = Cfg::THREADS;
  static_assert(TILE <= 65536, "tile-local positions are 16 bit");
  static_assert(TILE % (4 * THREADS) == 0, "cp.async chunks per thread");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S &sm = *reinterpret_cast<S *>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

  const long long t_start = clock64();
  if (tid == 0) sm.tile = atomicAdd(P.ticket, 1u);
  {
    uint4 *zw = reinterpret_cast<uint4 *>(sm.whist[warp * Cfg::CHAINS]), *zm = reinterpret_cast<uint4 *>(sm.match[warp * Cfg::CHAINS]);
#pragma unroll
    for (int j = 0; j < Cfg::CHAINS * RADIX / 4 / 32; ++j) { zw[j * 32 + lane] = make_uint4(0, 0, 0, 0); zm[j * 32 + lane] = make_uint4(0, 0, 0, 0); }
  }
  if (tid < RADIX) sm.next[tid] = 0;

  // global exclusive base of every bin for this pass (the histogram is L2-resident, 1 KB)
  uint32_t gcount = tid < RADIX ? P.hist_cur[tid] : 0;
  const int64_t n_in = FROM_SEQ ? P.sv.nstarts : (int64_t)*P.n_records;
  uint32_t n_total;
  const uint32_t gbase = bins_excl_scan<THREADS>(gcount, sm.scratch, n_total);   // syncs: sm.tile visible
  const uint32_t tile = sm.tile;
  const int64_t q0 = (int64_t)tile * TILE;
  if (q0 >= n_in) return;
  if (P.trace && tid == 0) P.trace[(size_t)tile * 8] = (unsigned long long)t_start;

  bool special = false;
  if constexpr (FROM_SEQ) special = tile_pack<TILE, THREADS>(P.sv, q0, sm.tc);
  const bool touches_end = FROM_SEQ && (P.sv.s0 + q0 + TILE + P.sv.k > P.sv.L);
  if (q0 + TILE <= n_in && !special && !touches_end)
    pass_tile<Cfg, FROM_SEQ, true, BinFn, NextFn, HAS_NEXT>(P, sm, tile, q0, n_in, gbase, special);
  else
    pass_tile<Cfg, FROM_SEQ, false, BinFn, NextFn, HAS_NEXT>(P, sm, tile, q0, n_in, gbase, special);

  if (HAS_NEXT) {
    __syncthreads();
    if (tid < RADIX) {
      uint32_t c = sm.next[tid];
      if (c) atomicAdd(P.hist_next + tid, c);
    }
  }
}

// ---- histogram of the first pass's bins -------------------------------------------------------------
// Sequence source: a persistent grid walks the tiles, encoding on the fly (reads L bytes).
template <int THREADS, int ITEMS, class BinFn>
__global__ void __launch_bounds__(THREADS)
hist_seq_kernel(const SeqView sv, uint32_t *hist, BinFn bin) {
  constexpr int TILE = THREADS * ITEMS;
  __shared__ TileCodes<TILE> tc;
  __shared__ uint32_t sh[RADIX];
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  for (int b = tid; b < RADIX; b += THREADS) sh[b] = 0;
  const int64_t tiles = ceil_div<int64_t>(sv.nstarts, TILE);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t q0 = tile * TILE;
    __syncthreads();                                   // previous tile's readers are done
    const bool special = tile_pack<TILE, THREADS>(sv, q0, tc);
    const int t0 = warp * (32 * ITEMS) + lane;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int t = t0 + i * 32;
      if (tile_valid<TILE>(sv, tc, q0, t, special)) atomicAdd(&sh[bin(tile_key<TILE>(tc, t, sv.k))], 1u);
    }
  }
  __syncthreads();
  for (int b = tid; b < RADIX; b += THREADS) {
    uint32_t c = sh[b];
    if (c) atomicAdd(hist + b, c);
  }
}

// Record source.
template <int THREADS, class BinFn>
__global__ void __launch_bounds__(THREADS)
hist_rec_kernel(const uint64_t *keys, int64_t n, uint32_t *hist, BinFn bin) {
  __shared__ uint32_t sh[RADIX];
  for (int b = threadIdx.x; b < RADIX; b += THREADS) sh[b] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS)
    atomicAdd(&sh[bin(ld_stream_u64(keys + i))], 1u);
  __syncthreads();
  for (int b = threadIdx.x; b < RADIX; b += THREADS) {
    uint32_t c = sh[b];
    if (c) atomicAdd(hist + b, c);
  }
}

}  // namespace kmg
