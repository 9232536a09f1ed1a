// sort.cuh -- the index build's core: a stable single-sweep ("onesweep") binning pass.
//
// Replaces kh_put + kv_push (src/kmer_pos.c:36-50, src/khash.h:307-348, src/kvec.h:74-80): instead
// of one random DRAM probe per k-mer, (key,pos) records are ordered by a stable LSD radix sort,
// 8 bits per pass.  Stability keeps each k-mer's positions ascending (what insertion order gives
// the reference, README "positions are sorted").
//
// One kernel does a whole pass:
//   - records come either from HBM arrays (a warp's keys by one 1-D bulk copy -- cp.async.bulk + mbarrier -- into the
//     shared-memory array that later holds the regrouped tile) or straight from the ASCII sequence (the 2-bit encoder
//     of windows.cuh is fused into the first pass: keys are never written unsorted, and windows
//     that contain an N are dropped by simply not being ranked);
//   - the rank of a record inside the tile comes from one shared-memory atomic per record on its warp's
//     bin counter (where the device applies colliding lanes in lane order, checked once; otherwise
//     from a per-warp bitmap match), so the pass is insensitive to skew (homopolymers, microsatellites);
//   - the tile is regrouped by bin in shared memory and written out in runs;
//   - tile offsets chain through a decoupled look-back, one status word per (tile, bin).
// Bin bases come from histograms that are complete before the pass starts: for a sorted build from the
// sequence all passes' histograms are taken up front by one kernel (hist_all_kernel); for the grouped
// build (digits of mix64(key)) and for builds from records (sharded build) the histogram of the NEXT
// pass's digit is taken while the regrouped keys stream out (HAS_NEXT).
// The same kernel with OwnerBin (key-range owner instead of digit) is the multi-GPU partitioner; with PEER
// it writes every record straight into its owner's arrays over NVLink, owner by owner in 32-record-aligned blocks
// (whole 128-byte lines per remote store instruction).
#pragma once
#include "common.cuh"
#include "lookback.cuh"
#include "windows.cuh"

namespace kmg {

struct DigitBin {
  static constexpr bool CHEAP = true;    // two instructions: recomputed wherever the bin is needed
  int shift;
  uint32_t mask = RADIX - 1;             // bins - 1 of the pass (8-, 9- or 10-bit digits)
  __device__ __forceinline__ uint32_t operator()(uint64_t key) const { return (uint32_t)(key >> shift) & mask; }
};
// owner r holds keys in [spl[r-1], spl[r]): bin = number of splitters <= key
struct OwnerBin {
  static constexpr bool CHEAP = false;   // a binary search: computed once per record and kept (packed bytes)
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
// owner of the mixed key (grouped sharded build: owners hold ranges of mix64(key))
struct HashOwnerBin {
  static constexpr bool CHEAP = false;
  OwnerBin ob;
  __device__ __forceinline__ uint32_t operator()(uint64_t key) const { return ob(mix64(key)); }
};
// digit of the mixed key (grouped build: histogram of pass 0 taken straight from the sequence)
struct HashDigitBin {
  static constexpr bool CHEAP = true;
  int shift;
  uint32_t mask = RADIX - 1;
  __device__ __forceinline__ uint32_t operator()(uint64_t key) const { return (uint32_t)(mix64(key) >> shift) & mask; }
};
// owner of a record of the grouped sharded build: G equal ranges of the mixed key (records carry mix64(key), which is
// uniform whatever the sequence's composition: no sampling, no splitters, balanced by construction)
struct RangeBin {
  static constexpr bool CHEAP = true;
  uint32_t nparts;
  __device__ __forceinline__ uint32_t operator()(uint64_t h) const { return (uint32_t)__umul64hi(h, (uint64_t)nparts); }
};
struct NoBin {
  static constexpr bool CHEAP = true;
  __device__ __forceinline__ uint32_t operator()(uint64_t) const { return 0; }
};

// PEER mode (sharded build): bin b's records go to arrays of the GPU that owns key range b, which may
// be peer memory mapped over NVLink; the pass writes them there directly (no staging, no all-to-all).
constexpr int MAX_PEERS = 16;
struct PeerTable {
  uint64_t *keys[MAX_PEERS];
  uint32_t *pos[MAX_PEERS];
  int64_t delta[MAX_PEERS];   // (this rank's first slot in owner b's arrays) - (first index of bin b in this rank's stream)
  uint64_t cap;               // records each destination array can hold (region mode: each source's region of it)
  // region mode (kmg_shard_scatter_ranges): every source has its own fixed region of every owner's arrays, so nothing has
  // to be counted or exchanged before the scatter; the last tile tells owner b how many records it got from this rank
  uint64_t *counts[MAX_PEERS];  // owner b's received-count array (indexed by source rank), or all null
  uint32_t rank, region;
};
// A record source (or destination) made of `nsegs` segments of `stride` slots each, segment r holding count[r] records from
// its start: what an owner receives from a region-mode scatter (segment = source rank), and what the first pass of a grouped
// build leaves when it writes every bin into its own over-provisioned region instead of needing the bins' sizes beforehand
// (segment = bin).  The pass that reads it walks the segments tile by tile: tile_prefix[r] = tiles (of `tile` records) before
// segment r, so a ticket maps to (segment, tile inside it) by one search, no tile straddles two segments, and every tile but
// a segment's last takes the predicate-free path.
constexpr int MAX_SEGS = 512;
struct TileSegs {
  uint32_t tile_prefix[MAX_SEGS + 1];
  uint32_t count[MAX_SEGS];
  uint64_t stride, total;       // slots per segment; records in all segments
  uint32_t nsegs, tile;         // tile = records per tile of the reading pass
  uint32_t overflow, pad;       // overflow: some segment was sent more than it holds (the excess was dropped)
};
template <class BinFn, class NextFn>
struct PassParams {
  SeqView sv;                 // FROM_SEQ source
  const uint64_t *keys_in;    // record source
  const uint32_t *pos_in;
  uint64_t *keys_out;
  uint32_t *pos_out;
  const uint32_t *gbase;      // [RADIX] exclusive scan of this pass's (complete) bin histogram
  uint32_t *hist_next;        // [RADIX] accumulates the next pass's histogram (HAS_NEXT), or nullptr
  uint64_t *status;           // [tiles][RADIX] look-back words
  uint32_t *ticket;           // tile id dispenser (zero before launch)
  const uint64_t *n_records;  // record source: exact number of records (device scalar)
  uint32_t epoch;
  uint32_t hashed;            // FROM_SEQ: records carry mix64(key) instead of the key (grouped build)
  uint32_t pos_add;           // FROM_SEQ: added to the 1-based start (k-1 turns it into the 1-based end of a query window)
  const PeerTable *peer;      // PEER: per-bin destination arrays (own or NVLink-mapped peer memory)
  const TileSegs *segs;       // segmented record source (see TileSegs), or nullptr
  uint64_t *bin_counts;       // FROM_SEQ, bins written into regions of bin_cap slots (gbase[b] = b * bin_cap): the last tile
  uint32_t bin_cap;           //   stores every bin's total here; a bin that outgrows its region sets *bin_overflow
  uint32_t *bin_overflow;
  unsigned long long *trace;  // tuning runs only: 8 clock64 stamps per tile, or nullptr
  uint32_t dbg;               // tuning runs only (wrong results): 1 no look-back wait, 2 no global stores
  BinFn bin;
  NextFn next;
};

// Tuning knobs of one pass (chosen per launch by the host, see api.cu).
//   RANK 3: ONE shared-memory atomicAdd per record on the warp's bin counter; its return value is the rank.
//           Needs colliding lanes of one instruction to be applied in ascending lane order, which the library
//           verifies per device before using it (lane_order_selftest_kernel) and re-verifies on every index it
//           builds (rle_kernel checks that positions ascend inside every k-mer); the default when that holds;
//   RANK 0: peers through a per-warp bitmap table (atomicOr the lane bit, read back, leader clears): assumes
//           nothing; the default otherwise.
//   LB    : status words fetched per look-back round trip.
//   RB    : bits of the digit (bins = 2^RB): 8, 9 or 10.  More bits = fewer passes over the records.
// Measured and dropped on B200 (profiles/r01_sort_pass_tuning.md): ranking by 8 ballots, __match_any_sync,
// several items per bitmap round trip, two interleaved rank chains, cp.async of the positions, persistent CTAs
// that claim the next ticket early and prefetch its records into L2.
template <int THREADS_, int ITEMS_, int MINBLOCKS_, int RANK_ = 3, int LB_ = 4, int RB_ = RADIX_BITS>
struct PassCfg {
  static constexpr int THREADS = THREADS_, ITEMS = ITEMS_, MINBLOCKS = MINBLOCKS_, RANK = RANK_, LB = LB_, RB = RB_;
  static constexpr int TILE = THREADS * ITEMS;
  static constexpr int NB = 1 << RB_;                                    // bins
  static constexpr int BPT = NB >= THREADS_ ? NB / THREADS_ : 1;         // consecutive bins owned by a thread
  static_assert(RANK_ == 0 || RANK_ == 3 || RANK_ == 4, "rank variants: 0 bitmap, 3 one atomic, 4 = 3 made unstable on purpose (tests)");
  static_assert(NB < THREADS_ || NB % THREADS_ == 0, "bins per thread");
  static_assert(BPT == 1 || BPT == 2 || BPT == 4, "a thread's bins are moved as one vector");
};

template <class Cfg, bool FROM_SEQ>
struct PassSmem {
  static constexpr int TILE = Cfg::TILE;
  static constexpr int WARPS = Cfg::THREADS / 32;
  static constexpr int NB = Cfg::NB;
  using PosT = typename std::conditional<FROM_SEQ, uint16_t, uint32_t>::type;
  uint64_t keys[TILE];
  PosT pos[TILE];
  alignas(16) uint32_t whist[WARPS][NB];                 // per-warp bin counts, later the warp's first slot of the bin
  uint32_t match[Cfg::RANK == 0 ? WARPS * NB : 1];       // RANK 0: lane bitmaps of the item being matched (self-clearing)
  int32_t goff[NB];                                      // global index of the bin's first record of this tile - its tile slot
  uint32_t next[NB];
  uint32_t scratch[32];
  uint32_t tile;
  uint32_t seg_tp[MAX_SEGS + 1];                         // segmented record source only: TileSegs::tile_prefix
  uint32_t skip_write;                                   // region output: a bin outgrew its region, drop this tile's records
  alignas(8) uint64_t bar[WARPS];                        // record source: one mbarrier per warp for the bulk copy of its keys
  alignas(8) uint64_t bar2[WARPS];                       //   and one for its positions
  uint32_t bstart[MAX_PEERS], bcnt[MAX_PEERS];           // PEER mode only: first tile slot and size of every owner's run
  PeerTable peer;                                        // PEER mode only
  TileCodes<FROM_SEQ ? TILE : 16> tc;
};

// Do the shared-memory atomics of one warp take effect in (instruction, lane) order, i.e. do colliding lanes of
// one instruction get their old values in ascending lane order and do back-to-back instructions stay in issue
// order?  (RANK 3 needs it.)  Several collision patterns, many repetitions, every SM; *bad counts failures.
__global__ void lane_order_selftest_kernel(uint32_t *bad) {
  // As in the pass: STEPS atomics per lane issued back to back (no result consumed in between) on the warp's
  // own table; then lane 0 replays them in (step, lane) order and compares every returned value.
  constexpr int STEPS = 8;
  __shared__ uint32_t tab[8][RADIX];
  __shared__ uint32_t sim[8][RADIX];
  __shared__ uint8_t dig[8][STEPS][32];
  __shared__ uint32_t got[8][STEPS][32];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t fails = 0;
  for (int rep = 0; rep < 32; ++rep) {
    for (int pattern = 0; pattern < 6; ++pattern) {
      for (int b = lane; b < RADIX; b += 32) { tab[warp][b] = 0; sim[warp][b] = 0; }
      __syncwarp();
      uint32_t d[STEPS], g[STEPS];
#pragma unroll
      for (int j = 0; j < STEPS; ++j) {
        switch (pattern) {
          case 0: d[j] = 7; break;                                                  // all lanes, all steps one address
          case 1: d[j] = (lane + j) & 1; break;                                     // two addresses, interleaved
          case 2: d[j] = (lane >> 3) + (j & 1); break;                              // runs
          case 3: d[j] = ((lane + 7 * j) * 2654435761u + rep * 40503u) >> 29; break;   // 8 addresses, scattered
          case 4: d[j] = ((lane + 32 * j) * 2246822519u + rep * 97u) & 255u; break;    // mostly distinct
          default: d[j] = (lane % 3 == 0) ? 200 : ((lane * 5 + j) & 255u); break;
        }
      }
#pragma unroll
      for (int j = 0; j < STEPS; ++j) g[j] = atomicAdd(&tab[warp][d[j]], 1u);
#pragma unroll
      for (int j = 0; j < STEPS; ++j) { dig[warp][j][lane] = (uint8_t)d[j]; got[warp][j][lane] = g[j]; }
      __syncwarp();
      if (lane == 0) {
        for (int j = 0; j < STEPS; ++j)
          for (int l = 0; l < 32; ++l)
            if (got[warp][j][l] != sim[warp][dig[warp][j][l]]++) ++fails;
      }
      __syncwarp();
    }
  }
  if (fails) atomicAdd(bad, fails);
}

// BPT consecutive 32-bit words as one access
template <int BPT> struct WordVec;
template <> struct WordVec<1> { using T = uint32_t; };
template <> struct WordVec<2> { using T = uint2; };
template <> struct WordVec<4> { using T = uint4; };
template <int BPT>
__device__ __forceinline__ void load_words(const uint32_t *p, uint32_t (&v)[BPT]) {
  const typename WordVec<BPT>::T t = *reinterpret_cast<const typename WordVec<BPT>::T *>(p);
  if constexpr (BPT == 1) v[0] = t;
  else if constexpr (BPT == 2) { v[0] = t.x; v[1] = t.y; }
  else { v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
}
template <int BPT>
__device__ __forceinline__ void store_words(uint32_t *p, const uint32_t (&v)[BPT]) {
  typename WordVec<BPT>::T t;
  if constexpr (BPT == 1) t = v[0];
  else if constexpr (BPT == 2) { t.x = v[0]; t.y = v[1]; }
  else { t.x = v[0]; t.y = v[1]; t.z = v[2]; t.w = v[3]; }
  *reinterpret_cast<typename WordVec<BPT>::T *>(p) = t;
}
// a thread's BPT status words (consecutive bins) of one tile: one 8- or 16-byte access each way where possible
template <int BPT>
__device__ __forceinline__ void ld_status(const uint64_t *p, uint64_t (&w)[BPT]) {
  if constexpr (BPT == 1) w[0] = ld_relaxed_u64(p);
  else {
#pragma unroll
    for (int q = 0; q < BPT; q += 2)
      asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(w[q]), "=l"(w[q + 1]) : "l"(p + q) : "memory");
  }
}
template <int BPT>
__device__ __forceinline__ void st_status(uint64_t *p, const uint64_t (&w)[BPT]) {
  if constexpr (BPT == 1) st_relaxed_u64(p, w[0]);
  else {
#pragma unroll
    for (int q = 0; q < BPT; q += 2)
      asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p + q), "l"(w[q]), "l"(w[q + 1]) : "memory");
  }
}

// ---- bulk copy (TMA, 1-D) of a warp's keys into shared memory ---------------------------------------------------
// A warp's ITEMS x 32 keys are one contiguous, 16-byte aligned piece of the record array: lane 0 arms the warp's mbarrier
// with the byte count and issues ONE cp.async.bulk; the copy engine moves the data without occupying registers, load/store
// slots or L1 lines while it is in flight.  The lanes then read their keys out of shared memory.
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load_arm(uint64_t *bar, void *dst, const void *src, uint32_t bytes) {
  const uint32_t b = smem_addr(bar), d = smem_addr(dst);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the initialised barrier is visible to the copy engine
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(d), "l"(src), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void bulk_load_wait(uint64_t *bar) {
  const uint32_t b = smem_addr(bar);
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "KMG_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@p bra KMG_DONE;\n\t"
      "bra KMG_WAIT;\n\t"
      "KMG_DONE:\n\t}" ::"r"(b) : "memory");
}

// One tile.  FULL: every slot of the tile holds a valid record (no predicates on the hot path).
template <class Cfg, bool FROM_SEQ, bool FULL, class BinFn, class NextFn, bool HAS_NEXT, bool PEER>
__device__ __forceinline__ void pass_tile(const PassParams<BinFn, NextFn> &P, PassSmem<Cfg, FROM_SEQ> &sm,
                                          const uint32_t tile, const int64_t q0, const int64_t n_in,
                                          const uint32_t (&gbase)[Cfg::BPT], const bool special) {
  using S = PassSmem<Cfg, FROM_SEQ>;
  constexpr int TILE = Cfg::TILE, THREADS = Cfg::THREADS, ITEMS = Cfg::ITEMS, WARPS = S::WARPS, NB = Cfg::NB, BPT = Cfg::BPT;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const int t0 = warp * (32 * ITEMS) + lane;
  const bool owns = NB >= THREADS || tid < NB;           // this thread owns bins [b0, b0 + BPT)
  const int b0 = tid * BPT;
#define KMG_STAMP(slot) do { if (P.trace && tid == 0) P.trace[(size_t)tile * 8 + (slot)] = clock64(); } while (0)
  KMG_STAMP(1);

  // ---- load ITEMS records per thread, warp-striped (item i of lane l = element i*32+l of the
  //      warp's chunk), which is memory order => the ranks below are stable
  uint64_t key[ITEMS];
  uint32_t valid = FULL ? 0xFFFFFFFFu : 0u;
  bool bulk_pos = false;                                 // record source: the positions come by bulk copy too
  if constexpr (FROM_SEQ) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int t = t0 + i * 32;
      key[i] = tile_key<TILE>(sm.tc, t, P.sv.k);
      if (P.hashed) key[i] = mix64(key[i]);
      if constexpr (!FULL)
        if (tile_valid<TILE>(P.sv, sm.tc, q0, t, special)) valid |= 1u << i;
    }
  } else {
    const uint64_t *ksrc = P.keys_in + q0 + t0;             // q0, n_in: physical indices (inside one segment of a segmented source)
    bool bulk = false;
    // (a segmented source's tiles start at segment base + k x TILE: the base, and so the address, may be only 8-byte aligned)
    if constexpr (FULL && !PEER) bulk = !(P.dbg & 8u) && (((uintptr_t)(P.keys_in + q0) & 15u) == 0);   // dbg 8 (tuning runs): register loads
    if constexpr (FULL && !PEER) bulk_pos = bulk && !(P.dbg & 16u) && (((uintptr_t)(P.pos_in + q0) & 15u) == 0);   // dbg 16: positions by register loads
    if (bulk) {
      if constexpr (FULL && !PEER) {
        uint64_t *stage = sm.keys + warp * (32 * ITEMS);
        if (lane == 0) {
          bulk_load_arm(&sm.bar[warp], stage, P.keys_in + q0 + warp * (32 * ITEMS), 32 * ITEMS * 8);
          if constexpr (!FROM_SEQ)                         // the positions travel meanwhile; they are needed after the ranking
            if (bulk_pos) bulk_load_arm(&sm.bar2[warp], sm.pos + warp * (32 * ITEMS), P.pos_in + q0 + warp * (32 * ITEMS), 32 * ITEMS * 4);
        }
        __syncwarp();
        bulk_load_wait(&sm.bar[warp]);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) key[i] = stage[i * 32 + lane];
      }
    } else
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      if constexpr (FULL) key[i] = ld_stream_u64(ksrc + i * 32);
      else {
        key[i] = 0;
        const int64_t q = q0 + t0 + i * 32;
        if (q < n_in) { key[i] = ld_stream_u64(ksrc + i * 32); valid |= 1u << i; }
      }
    }
  }

  // ---- rank inside the warp.  rk = bin << 16 | rank among the warp's earlier records of that bin.
  //      A warp's shared-memory atomics are applied in issue order, so step i sees steps < i.
  uint32_t rk[ITEMS];
  const unsigned lt = lanemask_lt();
  if (P.trace && tid == 0) { P.trace[(size_t)tile * 8 + 2] = (unsigned long long)(key[0] & 1) + clock64(); }   // first key has arrived
  uint32_t dpk[BinFn::CHEAP ? 1 : (ITEMS + 3) / 4] = {};   // expensive bins: the record's bin, one byte each
  if constexpr (!BinFn::CHEAP) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) dpk[i >> 2] |= P.bin(key[i]) << (8 * (i & 3));
  }
#define KMG_BIN(i) (BinFn::CHEAP ? P.bin(key[i]) : ((dpk[(i) >> 2] >> (8 * ((i) & 3))) & 0xFFu))
  if constexpr (Cfg::RANK >= 3) {
    // One shared-memory atomic per record: the value it returns is the record's rank among the warp's
    // earlier records of the bin, PROVIDED the hardware applies the lanes of one instruction that hit the
    // same address in ascending lane order.  PTX does not promise that; the library checks it once per
    // device (lane_order_selftest_kernel), never selects this variant unless the check passed, and checks every
    // finished index (rle_kernel: positions ascend inside each k-mer), rebuilding with the bitmap variant if not.
    // RANK 4 (tests only): the same with the items visited in reverse, which breaks stability on purpose.
#pragma unroll
    for (int ii = 0; ii < ITEMS; ++ii) {
      const int i = Cfg::RANK == 4 ? ITEMS - 1 - ii : ii;
      const uint32_t d = KMG_BIN(i);
      const bool ok = FULL || ((valid >> i) & 1u);
      uint32_t r = 0;
      if (ok) r = atomicAdd(&sm.whist[warp][d], 1u);
      rk[i] = (d << 16) | r;
      if constexpr (!FULL) __syncwarp();                 // keep the steps in order when the guard diverges
    }
  } else {
    const uint32_t lanebit = 1u << lane;
    uint32_t *mrow = sm.match + warp * NB;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const uint32_t d = KMG_BIN(i);
      const bool ok = FULL || ((valid >> i) & 1u);
      __syncwarp();                                        // previous step's clears are visible
      if (ok) atomicOr(&mrow[d], lanebit);
      __syncwarp();
      const uint32_t peers = ok ? mrow[d] : 0u;
      __syncwarp();                                        // everyone has read before the clears
      uint32_t old = 0;
      if (ok && (peers & lt) == 0) {
        mrow[d] = 0;
        old = atomicAdd(&sm.whist[warp][d], (uint32_t)__popc(peers));
      }
      const uint32_t base = __shfl_sync(FULL_MASK_, old, (FULL || peers) ? (__ffs(peers) - 1) : 0);
      rk[i] = (d << 16) | (base + __popc(peers & lt));
    }
  }
#undef KMG_BIN
  KMG_STAMP(3);                                          // this warp's ranking done
  __syncthreads();
  KMG_STAMP(4);

  // ---- per-bin totals of the tile; publish them, then turn whist[w][bin] into the first slot
  //      of warp w's records of that bin in the regrouped tile
  uint32_t cnt[BPT], lstart[BPT], tile_count;
  {
    uint32_t c[WARPS][BPT], mine = 0;
#pragma unroll
    for (int q = 0; q < BPT; ++q) cnt[q] = 0;
    if (owns) {
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        load_words<BPT>(&sm.whist[w][b0], c[w]);
#pragma unroll
        for (int q = 0; q < BPT; ++q) cnt[q] += c[w][q];
      }
      uint64_t sw[BPT];
#pragma unroll
      for (int q = 0; q < BPT; ++q) { sw[q] = st_pack(tile == 0 ? ST_INCL : ST_AGG, P.epoch, cnt[q]); mine += cnt[q]; }
      st_status<BPT>(P.status + (size_t)tile * NB + b0, sw);
    }
    const uint32_t incl = warp_incl_scan(mine);
    if (lane == 31) sm.scratch[warp] = incl;
    __syncthreads();
    uint32_t add = 0, tot = 0;
#pragma unroll
    for (int j = 0; j < WARPS; ++j) {
      const uint32_t s = sm.scratch[j];
      if (j < (int)warp) add += s;
      tot += s;
    }
    tile_count = tot;
    uint32_t run = incl - mine + add;
#pragma unroll
    for (int q = 0; q < BPT; ++q) { lstart[q] = run; run += cnt[q]; }
    if (owns) {
      uint32_t r2[BPT];
#pragma unroll
      for (int q = 0; q < BPT; ++q) r2[q] = lstart[q];
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        store_words<BPT>(&sm.whist[w][b0], r2);
#pragma unroll
        for (int q = 0; q < BPT; ++q) r2[q] += c[w][q];
      }
    }
  }
  __syncthreads();

#pragma unroll
  for (int i = 0; i < ITEMS; ++i)                        // rk becomes the slot in the regrouped tile
    rk[i] = sm.whist[warp][rk[i] >> 16] + (rk[i] & 0xFFFFu);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    if (FULL || ((valid >> i) & 1u)) sm.keys[rk[i]] = key[i];
  if constexpr (FROM_SEQ) {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
      if (FULL || ((valid >> i) & 1u)) sm.pos[rk[i]] = (uint16_t)(t0 + i * 32);
  } else {
    uint32_t val[ITEMS];
    if (bulk_pos) {                                        // (block-uniform) staged linearly in sm.pos, the array they are regrouped in:
      bulk_load_wait(&sm.bar2[warp]);                      //   every thread reads its own first, then a barrier, then the scatter
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) val[i] = sm.pos[t0 + i * 32];
      __syncthreads();
    } else
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      val[i] = (FULL || ((valid >> i) & 1u)) ? ld_stream_u32(P.pos_in + q0 + t0 + i * 32) : 0;
    }
#pragma unroll
    for (int i = 0; i < ITEMS; ++i)
      if (FULL || ((valid >> i) & 1u)) sm.pos[rk[i]] = val[i];
  }

  KMG_STAMP(5);                                          // regrouped in shared memory
  // ---- look back over earlier tiles (a thread's BPT bins in lockstep, LB tiles per round trip);
  //      predecessors have had the whole regrouping above to publish
  if (owns) {
    uint64_t excl[BPT];
    bool done[BPT];
#pragma unroll
    for (int q = 0; q < BPT; ++q) { excl[q] = 0; done[q] = false; }
    if (tile > 0 && !(P.dbg & 1u)) {
      constexpr int LB = Cfg::LB;
      int64_t t = (int64_t)tile - 1;
      bool all_done = false;
      while (!all_done) {
        uint64_t w[LB][BPT];
#pragma unroll
        for (int j = 0; j < LB; ++j) {
          if (t - j >= 0) ld_status<BPT>(P.status + (size_t)(t - j) * NB + b0, w[j]);
          else {
#pragma unroll
            for (int q = 0; q < BPT; ++q) w[j][q] = st_pack(ST_INCL, P.epoch, 0);
          }
        }
        int adv = 0;                                       // tiles consumed this round: all pending bins advance together
        bool stop = false;
#pragma unroll
        for (int j = 0; j < LB; ++j) {
          if (stop) break;
          bool ready = true;
#pragma unroll
          for (int q = 0; q < BPT; ++q) ready &= done[q] || st_flag(w[j][q], P.epoch) != 0;
          if (!ready) { stop = true; break; }              // not published yet: poll this tile again
          bool pending = false;
#pragma unroll
          for (int q = 0; q < BPT; ++q) {
            if (!done[q]) {
              excl[q] += st_value(w[j][q]);
              if (st_flag(w[j][q], P.epoch) == ST_INCL) done[q] = true; else pending = true;
            }
          }
          ++adv;
          if (!pending) { all_done = true; stop = true; }
        }
        t -= adv;
        if (adv == 0) __nanosleep(40);                     // predecessor not published yet: back off
      }
      uint64_t sw[BPT];
#pragma unroll
      for (int q = 0; q < BPT; ++q) sw[q] = st_pack(ST_INCL, P.epoch, excl[q] + cnt[q]);
      st_status<BPT>(P.status + (size_t)tile * NB + b0, sw);
    }
#pragma unroll
    for (int q = 0; q < BPT; ++q) {
      const int64_t go = (int64_t)gbase[q] + (int64_t)excl[q] - (int64_t)lstart[q];   // index in this rank's stream of bin b - tile slot
      sm.goff[b0 + q] = (int32_t)go;
      if constexpr (FROM_SEQ && !PEER) {                   // bins written into regions: totals from the last tile, overflow guard
        if (P.bin_counts) {
          if (excl[q] + cnt[q] > P.bin_cap) { *P.bin_overflow = 1; sm.skip_write = 1; }
          if (q0 + TILE >= n_in) P.bin_counts[b0 + q] = excl[q] + cnt[q];
        }
      }
      if constexpr (PEER) {                                // region mode: the last tile reports this rank's total to each owner
        if (sm.peer.region && q0 + TILE >= n_in && b0 + q < MAX_PEERS && sm.peer.counts[b0 + q])
          sm.peer.counts[b0 + q][sm.peer.rank] = excl[q] + cnt[q];
        if (b0 + q < MAX_PEERS) { sm.bstart[b0 + q] = lstart[q]; sm.bcnt[b0 + q] = cnt[q]; }
      }
    }
  }
  __syncthreads();
  KMG_STAMP(6);                                          // look-back finished for all bins

  // ---- stream the regrouped tile out: consecutive threads -> consecutive slots -> runs per bin
  bool drop = false;
  if constexpr (FROM_SEQ && !PEER) drop = sm.skip_write != 0;      // region output: a bin outgrew its region (the build is redone)
  bool aligned_out = false;
  if constexpr (PEER) aligned_out = !(P.dbg & 4u);                 // dbg 4 (tuning runs): the generic write-out
  if (aligned_out) {
   if constexpr (PEER) {
    // Owner by owner, every warp's 32 lanes covering one 32-record-ALIGNED block of the destination (256 bytes of keys, 128
    // of positions: whole 128-byte lines): a remote store instruction then becomes full-line NVLink write packets instead of
    // the 3 + 2 partial ones of an arbitrarily aligned run.  A handful of owners with long runs, so the per-owner loop and the
    // idle lanes at a run's ends cost little.
    constexpr int OWNERS = NB < MAX_PEERS ? NB : MAX_PEERS;
    for (int b = 0; b < OWNERS; ++b) {
      const int cb = (int)sm.bcnt[b];
      if (cb == 0) continue;
      const int st = (int)sm.bstart[b];
      const int64_t go = sm.goff[b], delta = sm.peer.delta[b];
      uint64_t *kout = sm.peer.keys[b];
      uint32_t *pout = sm.peer.pos[b];
      const int mis = (int)((go + st + delta) & 31);
      for (int s = st - mis + (int)tid; s < st + cb; s += THREADS) {
        if (s < st) continue;
        const int64_t local = go + s, dst = local + delta;
        const bool room = (uint64_t)(sm.peer.region ? local : dst) < sm.peer.cap;
        const uint64_t kk = sm.keys[s];
        if (room && !(P.dbg & 2u)) {
          kout[dst] = kk;
          if constexpr (FROM_SEQ) pout[dst] = (uint32_t)(P.sv.s0 + q0 + 1) + P.pos_add + sm.pos[s];
          else pout[dst] = sm.pos[s];
        }
        if constexpr (HAS_NEXT) atomicAdd(&sm.next[P.next(kk)], 1u);
      }
    }
   }
  } else
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const int s = j * THREADS + tid;
    if (FULL || s < (int)tile_count) {
      const uint64_t kk = sm.keys[s];
      const uint32_t b = P.bin(kk);
      int64_t dst = sm.goff[b] + s;
      uint64_t *kout = P.keys_out;
      uint32_t *pout = P.pos_out;
      bool room = true;
      if constexpr (PEER) {
        kout = sm.peer.keys[b]; pout = sm.peer.pos[b];
        const int64_t local = dst;
        dst += sm.peer.delta[b];
        room = (uint64_t)(sm.peer.region ? local : dst) < sm.peer.cap;
      }
      if constexpr (FROM_SEQ && !PEER) room = !drop;
      if (room && !(P.dbg & 2u)) {
        kout[dst] = kk;
        if constexpr (FROM_SEQ) pout[dst] = (uint32_t)(P.sv.s0 + q0 + 1) + P.pos_add + sm.pos[s];   // 1-based start (+ pos_add)
        else pout[dst] = sm.pos[s];
      }
      if constexpr (HAS_NEXT) atomicAdd(&sm.next[P.next(kk)], 1u);
    }
  }
  KMG_STAMP(7);                                          // stores issued
#undef KMG_STAMP
}

// SEGS: the record source is segmented (P.segs); compiled separately so that the dense record pass -- the dominant kernel,
// at its 128-register limit -- does not carry the ticket-to-segment search.
template <class Cfg, bool FROM_SEQ, class BinFn, class NextFn, bool HAS_NEXT, bool PEER = false, bool SEGS = false>
__global__ void __launch_bounds__(Cfg::THREADS, Cfg::MINBLOCKS)
scatter_pass_kernel(const PassParams<BinFn, NextFn> P) {
  using S = PassSmem<Cfg, FROM_SEQ>;
  constexpr int TILE = Cfg::TILE, THREADS = Cfg::THREADS, NB = Cfg::NB, BPT = Cfg::BPT;
  static_assert(TILE <= 65536, "tile-local positions are 16 bit");
  static_assert(!PEER || NB >= MAX_PEERS, "one bin per owner");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S &sm = *reinterpret_cast<S *>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

  const long long t_start = clock64();
  if (tid == 0) sm.tile = atomicAdd(P.ticket, 1u);
  {
    uint4 *zw = reinterpret_cast<uint4 *>(sm.whist[warp]);
#pragma unroll
    for (int j = 0; j < NB / 4; j += 32) if (j + (int)lane < NB / 4) zw[j + lane] = make_uint4(0, 0, 0, 0);
    if constexpr (Cfg::RANK == 0) {
      uint4 *zm = reinterpret_cast<uint4 *>(sm.match + warp * NB);
#pragma unroll
      for (int j = 0; j < NB / 4; j += 32) if (j + (int)lane < NB / 4) zm[j + lane] = make_uint4(0, 0, 0, 0);
    }
  }
  if constexpr (HAS_NEXT)
    for (int b = tid; b < NB; b += THREADS) sm.next[b] = 0;
  if constexpr (PEER) {
    static_assert(sizeof(PeerTable) % 8 == 0, "copied as 64-bit words");
    if (tid < sizeof(PeerTable) / 8) reinterpret_cast<uint64_t *>(&sm.peer)[tid] = reinterpret_cast<const uint64_t *>(P.peer)[tid];
  }
  uint32_t gbase[BPT];                                   // bin bases: a few KB, L2-resident
#pragma unroll
  for (int q = 0; q < BPT; ++q) gbase[q] = (NB >= THREADS || tid < NB) ? __ldg(P.gbase + tid * BPT + q) : 0;
  static_assert(!(SEGS && FROM_SEQ), "only a record source can be segmented");
  uint32_t nsegs = 0;
  if constexpr (SEGS) {
    nsegs = P.segs->nsegs;
    for (uint32_t i = tid; i <= nsegs; i += THREADS) sm.seg_tp[i] = P.segs->tile_prefix[i];
  }
  if constexpr (FROM_SEQ) { if (tid == 0) sm.skip_write = 0; }
  int64_t n_in = FROM_SEQ ? P.sv.nstarts : (SEGS ? 0 : (int64_t)*P.n_records);
  __syncthreads();
  const uint32_t tile = sm.tile;
  int64_t q0 = (int64_t)tile * TILE;
  if constexpr (SEGS) {
    {                                                      // ticket -> (segment, tile inside it); q0, n_in become physical indices
      if (tile >= sm.seg_tp[nsegs]) return;
      uint32_t lo = 0, hi = nsegs;                         // largest r with tile_prefix[r] <= tile
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sm.seg_tp[mid] <= tile) lo = mid; else hi = mid;
      }
      const int64_t base = (int64_t)lo * (int64_t)P.segs->stride;
      q0 = base + (int64_t)(tile - sm.seg_tp[lo]) * TILE;
      n_in = base + (int64_t)P.segs->count[lo];
    }
  }
  if (q0 >= n_in) return;
  if (P.trace && tid == 0) P.trace[(size_t)tile * 8] = (unsigned long long)t_start;

  bool special = false;
  if constexpr (FROM_SEQ) special = tile_pack<TILE, THREADS>(P.sv, q0, sm.tc);
  const bool touches_end = FROM_SEQ && (P.sv.s0 + q0 + TILE + P.sv.k > P.sv.L);
  if (q0 + TILE <= n_in && !special && !touches_end)
    pass_tile<Cfg, FROM_SEQ, true, BinFn, NextFn, HAS_NEXT, PEER>(P, sm, tile, q0, n_in, gbase, special);
  else
    pass_tile<Cfg, FROM_SEQ, false, BinFn, NextFn, HAS_NEXT, PEER>(P, sm, tile, q0, n_in, gbase, special);

  if constexpr (HAS_NEXT) {
    __syncthreads();
    for (int b = tid; b < NB; b += THREADS) {
      const uint32_t c = sm.next[b];
      if (c) atomicAdd(P.hist_next + b, c);
    }
  }
}

// ---- bin bases -----------------------------------------------------------------------------------------
// gbase = exclusive scan of one histogram of NB bins; *n (if given) = its total.  One block of NB threads.
template <int NB>
__global__ void __launch_bounds__(NB)
scan_hist_kernel(const uint32_t *__restrict__ hist, uint32_t *__restrict__ gbase, uint64_t *n) {
  __shared__ uint32_t part[NB / 32];
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t v = hist[tid];
  const uint32_t incl = warp_incl_scan(v);
  if (lane == 31) part[warp] = incl;
  __syncthreads();
  uint32_t add = 0;
  uint64_t tot = 0;
#pragma unroll
  for (int j = 0; j < NB / 32; ++j) {
    if (j < (int)warp) add += part[j];
    tot += part[j];
  }
  gbase[tid] = incl - v + add;
  if (n && tid == 0) *n = tot;
}

// ---- sharded build: where this rank's records of every owner go -------------------------------------
// matrix[src][owner] = records rank `src` holds for `owner` (all-gathered).  One block of RADIX threads.
//   gbase[b]  = first index of bin b in this rank's partitioned stream (exclusive scan of its own row);
//   delta[b]  = (records ranks < rank send to b) - gbase[b]: the pass adds it to the stream index;
//   info[0]   = records this rank receives, info[1] = 1 if some owner receives more than `cap`.
struct PeerPtrs { uint64_t *keys[MAX_PEERS]; uint32_t *pos[MAX_PEERS]; };
__global__ void __launch_bounds__(RADIX)
owner_offsets_kernel(const uint64_t *__restrict__ matrix, int nparts, int rank, uint64_t cap, PeerPtrs ptrs,
                     uint32_t *gbase, PeerTable *tab, uint64_t *info) {
  const int b = threadIdx.x;
  if (b == 0) {
    info[1] = 0;
    uint64_t run = 0;
    for (int o = 0; o < RADIX; ++o) { gbase[o] = (uint32_t)run; if (o < nparts) run += matrix[(size_t)rank * nparts + o]; }
  }
  __syncthreads();
  if (b < MAX_PEERS) {
    uint64_t before = 0, total = 0;
    if (b < nparts)
      for (int src = 0; src < nparts; ++src) {
        const uint64_t c = matrix[(size_t)src * nparts + b];
        if (src < rank) before += c;
        total += c;
      }
    tab->keys[b] = b < nparts ? ptrs.keys[b] : nullptr;
    tab->pos[b] = b < nparts ? ptrs.pos[b] : nullptr;
    tab->delta[b] = (int64_t)before - (int64_t)gbase[b];
    tab->counts[b] = nullptr;
    if (b == rank) info[0] = total;
    if (total > cap) info[1] = 1;
  }
  if (b == 0) { tab->cap = cap; tab->rank = (uint32_t)rank; tab->region = 0; }
}

// ---- all passes' histograms from the sequence, in one sweep -------------------------------------------
// Digit r of the window that starts at p is the code of bases [p + lo_r, p + hi_r], lo_r = max(0, k-4r-4),
// hi_r = k-4r-1: a whole 4-mer for every digit but a possibly shorter top one.  So over a stretch of
// valid window starts [a, b) the histogram of digit r is the 4-mer histogram of sequence positions
// [a + lo_r, b + lo_r): the same for all r up to lo_r positions at either end.  A tile without breakers
// therefore counts each of its positions ONCE into `common` and fixes up the <= 28 positions at each
// end per digit (ind[r], modulo 2^32); tiles with breakers or at the end of the data count every
// window's digits directly into ind[r].  hist_finish_kernel adds the two.
template <int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS)
hist_all_kernel(const SeqView sv, uint32_t *common /* [RADIX] */, uint32_t *ind /* [R][stride] */, const int stride) {
  constexpr int TILE = THREADS * ITEMS;
  __shared__ TileCodes<TILE> tc;
  __shared__ uint32_t sh_common[RADIX];
  __shared__ uint32_t sh_ind[MAX_PASSES][RADIX];
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const int k = sv.k, R = num_passes(k);
  for (int b = tid; b < RADIX; b += THREADS) sh_common[b] = 0;
  for (int b = tid; b < R * RADIX; b += THREADS) (&sh_ind[0][0])[b] = 0;
  const int64_t tiles = ceil_div<int64_t>(sv.nstarts, TILE);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t q0 = tile * TILE;
    __syncthreads();                                   // previous tile's readers are done (and the zeroing)
    const bool special = tile_pack<TILE, THREADS>(sv, q0, tc);
    const int t0 = warp * (32 * ITEMS) + lane;
    if (!special && q0 + TILE <= sv.nstarts) {
      // every window of the tile is valid (a tile that touches the end of the data is `special`)
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) atomicAdd(&sh_common[tile_c4<TILE>(tc, t0 + i * 32)], 1u);
      // ends: digit r uses positions [q0 + lo_r, q0 + TILE + lo_r) instead of [q0, q0 + TILE)
      for (int e = tid; e < R * 32; e += THREADS) {
        const int r = e >> 5, t = e & 31, lo = k - 4 * r - 4;
        if (t < lo) {
          atomicSub(&sh_ind[r][tile_c4<TILE>(tc, t)], 1u);
          atomicAdd(&sh_ind[r][tile_c4<TILE>(tc, TILE + t)], 1u);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        const int t = t0 + i * 32;
        if (tile_valid<TILE>(sv, tc, q0, t, special)) {
          const uint64_t key = tile_key<TILE>(tc, t, k);
          for (int r = 0; r < R; ++r) atomicAdd(&sh_ind[r][(uint32_t)(key >> (r * RADIX_BITS)) & (RADIX - 1)], 1u);
        }
      }
    }
  }
  __syncthreads();
  for (int b = tid; b < RADIX; b += THREADS) {
    const uint32_t c = sh_common[b];
    if (c) atomicAdd(common + b, c);
  }
  for (int b = tid; b < R * RADIX; b += THREADS) {
    const uint32_t c = (&sh_ind[0][0])[b];
    if (c) atomicAdd(ind + (size_t)(b / RADIX) * stride + (b % RADIX), c);
  }
}

// hist[r] = ind[r] + common folded to digit r's width; gbase[r] = its exclusive scan; *n = windows.
// Launched with R blocks of RADIX threads.
__global__ void __launch_bounds__(RADIX)
hist_finish_kernel(int k, const uint32_t *__restrict__ common, uint32_t *hist /* [R][stride], holds ind */,
                   uint32_t *gbase /* [R][stride] */, const int stride, uint64_t *n) {
  __shared__ uint32_t part[RADIX / 32];
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const int r = blockIdx.x;
  const int m = min(4, k - 4 * r);                      // bases in digit r
  const int s = 2 * (4 - m);
  uint32_t v = hist[(size_t)r * stride + tid];
  if (tid < (1u << (2 * m))) {
    for (uint32_t c = tid << s; c < ((tid + 1u) << s); ++c) v += common[c];
  }
  hist[(size_t)r * stride + tid] = v;
  const uint32_t incl = warp_incl_scan(v);
  if (lane == 31) part[warp] = incl;
  __syncthreads();
  uint32_t add = 0;
  uint64_t tot = 0;
#pragma unroll
  for (int j = 0; j < RADIX / 32; ++j) {
    if (j < (int)warp) add += part[j];
    tot += part[j];
  }
  gbase[(size_t)r * stride + tid] = incl - v + add;
  if (r == 0 && tid == 0) *n = tot;
}

// ---- histogram of one pass's bins (sharded build: owner bins from the sequence, digit 0 of records) ----
template <int THREADS, int ITEMS, class BinFn>
__global__ void __launch_bounds__(THREADS)
hist_seq_kernel(const SeqView sv, uint32_t *hist, BinFn bin) {
  constexpr int TILE = THREADS * ITEMS;
  __shared__ TileCodes<TILE> tc;
  __shared__ uint32_t sh[MAX_NB];
  const unsigned tid = threadIdx.x;
  for (int b = tid; b < MAX_NB; b += THREADS) sh[b] = 0;
  const int64_t tiles = ceil_div<int64_t>(sv.nstarts, TILE);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t q0 = tile * TILE;
    __syncthreads();                                   // previous tile's readers are done
    const bool special = tile_pack<TILE, THREADS>(sv, q0, tc);
    // a histogram does not care about order: a thread takes ITEMS consecutive windows and rolls the key from
    // one to the next (one base in) instead of extracting every window from the packed tile
    const int t0 = tid * ITEMS;
    const uint64_t kmask = key_mask(sv.k);
    uint64_t key = tile_key<TILE>(tc, t0, sv.k);
    // the ITEMS - 1 bases that roll in (tile positions t0 + k .. t0 + k + ITEMS - 2) span at most two packed words: take
    // them into a register once instead of one shared-memory read per window
    static_assert(ITEMS <= 17, "incoming bases fit two 16-base words");
    const int p0 = t0 + sv.k;
    const uint64_t inc = ((uint64_t(tc.codes[p0 >> 4]) << 32) | tc.codes[(p0 >> 4) + 1]) << (2 * (p0 & 15));
    if (!special && q0 + TILE <= sv.nstarts && sv.s0 + q0 + TILE + sv.k <= sv.L) {     // every window of the tile is valid
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        if (i > 0) key = ((key << 2) | ((inc >> (64 - 2 * i)) & 3u)) & kmask;
        atomicAdd(&sh[bin(key)], 1u);
      }
    } else {
#pragma unroll
      for (int i = 0; i < ITEMS; ++i) {
        if (i > 0) key = ((key << 2) | ((inc >> (64 - 2 * i)) & 3u)) & kmask;
        if (tile_valid<TILE>(sv, tc, q0, t0 + i, special)) atomicAdd(&sh[bin(key)], 1u);
      }
    }
  }
  __syncthreads();
  for (int b = tid; b < MAX_NB; b += THREADS) {
    uint32_t c = sh[b];
    if (c) atomicAdd(hist + b, c);
  }
}

// Record source.
template <int THREADS, class BinFn>
__global__ void __launch_bounds__(THREADS)
hist_rec_kernel(const uint64_t *keys, int64_t n_host, const uint64_t *n_dev, uint32_t *hist, BinFn bin) {
  __shared__ uint32_t sh[MAX_NB];
  const int64_t n = n_dev ? (int64_t)*n_dev : n_host;   // the count may only exist on the device
  for (int b = threadIdx.x; b < MAX_NB; b += THREADS) sh[b] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS)
    atomicAdd(&sh[bin(ld_stream_u64(keys + i))], 1u);
  __syncthreads();
  for (int b = threadIdx.x; b < MAX_NB; b += THREADS) {
    uint32_t c = sh[b];
    if (c) atomicAdd(hist + b, c);
  }
}

// The same over a segmented record source (see TileSegs).
template <int THREADS, class BinFn>
__global__ void __launch_bounds__(THREADS)
hist_seg_kernel(const uint64_t *keys, const TileSegs *seg, uint32_t *hist, BinFn bin) {
  __shared__ uint32_t sh[MAX_NB];
  for (int b = threadIdx.x; b < MAX_NB; b += THREADS) sh[b] = 0;
  __syncthreads();
  const uint32_t nsegs = seg->nsegs;
  for (uint32_t r = 0; r < nsegs; ++r) {
    const int64_t n = (int64_t)seg->count[r];
    const uint64_t *src = keys + (uint64_t)r * seg->stride;
    for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += (int64_t)gridDim.x * THREADS)
      atomicAdd(&sh[bin(ld_stream_u64(src + i))], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < MAX_NB; b += THREADS) {
    uint32_t c = sh[b];
    if (c) atomicAdd(hist + b, c);
  }
}

// counts (per segment) -> TileSegs for a reading pass with tiles of `tile` records; *n = records held; counts above the
// segment size are clamped and flagged.  One block of MAX_SEGS threads.
__global__ void __launch_bounds__(MAX_SEGS)
tile_segs_kernel(const uint64_t *__restrict__ counts, int nsegs, uint64_t stride, uint32_t tile, TileSegs *seg, uint64_t *n) {
  __shared__ uint32_t part_t[MAX_SEGS / 32];
  __shared__ uint64_t part_c[MAX_SEGS / 32];
  __shared__ uint32_t over;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) over = 0;
  __syncthreads();
  uint64_t c = (int)tid < nsegs ? counts[tid] : 0;
  if (c > stride) over = 1;
  __syncthreads();
  if (over) c = 0;      // a segment was sent more than it holds (the excess was dropped): hand out an EMPTY source, so that every
                        // kernel queued behind this one is a no-op on well-defined data; the host sees the flag and takes another path
  const uint32_t t = (uint32_t)((c + tile - 1) / tile);
  const uint32_t it = warp_incl_scan(t);
  const uint64_t ic = warp_incl_scan64(c);
  if (lane == 31) { part_t[warp] = it; part_c[warp] = ic; }
  __syncthreads();
  uint32_t bt = 0, tt = 0;
  uint64_t tc = 0;
  for (int w = 0; w < MAX_SEGS / 32; ++w) { if (w < (int)warp) bt += part_t[w]; tt += part_t[w]; tc += part_c[w]; }
  seg->tile_prefix[tid] = bt + it - t;
  seg->count[tid] = (uint32_t)c;
  if (tid == 0) {
    seg->tile_prefix[MAX_SEGS] = tt;
    seg->stride = stride; seg->total = tc; seg->nsegs = (uint32_t)nsegs; seg->tile = tile; seg->overflow = over; seg->pad = 0;
    *n = tc;
  }
  if ((int)tid == nsegs) seg->tile_prefix[nsegs] = tt;      // (also covers nsegs < MAX_SEGS: the entry the reader bounds by)
}
// dense copy of a segmented (key, payload) source
__global__ void seg_compact_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pay, const TileSegs *seg,
                                   uint64_t *__restrict__ okeys, uint32_t *__restrict__ opay) {
  const uint32_t nsegs = seg->nsegs;
  uint64_t dst = 0;
  for (uint32_t r = 0; r < nsegs; ++r) {
    const int64_t n = (int64_t)seg->count[r];
    const uint64_t src = (uint64_t)r * seg->stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      okeys[dst + i] = ld_stream_u64(keys + src + i);
      opay[dst + i] = ld_stream_u32(pay + src + i);
    }
    dst += (uint64_t)n;
  }
}
// Is the mixed digit's heaviest bin safely below a region's size?  Every `stride`-th window (a prime, so that tandem arrays
// are sampled in all their phases) is hashed and binned; est[0] = stride * largest sampled bin.  (Breakers are encoded like
// any byte -- this only steers a choice -- but windows that contain one are skipped, as the index skips them.)
__global__ void region_sample_kernel(const SeqView sv, int stride, uint32_t mask, uint32_t *hist /* [MAX_NB], zero */) {
  __shared__ uint32_t sh[MAX_NB];
  for (int b = threadIdx.x; b < MAX_NB; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  const int64_t nsamp = sv.nstarts / stride;
  const uint64_t kmask = key_mask(sv.k);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nsamp; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t *p = sv.base + i * stride;
    uint64_t w = 0;
    bool clean = true;                                     // windows with a breaker are not indexed: an N gap must not look like a repeat
    for (int j = 0; j < sv.k; ++j) { const uint8_t c = p[j]; clean &= (c | 0x20) != 'n'; w = (w << 2) | ((c >> 1) & 3u); }
    if (clean) atomicAdd(&sh[(uint32_t)mix64(w & kmask) & mask], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < MAX_NB; b += blockDim.x) {
    const uint32_t c = sh[b];
    if (c) atomicAdd(hist + b, c);
  }
}
__global__ void region_estimate_kernel(const uint32_t *hist, int nb, int stride, uint32_t *est) {
  __shared__ uint32_t m;
  if (threadIdx.x == 0) m = 0;
  __syncthreads();
  uint32_t mine = 0;
  for (int b = threadIdx.x; b < nb; b += blockDim.x) mine = max(mine, hist[b]);
  atomicMax(&m, mine);
  __syncthreads();
  if (threadIdx.x == 0) est[0] = m * (uint32_t)stride;
}
// gbase[b] = b * cap (bins written into their own regions) and a zeroed overflow word
__global__ void region_gbase_kernel(uint32_t *gbase, int nb, uint32_t cap, uint32_t *overflow) {
  for (int b = threadIdx.x; b < MAX_NB; b += blockDim.x) gbase[b] = b < nb ? (uint32_t)b * cap : 0;
  if (threadIdx.x == 0) *overflow = 0;
}

}  // namespace kmg
