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
  uint32_t epoch;
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

template <int THREADS, int ITEMS, bool FROM_SEQ>
struct PassSmem {
  static constexpr int TILE = THREADS * ITEMS;
  static constexpr int WARPS = THREADS / 32;
  using PosT = typename std::conditional<FROM_SEQ, uint16_t, uint32_t>::type;
  uint64_t keys[TILE];
  PosT pos[TILE];
  uint16_t whist[WARPS][RADIX];
  int64_t goff[RADIX];
  uint32_t start[RADIX];
  uint32_t next[RADIX];
  uint32_t scratch[8];
  uint32_t tile;
  TileCodes<FROM_SEQ ? TILE : 16> tc;
};

template <int THREADS, int ITEMS, bool FROM_SEQ, class BinFn, class NextFn, bool HAS_NEXT>
__global__ void __launch_bounds__(THREADS)
scatter_pass_kernel(const PassParams<BinFn, NextFn> P) {
  using S = PassSmem<THREADS, ITEMS, FROM_SEQ>;
  constexpr int TILE = S::TILE;
  static_assert(TILE <= 65536, "tile-local positions are 16 bit");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S &sm = *reinterpret_cast<S *>(smem_raw);
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

  if (tid == 0) sm.tile = atomicAdd(P.ticket, 1u);
  for (int b = lane; b < RADIX; b += 32) sm.whist[warp][b] = 0;
  if (tid < RADIX) sm.next[tid] = 0;

  // global exclusive base of every bin for this pass (the histogram is L2-resident, 1 KB)
  uint32_t gcount = tid < RADIX ? P.hist_cur[tid] : 0;
  uint32_t n_total;
  uint32_t gbase = bins_excl_scan<THREADS>(gcount, sm.scratch, n_total);   // syncs: sm.tile visible
  const uint32_t tile = sm.tile;
  const int64_t q0 = (int64_t)tile * TILE;
  const int64_t n_in = FROM_SEQ ? P.sv.nstarts : (int64_t)n_total;
  if (q0 >= n_in) return;

  // ---- load ITEMS records per thread, warp-striped (item i of lane l = element i*32+l of the
  //      warp's chunk), which is memory order => ranks below are stable
  uint64_t key[ITEMS];
  uint32_t val[ITEMS];
  uint32_t valid = 0;
  const int t0 = warp * (32 * ITEMS) + lane;
  if constexpr (FROM_SEQ) {
    const bool special = tile_pack<TILE, THREADS>(P.sv, q0, sm.tc);
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int t = t0 + i * 32;
      key[i] = tile_key<TILE>(sm.tc, t, P.sv.k);
      val[i] = (uint32_t)t;
      if (tile_valid<TILE>(P.sv, sm.tc, q0, t, special)) valid |= 1u << i;
    }
  } else {
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const int64_t idx = q0 + t0 + i * 32;
      key[i] = 0; val[i] = 0;
      if (idx < n_in) {
        key[i] = ld_stream_u64(P.keys_in + idx);
        val[i] = ld_stream_u32(P.pos_in + idx);
        valid |= 1u << i;
      }
    }
  }

  // ---- rank inside the warp by ballot matching on the bin bits
  uint16_t rank[ITEMS];
  const unsigned lt = lanemask_lt();
  __syncwarp();
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const bool ok = (valid >> i) & 1u;
    const uint32_t d = P.bin(key[i]);
    unsigned peers = __ballot_sync(FULL, ok);
#pragma unroll
    for (int b = 0; b < RADIX_BITS; ++b) {
      const bool bit = (d >> b) & 1u;
      const unsigned bal = __ballot_sync(FULL, bit);
      peers &= bit ? bal : ~bal;
    }
    const int leader = __ffs(peers) - 1;
    uint32_t old = 0;
    if (ok && (int)lane == leader) {
      old = sm.whist[warp][d];
      sm.whist[warp][d] = (uint16_t)(old + __popc(peers));
    }
    __syncwarp();
    old = __shfl_sync(FULL, old, leader < 0 ? 0 : leader);
    rank[i] = (uint16_t)(old + __popc(peers & lt));
  }
  __syncthreads();

  // ---- per-bin totals of the tile; publish them, then place the tile's bins in shared memory
  uint32_t cnt = 0;
  if (tid < RADIX) {
#pragma unroll
    for (int w = 0; w < S::WARPS; ++w) {
      uint32_t c = sm.whist[w][tid];
      sm.whist[w][tid] = (uint16_t)cnt;      // becomes the warp's offset inside the bin
      cnt += c;
    }
    uint64_t *mine = P.status + (size_t)tile * RADIX + tid;
    st_relaxed_u64(mine, st_pack(tile == 0 ? ST_INCL : ST_AGG, P.epoch, cnt));
  }
  uint32_t tile_count;
  const uint32_t lstart = bins_excl_scan<THREADS>(cnt, sm.scratch, tile_count);
  if (tid < RADIX) sm.start[tid] = lstart;
  __syncthreads();

#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    if ((valid >> i) & 1u) {
      const uint32_t d = P.bin(key[i]);
      const uint32_t slot = sm.start[d] + sm.whist[warp][d] + rank[i];
      sm.keys[slot] = key[i];
      sm.pos[slot] = (typename S::PosT)val[i];
    }
  }

  // ---- look back over earlier tiles (one bin per thread); predecessors have had the whole
  //      regrouping above to publish
  if (tid < RADIX) {
    uint64_t excl = 0;
    if (tile > 0) {
      for (int64_t t = (int64_t)tile - 1; t >= 0; --t) {
        const uint64_t *p = P.status + (size_t)t * RADIX + tid;
        uint64_t w, f;
        do { w = ld_relaxed_u64(p); f = st_flag(w, P.epoch); } while (f == 0);
        excl += st_value(w);
        if (f == ST_INCL) break;
      }
      st_relaxed_u64(P.status + (size_t)tile * RADIX + tid, st_pack(ST_INCL, P.epoch, excl + cnt));
    }
    sm.goff[tid] = (int64_t)gbase + (int64_t)excl - (int64_t)lstart;
  }
  __syncthreads();

  // ---- stream the regrouped tile out: consecutive threads -> consecutive slots -> runs per bin
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t s = j * THREADS + tid;
    if (s < tile_count) {
      const uint64_t kk = sm.keys[s];
      const uint32_t d = P.bin(kk);
      const int64_t dst = sm.goff[d] + (int64_t)s;
      P.keys_out[dst] = kk;
      if constexpr (FROM_SEQ) P.pos_out[dst] = (uint32_t)(P.sv.s0 + q0 + (int64_t)sm.pos[s] + 1);   // 1-based start
      else P.pos_out[dst] = sm.pos[s];
      if (HAS_NEXT) atomicAdd(&sm.next[P.next(kk)], 1u);
    }
  }
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
