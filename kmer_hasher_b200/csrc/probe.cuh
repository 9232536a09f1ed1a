// probe.cuh -- seq.kmer.pos: match query windows against the sorted distinct keys.
//
// Replaces seq_kmer_positions (src/kmer_pos.c:110-136): kmer_pos/kh_get (:55-60) becomes a prefix
// table + short binary search over ukeys[]; pair_positions_push (:101-108) becomes a count pass, a
// chained 64-bit scan and a load-balanced emit, so rows come out ordered by query position then
// index position exactly as the reference pushes them.
#pragma once
#include "common.cuh"
#include "lookback.cuh"
#include "windows.cuh"
#include "csr.cuh"

namespace kmg {

struct QueryStats {
  uint64_t H;   // query windows that hit
  uint64_t M;   // result rows
};

// Prefix table over the distinct keys: lut[b] = first u whose key >> shift is >= b, b in [0, 2^B].
__global__ void lut_kernel(const uint64_t *__restrict__ ukeys, uint64_t U, int shift, uint64_t nbuckets,
                           uint32_t *__restrict__ lut) {
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < U; u += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t p = ukeys[u] >> shift;
    const uint64_t from = u == 0 ? 0 : (ukeys[u - 1] >> shift) + 1;
    for (uint64_t b = from; b <= p; ++b) lut[b] = (uint32_t)u;
    if (u == U - 1)
      for (uint64_t b = p + 1; b <= nbuckets; ++b) lut[b] = (uint32_t)U;
  }
}

struct KeyTable {
  const uint64_t *ukeys;
  const uint32_t *ustart;
  const uint32_t *lut;
  uint64_t U;
  uint64_t nbuckets;   // 2^B
  int shift;
};

// index of `key` among the distinct keys, or 0xFFFFFFFF
__device__ __forceinline__ uint32_t find_key(const KeyTable &kt, uint64_t key) {
  const uint64_t b = kt.shift >= 64 ? 0 : (key >> kt.shift);
  if (b >= kt.nbuckets) return 0xFFFFFFFFu;
  uint32_t lo = __ldg(kt.lut + b), hi = __ldg(kt.lut + b + 1);
  const uint32_t end = hi;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint64_t v = __ldg(kt.ukeys + mid);
    if (v < key) lo = mid + 1; else hi = mid;
  }
  if (lo < end && __ldg(kt.ukeys + lo) == key) return lo;
  return 0xFFFFFFFFu;
}

// Count pass + ordered compaction of the hits + chained scan of their row counts.
//   FROM_SEQ: windows of the query sequence, coordinate i = 1-based END of the window
//             (src/kmer_pos.c:127,132: `i` is one past the window);
//   else    : pre-encoded (key, i) records.
template <int THREADS, int ITEMS, bool FROM_SEQ>
__global__ void __launch_bounds__(THREADS)
probe_match_kernel(const SeqView sv, const uint64_t *__restrict__ keys_in, const int32_t *__restrict__ i_in,
                   int64_t n_in, const uint64_t *__restrict__ n_dev, const KeyTable kt, int32_t *__restrict__ hit_i, uint32_t *__restrict__ hit_u,
                   uint64_t *__restrict__ row_off, QueryStats *qs, Pair64 *status, uint32_t *ticket) {
  constexpr int TILE = THREADS * ITEMS, WARPS = THREADS / 32;
  __shared__ TileCodes<FROM_SEQ ? TILE : 16> tc;
  __shared__ uint32_t s_tile, s_wh[WARPS];
  __shared__ uint64_t s_wr[WARPS], s_bh, s_br;
  const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const int64_t q0 = (int64_t)tile * TILE;
  const int64_t total = FROM_SEQ ? sv.nstarts : (n_dev ? min((int64_t)*n_dev, n_in) : n_in);   // the count may only exist on the device
  if (q0 >= total) return;

  uint32_t u[ITEMS], cnt[ITEMS];
  int32_t coord[ITEMS];
  const int t0 = warp * (32 * ITEMS) + lane;
  bool special = false;
  if constexpr (FROM_SEQ) special = tile_pack<TILE, THREADS>(sv, q0, tc);
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const int t = t0 + i * 32;
    u[i] = 0xFFFFFFFFu;
    if constexpr (FROM_SEQ) {
      coord[i] = (int32_t)(sv.s0 + q0 + t + sv.k);
      if (tile_valid<TILE>(sv, tc, q0, t, special)) u[i] = find_key(kt, tile_key<TILE>(tc, t, sv.k));
    } else {
      coord[i] = 0;
      if (q0 + t < total) { coord[i] = i_in[q0 + t]; u[i] = find_key(kt, ld_stream_u64(keys_in + q0 + t)); }
    }
  }
#pragma unroll
  for (int i = 0; i < ITEMS; ++i)
    cnt[i] = u[i] != 0xFFFFFFFFu ? __ldg(kt.ustart + u[i] + 1) - __ldg(kt.ustart + u[i]) : 0;

  uint32_t hb[ITEMS], hrun = 0;
  uint64_t rb[ITEMS], rrun = 0;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    const unsigned bal = __ballot_sync(FULL, cnt[i] != 0);
    hb[i] = hrun + __popc(bal & lanemask_lt());
    hrun += __popc(bal);
    const uint64_t inc = warp_incl_scan64(cnt[i]);
    rb[i] = rrun + inc - cnt[i];
    rrun += __shfl_sync(FULL, inc, 31);
  }
  if (lane == 0) { s_wh[warp] = hrun; s_wr[warp] = rrun; }
  __syncthreads();
  uint64_t bh = 0, br = 0, th = 0, tr = 0;
#pragma unroll
  for (int w = 0; w < WARPS; ++w) {
    if (w < (int)warp) { bh += s_wh[w]; br += s_wr[w]; }
    th += s_wh[w]; tr += s_wr[w];
  }
  if (warp == 0) {
    uint64_t ea, eb;
    pair_lookback(status, tile, th, tr, ea, eb);
    if (lane == 0) { s_bh = ea; s_br = eb; }
  }
  __syncthreads();
  bh += s_bh; br += s_br;
#pragma unroll
  for (int i = 0; i < ITEMS; ++i) {
    if (cnt[i]) {
      const uint64_t h = bh + hb[i];
      hit_i[h] = coord[i];
      hit_u[h] = u[i];
      row_off[h] = br + rb[i];
    }
  }
  if (q0 + TILE >= total && tid == 0) { qs->H = s_bh + th; qs->M = s_br + tr; }
}

// Emit rows [first, first+nrows): row r belongs to hit h = largest h with row_off[h] <= r; it pairs
// the hit's query coordinate with the (r - row_off[h])-th position of its k-mer.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
probe_emit_kernel(const int32_t *__restrict__ hit_i, const uint32_t *__restrict__ hit_u,
                  const uint64_t *__restrict__ row_off, uint64_t H, const uint32_t *__restrict__ ustart,
                  const uint32_t *__restrict__ pos, uint64_t first, uint64_t nrows, int2 *__restrict__ out) {
  constexpr int PER = 8, T = THREADS * PER;
  __shared__ __align__(16) uint8_t s_flag[T];
  __shared__ __align__(16) uint32_t s_seg[T];
  __shared__ uint32_t s_warp[THREADS / 32];
  __shared__ uint64_t s_first;
  const uint64_t b0 = (uint64_t)blockIdx.x * T;
  if (b0 >= nrows) return;
  const uint64_t r0 = first + b0;
  const uint64_t h0 = block_segments<THREADS, PER, uint64_t>(row_off, H, r0, s_flag, s_seg, s_warp, &s_first);
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const uint32_t s = j * THREADS + threadIdx.x;
    if (b0 + s >= nrows) continue;
    const uint64_t h = h0 + s_seg[s];
    const uint32_t u = hit_u[h];
    const uint64_t within = r0 + s - row_off[h];
    out[b0 + s] = make_int2(hit_i[h], (int)pos[ustart[u] + within]);
  }
}

}  // namespace kmg
