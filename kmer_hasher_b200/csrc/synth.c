/*
 * synth.c -- deterministic synthetic sequences for the BASELINE.json configurations
 * (SURVEY.md section 8d).  Host-only C so that the CUDA library, the oracle and the reference
 * engine all consume byte-identical inputs on any machine.  There is no network for real
 * genomes; bench.py says "data": "synthetic".
 *
 * The mixture imitates what stresses a k-mer position index: unique sequence, interspersed
 * repeat families with diverged copies, tandem arrays (microsatellites, telomere-like
 * (CCCTAA)n), homopolymer runs, soft-masked (lower-case) blocks, N gaps and isolated N/n.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r) {          /* xorshift64* */
  uint64_t x = r->s;
  x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
  r->s = x;
  return x * 0x2545F4914F6CDD1DULL;
}
static inline uint64_t rng_below(rng_t *r, uint64_t n) { return n ? rng_next(r) % n : 0; }
static inline double rng_unit(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline int64_t rng_loguniform(rng_t *r, int64_t lo, int64_t hi) {
  if (hi <= lo) return lo;
  double v = exp(log((double)lo) + rng_unit(r) * (log((double)hi) - log((double)lo)));
  int64_t x = (int64_t)v;
  return x < lo ? lo : (x > hi ? hi : x);
}
static const char ACGT[4] = {'A', 'C', 'G', 'T'};

static void fill_random(uint8_t *out, int64_t n, rng_t *r) {
  int64_t i = 0;
  while (i < n) {
    uint64_t x = rng_next(r);
    for (int j = 0; j < 32 && i < n; ++j, ++i) { out[i] = (uint8_t)ACGT[x & 3]; x >>= 2; }
  }
}

/*
 * params (doubles, so ctypes can pass one array):
 *  [0] repeat_frac   fraction of bases covered by interspersed repeat copies
 *  [1] tandem_frac   fraction covered by tandem arrays
 *  [2] homo_frac     fraction covered by homopolymer runs
 *  [3] lower_frac    fraction lower-cased in blocks
 *  [4] n_gaps        number of N gaps
 *  [5] gap_min  [6] gap_max   gap length range (log-uniform)
 *  [7] n_single      isolated single N / n
 *  [8] tail_k        if > 0, put an N so that the final N-free run has length exactly tail_k
 *  [9] tandem_unit_max  (default 60)   [10] tandem_len_max (default 200000)
 *  [11] n_families (default 200)
 *  [12] tandem_len_min (default 1000)   [13] homo_len_max (default 500)
 */
int kms_generate(uint8_t *out, int64_t L, uint64_t seed, const double *p, int np) {
  if (!out || L < 0 || np < 14) return -1;
  rng_t r = {seed ? seed : 0x9E3779B97F4A7C15ULL};
  for (int i = 0; i < 8; ++i) rng_next(&r);
  fill_random(out, L, &r);
  if (L < 1000) goto tail;
  {
    const double repeat_frac = p[0], tandem_frac = p[1], homo_frac = p[2], lower_frac = p[3];
    const int n_gaps = (int)p[4];
    const int64_t gap_min = (int64_t)p[5], gap_max = (int64_t)p[6];
    const int n_single = (int)p[7];
    const int unit_max = p[9] > 1 ? (int)p[9] : 60;
    const int64_t tlen_max = p[10] > 1 ? (int64_t)p[10] : 200000;
    const int n_fam = p[11] >= 1 ? (int)p[11] : 200;
    const int64_t tlen_min = p[12] >= 1 ? (int64_t)p[12] : 1000;
    const int64_t hlen_max = p[13] >= 21 ? (int64_t)p[13] : 500;

    /* interspersed repeat families */
    if (repeat_frac > 0) {
      int64_t target = (int64_t)(repeat_frac * (double)L), covered = 0;
      uint8_t **cons = malloc((size_t)n_fam * sizeof(uint8_t *));
      int64_t *clen = malloc((size_t)n_fam * sizeof(int64_t));
      for (int f = 0; f < n_fam; ++f) {
        clen[f] = 300 + (int64_t)rng_below(&r, 5701);
        if (clen[f] > L / 4) clen[f] = L / 4 > 0 ? L / 4 : 1;
        cons[f] = malloc((size_t)clen[f]);
        fill_random(cons[f], clen[f], &r);
      }
      while (covered < target) {
        int f = (int)rng_below(&r, (uint64_t)n_fam);
        double div = 0.02 + 0.13 * rng_unit(&r);
        uint64_t thr = (uint64_t)(div * 4294967296.0);
        int64_t at = (int64_t)rng_below(&r, (uint64_t)(L - clen[f] + 1));
        for (int64_t j = 0; j < clen[f]; ++j) {
          uint64_t x = rng_next(&r);
          out[at + j] = ((x & 0xFFFFFFFFu) < thr) ? (uint8_t)ACGT[(x >> 40) & 3] : cons[f][j];
        }
        covered += clen[f];
      }
      for (int f = 0; f < n_fam; ++f) free(cons[f]);
      free(cons); free(clen);
    }
    /* tandem arrays: the first two are (CA)n and (CCCTAA)n */
    if (tandem_frac > 0) {
      int64_t target = (int64_t)(tandem_frac * (double)L), covered = 0;
      int a = 0;
      while (covered < target) {
        uint8_t unit[64];
        int ul;
        if (a == 0) { memcpy(unit, "CA", 2); ul = 2; }
        else if (a == 1) { memcpy(unit, "CCCTAA", 6); ul = 6; }
        else { ul = 2 + (int)rng_below(&r, (uint64_t)(unit_max - 1)); fill_random(unit, ul, &r); }
        int64_t alen = rng_loguniform(&r, tlen_min < tlen_max ? tlen_min : tlen_max, tlen_max);
        if (alen > L / 2) alen = L / 2;
        int64_t at = (int64_t)rng_below(&r, (uint64_t)(L - alen + 1));
        for (int64_t j = 0; j < alen; ++j) out[at + j] = unit[j % ul];
        covered += alen;
        ++a;
      }
    }
    /* homopolymer / low complexity */
    if (homo_frac > 0) {
      int64_t target = (int64_t)(homo_frac * (double)L), covered = 0;
      while (covered < target) {
        int64_t rl = 20 + (int64_t)rng_below(&r, (uint64_t)(hlen_max - 19));
        int64_t at = (int64_t)rng_below(&r, (uint64_t)(L - rl + 1));
        memset(out + at, ACGT[rng_below(&r, 4)], (size_t)rl);
        covered += rl;
      }
    }
    /* soft-masked blocks */
    if (lower_frac > 0) {
      int64_t target = (int64_t)(lower_frac * (double)L), covered = 0;
      while (covered < target) {
        int64_t bl = 100 + (int64_t)rng_below(&r, 20000);
        if (bl > L) bl = L;
        int64_t at = (int64_t)rng_below(&r, (uint64_t)(L - bl + 1));
        for (int64_t j = 0; j < bl; ++j) out[at + j] |= 0x20;
        covered += bl;
      }
    }
    /* N gaps and isolated N / n */
    for (int g = 0; g < n_gaps; ++g) {
      int64_t gl = rng_loguniform(&r, gap_min > 0 ? gap_min : 1, gap_max > 0 ? gap_max : 1);
      if (gl > L / 8) gl = L / 8;
      int64_t at = (int64_t)rng_below(&r, (uint64_t)(L - gl + 1));
      memset(out + at, 'N', (size_t)gl);
    }
    for (int g = 0; g < n_single; ++g) out[rng_below(&r, (uint64_t)L)] = (g & 1) ? 'n' : 'N';
  }
tail:
  {
    const int tail_k = (int)p[8];
    if (tail_k > 0 && L > tail_k + 1) {
      for (int j = 1; j <= tail_k; ++j)
        if ((out[L - j] | 0x20) == 'n') out[L - j] = 'A';
      out[L - tail_k - 1] = 'N';
    }
  }
  return 0;
}

/*
 * Query for the dot-plot configuration (C4): segments copied from `ref` (log-uniform lengths in
 * [seg_min, seg_max], random order) with substitutions and short indels, interleaved with
 * unrelated random sequence.  Forward strand only (the reference never reverse-complements on
 * this path).
 */
int kms_make_query(const uint8_t *ref, int64_t L, uint8_t *out, int64_t Lq, uint64_t seed,
                   double unrelated_frac, double sub_rate, double indel_rate, int64_t seg_min,
                   int64_t seg_max) {
  if (!ref || !out || L < 1 || Lq < 0) return -1;
  rng_t r = {seed ? seed : 0xC4C4C4C4ULL};
  for (int i = 0; i < 8; ++i) rng_next(&r);
  const uint64_t sub_thr = (uint64_t)(sub_rate * 4294967296.0), indel_thr = (uint64_t)(indel_rate * 4294967296.0);
  int64_t w = 0;
  while (w < Lq) {
    int64_t sl = rng_loguniform(&r, seg_min, seg_max);
    if (sl > Lq - w) sl = Lq - w;
    if (rng_unit(&r) < unrelated_frac) {
      fill_random(out + w, sl, &r);
      w += sl;
      continue;
    }
    if (sl > L) sl = L;
    int64_t src = (int64_t)rng_below(&r, (uint64_t)(L - sl + 1));
    int64_t end = w + sl;
    while (w < end && src < L) {
      uint64_t x = rng_next(&r);
      uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
      if (lo < indel_thr) {
        if (hi & 1) { ++src; continue; }                /* deletion */
        out[w++] = (uint8_t)ACGT[(hi >> 1) & 3];         /* insertion */
        continue;
      }
      out[w++] = (hi < sub_thr) ? (uint8_t)ACGT[(x >> 20) & 3] : ref[src];
      ++src;
    }
  }
  return 0;
}
