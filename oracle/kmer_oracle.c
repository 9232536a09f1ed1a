/*
 * oracle/kmer_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A plain-C CPU restatement of kmer_hasheR's k-mer position-index path, used
 * only as the checker for the CUDA library (tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg).  Nothing under kmer_hasher_b200/ may
 * import, link or execute it.
 *
 * Parity pin: this file is validated against the UNMODIFIED reference engine
 * (oracle/_ref/libkmer_ref.so = /root/reference/src/kmer_pos.c + kmer_util.c
 * compiled as they lie) by tests/test_oracle.py, and against fixtures under
 * tests/golden/ that were generated from that engine by
 * tests/golden/make_golden.py.  The reference ships no golden vectors of its
 * own for this path.
 *
 * What is restated, and from where (paths relative to /root/reference):
 *   base code (c>>1)&3, roll off<<2|code      src/kmer_util.h:8
 *   window breaker (c|0x20)=='n'              src/kmer_util.h:10, kmer_util.c:5,22
 *   prime-a-window / skip-N control flow      src/kmer_util.c:4-8, 18-32
 *   stream of (key, 1-based start) windows    src/kmer_pos.c:66-98
 *   mask special case at k=32                 src/kmer_pos.c:77
 *   probe stream, i = 1-based END of window   src/kmer_pos.c:110-136
 *   extraction (i,pos) / (i,x,y) / counts     src/kmer_hash.c:1095-1124
 *   decoder, alphabet A,C,T,G                 src/kmer_hash.c:21, 123-133
 *   argument guards                           src/kmer_hash.c:515-520, 1163
 *
 * The reference keeps k-mers in a khash table whose iteration order is not
 * semantic; this oracle keeps them ordered by ascending uint64 key (the
 * canonical order every comparison uses).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int k;
  uint64_t U, N, P;
  uint64_t *ukeys;   /* [U]   distinct keys, ascending          */
  uint32_t *start;   /* [U+1] first slot of each k-mer in pos[] */
  int32_t *pos;      /* [N]   1-based starts, ascending per k-mer */
} ko_index;

static inline int is_breaker(char c) { return (c | 0x20) == 'n'; }
static inline uint64_t roll(uint64_t w, char c) { return (w << 2) | (uint64_t)((c >> 1) & 3); }
static inline uint64_t key_mask(int k) { return k < 32 ? (((uint64_t)1 << (2 * k)) - 1) : ~(uint64_t)0; }

/*
 * Walk the sequence exactly like seq_to_hash / seq_kmer_positions do and hand
 * every emitted window to `sink(key, end, ctx)`, where `end` is the index one
 * past the window (the reference's `i`).  Index position = end+1-k, probe
 * coordinate = end.  The NUL test of the reference becomes `i < len`.
 */
typedef void (*window_sink)(uint64_t key, int64_t end, void *ctx);

static void walk_windows(const char *s, int64_t len, int k, window_sink sink, void *ctx) {
  const uint64_t mask = key_mask(k);
  int64_t i = 0;
  uint64_t w = 0;
  while (i < len) {
    /* prime: find the next k consecutive non-breaker bytes starting at i
       (init_kmer, kmer_util.c:18-32; skip_n, kmer_util.c:4-8) */
    int64_t j = 0;
    while (i < len) {
      w = 0;
      for (j = 0; j < k && i + j < len && !is_breaker(s[i + j]); ++j) w = roll(w, s[i + j]);
      if (i + j >= len || j == k) break;
      i += j;
      while (i < len && is_breaker(s[i])) ++i;
      j = 0;
    }
    i += j;
    /* a freshly primed window that ends exactly at the terminator is NOT
       emitted (kmer_pos.c:81-83 and :123-125) */
    if (i >= len) break;
    sink(w & mask, i, ctx);
    /* roll until a breaker or the end (kmer_pos.c:88-95) */
    while (i < len && !is_breaker(s[i])) {
      w = roll(w, s[i]);
      ++i;
      sink(w & mask, i, ctx);
    }
  }
}

typedef struct { uint64_t *keys; int32_t *pos; int64_t n; int k; } collect_ctx;
static void collect_sink(uint64_t key, int64_t end, void *vctx) {
  collect_ctx *c = (collect_ctx *)vctx;
  if (c->keys) c->keys[c->n] = key;
  if (c->pos) c->pos[c->n] = (int32_t)(end + 1 - c->k);
  c->n++;
}

/* The (key, 1-based start) stream in emission order.  keys/pos may be NULL to
 * only count.  Capacity needed: max(0, len-k+1). */
int64_t ko_windows(const char *seq, int64_t len, int k, uint64_t *keys, int32_t *pos) {
  collect_ctx c = {keys, pos, 0, k};
  walk_windows(seq, len, k, collect_sink, &c);
  return c.n;
}

/* stable LSD radix sort of (key,pos) by key, 16-bit digits */
static void sort_records(uint64_t *keys, int32_t *pos, int64_t n, int k) {
  if (n < 2) return;
  uint64_t *k2 = malloc((size_t)n * sizeof(uint64_t));
  int32_t *p2 = malloc((size_t)n * sizeof(int32_t));
  size_t *cnt = malloc(65537 * sizeof(size_t));
  uint64_t *ka = keys, *kb = k2;
  int32_t *pa = pos, *pb = p2;
  for (int shift = 0; shift < 2 * k; shift += 16) {
    memset(cnt, 0, 65537 * sizeof(size_t));
    for (int64_t i = 0; i < n; ++i) cnt[((ka[i] >> shift) & 0xFFFF) + 1]++;
    for (int d = 0; d < 65536; ++d) cnt[d + 1] += cnt[d];
    for (int64_t i = 0; i < n; ++i) {
      size_t dst = cnt[(ka[i] >> shift) & 0xFFFF]++;
      kb[dst] = ka[i];
      pb[dst] = pa[i];
    }
    uint64_t *tk = ka; ka = kb; kb = tk;
    int32_t *tp = pa; pa = pb; pb = tp;
  }
  if (ka != keys) { memcpy(keys, ka, (size_t)n * sizeof(uint64_t)); memcpy(pos, pa, (size_t)n * sizeof(int32_t)); }
  free(k2); free(p2); free(cnt);
}

static ko_index *index_from_records(uint64_t *keys, int32_t *pos, int64_t n, int k) {
  ko_index *ix = calloc(1, sizeof(ko_index));
  ix->k = k;
  ix->N = (uint64_t)n;
  sort_records(keys, pos, n, k);
  uint64_t U = 0;
  for (int64_t i = 0; i < n; ++i) if (i == 0 || keys[i] != keys[i - 1]) ++U;
  ix->U = U;
  ix->ukeys = malloc((U ? U : 1) * sizeof(uint64_t));
  ix->start = malloc((U + 1) * sizeof(uint32_t));
  uint64_t u = 0;
  for (int64_t i = 0; i < n; ++i)
    if (i == 0 || keys[i] != keys[i - 1]) { ix->ukeys[u] = keys[i]; ix->start[u] = (uint32_t)i; ++u; }
  ix->start[U] = (uint32_t)n;
  ix->pos = pos;
  for (u = 0; u < U; ++u) { uint64_t c = ix->start[u + 1] - ix->start[u]; ix->P += c * (c - 1) / 2; }
  free(keys);
  return ix;
}

/* make.kmer.hash.  err: 0 ok, 1 = k outside [1,32] (kmer_hash.c:515),
 * 2 = len <= k (kmer_hash.c:519).  `guard`=0 skips the length guard so the C
 * core can be exercised on short strings like the reference's core can. */
ko_index *ko_build(const char *seq, int64_t len, int k, int guard, int *err) {
  if (err) *err = 0;
  if (k < 1 || k > 32) { if (err) *err = 1; return NULL; }
  if (guard && len <= k) { if (err) *err = 2; return NULL; }
  int64_t cap = len - k + 1 > 0 ? len - k + 1 : 1;
  uint64_t *keys = malloc((size_t)cap * sizeof(uint64_t));
  int32_t *pos = malloc((size_t)cap * sizeof(int32_t));
  int64_t n = ko_windows(seq, len, k, keys, pos);
  return index_from_records(keys, pos, n, k);
}

/* The same index from an explicit record stream (used by the multi-GPU host
 * tests: records routed to one owner, in source order). Takes copies. */
ko_index *ko_build_from_records(const uint64_t *keys_in, const int32_t *pos_in, int64_t n, int k) {
  uint64_t *keys = malloc((size_t)(n ? n : 1) * sizeof(uint64_t));
  int32_t *pos = malloc((size_t)(n ? n : 1) * sizeof(int32_t));
  memcpy(keys, keys_in, (size_t)n * sizeof(uint64_t));
  memcpy(pos, pos_in, (size_t)n * sizeof(int32_t));
  return index_from_records(keys, pos, n, k);
}

void ko_free(ko_index *ix) {
  if (!ix) return;
  free(ix->ukeys); free(ix->start); free(ix->pos); free(ix);
}

void ko_sizes(const ko_index *ix, uint64_t *U, uint64_t *N, uint64_t *P) {
  if (U) *U = ix->U;
  if (N) *N = ix->N;
  if (P) *P = ix->P;
}

/* kmer.pos in canonical order; any pointer may be NULL. Layouts as in
 * kmer_positions: kmers U*(k+1) chars, pos 2N ints (i,pos), pairs 3P ints
 * (i,x,y) with the first member of the pair varying slowest, counts U ints. */
void ko_extract(const ko_index *ix, uint64_t *keys, char *kmers, int32_t *pos, int32_t *pairs,
                int32_t *counts) {
  static const char alphabet[4] = {'A', 'C', 'T', 'G'};
  const int k = ix->k;
  size_t np = 0, npp = 0;
  for (uint64_t u = 0; u < ix->U; ++u) {
    const uint32_t a0 = ix->start[u], a1 = ix->start[u + 1];
    const int32_t i = (int32_t)(u + 1);
    if (keys) keys[u] = ix->ukeys[u];
    if (kmers) {
      char *dst = kmers + u * (size_t)(k + 1);
      uint64_t key = ix->ukeys[u];
      dst[k] = 0;
      for (int b = k - 1; b >= 0; --b) { dst[b] = alphabet[key & 3u]; key >>= 2; }
    }
    if (counts) counts[u] = (int32_t)(a1 - a0);
    for (uint32_t a = a0; a < a1; ++a) {
      if (pos) { pos[np++] = i; pos[np++] = ix->pos[a]; }
      if (pairs)
        for (uint32_t b = a + 1; b < a1; ++b) { pairs[npp++] = i; pairs[npp++] = ix->pos[a]; pairs[npp++] = ix->pos[b]; }
    }
  }
}

typedef struct { const ko_index *ix; int32_t *rows; int64_t n, cap; int count_only; } query_ctx;
static void query_sink(uint64_t key, int64_t end, void *vctx) {
  query_ctx *q = (query_ctx *)vctx;
  const ko_index *ix = q->ix;
  uint64_t lo = 0, hi = ix->U;                   /* stand-in for kh_get (kmer_pos.c:55-60) */
  while (lo < hi) { uint64_t mid = (lo + hi) >> 1; if (ix->ukeys[mid] < key) lo = mid + 1; else hi = mid; }
  if (lo == ix->U || ix->ukeys[lo] != key) return;
  for (uint32_t a = ix->start[lo]; a < ix->start[lo + 1]; ++a) {   /* pair_positions_push, kmer_pos.c:101-108 */
    if (!q->count_only) {
      if (q->n + 1 > q->cap) { q->cap = q->cap ? q->cap * 2 : 1024; q->rows = realloc(q->rows, (size_t)q->cap * 2 * sizeof(int32_t)); }
      q->rows[2 * q->n] = (int32_t)end;
      q->rows[2 * q->n + 1] = ix->pos[a];
    }
    q->n++;
  }
}

/* seq.kmer.pos at the C level (k up to 32, no guard; the R entry's guard is
 * kmer_hash.c:1163).  Returns the number of (i,j) rows; *rows is malloc'd. */
int64_t ko_query(const ko_index *ix, const char *seq, int64_t len, int k, int32_t **rows) {
  query_ctx q = {ix, NULL, 0, 0, rows == NULL};
  walk_windows(seq, len, k, query_sink, &q);
  if (rows) *rows = q.rows;
  return q.n;
}
void ko_free_buf(void *p) { free(p); }
