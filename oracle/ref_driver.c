/*
 * oracle/ref_driver.c -- TEST INFRASTRUCTURE, not product code.
 *
 * A thin C-callable driver around the UNMODIFIED reference engine.  It is
 * compiled together with /root/reference/src/kmer_pos.c and kmer_util.c (the
 * sources are referenced where they lie, never copied) into
 * oracle/_ref/libkmer_ref.so by oracle/Makefile.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load it.
 *
 * What is called straight from the reference:
 *   seq_to_hash          kmer_pos.c:66-98   (index build)
 *   seq_kmer_positions   kmer_pos.c:110-136 (probe)
 *   sort_kmer_pos        kmer_pos.c:21-33   (do.sort)
 *   clear_kmer_h         kmer_pos.c:10-19
 * What has to be restated because it lives in the R glue (needs R.h, absent
 * here): the extraction loop of kmer_positions, kmer_hash.c:1095-1124, and the
 * decoder kmer_seq, kmer_hash.c:123-133 with NUC = {A,C,T,G} (kmer_hash.c:21).
 *
 * khash bucket order is not semantic, so every extraction is also offered in
 * CANONICAL form: k-mers ordered by ascending uint64 key, the 1-based k-mer
 * index i remapped to that rank, rows grouped by i (positions / pairs keep
 * their within-k-mer order, which is deterministic).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "kmer_pos.h"   /* from -I/root/reference/src */
#include "kmer_util.h"

typedef struct {
  khash_ptr hp;          /* the reference's own handle struct (kmer_pos.h:43-48) */
  uint64_t n_pos;        /* sum of list lengths */
  uint64_t n_pairs;      /* sum n(n-1)/2 */
} ref_index;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* make_kmer_h_index minus the SEXP handling (kmer_hash.c:506-540).
 * Returns NULL when the R entry would have raised an error (guards at
 * kmer_hash.c:515-520); *err says which: 1 = k out of range, 2 = seq too short. */
ref_index *ref_build(const char *seq, int k, int do_sort, int *err, double *seconds) {
  if (err) *err = 0;
  if (k < 1 || k > 32) { if (err) *err = 1; return NULL; }
  if ((int64_t)strlen(seq) <= (int64_t)k) { if (err) *err = 2; return NULL; }
  ref_index *ri = calloc(1, sizeof(ref_index));
  ri->hp.k = k;
  ri->hp.hash = kh_init(kmer_h);
  double t0 = now_s();
  ri->hp.kmer_count = (size_t)seq_to_hash(seq, k, ri->hp.hash);
  if (do_sort) sort_kmer_pos(&ri->hp);
  double t1 = now_s();
  if (seconds) *seconds = t1 - t0;
  khash_t(kmer_h) *h = ri->hp.hash;
  for (khiter_t it = kh_begin(h); it != kh_end(h); ++it) {
    if (!kh_exist(h, it)) continue;
    uint64_t n = kh_val(h, it).v.n;
    ri->n_pos += n;
    ri->n_pairs += n * (n - 1) / 2;
  }
  return ri;
}

/* Same, but without the R-level length guard: lets tests drive the C core on
 * the short adversarial strings of SURVEY Appendix B. */
ref_index *ref_build_core(const char *seq, int k) {
  ref_index *ri = calloc(1, sizeof(ref_index));
  ri->hp.k = k;
  ri->hp.hash = kh_init(kmer_h);
  ri->hp.kmer_count = (size_t)seq_to_hash(seq, k, ri->hp.hash);
  khash_t(kmer_h) *h = ri->hp.hash;
  for (khiter_t it = kh_begin(h); it != kh_end(h); ++it) {
    if (!kh_exist(h, it)) continue;
    uint64_t n = kh_val(h, it).v.n;
    ri->n_pos += n;
    ri->n_pairs += n * (n - 1) / 2;
  }
  return ri;
}

void ref_free(ref_index *ri) {
  if (!ri) return;
  if (ri->hp.hash) clear_kmer_h(ri->hp.hash);
  free(ri);
}

void ref_sizes(const ref_index *ri, uint64_t *U, uint64_t *N, uint64_t *P, uint64_t *buckets,
               uint64_t *new_kmers) {
  if (U) *U = kh_size(ri->hp.hash);
  if (N) *N = ri->n_pos;
  if (P) *P = ri->n_pairs;
  if (buckets) *buckets = kh_n_buckets(ri->hp.hash);
  if (new_kmers) *new_kmers = ri->hp.kmer_count;
}

/* decoder restated from kmer_seq (kmer_hash.c:123-133): last base in the low
 * two bits, alphabet A,C,T,G. */
static void decode_kmer(char *dst, int k, uint64_t key) {
  static const char alphabet[4] = {'A', 'C', 'T', 'G'};
  dst[k] = 0;
  for (int b = k - 1; b >= 0; --b) { dst[b] = alphabet[key & 3u]; key >>= 2; }
}

/*
 * The kmer_positions loop (kmer_hash.c:1095-1124) restated, writing into
 * caller-provided arrays instead of kvecs.  RAW = bucket order, exactly what R
 * would receive.  Any output pointer may be NULL (its opt.flag bit is off).
 *   keys   [U]        uint64 key per k-mer, in emission order (not in R's
 *                     result, needed for canonicalisation)
 *   kmers  [U*(k+1)]  NUL-terminated upper-case strings      (flag 1)
 *   pos    [2*N]      interleaved (i,pos)                    (flag 2)
 *   pairs  [3*P]      interleaved (i,x,y), x before y        (flag 4)
 *   counts [U]                                               (flag 8)
 */
void ref_extract_raw(const ref_index *ri, uint64_t *keys, char *kmers, int *pos, int *pairs,
                     int *counts, double *seconds) {
  khash_t(kmer_h) *h = ri->hp.hash;
  const int k = ri->hp.k;
  size_t np = 0, npp = 0;
  int i = 0;
  double t0 = now_s();
  for (khiter_t it = kh_begin(h); it != kh_end(h); ++it) {
    if (!kh_exist(h, it)) continue;
    const kmer_pos_t *kv = &kh_val(h, it);
    if (keys) keys[i] = kv->kmer;
    if (kmers) decode_kmer(kmers + (size_t)i * (size_t)(k + 1), k, kv->kmer);
    if (counts) counts[i] = (int)kv->v.n;
    ++i;                                   /* 1-based from here on (kmer_hash.c:1105) */
    if (!pos && !pairs) continue;
    for (size_t a = 0; a < kv->v.n; ++a) {
      if (pos) { pos[np++] = i; pos[np++] = kv->v.a[a]; }
      if (pairs)
        for (size_t b = a + 1; b < kv->v.n; ++b) {
          pairs[npp++] = i; pairs[npp++] = kv->v.a[a]; pairs[npp++] = kv->v.a[b];
        }
    }
  }
  if (seconds) *seconds = now_s() - t0;
}

typedef struct { uint64_t key; uint32_t slot; } key_slot;
static int cmp_key_slot(const void *a, const void *b) {
  uint64_t x = ((const key_slot *)a)->key, y = ((const key_slot *)b)->key;
  return (x > y) - (x < y);
}

/* Canonical extraction: same fields, k-mers by ascending key. */
void ref_extract_canonical(const ref_index *ri, uint64_t *keys, char *kmers, int *pos, int *pairs,
                           int *counts) {
  khash_t(kmer_h) *h = ri->hp.hash;
  const int k = ri->hp.k;
  size_t U = kh_size(h);
  key_slot *ord = malloc((U ? U : 1) * sizeof(key_slot));
  size_t u = 0;
  for (khiter_t it = kh_begin(h); it != kh_end(h); ++it)
    if (kh_exist(h, it)) { ord[u].key = kh_val(h, it).kmer; ord[u].slot = (uint32_t)it; ++u; }
  qsort(ord, U, sizeof(key_slot), cmp_key_slot);
  size_t np = 0, npp = 0;
  for (u = 0; u < U; ++u) {
    const kmer_pos_t *kv = &kh_val(h, ord[u].slot);
    int i = (int)(u + 1);
    if (keys) keys[u] = kv->kmer;
    if (kmers) decode_kmer(kmers + u * (size_t)(k + 1), k, kv->kmer);
    if (counts) counts[u] = (int)kv->v.n;
    for (size_t a = 0; a < kv->v.n; ++a) {
      if (pos) { pos[np++] = i; pos[np++] = kv->v.a[a]; }
      if (pairs)
        for (size_t b = a + 1; b < kv->v.n; ++b) {
          pairs[npp++] = i; pairs[npp++] = kv->v.a[a]; pairs[npp++] = kv->v.a[b];
        }
    }
  }
  free(ord);
}

/* ---- order-sensitive digests of the canonical outputs (full-size golden fixtures) -----------------
 * The BASELINE-size outputs (2N ints of pos at 250 Mbp, 3P ints of pair.pos at P > 10^9) are too large to
 * commit, so tests/golden/make_fullsize.py records digests of them instead: for the flattened value stream
 * v_0, v_1, ... : n, sum v_t and sum v_t * (2t + 1), all mod 2^64.  tests/ recompute the same three numbers
 * from the CUDA path's output (on the device).  Nothing is materialised here: the streams are generated
 * straight from the reference's own tables. */
typedef struct { uint64_t n, sum, wsum; } ref_dig;
static inline void dig_add(ref_dig *d, uint64_t v) { d->sum += v; d->wsum += v * (2 * d->n + 1); d->n++; }

/* bind[0] = sum key * count, bind[1] = sum key * pos over all (k-mer, position) pairs, mod 2^64: ORDER-INDEPENDENT sums
 * that tie counts and positions to their k-mer.  A sharded index (k-mers spread over several GPUs in no global order)
 * is checked against them together with the n / sum parts of the digests above (bench.py --gpus N, tests). */
void ref_digest_canonical(const ref_index *ri, ref_dig *keys, ref_dig *counts, ref_dig *pos, ref_dig *pairs, uint64_t *bind) {
  khash_t(kmer_h) *h = ri->hp.hash;
  size_t U = kh_size(h);
  key_slot *ord = malloc((U ? U : 1) * sizeof(key_slot));
  size_t u = 0;
  for (khiter_t it = kh_begin(h); it != kh_end(h); ++it)
    if (kh_exist(h, it)) { ord[u].key = kh_val(h, it).kmer; ord[u].slot = (uint32_t)it; ++u; }
  qsort(ord, U, sizeof(key_slot), cmp_key_slot);
  if (keys) memset(keys, 0, sizeof *keys);
  if (counts) memset(counts, 0, sizeof *counts);
  if (pos) memset(pos, 0, sizeof *pos);
  if (pairs) memset(pairs, 0, sizeof *pairs);
  if (bind) bind[0] = bind[1] = 0;
  for (u = 0; u < U; ++u) {
    const kmer_pos_t *kv = &kh_val(h, ord[u].slot);
    const uint64_t i = u + 1;                       /* canonical 1-based k-mer number */
    if (keys) dig_add(keys, kv->kmer);
    if (counts) dig_add(counts, (uint64_t)kv->v.n);
    if (bind) bind[0] += kv->kmer * (uint64_t)kv->v.n;
    for (size_t a = 0; a < kv->v.n; ++a) {          /* same loops as kmer_hash.c:1108-1120 */
      if (bind) bind[1] += kv->kmer * (uint64_t)(int64_t)kv->v.a[a];
      if (pos) { dig_add(pos, i); dig_add(pos, (uint64_t)(int64_t)kv->v.a[a]); }
      if (pairs)
        for (size_t b = a + 1; b < kv->v.n; ++b) {
          dig_add(pairs, i); dig_add(pairs, (uint64_t)(int64_t)kv->v.a[a]); dig_add(pairs, (uint64_t)(int64_t)kv->v.a[b]);
        }
    }
  }
  free(ord);
}

/* digest of the rows seq_kmer_positions (kmer_pos.c:110-136) returns, in its own order */
int64_t ref_query_digest(const ref_index *ri, const char *seq, int k, ref_dig *rows, double *seconds) {
  double t0 = now_s();
  kmer_ppos pp = seq_kmer_positions(ri->hp.hash, seq, k);
  if (seconds) *seconds = now_s() - t0;
  memset(rows, 0, sizeof *rows);
  for (size_t t = 0; t < pp.n; ++t) dig_add(rows, (uint64_t)(int64_t)pp.a[t]);
  free(pp.a);
  return (int64_t)(pp.n / 2);
}

/* seq_kmer_positions (kmer_pos.c:110-136) called as is; rows are (i,j)
 * interleaved, already in a deterministic order.  The result is malloc'd by the
 * reference's kvec; ref_query_free releases it. */
int64_t ref_query(const ref_index *ri, const char *seq, int k, int **rows, double *seconds) {
  double t0 = now_s();
  kmer_ppos pp = seq_kmer_positions(ri->hp.hash, seq, k);
  if (seconds) *seconds = now_s() - t0;
  *rows = pp.a;
  return (int64_t)(pp.n / 2);
}
void ref_query_free(int *rows) { free(rows); }

/* kmer_pair_pos restated from the R glue (kmer_hash.c:1174-1203) on the reference's own tables: for
 * every k-mer of a that b holds (kh_get, :1185), rows (a_pos, b_pos), a position outer, b position
 * inner (:1190-1195).  Two repairs, both stated in SURVEY.md 8f: buckets of a that hold nothing are
 * skipped (the original omits kh_exist on it_a and crashes, test.R:330-331) and kh_get's "not found"
 * is tested against kh_end.  Canonical form: a's k-mers in ascending key order. */
int64_t ref_pairs_join(const ref_index *ra, const ref_index *rb, int **rows_out) {
  khash_t(kmer_h) *ha = ra->hp.hash, *hb = rb->hp.hash;
  const size_t U = kh_size(ha);
  key_slot *ord = malloc((U ? U : 1) * sizeof(key_slot));
  size_t w = 0;
  for (khiter_t it = kh_begin(ha); it != kh_end(ha); ++it)
    if (kh_exist(ha, it)) { ord[w].key = kh_key(ha, it); ord[w].slot = (uint32_t)it; ++w; }
  qsort(ord, U, sizeof(key_slot), cmp_key_slot);
  int64_t n = 0;
  for (int pass = 0; pass < 2; ++pass) {
    int *rows = pass ? malloc((size_t)(n ? n : 1) * 2 * sizeof(int)) : NULL;
    int64_t r = 0;
    for (size_t u = 0; u < U; ++u) {
      khiter_t ib = kh_get(kmer_h, hb, ord[u].key);
      if (ib == kh_end(hb)) continue;
      const kmer_pos_t *av = &kh_val(ha, ord[u].slot), *bv = &kh_val(hb, ib);
      if (rows)
        for (size_t i = 0; i < av->v.n; ++i)
          for (size_t j = 0; j < bv->v.n; ++j) { rows[2 * r] = av->v.a[i]; rows[2 * r + 1] = bv->v.a[j]; ++r; }
      else
        r += (int64_t)(av->v.n * bv->v.n);
    }
    if (!pass) n = r; else *rows_out = rows;
  }
  free(ord);
  return n;
}

/* ---- count.kmers ---------------------------------------------------------------
 * seq_to_counts / kmer_count_insert live in the R glue file (kmer_hash.c:185-252, which needs R.h for its
 * warning()), so they are RESTATED here, on the reference's own khash type and with the reference's own
 * init_kmer / UPDATE_OFFSET / LC (kmer_util.h, kmer_util.c: called, not restated): per distinct k-mer an array
 * of source_n ints in kmer_pos_t.v (kmer_hash.c:196-203), column `source` bumped once per window (:205), the
 * window loop that of seq_to_hash (:236-250).  Pinned by tests/test_oracle.py: with one source the counts must
 * equal the list lengths of the index the UNMODIFIED seq_to_hash builds from the same sequence.
 * The result is read back through ref_extract_canonical, exactly as R's kmer.pos reads a count table (it walks
 * v.a[0..v.n), which now holds the counters). */
static int ref_count_insert(uint64_t kmer, khash_t(kmer_h) *hash, size_t source, size_t source_n) {
  if (source >= source_n) return -1;
  int ret = 0, new_kmer = 0;
  khiter_t it = kh_get(kmer_h, hash, kmer);
  if (it == kh_end(hash)) {
    it = kh_put(kmer_h, hash, kmer, &ret);
    if (it == kh_end(hash)) return -1;
    kv_init(kh_val(hash, it).v);
    kh_val(hash, it).kmer = kmer;
    kh_val(hash, it).v.a = calloc(source_n, sizeof(int));
    kh_val(hash, it).v.m = source_n;
    kh_val(hash, it).v.n = source_n;
    new_kmer = 1;
  }
  kh_val(hash, it).v.a[source]++;
  return new_kmer;
}

ref_index *ref_count_new(int k) {
  ref_index *ri = calloc(1, sizeof(ref_index));
  ri->hp.k = k;
  ri->hp.hash = kh_init(kmer_h);
  return ri;
}

/* one sequence of count_kmers' loop (kmer_hash.c:580-587): sequences with length <= k are skipped there */
int ref_count_add(ref_index *ri, const char *seq, int source, int source_n) {
  const int k = ri->hp.k;
  if ((int64_t)strlen(seq) <= (int64_t)k) return 0;
  khash_t(kmer_h) *hash = ri->hp.hash;
  size_t i = 0;
  uint64_t offset = 0;
  int word_count = 0;
  const uint64_t mask = k < 32 ? (((uint64_t)1) << (2 * k)) - 1 : ~(uint64_t)0;
  while (seq[i]) {
    i = init_kmer(seq, i, &offset, k);
    if (!seq[i]) break;
    int r = ref_count_insert(offset & mask, hash, (size_t)source, (size_t)source_n);
    if (r < 0) return r;
    word_count += r;
    while (seq[i] && LC(seq[i]) != 'n') {
      offset = UPDATE_OFFSET(offset, seq[i]);
      ++i;
      r = ref_count_insert(offset & mask, hash, (size_t)source, (size_t)source_n);
      if (r < 0) return r;
      word_count += r;
    }
  }
  if (word_count > 0) ri->hp.kmer_count += (size_t)word_count;
  ri->n_pos = (uint64_t)kh_size(hash) * (uint64_t)source_n;                       /* rows kmer.pos(ptr, 2) returns */
  ri->n_pairs = (uint64_t)kh_size(hash) * (uint64_t)source_n * (uint64_t)(source_n - 1) / 2;
  return word_count;
}

/* The insertion stream of seq_to_hash, observed without touching the reference:
 * build an index of the sequence, then read every (key,pos) back and order by
 * position.  Gives the exact multiset of windows the reference emits. */
int64_t ref_window_stream(const char *seq, int k, uint64_t **keys_out, int **pos_out) {
  ref_index *ri = ref_build_core(seq, k);
  khash_t(kmer_h) *h = ri->hp.hash;
  size_t n = ri->n_pos, w = 0;
  key_slot *tmp = malloc((n ? n : 1) * sizeof(key_slot));
  for (khiter_t it = kh_begin(h); it != kh_end(h); ++it) {
    if (!kh_exist(h, it)) continue;
    const kmer_pos_t *kv = &kh_val(h, it);
    for (size_t a = 0; a < kv->v.n; ++a) { tmp[w].key = (uint64_t)(uint32_t)kv->v.a[a]; tmp[w].slot = (uint32_t)it; ++w; }
  }
  qsort(tmp, n, sizeof(key_slot), cmp_key_slot);      /* by position (unique) */
  uint64_t *keys = malloc((n ? n : 1) * sizeof(uint64_t));
  int *pos = malloc((n ? n : 1) * sizeof(int));
  for (w = 0; w < n; ++w) { pos[w] = (int)tmp[w].key; keys[w] = kh_val(h, tmp[w].slot).kmer; }
  free(tmp);
  ref_free(ri);
  *keys_out = keys; *pos_out = pos;
  return (int64_t)n;
}
void ref_free_buf(void *p) { free(p); }
