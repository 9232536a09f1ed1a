/*
 * kmer_hash.c -- R glue of the B200 k-mer position index: the drop-in replacement for the three
 * index entry points of the reference's src/kmer_hash.c.  It exports the same .Call names with the
 * same arity, argument types, return layouts, external-pointer tag and finaliser convention, so the
 * reference's kmer_hash.R (make.kmer.hash / kmer.pos / seq.kmer.pos, kmer_hash.R:5-28) works
 * unchanged when this file is built as src/kmer_hash.so.  All work is done by libkmergpu
 * (include/kmergpu.h); nothing here builds a hash table and there is no CPU path.
 *
 *   make_kmer_h_index(seq, k, do.sort)      replaces src/kmer_hash.c:506-540
 *   kmer_positions(ptr, opt.flag)           replaces src/kmer_hash.c:1054-1147
 *   sequence_kmer_positions(ptr, seq, k)    replaces src/kmer_hash.c:1151-1172
 *   kmer_pair_pos(ptr.a, ptr.b)             replaces src/kmer_hash.c:1174-1203 (which crashes: test.R:330-331)
 *   R_init_kmer_hash                        replaces src/kmer_hash.c:1221-1224
 *
 * Differences a user can observe: k-mers come out in the order of a mix of their 2-bit key (ascending key
 * with do.sort = TRUE) instead of khash bucket order (not semantic); results too large for an R matrix raise a clean error before anything is
 * allocated (the reference overflows or leaks, README "pair.pos"); do.sort is a no-op (lists are
 * always ascending).  Set KMERGPU_ALLOW_K32=1 to lift the reference's k <= 31 limit of seq.kmer.pos
 * (its C core handles 32; the limit is only in its R entry, src/kmer_hash.c:1163).
 *
 * Built against R's headers on a machine with R, or against rstub/ (a minimal R C-API stand-in)
 * where R is absent, which is how tests/test_rglue.py drives it.
 */
#include <R.h>
#include <Rinternals.h>
#include <limits.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "kmergpu.h"

#define KMER_HASH_TAG "kmer_hash_250930" /* src/kmer_hash.c:22 */
enum { F_KMER = 1, F_POS = 2, F_PAIRS = 4, F_COUNT = 8 }; /* pos_opt_flags, src/kmer_hash.c:17 */
static const char *const field_names[4] = {"kmer", "pos", "pair.pos", "count"}; /* :18 */

/* what the external pointer addresses (the reference's khash_ptr, src/kmer_pos.h:43-48) */
typedef struct {
  kmg_index *index;
  int k;
} kmer_handle;

/* Tag-checked unwrap that never raises: safe inside the finaliser. */
static kmer_handle *handle_or_null(SEXP ptr) {
  if (TYPEOF(ptr) != EXTPTRSXP) return NULL;
  SEXP tag = R_ExternalPtrTag(ptr);
  if (TYPEOF(tag) != STRSXP || length(tag) != 1 || strcmp(CHAR(STRING_ELT(tag, 0)), KMER_HASH_TAG) != 0) return NULL;
  return (kmer_handle *)R_ExternalPtrAddr(ptr);
}

/* The same with the reference's errors (extract_khash_ptr, src/kmer_hash.c:491-503). */
static kmer_handle *handle_or_error(SEXP ptr) {
  if (TYPEOF(ptr) != EXTPTRSXP) error("ptr_r should be an external pointer");
  kmer_handle *h = handle_or_null(ptr);
  if (!h) {
    SEXP tag = R_ExternalPtrTag(ptr);
    if (TYPEOF(tag) != STRSXP || length(tag) != 1 || strcmp(CHAR(STRING_ELT(tag, 0)), KMER_HASH_TAG) != 0)
      error("External pointer has incorrect tag");
    error("external pointer is NULL");
  }
  if (!h->index) error("the k-mer index has been released");
  return h;
}

/* Runs from the garbage collector or at exit: frees the device index once, tolerates a cleared
 * pointer, never calls error() (convention stated at src/kmer_hash.c:37-40). */
static void finalise_handle(SEXP ptr) {
  kmer_handle *h = handle_or_null(ptr);
  if (!h) return;
  if (h->index) {
    kmg_free(h->index);
    h->index = NULL;
  }
  free(h);
  R_ClearExternalPtr(ptr);
}

SEXP make_kmer_h_index(SEXP seq_r, SEXP k_r, SEXP sort_pos_r) {
  if (TYPEOF(seq_r) != STRSXP || length(seq_r) < 1) error("seq_r should be a character vector of length at least one");
  if (TYPEOF(k_r) != INTSXP || length(k_r) < 1) error("k_r must be an integer vector of length at least one");
  if (TYPEOF(sort_pos_r) != INTSXP || length(sort_pos_r) < 1) error("sort_pos_r must be an integer vector of length at least one");
  const int k = INTEGER(k_r)[0];
  if (k < 1 || k > KMG_MAX_K) error("k must be a positive integer less than 1+MAX_K");
  SEXP s = STRING_ELT(seq_r, 0); /* like the reference, only the first element is indexed */
  const int len = length(s);
  if (len <= k) error("the length of the sequence must be at least k");

  kmer_handle *h = (kmer_handle *)calloc(1, sizeof *h);
  if (!h) error("out of memory");
  h->k = k;
  /* do.sort: position lists are always ascending, so there is nothing to sort there (src/kmer_pos.c:21-33 is a
   * no-op in the reference too); it selects the order of the k-mers instead: TRUE = ascending key, FALSE (the R
   * default) = the faster grouped build.  Neither order is semantic (the reference's is khash bucket order). */
  const int order = INTEGER(sort_pos_r)[0] ? KMG_ORDER_SORTED : KMG_ORDER_GROUPED;
  if (kmg_build_ordered(CHAR(s), (int64_t)len, k, order, &h->index) != KMG_OK) {
    free(h);
    error("make.kmer.hash failed: %s", kmg_last_error());
  }
  SEXP tag = PROTECT(allocVector(STRSXP, 1));
  SET_STRING_ELT(tag, 0, mkChar(KMER_HASH_TAG));
  SEXP ptr = PROTECT(R_MakeExternalPtr(h, tag, R_NilValue));
  R_RegisterCFinalizerEx(ptr, finalise_handle, TRUE);
  UNPROTECT(2);
  return ptr;
}

SEXP kmer_positions(SEXP ptr_r, SEXP opt_flag_r) {
  kmer_handle *h = handle_or_error(ptr_r);
  if (TYPEOF(opt_flag_r) != INTSXP || length(opt_flag_r) != 1) error("opt_flag_r should be an integer vector of length 1");
  const unsigned flag = (unsigned)asInteger(opt_flag_r);

  uint64_t U = 0, N = 0, P = 0;
  if (kmg_sizes(h->index, &U, &N, &P) != KMG_OK) error("kmer.pos failed: %s", kmg_last_error());
  /* size checks BEFORE any allocation: R vectors/matrix extents are int */
  if ((flag & (F_KMER | F_COUNT)) && U > (uint64_t)INT_MAX) error("%llu distinct k-mers do not fit an R vector", (unsigned long long)U);
  if ((flag & F_POS) && N > (uint64_t)INT_MAX) error("%llu positions do not fit an R matrix", (unsigned long long)N);
  if ((flag & F_PAIRS) && P > (uint64_t)INT_MAX)
    error("pair.pos would have %llu rows, more than an R matrix can hold (2^31-1); index a shorter region", (unsigned long long)P);

  SEXP ret = PROTECT(allocVector(VECSXP, 4));
  SEXP names = PROTECT(allocVector(STRSXP, 4));
  for (int i = 0; i < 4; ++i) SET_STRING_ELT(names, i, mkChar(field_names[i]));
  setAttrib(ret, R_NamesSymbol, names);
  UNPROTECT(1);

  const char *failed = NULL;
  if (flag & F_KMER) {
    const size_t stride = (size_t)h->k + 1;
    char *buf = R_alloc(U ? U : 1, (int)stride);          /* transient: R reclaims it, also on error() */
    if (kmg_kmers_ascii(h->index, buf) != KMG_OK) failed = "kmer";
    else {
      SEXP kmers = allocVector(STRSXP, (R_xlen_t)U);
      SET_VECTOR_ELT(ret, 0, kmers);
      for (uint64_t u = 0; u < U; ++u) SET_STRING_ELT(kmers, (R_xlen_t)u, mkCharLen(buf + u * stride, h->k));
    }
  }
  if (!failed && (flag & F_POS)) {
    SEXP m = allocMatrix(INTSXP, 2, (int)N);          /* rows (i,pos); kmer.pos() transposes it */
    SET_VECTOR_ELT(ret, 1, m);
    if (kmg_positions(h->index, INTEGER(m)) != KMG_OK) failed = "pos";
  }
  if (!failed && (flag & F_PAIRS)) {
    SEXP m = allocMatrix(INTSXP, 3, (int)P);          /* rows (i,x,y) */
    SET_VECTOR_ELT(ret, 2, m);
    if (kmg_pairs(h->index, INTEGER(m)) != KMG_OK) failed = "pair.pos";
  }
  if (!failed && (flag & F_COUNT)) {
    SEXP v = allocVector(INTSXP, (R_xlen_t)U);
    SET_VECTOR_ELT(ret, 3, v);
    if (kmg_counts(h->index, INTEGER(v)) != KMG_OK) failed = "count";
  }
  UNPROTECT(1);
  if (failed) error("kmer.pos failed while extracting %s: %s", failed, kmg_last_error());
  return ret;
}

SEXP sequence_kmer_positions(SEXP ptr_r, SEXP seq_r, SEXP k_r) {
  kmer_handle *h = handle_or_error(ptr_r);
  if (TYPEOF(seq_r) != STRSXP || length(seq_r) != 1) error("seq_r should be a single sequence");
  if (TYPEOF(k_r) != INTSXP || length(k_r) != 1) error("k should be an integer of length 1");
  const int k = INTEGER(k_r)[0];
  SEXP s = STRING_ELT(seq_r, 0);
  const char *allow = getenv("KMERGPU_ALLOW_K32");
  const int kmax = (allow && allow[0] == '1') ? 32 : 31;
  if (length(s) <= k || k > kmax || k < 1) error("the sequence should be longer than k and k should not be longer than 31");

  kmg_query *q = NULL;
  uint64_t M = 0;
  if (kmg_query_begin(h->index, CHAR(s), (int64_t)length(s), k, &q, &M) != KMG_OK) error("seq.kmer.pos failed: %s", kmg_last_error());
  if (M > (uint64_t)INT_MAX) {
    kmg_query_free(q);
    error("seq.kmer.pos would return %llu rows, more than an R matrix can hold (2^31-1)", (unsigned long long)M);
  }
  SEXP m = PROTECT(allocMatrix(INTSXP, 2, (int)M)); /* rows (i,j); seq.kmer.pos() names and transposes */
  const int rc = kmg_query_emit(q, INTEGER(m));
  kmg_query_free(q);
  UNPROTECT(1);
  if (rc != KMG_OK) error("seq.kmer.pos failed: %s", kmg_last_error());
  return m;
}

/* kmer.pairs(ptr.a, ptr.b) -> .Call("kmer_pair_pos", ...), kmer_hash.R:30-34; replaces src/kmer_hash.c:1174-1203:
 * rows (a, b) for every k-mer the two indexes share; kmer.pairs() names and transposes. */
SEXP kmer_pair_pos(SEXP ptr_a, SEXP ptr_b) {
  kmer_handle *a = handle_or_error(ptr_a);
  kmer_handle *b = handle_or_error(ptr_b);
  kmg_join *j = NULL;
  uint64_t M = 0;
  if (kmg_join_begin(a->index, b->index, &j, &M) != KMG_OK) error("kmer.pairs failed: %s", kmg_last_error());
  if (M > (uint64_t)INT_MAX) {
    kmg_join_free(j);
    error("kmer.pairs would return %llu rows, more than an R matrix can hold (2^31-1)", (unsigned long long)M);
  }
  SEXP m = PROTECT(allocMatrix(INTSXP, 2, (int)M));
  const int rc = kmg_join_emit(j, INTEGER(m));
  kmg_join_free(j);
  UNPROTECT(1);
  if (rc != KMG_OK) error("kmer.pairs failed: %s", kmg_last_error());
  return m;
}

static const R_CallMethodDef call_methods[] = {
    {"make_kmer_h_index", (DL_FUNC)&make_kmer_h_index, 3},
    {"kmer_positions", (DL_FUNC)&kmer_positions, 2},
    {"sequence_kmer_positions", (DL_FUNC)&sequence_kmer_positions, 3},
    {"kmer_pair_pos", (DL_FUNC)&kmer_pair_pos, 2},
    {NULL, NULL, 0}};

void R_init_kmer_hash(DllInfo *info) { R_registerRoutines(info, NULL, call_methods, NULL, NULL); }
