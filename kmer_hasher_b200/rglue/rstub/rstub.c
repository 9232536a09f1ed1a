/* rstub.c -- implementation of the R C-API stand-in declared in Rinternals.h (tests only). */
#include "Rinternals.h"

#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct rstub_sexp {
  int type;
  R_xlen_t len;
  void *data;            /* int[], double[], SEXP[], char[] */
  SEXP names, dim;       /* the two attributes the glue sets */
  void *ext_addr;        /* EXTPTRSXP */
  SEXP ext_tag;
  R_CFinalizer_t fin;
};

static struct rstub_sexp nil_obj = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL, NULL};
static struct rstub_sexp names_sym = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL, NULL};
static struct rstub_sexp dim_sym = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL, NULL};
SEXP R_NilValue = &nil_obj, R_NamesSymbol = &names_sym, R_DimSymbol = &dim_sym;

static long live = 0;
static int protect_depth = 0;
static jmp_buf *top = NULL;
static char errmsg[1024];
static const R_CallMethodDef *routines = NULL;

struct transient { struct transient *next; size_t bytes; };
static struct transient *transients = NULL;
static size_t transient_bytes = 0;

static SEXP new_obj(int type, R_xlen_t n, size_t elt) {
  SEXP x = (SEXP)calloc(1, sizeof *x);
  if (!x) error("rstub: out of memory");
  x->type = type; x->len = n;
  x->names = x->dim = R_NilValue; x->ext_tag = R_NilValue;
  if (elt) {
    x->data = calloc((size_t)(n > 0 ? n : 1), elt);
    if (!x->data) { free(x); error("cannot allocate vector of size %.1f Gb", (double)n * elt / 1073741824.0); }
  }
  ++live;
  return x;
}

int TYPEOF(SEXP x) { return x->type; }
int length(SEXP x) { return (int)x->len; }
int asInteger(SEXP x) { return x->type == INTSXP && x->len > 0 ? ((int *)x->data)[0] : (int)0x80000000; }
int *INTEGER(SEXP x) { return (int *)x->data; }
double *REAL(SEXP x) { return (double *)x->data; }
const char *CHAR(SEXP x) { return (const char *)x->data; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP *)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP *)x->data)[i]; }
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; return v; }
SEXP mkCharLen(const char *s, int n) {
  SEXP x = new_obj(CHARSXP, n, 0);
  x->data = malloc((size_t)n + 1);
  memcpy(x->data, s, (size_t)n);
  ((char *)x->data)[n] = 0;
  return x;
}
SEXP mkChar(const char *s) { return mkCharLen(s, (int)strlen(s)); }
SEXP allocVector(int type, R_xlen_t n) {
  switch (type) {
    case INTSXP: return new_obj(type, n, sizeof(int));
    case REALSXP: return new_obj(type, n, sizeof(double));
    case STRSXP: case VECSXP: {
      SEXP x = new_obj(type, n, sizeof(SEXP));
      for (R_xlen_t i = 0; i < n; ++i) ((SEXP *)x->data)[i] = R_NilValue;
      return x;
    }
    default: error("rstub: allocVector type %d not supported", type);
  }
}
SEXP allocMatrix(int type, int nrow, int ncol) {
  if (nrow < 0 || ncol < 0) error("negative extents to matrix");
  SEXP x = allocVector(type, (R_xlen_t)nrow * ncol);
  SEXP d = allocVector(INTSXP, 2);
  INTEGER(d)[0] = nrow; INTEGER(d)[1] = ncol;
  x->dim = d;
  return x;
}
SEXP setAttrib(SEXP x, SEXP name, SEXP v) {
  if (name == R_NamesSymbol) x->names = v;
  else if (name == R_DimSymbol) x->dim = v;
  return v;
}
SEXP getAttrib(SEXP x, SEXP name) { return name == R_NamesSymbol ? x->names : name == R_DimSymbol ? x->dim : R_NilValue; }
SEXP Rf_protect(SEXP x) { ++protect_depth; return x; }
void Rf_unprotect(int n) { protect_depth -= n; }
char *R_alloc(size_t n, int size) {
  size_t bytes = n * (size_t)size;
  struct transient *t = (struct transient *)malloc(sizeof *t + bytes);
  if (!t) error("cannot allocate memory block of size %.1f Gb", (double)bytes / 1073741824.0);
  t->next = transients; t->bytes = bytes; transients = t; transient_bytes += bytes;
  return (char *)(t + 1);
}
static void free_transients(void) {
  while (transients) { struct transient *t = transients; transients = t->next; free(t); }
  transient_bytes = 0;
}

SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot) {
  (void)prot;
  SEXP x = new_obj(EXTPTRSXP, 1, 0);
  x->ext_addr = p; x->ext_tag = tag;
  return x;
}
void *R_ExternalPtrAddr(SEXP s) { return s->ext_addr; }
SEXP R_ExternalPtrTag(SEXP s) { return s->ext_tag; }
void R_ClearExternalPtr(SEXP s) { s->ext_addr = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit) { (void)onexit; s->fin = fun; }
int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call, const void *f, const void *e) {
  (void)info; (void)c; (void)f; (void)e;
  routines = call;
  return 1;
}

void error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(errmsg, sizeof errmsg, fmt, ap);
  va_end(ap);
  if (top) longjmp(*top, 1);
  fprintf(stderr, "rstub: error() outside rstub_call: %s\n", errmsg);
  abort();
}
void warning(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
void Rprintf(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); }

/* ---- harness ---------------------------------------------------------------------------------- */
SEXP rstub_string_vector(int n, const char *const *strings, const long *lengths) {
  SEXP x = allocVector(STRSXP, n);
  for (int i = 0; i < n; ++i) SET_STRING_ELT(x, i, mkCharLen(strings[i], (int)lengths[i]));
  return x;
}
SEXP rstub_int_vector(int n, const int *values) {
  SEXP x = allocVector(INTSXP, n);
  memcpy(INTEGER(x), values, (size_t)n * sizeof(int));
  return x;
}
typedef SEXP (*fn2)(SEXP, SEXP);
typedef SEXP (*fn3)(SEXP, SEXP, SEXP);
typedef SEXP (*fn4)(SEXP, SEXP, SEXP, SEXP);
SEXP rstub_call(const char *name, int nargs, SEXP *args, char *errbuf, int errlen) {
  if (errbuf && errlen) errbuf[0] = 0;
  const R_CallMethodDef *r = routines;
  for (; r && r->name; ++r) if (!strcmp(r->name, name)) break;
  if (!r || !r->name) { snprintf(errbuf, errlen, "C symbol name \"%s\" not in load table", name); return NULL; }
  if (r->numArgs != nargs) { snprintf(errbuf, errlen, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, r->numArgs, name); return NULL; }
  jmp_buf here, *saved = top;
  const int depth = protect_depth;
  SEXP out = NULL;
  top = &here;
  if (setjmp(here) == 0) {
    out = nargs == 2 ? ((fn2)r->fun)(args[0], args[1]) : nargs == 3 ? ((fn3)r->fun)(args[0], args[1], args[2]) : ((fn4)r->fun)(args[0], args[1], args[2], args[3]);
  } else {
    if (errbuf && errlen) snprintf(errbuf, errlen, "%s", errmsg);
    protect_depth = depth;        /* R unwinds the protect stack on error */
    out = NULL;
  }
  top = saved;
  free_transients();              /* R reclaims R_alloc memory when .Call returns or errors */
  return out;
}
int rstub_nrow(SEXP x) { return x->dim != R_NilValue ? INTEGER(x->dim)[0] : -1; }
int rstub_ncol(SEXP x) { return x->dim != R_NilValue ? INTEGER(x->dim)[1] : -1; }
void rstub_finalize(SEXP p) { if (p->type == EXTPTRSXP && p->fin) { R_CFinalizer_t f = p->fin; p->fin = NULL; f(p); } }
void rstub_release(SEXP x) {
  if (!x || x == R_NilValue || x == R_NamesSymbol || x == R_DimSymbol) return;
  if (x->type == STRSXP || x->type == VECSXP)
    for (R_xlen_t i = 0; i < x->len; ++i) rstub_release(((SEXP *)x->data)[i]);
  if (x->type == EXTPTRSXP) rstub_release(x->ext_tag);
  rstub_release(x->names);
  rstub_release(x->dim);
  free(x->data);
  free(x);
  --live;
}
int rstub_protect_depth(void) { return protect_depth; }
long rstub_live_objects(void) { return live; }
size_t rstub_transient_bytes(void) { return transient_bytes; }
