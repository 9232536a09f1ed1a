/*
 * rstub/Rinternals.h -- a minimal stand-in for the part of R's C API that the index glue uses
 * (SURVEY.md Appendix C), so kmer_hash.c can be compiled and exercised in an image without R.
 * It is NOT R: vectors are malloc'd structs, PROTECT is a counter, the garbage collector is the
 * test harness calling rstub_finalize().  error() longjmps to the innermost rstub_call(), like
 * R's error() longjmps to top level.
 */
#ifndef RSTUB_RINTERNALS_H
#define RSTUB_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct rstub_sexp *SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif

enum { NILSXP = 0, CHARSXP = 9, INTSXP = 13, REALSXP = 14, STRSXP = 16, VECSXP = 19, EXTPTRSXP = 22 };

extern SEXP R_NilValue, R_NamesSymbol, R_DimSymbol;

/* Like R (without R_NO_REMAP) the short names are macros for Rf_ functions; this also keeps them
 * from colliding with libc symbols such as error(3). */
#define length Rf_length
#define asInteger Rf_asInteger
#define mkChar Rf_mkChar
#define mkCharLen Rf_mkCharLen
#define allocVector Rf_allocVector
#define allocMatrix Rf_allocMatrix
#define setAttrib Rf_setAttrib
#define getAttrib Rf_getAttrib
#define error Rf_error
#define warning Rf_warning
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)

int TYPEOF(SEXP x);
int length(SEXP x);
int asInteger(SEXP x);
int *INTEGER(SEXP x);
double *REAL(SEXP x);
const char *CHAR(SEXP x);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP mkChar(const char *s);
SEXP mkCharLen(const char *s, int n);
SEXP allocVector(int type, R_xlen_t n);
SEXP allocMatrix(int type, int nrow, int ncol);
SEXP setAttrib(SEXP x, SEXP name, SEXP v);
SEXP getAttrib(SEXP x, SEXP name);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
char *R_alloc(size_t n, int size);

typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void *p, SEXP tag, SEXP prot);
void *R_ExternalPtrAddr(SEXP s);
SEXP R_ExternalPtrTag(SEXP s);
void R_ClearExternalPtr(SEXP s);
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t fun, Rboolean onexit);

typedef void *(*DL_FUNC)(void);
typedef struct { const char *name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct rstub_dllinfo DllInfo;
int R_registerRoutines(DllInfo *info, const void *c, const R_CallMethodDef *call, const void *f, const void *e);

void error(const char *fmt, ...) __attribute__((noreturn, format(printf, 1, 2)));
void warning(const char *fmt, ...);
void Rprintf(const char *fmt, ...);

/* ---- harness side (what an R session would do) ------------------------------------------------- */
SEXP rstub_string_vector(int n, const char *const *strings, const long *lengths); /* STRSXP        */
SEXP rstub_int_vector(int n, const int *values);                                /* INTSXP        */
/* .Call(name, args...): returns NULL and fills errbuf when the routine raised error() */
SEXP rstub_call(const char *name, int nargs, SEXP *args, char *errbuf, int errlen);
int rstub_nrow(SEXP x);
int rstub_ncol(SEXP x);
void rstub_finalize(SEXP extptr);   /* the garbage collector collecting an external pointer         */
void rstub_release(SEXP x);         /* free a value made by the stub (not external-pointer targets)  */
int rstub_protect_depth(void);
long rstub_live_objects(void);
size_t rstub_transient_bytes(void); /* R_alloc memory still held (0 between calls)                    */

#ifdef __cplusplus
}
#endif
#endif
