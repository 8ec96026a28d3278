/* Minimal DECLARATIONS of the R C API used by r/src/ca_shim.c -- for `gcc -fsyntax-only` in an image without R
 * (tests/test_host.py::test_r_shim_compiles_against_stub_headers).  Signatures follow R's Rinternals.h; nothing here
 * is ever linked or run.  It keeps the shim in step with include/clonealign_b200.h (argument counts and types). */
#ifndef CA_R_STUB_RINTERNALS_H
#define CA_R_STUB_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef int Rboolean;
#define TRUE 1
#define FALSE 0
#define INTSXP 13
#define REALSXP 14
#define VECSXP 19
extern SEXP R_NilValue, R_DimSymbol;
extern double R_NaReal;
#define NA_REAL R_NaReal
int TYPEOF(SEXP);
int* INTEGER(SEXP);
double* REAL(SEXP);
R_xlen_t XLENGTH(SEXP);
int Rf_isNull(SEXP);
int Rf_asInteger(SEXP);
double Rf_asReal(SEXP);
int Rf_ncols(SEXP);
int Rf_nrows(SEXP);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_allocVector(unsigned int, R_xlen_t);
SEXP Rf_allocMatrix(unsigned int, int, int);
SEXP Rf_mkNamed(unsigned int, const char**);
SEXP Rf_ScalarReal(double);
SEXP Rf_ScalarInteger(int);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
void Rf_error(const char*, ...) __attribute__((noreturn));
char* R_alloc(size_t, int);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
void R_ClearExternalPtr(SEXP);
typedef void (*R_CFinalizer_t)(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
