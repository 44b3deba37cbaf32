/* Minimal DECLARATIONS of the R C API subset that r_shim/src/gpv_shim.c uses, written from the public
 * "Writing R Extensions" manual: enough for `gcc -fsyntax-only` to type-check the shim in an image without
 * R (tests/test_capi_cpu.py), and -- with the definitions in mock_runtime.c -- to build and RUN it
 * (tests/test_r_shim_mock.py).  Test infrastructure only. */
#ifndef GPV_R_API_MOCK_RINTERNALS_H
#define GPV_R_API_MOCK_RINTERNALS_H
#include <stddef.h>
typedef struct SEXPREC* SEXP;
typedef ptrdiff_t R_xlen_t;
typedef enum { FALSE = 0, TRUE } Rboolean;
typedef unsigned int SEXPTYPE;
#define NILSXP 0
#define LGLSXP 10
#define INTSXP 13
#define REALSXP 14
#define STRSXP 16
#define VECSXP 19
extern SEXP R_NilValue, R_NamesSymbol;
extern int R_NaInt;
#define NA_INTEGER R_NaInt
SEXP Rf_protect(SEXP);
void Rf_unprotect(int);
#define PROTECT(s) Rf_protect(s)
#define UNPROTECT(n) Rf_unprotect(n)
double* REAL(SEXP);
int* INTEGER(SEXP);
int* LOGICAL(SEXP);
R_xlen_t XLENGTH(SEXP);
int LENGTH(SEXP);
int TYPEOF(SEXP);
SEXP Rf_allocVector(SEXPTYPE, R_xlen_t);
SEXP Rf_allocMatrix(SEXPTYPE, int, int);
SEXP Rf_coerceVector(SEXP, SEXPTYPE);
SEXP Rf_duplicate(SEXP);
SEXP SET_VECTOR_ELT(SEXP, R_xlen_t, SEXP);
SEXP VECTOR_ELT(SEXP, R_xlen_t);
SEXP STRING_ELT(SEXP, R_xlen_t);
void SET_STRING_ELT(SEXP, R_xlen_t, SEXP);
const char* CHAR(SEXP);
SEXP Rf_mkChar(const char*);
SEXP Rf_install(const char*);
SEXP Rf_setAttrib(SEXP, SEXP, SEXP);
SEXP Rf_getAttrib(SEXP, SEXP);
SEXP Rf_GetOption1(SEXP);
SEXP Rf_ScalarReal(double);
int Rf_ncols(SEXP);
int Rf_nrows(SEXP);
double Rf_asReal(SEXP);
int Rf_asInteger(SEXP);
Rboolean Rf_isNull(SEXP);
Rboolean Rf_isMatrix(SEXP);
Rboolean Rf_isReal(SEXP);
Rboolean Rf_isLogical(SEXP);
void Rf_error(const char*, ...) __attribute__((noreturn));
void Rf_warning(const char*, ...);
typedef void (*R_CFinalizer_t)(SEXP);
SEXP R_MakeExternalPtr(void*, SEXP, SEXP);
void* R_ExternalPtrAddr(SEXP);
SEXP R_ExternalPtrTag(SEXP);
void R_ClearExternalPtr(SEXP);
void R_RegisterCFinalizerEx(SEXP, R_CFinalizer_t, Rboolean);
#endif
