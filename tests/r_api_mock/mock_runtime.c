/* mock_runtime.c -- an EXECUTABLE stand-in for the subset of R's C API that r_shim/src/gpv_shim.c uses, so that
 * the shim -- the code a GPvecchia maintainer would compile into the package -- can be run in an image without R:
 * tests/test_r_shim_mock.py builds gpv_shim.c against this file, calls the routines it registers BY NAME through
 * the table it hands to R_registerRoutines (what `.Call("_GPvecchia_U_NZentries", ...)` resolves), and compares
 * the results with the oracle and with the ctypes front end.
 *
 * Semantics follow the "Writing R Extensions" manual where the shim depends on them:
 *   - Rf_coerceVector returns its ARGUMENT when the type already matches (no copy);
 *   - NA_INTEGER = NA_LOGICAL = INT_MIN; integer/logical NA become NA_real_ and back;
 *   - Rf_error does not return: it longjmps to the frame mock_call set up (R's top level);
 *   - Rf_allocMatrix = vector + integer `dim` attribute; attributes are (symbol, value) pairs;
 *   - external pointers carry address, tag and a finalizer (run by mock_run_finalizer, as R's GC would).
 * Nothing is ever freed (test process).  Test infrastructure only: nothing in the product links this. */
#include <limits.h>
#include <math.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "Rinternals.h"
#include "R_ext/Rdynload.h"
#include "R_ext/Rallocators.h"

#define CHARSXP 9
#define SYMSXP 1
#define EXTPTRSXP 22

struct attr { SEXP sym, val; struct attr* next; };
struct SEXPREC {
  int type;
  R_xlen_t len;
  void* data;
  struct attr* attrs;
  void* addr;             /* external pointer */
  SEXP tag, prot;
  R_CFinalizer_t fin;
  R_allocator_t custom;   /* Rf_allocVector3: R's own copy of the allocator ... */
  void* custom_block;     /* ... and the block it returned (header + data) */
};

static struct SEXPREC nil_rec = {NILSXP, 0, NULL, NULL, NULL, NULL, NULL, NULL, {NULL, NULL, NULL, NULL}, NULL};
SEXP R_NilValue = &nil_rec;
SEXP R_NamesSymbol = NULL;
int R_NaInt = INT_MIN;
static SEXP dim_symbol = NULL;

static SEXP new_rec(int type, R_xlen_t len, size_t elt) {
  SEXP s = (SEXP)calloc(1, sizeof(struct SEXPREC));
  s->type = type; s->len = len; s->tag = s->prot = R_NilValue;
  s->data = calloc(len > 0 ? (size_t)len : 1, elt ? elt : 1);
  return s;
}
static double na_real(void) { union { double d; uint64_t u; } x; x.u = 0x7FF00000000007A2ull; return x.d; }   /* low word 1954 */
static int is_na_real(double v) { union { double d; uint64_t u; } x; x.d = v; return isnan(v) && (uint32_t)x.u == 1954u; }

/* ---- symbols ----------------------------------------------------------------------------------------- */
static SEXP symtab[256];
static int nsym = 0;
SEXP Rf_install(const char* name) {
  for (int i = 0; i < nsym; ++i) if (strcmp((const char*)symtab[i]->data, name) == 0) return symtab[i];
  SEXP s = new_rec(SYMSXP, (R_xlen_t)strlen(name), 1);
  free(s->data); s->data = strdup(name);
  symtab[nsym++] = s;
  return s;
}
static void init_symbols(void) {
  if (!R_NamesSymbol) { R_NamesSymbol = Rf_install("names"); dim_symbol = Rf_install("dim"); }
}

/* ---- protection: a counter, checked for balance by the tests ------------------------------------------- */
static int protect_depth = 0;
SEXP Rf_protect(SEXP s) { ++protect_depth; return s; }
void Rf_unprotect(int n) { protect_depth -= n; }
int mock_protect_depth(void) { return protect_depth; }

/* ---- vectors ------------------------------------------------------------------------------------------- */
SEXP Rf_allocVector(SEXPTYPE t, R_xlen_t n) {
  init_symbols();
  switch (t) {
    case LGLSXP: case INTSXP: return new_rec((int)t, n, sizeof(int));
    case REALSXP: return new_rec((int)t, n, sizeof(double));
    case STRSXP: case VECSXP: {
      SEXP s = new_rec((int)t, n, sizeof(SEXP));
      for (R_xlen_t i = 0; i < n; ++i) ((SEXP*)s->data)[i] = R_NilValue;
      return s;
    }
    default: fprintf(stderr, "mock R: allocVector type %u not supported\n", t); abort();
  }
}
/* vector whose storage comes from a custom allocator: one block for header + data, as R does (the data start some
 * bytes into the block); mock_release plays the garbage collector */
#define MOCK_HEADER_BYTES 64
SEXP Rf_allocVector3(SEXPTYPE t, R_xlen_t n, R_allocator_t* al) {
  if (!al) return Rf_allocVector(t, n);
  if (t != REALSXP && t != INTSXP && t != LGLSXP) { fprintf(stderr, "mock R: allocVector3 type %u not supported\n", t); abort(); }
  const size_t elt = (t == REALSXP) ? sizeof(double) : sizeof(int);
  SEXP s = new_rec((int)t, n, 1);
  free(s->data);
  s->custom = *al;
  s->custom_block = al->mem_alloc(&s->custom, MOCK_HEADER_BYTES + (size_t)(n > 0 ? n : 1) * elt);
  if (!s->custom_block) Rf_error("cannot allocate vector of size %.1f Mb", (double)n * elt / 1048576.0);
  s->data = (char*)s->custom_block + MOCK_HEADER_BYTES;
  return s;
}
void mock_release(SEXP s) {
  if (s->custom_block) { s->custom.mem_free(&s->custom, s->custom_block); s->custom_block = NULL; s->data = NULL; }
}
SEXP Rf_allocMatrix(SEXPTYPE t, int nr, int nc) {
  SEXP s = Rf_allocVector(t, (R_xlen_t)nr * nc);
  SEXP d = Rf_allocVector(INTSXP, 2);
  ((int*)d->data)[0] = nr; ((int*)d->data)[1] = nc;
  Rf_setAttrib(s, dim_symbol, d);
  return s;
}
static void type_check(SEXP s, int t, const char* what) {
  if (s->type != t) { fprintf(stderr, "mock R: %s() applied to an object of type %d\n", what, s->type); abort(); }   /* R: error */
}
double* REAL(SEXP s) { type_check(s, REALSXP, "REAL"); return (double*)s->data; }
int* INTEGER(SEXP s) { if (s->type != INTSXP && s->type != LGLSXP) type_check(s, INTSXP, "INTEGER"); return (int*)s->data; }
int* LOGICAL(SEXP s) { type_check(s, LGLSXP, "LOGICAL"); return (int*)s->data; }
R_xlen_t XLENGTH(SEXP s) { return s->len; }
int LENGTH(SEXP s) { return (int)s->len; }
int TYPEOF(SEXP s) { return s->type; }
SEXP SET_VECTOR_ELT(SEXP v, R_xlen_t i, SEXP x) { type_check(v, VECSXP, "SET_VECTOR_ELT"); ((SEXP*)v->data)[i] = x; return x; }
SEXP VECTOR_ELT(SEXP v, R_xlen_t i) { type_check(v, VECSXP, "VECTOR_ELT"); return ((SEXP*)v->data)[i]; }
SEXP STRING_ELT(SEXP v, R_xlen_t i) { type_check(v, STRSXP, "STRING_ELT"); return ((SEXP*)v->data)[i]; }
void SET_STRING_ELT(SEXP v, R_xlen_t i, SEXP x) { type_check(v, STRSXP, "SET_STRING_ELT"); ((SEXP*)v->data)[i] = x; }
const char* CHAR(SEXP s) { type_check(s, CHARSXP, "CHAR"); return (const char*)s->data; }
SEXP Rf_mkChar(const char* c) {
  SEXP s = new_rec(CHARSXP, (R_xlen_t)strlen(c), 1);
  free(s->data); s->data = strdup(c);
  return s;
}
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); ((double*)s->data)[0] = v; return s; }

/* ---- attributes ------------------------------------------------------------------------------------------ */
SEXP Rf_setAttrib(SEXP s, SEXP sym, SEXP val) {
  for (struct attr* a = s->attrs; a; a = a->next) if (a->sym == sym) { a->val = val; return val; }
  struct attr* a = (struct attr*)calloc(1, sizeof(struct attr));
  a->sym = sym; a->val = val; a->next = s->attrs; s->attrs = a;
  return val;
}
SEXP Rf_getAttrib(SEXP s, SEXP sym) {
  for (struct attr* a = s->attrs; a; a = a->next) if (a->sym == sym) return a->val;
  return R_NilValue;
}
Rboolean Rf_isNull(SEXP s) { return s == R_NilValue || s->type == NILSXP; }
Rboolean Rf_isReal(SEXP s) { return s->type == REALSXP; }
Rboolean Rf_isLogical(SEXP s) { return s->type == LGLSXP; }
Rboolean Rf_isMatrix(SEXP s) { init_symbols(); SEXP d = Rf_getAttrib(s, dim_symbol); return d != R_NilValue && d->len == 2; }
int Rf_nrows(SEXP s) { init_symbols(); SEXP d = Rf_getAttrib(s, dim_symbol); return d != R_NilValue ? ((int*)d->data)[0] : (int)s->len; }
int Rf_ncols(SEXP s) { init_symbols(); SEXP d = Rf_getAttrib(s, dim_symbol); return (d != R_NilValue && d->len >= 2) ? ((int*)d->data)[1] : 1; }

/* ---- coercion and copies ---------------------------------------------------------------------------------- */
SEXP Rf_duplicate(SEXP s) {
  if (s == R_NilValue) return s;
  size_t elt = (s->type == REALSXP) ? sizeof(double) : (s->type == STRSXP || s->type == VECSXP) ? sizeof(SEXP) : sizeof(int);
  SEXP c = new_rec(s->type, s->len, elt);
  memcpy(c->data, s->data, (size_t)s->len * elt);
  for (struct attr* a = s->attrs; a; a = a->next) Rf_setAttrib(c, a->sym, a->val);
  return c;
}
SEXP Rf_coerceVector(SEXP s, SEXPTYPE t) {
  if ((SEXPTYPE)s->type == t) return s;                        /* no copy: the caller's object itself */
  SEXP c = Rf_allocVector(t, s->len);
  for (R_xlen_t i = 0; i < s->len; ++i) {
    double v; int na = 0;
    if (s->type == REALSXP) { v = ((double*)s->data)[i]; na = isnan(v); }
    else if (s->type == INTSXP || s->type == LGLSXP) { const int k = ((int*)s->data)[i]; na = (k == INT_MIN); v = k; }
    else { fprintf(stderr, "mock R: coerceVector from type %d not supported\n", s->type); abort(); }
    if (t == REALSXP) ((double*)c->data)[i] = na ? na_real() : v;
    else if (t == INTSXP) ((int*)c->data)[i] = na ? INT_MIN : (int)v;           /* truncation towards zero, like R */
    else if (t == LGLSXP) ((int*)c->data)[i] = na ? INT_MIN : (v != 0.0);
    else { fprintf(stderr, "mock R: coerceVector to type %u not supported\n", t); abort(); }
  }
  for (struct attr* a = s->attrs; a; a = a->next) Rf_setAttrib(c, a->sym, a->val);
  return c;
}
double Rf_asReal(SEXP s) {
  if (s->len < 1) return na_real();
  if (s->type == REALSXP) return ((double*)s->data)[0];
  if (s->type == INTSXP || s->type == LGLSXP) { const int k = ((int*)s->data)[0]; return k == INT_MIN ? na_real() : (double)k; }
  return na_real();
}
int Rf_asInteger(SEXP s) {
  if (s->len < 1) return INT_MIN;
  if (s->type == INTSXP || s->type == LGLSXP) return ((int*)s->data)[0];
  if (s->type == REALSXP) { const double v = ((double*)s->data)[0]; return isnan(v) ? INT_MIN : (int)v; }
  return INT_MIN;
}

/* ---- options ------------------------------------------------------------------------------------------------ */
static struct attr* options = NULL;
SEXP Rf_GetOption1(SEXP sym) {
  for (struct attr* a = options; a; a = a->next) if (a->sym == sym) return a->val;
  return R_NilValue;
}
void mock_set_option(const char* name, SEXP val) {
  SEXP sym = Rf_install(name);
  for (struct attr* a = options; a; a = a->next) if (a->sym == sym) { a->val = val; return; }
  struct attr* a = (struct attr*)calloc(1, sizeof(struct attr));
  a->sym = sym; a->val = val; a->next = options; options = a;
}

/* ---- conditions ------------------------------------------------------------------------------------------------ */
static jmp_buf* top_level = NULL;
static char last_error[1024], last_warning[1024];
static int warning_count = 0;
void Rf_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(last_error, sizeof last_error, fmt, ap); va_end(ap);
  if (!top_level) { fprintf(stderr, "mock R: Rf_error outside mock_call: %s\n", last_error); abort(); }
  longjmp(*top_level, 1);
}
void Rf_warning(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(last_warning, sizeof last_warning, fmt, ap); va_end(ap);
  ++warning_count;
}
const char* mock_last_error(void) { return last_error; }
const char* mock_last_warning(void) { return last_warning; }
int mock_warning_count(void) { return warning_count; }

/* ---- external pointers ------------------------------------------------------------------------------------------ */
SEXP R_MakeExternalPtr(void* p, SEXP tag, SEXP prot) {
  SEXP s = new_rec(EXTPTRSXP, 0, 1);
  s->addr = p; s->tag = tag; s->prot = prot;
  return s;
}
void* R_ExternalPtrAddr(SEXP s) { type_check(s, EXTPTRSXP, "R_ExternalPtrAddr"); return s->addr; }
SEXP R_ExternalPtrTag(SEXP s) { type_check(s, EXTPTRSXP, "R_ExternalPtrTag"); return s->tag; }
void R_ClearExternalPtr(SEXP s) { type_check(s, EXTPTRSXP, "R_ClearExternalPtr"); s->addr = NULL; }
void R_RegisterCFinalizerEx(SEXP s, R_CFinalizer_t f, Rboolean onexit) { (void)onexit; s->fin = f; }
void mock_run_finalizer(SEXP s) { if (s->fin) s->fin(s); }                     /* what the GC does when the object dies */
void mock_null_extptr(SEXP s) { s->addr = NULL; }                              /* what readRDS() gives back */

/* ---- routine registration and .Call ------------------------------------------------------------------------------ */
static const R_CallMethodDef* call_table = NULL;
int R_registerRoutines(DllInfo* dll, const R_CMethodDef* c, const R_CallMethodDef* call, const R_FortranMethodDef* f,
                       const R_ExternalMethodDef* e) {
  (void)dll; (void)c; (void)f; (void)e;
  call_table = call;
  return 1;
}
int R_useDynamicSymbols(DllInfo* dll, int v) { (void)dll; (void)v; return 1; }
int mock_routine_arity(const char* name) {
  if (!call_table) return -2;
  for (const R_CallMethodDef* m = call_table; m->name; ++m) if (strcmp(m->name, name) == 0) return m->numArgs;
  return -1;
}
int mock_routine_count(void) {
  int n = 0;
  if (call_table) for (const R_CallMethodDef* m = call_table; m->name; ++m) ++n;
  return n;
}
const char* mock_routine_name(int i) { return call_table[i].name; }

typedef SEXP (*F0)(void);
typedef SEXP (*F1)(SEXP);
typedef SEXP (*F2)(SEXP, SEXP);
typedef SEXP (*F3)(SEXP, SEXP, SEXP);
typedef SEXP (*F4)(SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F5)(SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F6)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
typedef SEXP (*F9)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
/* .Call(name, args...): NULL + mock_last_error() when the routine raised an R error, is not registered or is
 * called with the wrong number of arguments (R checks the registered arity) */
SEXP mock_call(const char* name, int nargs, SEXP* a) {
  init_symbols();
  last_error[0] = 0;
  const R_CallMethodDef* m = call_table;
  for (; m && m->name; ++m) if (strcmp(m->name, name) == 0) break;
  if (!m || !m->name) { snprintf(last_error, sizeof last_error, "\"%s\" not available for .Call()", name); return NULL; }
  if (m->numArgs != nargs) {
    snprintf(last_error, sizeof last_error, "Incorrect number of arguments (%d), expecting %d for '%s'", nargs, m->numArgs, name);
    return NULL;
  }
  jmp_buf here;
  jmp_buf* outer = top_level;
  const int depth = protect_depth;
  SEXP r = NULL;
  top_level = &here;
  if (setjmp(here) == 0) {
    switch (nargs) {
      case 0: r = ((F0)m->fun)(); break;
      case 1: r = ((F1)m->fun)(a[0]); break;
      case 2: r = ((F2)m->fun)(a[0], a[1]); break;
      case 3: r = ((F3)m->fun)(a[0], a[1], a[2]); break;
      case 4: r = ((F4)m->fun)(a[0], a[1], a[2], a[3]); break;
      case 5: r = ((F5)m->fun)(a[0], a[1], a[2], a[3], a[4]); break;
      case 6: r = ((F6)m->fun)(a[0], a[1], a[2], a[3], a[4], a[5]); break;
      case 7: r = ((F7)m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6]); break;
      case 9: r = ((F9)m->fun)(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]); break;
      default: snprintf(last_error, sizeof last_error, "mock R: arity %d not wired", nargs); r = NULL;
    }
  } else {
    r = NULL;
    protect_depth = depth;                                     /* R unwinds the protect stack on error */
  }
  top_level = outer;
  return r;
}

/* ---- object construction / inspection for the Python side ---------------------------------------------------------- */
SEXP mock_nil(void) { return R_NilValue; }
SEXP mock_alloc(int type, int64_t n) { return Rf_allocVector((SEXPTYPE)type, (R_xlen_t)n); }
void mock_set_dim(SEXP s, int nr, int nc) {
  init_symbols();
  SEXP d = Rf_allocVector(INTSXP, 2);
  ((int*)d->data)[0] = nr; ((int*)d->data)[1] = nc;
  Rf_setAttrib(s, dim_symbol, d);
}
SEXP mock_string(const char* c) { SEXP s = Rf_allocVector(STRSXP, 1); ((SEXP*)s->data)[0] = Rf_mkChar(c); return s; }
void* mock_data(SEXP s) { return s->data; }
int64_t mock_length(SEXP s) { return (int64_t)s->len; }
int mock_type(SEXP s) { return s->type; }
SEXP mock_elt(SEXP s, int64_t i) { return ((SEXP*)s->data)[i]; }
const char* mock_chars(SEXP s) { return (const char*)s->data; }
SEXP mock_names(SEXP s) { init_symbols(); return Rf_getAttrib(s, R_NamesSymbol); }
void* mock_extptr_addr(SEXP s) { return s->addr; }
int mock_is_na_real(double v) { return is_na_real(v); }
