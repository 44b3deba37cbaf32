/* see ../Rinternals.h.  Custom allocators for vectors (R >= 3.1.0, "Writing R Extensions" / R_ext/Rallocators.h):
 * Rf_allocVector3 asks mem_alloc for header + data in one block and calls mem_free when the vector is collected;
 * R keeps its own copy of the R_allocator_t. */
#ifndef GPV_R_API_MOCK_RALLOCATORS_H
#define GPV_R_API_MOCK_RALLOCATORS_H
#include <stddef.h>
typedef struct R_allocator R_allocator_t;
typedef void* (*custom_alloc_t)(R_allocator_t* allocator, size_t);
typedef void (*custom_free_t)(R_allocator_t* allocator, void*);
struct R_allocator {
  custom_alloc_t mem_alloc;
  custom_free_t mem_free;
  void* res;    /* reserved, must be NULL */
  void* data;   /* custom data */
};
SEXP Rf_allocVector3(SEXPTYPE, R_xlen_t, R_allocator_t*);
#endif
