/* see ../Rinternals.h: declarations only, for a syntax check of the R shim */
#ifndef GPV_R_API_MOCK_RDYNLOAD_H
#define GPV_R_API_MOCK_RDYNLOAD_H
typedef void* (*DL_FUNC)(void);
typedef struct { const char* name; DL_FUNC fun; int numArgs; } R_CallMethodDef;
typedef struct _DllInfo DllInfo;
typedef void R_CMethodDef;
typedef void R_FortranMethodDef;
typedef void R_ExternalMethodDef;
int R_registerRoutines(DllInfo*, const R_CMethodDef*, const R_CallMethodDef*, const R_FortranMethodDef*,
                       const R_ExternalMethodDef*);
int R_useDynamicSymbols(DllInfo*, int);
#endif
