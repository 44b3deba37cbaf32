/* see Rinternals.h in this directory: declarations only, for a syntax check of the R shim */
#ifndef GPV_R_API_MOCK_R_H
#define GPV_R_API_MOCK_R_H
#include <stdlib.h>
#include <stdint.h>
#endif
