#define GPV_INST_P 31
#include "u_inst_quad.inc"
