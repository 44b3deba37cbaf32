#define GPV_INST_P 32
#include "u_inst_quad.inc"
