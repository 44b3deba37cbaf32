#define GPV_INST_P 11
#define GPV_INST_G 8
#include "u_inst.inc"
