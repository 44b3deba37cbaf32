#define GPV_INST_P 16
#define GPV_INST_G 8
#include "u_inst.inc"
