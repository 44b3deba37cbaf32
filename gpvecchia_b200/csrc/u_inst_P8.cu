#define GPV_INST_P 8
#define GPV_INST_G 4
#include "u_inst.inc"
