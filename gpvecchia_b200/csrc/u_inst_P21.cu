#define GPV_INST_P 21
#define GPV_INST_G 16
#include "u_inst.inc"
