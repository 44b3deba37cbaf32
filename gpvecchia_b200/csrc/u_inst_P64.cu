#define GPV_INST_P 64
#define GPV_INST_G 32
#include "u_inst.inc"
