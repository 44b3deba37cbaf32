#define GPV_INST_P 32
#define GPV_INST_G 16
#include "u_inst.inc"
