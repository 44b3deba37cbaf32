#define GPV_INST_P 26
#define GPV_INST_G 16
#include "u_inst.inc"
