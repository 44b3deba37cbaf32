#define GPV_INST_P 51
#define GPV_INST_G 32
#include "u_inst.inc"
