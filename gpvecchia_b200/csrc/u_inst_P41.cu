#define GPV_INST_P 41
#define GPV_INST_G 32
#include "u_inst.inc"
