#define GPV_INST_G 8
#define GPV_INST_P 21
#define GPV_INST_D3 1
#include "u_inst_band.inc"
