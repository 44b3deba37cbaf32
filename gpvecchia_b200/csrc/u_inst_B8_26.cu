#define GPV_INST_G 8
#define GPV_INST_P 26
#define GPV_INST_D3 0
#include "u_inst_band.inc"
