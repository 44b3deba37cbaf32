#define GPV_INST_G 16
#define GPV_INST_P 41
#define GPV_INST_D3 1
#include "u_inst_band.inc"
