#define KFB_M 12
#include "kf_coopT_inst.inc"
