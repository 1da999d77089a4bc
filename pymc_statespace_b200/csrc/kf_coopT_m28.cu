#define KFB_M 28
#include "kf_coopT_inst.inc"
