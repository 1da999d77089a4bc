#define KFB_M 18
#include "kf_coopT_inst.inc"
