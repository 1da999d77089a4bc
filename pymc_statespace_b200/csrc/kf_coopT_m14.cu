#define KFB_M 14
#include "kf_coopT_inst.inc"
