#define KFB_M 8
#include "kf_coopT_inst.inc"
