#define KFB_M 7
#include "kf_coopT_inst.inc"
