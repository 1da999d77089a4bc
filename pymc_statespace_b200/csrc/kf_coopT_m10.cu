#define KFB_M 10
#include "kf_coopT_inst.inc"
