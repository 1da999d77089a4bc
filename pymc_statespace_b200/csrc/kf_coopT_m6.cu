#define KFB_M 6
#include "kf_coopT_inst.inc"
