#define KFB_M 5
#include "kf_coopT_inst.inc"
