#define KFB_M 26
#include "kf_coopT_inst.inc"
