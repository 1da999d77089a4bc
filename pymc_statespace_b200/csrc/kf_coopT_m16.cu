#define KFB_M 16
#include "kf_coopT_inst.inc"
