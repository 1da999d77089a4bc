#define KFB_M 20
#include "kf_coopT_inst.inc"
