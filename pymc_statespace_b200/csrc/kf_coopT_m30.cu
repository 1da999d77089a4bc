#define KFB_M 30
#include "kf_coopT_inst.inc"
