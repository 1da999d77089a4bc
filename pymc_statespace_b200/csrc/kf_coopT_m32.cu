#define KFB_M 32
#include "kf_coopT_inst.inc"
