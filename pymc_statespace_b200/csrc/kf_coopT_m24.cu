#define KFB_M 24
#include "kf_coopT_inst.inc"
