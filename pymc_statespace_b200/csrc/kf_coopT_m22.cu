#define KFB_M 22
#include "kf_coopT_inst.inc"
