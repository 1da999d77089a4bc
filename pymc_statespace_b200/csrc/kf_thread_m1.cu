#define KFB_M 1
#include "kf_thread_inst.inc"
