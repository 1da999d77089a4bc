#define KFB_M 3
#include "kf_thread_inst.inc"
