#define KFB_M 2
#include "kf_thread_inst.inc"
