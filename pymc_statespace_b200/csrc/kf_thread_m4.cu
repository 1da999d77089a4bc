#define KFB_M 4
#include "kf_thread_inst.inc"
