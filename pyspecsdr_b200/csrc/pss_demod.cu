#include "pss_common.cuh"
void pss_demod_release(pss_ctx*) {}
