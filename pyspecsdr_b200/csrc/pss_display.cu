#include "pss_common.cuh"
void pss_display_release(pss_ctx*) {}
