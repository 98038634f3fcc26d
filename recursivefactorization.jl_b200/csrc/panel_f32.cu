// K1 panel kernel, Float32 instantiation (see panel_impl.cuh).
#define RFB_PANEL_T float
#include "panel_impl.cuh"
