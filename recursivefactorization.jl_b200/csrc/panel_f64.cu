// K1 panel kernel, Float64 instantiation (see panel_impl.cuh).
#define RFB_PANEL_T double
#include "panel_impl.cuh"
