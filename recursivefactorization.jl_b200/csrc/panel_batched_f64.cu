// Batched small-matrix LU (one CTA per matrix), Float64 instantiation (see panel_impl.cuh).
#define RFB_PANEL_T double
#define RFB_PANEL_BATCHED 1
#include "panel_impl.cuh"
