// Fused multi-step lattice kernels (filled in below).
#pragma once
#include "dw_common.cuh"

struct dw_handle;
// placeholders until the fused lattice kernel lands: everything runs through the materialising kernels
static inline bool dw_fused_supported(const dw_handle *) { return false; }
static inline int run_steps_fused(dw_handle *, int, int, const int8_t *, unsigned long long) { return DW_E_UNSUPPORTED; }
