// static_kernels.h — host launchers of the ahead-of-time kernels (static_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace linon {

void launch_overlay(cudaStream_t stream, int W, int H, int overlay_mode, int delta, float scaling, const void* raycast,
                    const void* endpoint, void* final_rgba, float* exponents, int variant);  // variant 1: overlay_direct_kernel
void launch_outline_search(cudaStream_t stream, int W, int H, const float* exponents, unsigned long long* keys,
                           float* pos16, int n_sm);
void launch_click_positions(cudaStream_t stream, float mouse_x, float mouse_y, float* pos16);
double run_fp32_peak(int n_sm, int iters, float* d_scratch, float* ms_out);

}  // namespace linon
