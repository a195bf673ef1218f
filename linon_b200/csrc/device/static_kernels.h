// static_kernels.h — host launchers of the ahead-of-time kernels (static_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace linon {

// variant 0: overlay_kernel (hand-staged tile), 1: overlay_direct_kernel, 2: overlay_tile_kernel (needs endpoint_map, a
// CUtensorMap over the endpoint image made with the box of overlay_tile_box(delta, tile_rows), tile_rows from
// overlay_tile_rows(); without a map, or with a delta outside 1..10, variant 2 runs the direct kernel)
void launch_overlay(cudaStream_t stream, int W, int H, int overlay_mode, int delta, float scaling, const void* raycast,
                    const void* endpoint, void* final_rgba, float* exponents, int variant, const void* endpoint_map = nullptr,
                    int tile_rows = 0);
int overlay_tile_rows(int W, int H, int delta, int n_sm);
void overlay_tile_box(int delta, int tile_rows, unsigned int* box_floats, unsigned int* box_rows);
void launch_outline_search(cudaStream_t stream, int W, int H, const float* exponents, unsigned long long* keys,
                           float* pos16, int n_sm);
void launch_click_positions(cudaStream_t stream, float mouse_x, float mouse_y, float* pos16);
double run_fp32_peak(int n_sm, int iters, float* d_scratch, float* ms_out);

}  // namespace linon
