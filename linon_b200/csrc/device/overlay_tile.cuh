// overlay_tile.cuh — per-thread body of the TILE overlay kernel (src/overlay.wgsl:99-146, overlay_desktop):
// LINON_OPT_OVERLAY_KERNEL = 2, the default.
//
// The endpoint tile of a CTA - 32 x TH pixels plus a halo of `delta` on every side - arrives in shared memory by ONE
// cp.async.bulk.tensor.2d (TMA) issued by one thread: the tensor map describes the endpoint image as a (4 W, H) fp32 tensor,
// the box is (4 (32 + 2 delta), TH + 2 delta), texels outside the frame are zero-filled by the hardware (pixels inside the
// 2 delta border never use their taps, src/overlay.wgsl:118-123).  No thread computes a global address for a tap: the four
// taps of a pixel are four LDS.128 at compile-time offsets from one shared-memory address (delta is a template parameter),
// conflict-free (a warp reads 512 contiguous bytes).  Colours travel as packed 32-bit words through plain coalesced loads
// and stores, issued before the thread waits for the tile.  Against the direct kernel (overlay_direct.cuh: four LDG.128 per
// pixel with 64-bit address arithmetic, 134 instructions per pixel, 75 % of the issue slots, 53 us at 3840 x 2160) this is
// ~93 instructions per pixel on tiles that touch neither the frame's edge nor its border, which take a loop without a
// single bounds test (the flag is uniform per CTA).
//
// The arithmetic is ovd_pixel of overlay_direct.cuh, expression by expression (--fmad=false): bit-identical results.
// The body is written against names the including file provides (LINON_OV_HD, float4; wait_for_tile is the mbarrier wait on the device and a no-op on the host), so that
// tests/emu/overlay_tile_driver.cpp can compile the same text for the host with a tile it fills like the TMA does.
#pragma once

#include "overlay_direct.cuh"

namespace linon {

constexpr int OVT_TX = 32;     // tile width  = one warp
constexpr int OVT_WARPS = 8;   // 256 threads, TH / 8 rows per thread
constexpr int OVT_TH = 32;       // tile height ...
constexpr int OVT_TH_TALL = 64;  // ... and the one of large frames (static_kernels.cu: overlay_tile_rows)
constexpr int OVT_TALL_MAX_DELTA = 4;  // (32 + 2 d) (64 + 2 d) float4 <= 48 KB of static shared memory
constexpr int OVT_MAX_DELTA = 10;  // GUI slider range, src/application.rs:297

// tile: (32 + 2 D) x (TH + 2 D) float4, row-major, origin at pixel (x0 - D, y0 - D) of the frame; wait_for_tile() returns
// once the tile is complete (called after the thread's colour loads are in flight, only when overlay_mode == 1)
template <int TH, int D, class Wait>
LINON_OV_HD void overlay_tile_thread(int tx, int ty, int x0, int y0, int W, int H, int overlay_mode, float scaling,
                                     const unsigned int* __restrict__ raycast, const float4* __restrict__ tile,
                                     unsigned int* __restrict__ final_rgba, float* __restrict__ exponents, Wait wait_for_tile) {
    constexpr int ROWS = TH / OVT_WARPS;
    constexpr int bw = OVT_TX + 2 * D, padding = 2 * D;
    const bool stencil = overlay_mode == 1;
    const int x = x0 + tx;
    const float4* t0 = tile + (ty + D) * bw + (tx + D);
    if (stencil && x0 >= padding && x0 + OVT_TX <= W - padding && y0 >= padding && y0 + TH <= H - padding) {
        // interior tile: every pixel is inside the frame and outside the border
        const int idx0 = (y0 + ty) * W + x;
        unsigned int c[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) c[r] = raycast[idx0 + OVT_WARPS * r * W];
        wait_for_tile();
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float4* t = t0 + OVT_WARPS * r * bw;
            float scaled;
            ovd_pixel(t[D], t[-D], t[D * bw], t[-D * bw], scaling, c[r], scaled);
            exponents[idx0 + OVT_WARPS * r * W] = scaled;  // stride W (the reference uses size.y)
            final_rgba[idx0 + OVT_WARPS * r * W] = c[r];
        }
        return;
    }
    const bool inner_x = x >= padding && x < W - padding;
    unsigned int c[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int y = y0 + ty + OVT_WARPS * r;
        c[r] = (x < W && y < H) ? raycast[y * W + x] : 0u;
    }
    if (stencil) wait_for_tile();
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int y = y0 + ty + OVT_WARPS * r;
        if (x >= W || y >= H) continue;
        const int idx = y * W + x;
        if (stencil && inner_x && y >= padding && y < H - padding) {
            const float4* t = t0 + OVT_WARPS * r * bw;
            float scaled;
            ovd_pixel(t[D], t[-D], t[D * bw], t[-D * bw], scaling, c[r], scaled);
            exponents[idx] = scaled;
        }
        final_rgba[idx] = c[r];
    }
}

}  // namespace linon
