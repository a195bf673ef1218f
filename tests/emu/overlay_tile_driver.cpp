// overlay_tile_driver.cpp — TEST INFRASTRUCTURE: runs the per-thread body of the TILE overlay kernel
// (linon_b200/csrc/device/overlay_tile.cuh, the text the product compiles with nvcc) on the host over the same grid and block
// shape as linon::launch_overlay uses for variant 2.  The tile the device receives from ONE cp.async.bulk.tensor.2d is filled
// here the way the TMA fills it: the box (32 + 2 delta) x (TH + 2 delta) at (x0 - delta, y0 - delta) of the endpoint image,
// texels outside the frame as zeros; the mbarrier wait is a no-op.  IEEE sqrt / division and --fmad=false on the device,
// -ffp-contract=off here: bit-identical up to expf (libm here, like the oracle).  Only tests build this.
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <vector>

struct float4 { float x, y, z, w; };
#define LINON_OV_HD static inline
#define __restrict__
static inline float4 __ldg(const float4* p) { return *p; }
static inline int __float2int_rn(float x) { return (int)std::nearbyint(x); }  // round to nearest even, the default mode

#include "overlay_tile.cuh"

template <int TH, int D>
static void run(int W, int H, int overlay_mode, float scaling, const uint32_t* raycast, const float4* endpoint, uint32_t* final_rgba,
                float* exponents) {
    constexpr int bw = linon::OVT_TX + 2 * D, bh = TH + 2 * D;
    const int gx = (W + linon::OVT_TX - 1) / linon::OVT_TX, gy = (H + TH - 1) / TH;
#pragma omp parallel for collapse(2) schedule(static)
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            const int x0 = bx * linon::OVT_TX, y0 = by * TH;
            std::vector<float4> tile((size_t)bw * bh, float4{0.0f, 0.0f, 0.0f, 0.0f});
            if (overlay_mode == 1)  // the kernel issues the load in mode 1 only
                for (int r = 0; r < bh; ++r)
                    for (int c = 0; c < bw; ++c) {
                        const int x = x0 - D + c, y = y0 - D + r;
                        if (x >= 0 && x < W && y >= 0 && y < H) tile[(size_t)r * bw + c] = endpoint[(size_t)y * W + x];
                    }
            for (int ty = 0; ty < linon::OVT_WARPS; ++ty)
                for (int tx = 0; tx < linon::OVT_TX; ++tx)
                    linon::overlay_tile_thread<TH, D>(tx, ty, x0, y0, W, H, overlay_mode, scaling, raycast, tile.data(), final_rgba,
                                                      exponents, [] {});
        }
}

// tall = 1: the 32 x 64 tiles the launcher picks for large frames (delta <= OVT_TALL_MAX_DELTA)
extern "C" int emu_overlay_tile(int W, int H, int overlay_mode, int delta, float scaling, const uint32_t* raycast,
                                const float* endpoint, uint32_t* final_rgba, float* exponents, int tall) {
    const float4* ep = (const float4*)endpoint;
    if (tall) {
        switch (delta) {
#define CASE(D) case D: run<linon::OVT_TH_TALL, D>(W, H, overlay_mode, scaling, raycast, ep, final_rgba, exponents); return 0;
            CASE(1) CASE(2) CASE(3) CASE(4)
#undef CASE
        }
        return 1;
    }
    switch (delta) {
#define CASE(D) case D: run<linon::OVT_TH, D>(W, H, overlay_mode, scaling, raycast, ep, final_rgba, exponents); return 0;
        CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10)
#undef CASE
    }
    return 1;  // the launcher runs the direct kernel for such a delta
}
