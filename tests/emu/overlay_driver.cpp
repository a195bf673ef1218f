// overlay_driver.cpp — TEST INFRASTRUCTURE: runs the per-thread body of the direct overlay kernel
// (linon_b200/csrc/device/overlay_direct.cuh, the text the product compiles with nvcc) on the host, thread by thread over
// the same grid and block shape as linon::launch_overlay uses for variant 1.  The kernel has no shared memory and no
// barrier, so a plain loop over (block, thread) is a faithful execution.  IEEE sqrt / division and --fmad=false on the
// device, -ffp-contract=off here: bit-identical up to expf (libm here, like the oracle).  Only tests build this.
#include <cmath>
#include <cstddef>
#include <cstdint>

struct float4 { float x, y, z, w; };
#define LINON_OV_HD static inline
#define __restrict__
static inline float4 __ldg(const float4* p) { return *p; }
static inline int __float2int_rn(float x) { return (int)std::nearbyint(x); }  // round to nearest even, the default mode

#include "overlay_direct.cuh"

extern "C" void emu_overlay_direct(int W, int H, int overlay_mode, int delta, float scaling, const uint32_t* raycast,
                                   const float* endpoint, uint32_t* final_rgba, float* exponents) {
    const int gx = (W + linon::OVD_TX - 1) / linon::OVD_TX, gy = (H + linon::OVD_TY - 1) / linon::OVD_TY;
#pragma omp parallel for collapse(2) schedule(static)
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx)
            for (int ty = 0; ty < linon::OVD_WARPS; ++ty)
                for (int tx = 0; tx < linon::OVD_TX; ++tx)
                    linon::overlay_direct_thread(tx, ty, bx, by, W, H, overlay_mode, delta, scaling, raycast,
                                                 (const float4*)endpoint, final_rgba, exponents);
}
