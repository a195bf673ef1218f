// overlay_staged_driver.cpp — TEST INFRASTRUCTURE: the measured default overlay kernel (overlay_kernel of
// linon_b200/csrc/device/static_kernels.cu; its text is cut out of that file by the test and passed as LINON_OV_TEXT) on the
// host.  A CTA is 32 x 8 threads with ONE barrier between the shared-memory stage and the per-pixel work, so a CTA is run as
// two passes over its threads: in pass 1 __syncthreads() ends the thread (its part of the stage is written), in pass 2 it
// is a no-op - each thread re-writes the same staged values and carries on with a complete tile in shared memory.
// Only tests build this.
#include <cmath>
#include <cstddef>
#include <cstdint>

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct dim3e { unsigned int x, y, z; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return {x, y, z, w}; }
static inline float4 __ldg(const float4* p) { return *p; }
static inline int __float2int_rn(float x) { return (int)std::nearbyint(x); }
static dim3e threadIdx, blockIdx;
static int g_pass;
struct StageDone {};
static inline void __syncthreads() { if (g_pass == 1) throw StageDone(); }
#define __device__
#define __global__
#define __forceinline__ inline
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)

namespace linon {
#include LINON_OV_TEXT
}

extern "C" void emu_overlay_staged(int W, int H, int overlay_mode, int delta, float scaling, const uint8_t* raycast,
                                   const float* endpoint, uint8_t* final_rgba, float* exponents) {
    const int gx = (W + linon::OV_TX - 1) / linon::OV_TX, gy = (H + linon::OV_TY - 1) / linon::OV_TY;
    for (unsigned by = 0; by < (unsigned)gy; ++by)
        for (unsigned bx = 0; bx < (unsigned)gx; ++bx)
            for (g_pass = 1; g_pass <= 2; ++g_pass)
                for (unsigned ty = 0; ty < 8; ++ty)
                    for (unsigned tx = 0; tx < (unsigned)linon::OV_TX; ++tx) {
                        threadIdx = {tx, ty, 0};
                        blockIdx = {bx, by, 0};
                        try {
                            linon::overlay_kernel(W, H, overlay_mode, delta, scaling, (const uchar4*)raycast,
                                                  (const float4*)endpoint, (uchar4*)final_rgba, exponents);
                        } catch (const StageDone&) {
                        }
                    }
}
