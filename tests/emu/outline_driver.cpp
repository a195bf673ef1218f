// outline_driver.cpp — TEST INFRASTRUCTURE: the outline search's key logic (linon_b200/csrc/device/outline_keys.cuh, the text
// static_kernels.cu compiles with nvcc) on the host.  The kernel is a maximum over per-pixel 64-bit keys - by grid-stride
// loop, warp shuffles, shared memory and atomicMax, all of them order-independent - followed by outline_finalize_kernel;
// here the maximum is taken over the pixels in a scrambled order (`stride` coprime to W*H) to show the order does not
// matter, and the two lines of outline_finalize_kernel's arithmetic are restated.  Only tests build this.
#include <cstddef>
#include <cstdint>

#define LINON_OK_HD static inline
#include "outline_keys.cuh"

extern "C" void emu_outline_search(int W, int H, const float* exponents, float* pos16, unsigned long long stride) {
    unsigned long long best[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const unsigned long long n = (unsigned long long)W * H;
    for (unsigned long long t = 0, i = 0; t < n; ++t, i = (i + stride) % n) {
        if (!(exponents[i] >= 0.8f)) continue;
        const unsigned int y = (unsigned int)(i / W), x = (unsigned int)(i - (unsigned long long)y * W);
        unsigned long long k[8];
        linon::outline_keys(x, y, (unsigned int)W, (unsigned int)H, k);
        for (int j = 0; j < 8; ++j) best[j] = best[j] > k[j] ? best[j] : k[j];
    }
    const float width = (float)W, height = (float)H;
    for (int j = 0; j < 8; ++j) {
        float px = 0.0f, py = 0.0f;
        linon::outline_decode(j, best[j], W, H, px, py);
        pos16[2 * j] = px / (width - 1.0f);               // outline_finalize_kernel, src/main_view.wgsl:467-468
        pos16[2 * j + 1] = (height - 1.0f - py) / (height - 1.0f);
    }
}
