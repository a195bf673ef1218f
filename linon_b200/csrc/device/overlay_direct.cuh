// overlay_direct.cuh — per-thread body of the DIRECT overlay kernel (src/overlay.wgsl:99-146, overlay_desktop).
//
// LINON_OPT_OVERLAY_KERNEL = 1 (the fallback of the tile kernel, overlay_tile.cuh, when no tensor map can be made).  The
// round-1 kernel (static_kernels.cu: overlay_kernel) stages a 32x16 tile + halo of the endpoints into shared memory by hand
// and is bound by instruction issue, not by HBM (profiles/r1_overlay_c3.md: 255 thread-instructions per pixel, about 100 of
// them the staging loop, 79 % of the issue slots, 39 % of the measured HBM bandwidth; this kernel: 134 instructions per pixel,
// 53 us against 91 us at 3840 x 2160, profiles/r2_overlay_c3.md).  Here every pixel fetches its four taps E[x+-delta], E[y+-delta] straight from
// global memory through the read-only path: a warp's tap is 512 contiguous bytes, the y+-delta rows are the rows of the
// neighbouring warps of the CTA (L1 hits), the halo rows of neighbouring CTAs come from L2.  No shared memory, no barrier,
// no limit on delta; colours travel as packed 32-bit words and are unpacked only where the overlay blends.
// The arithmetic is the reference's, expression by expression, in the same order as overlay_kernel (compiled with
// --fmad=false): results are bit-identical with it.
//
// The body is written against four names the including file provides, so that tests/emu/overlay_driver.cpp can compile
// the same text for the host (LINON_OV_HD, float4, __ldg, __float2int_rn): TEST builds only, the product includes this
// file from static_kernels.cu alone.
#pragma once

#ifndef LINON_OV_HD
#define LINON_OV_HD __device__ __forceinline__
#endif

namespace linon {

constexpr int OVD_TX = 32;    // tile width  = one warp
constexpr int OVD_TY = 16;    // tile height = 8 warps x 2 rows
constexpr int OVD_WARPS = 8;

LINON_OV_HD unsigned int ovd_unorm8(float x) {
    x = fminf(fmaxf(x, 0.0f), 1.0f);
    return (unsigned int)__float2int_rn(x * 255.0f) & 0xffu;
}

// lyapunov_exponent (src/overlay.wgsl:99-112) from the four taps, the scaled exponent and the blend of overlay_desktop
// (:124-146) for one pixel; c is the packed rgba8 colour (updated only where the overlay blends).
// vec2squareroots (src/overlay.wgsl:15-35) on a = (det, -trace) is restated with ONE square root for both signs of the
// discriminant: sqrt(|discrim|) is the reference's sqrt(discrim) for discrim >= 0 and its sqrt(-discrim) otherwise, bit for
// bit (discrim = -0.0 gives -0.0 there and +0.0 here, and the two sums it enters, -a1 - root and -a1 + root with
// -a1 = trace >= +0, come out the same either way); the branch becomes two selects.
LINON_OV_HD void ovd_pixel(const float4 xp, const float4 xm, const float4 yp, const float4 ym, float scaling, unsigned int& c,
                           float& scaled) {
    const float gxx = 0.5f * (xp.x - xm.x);
    const float gxy = 0.5f * (xp.y - xm.y);
    const float gxz = 0.5f * (xp.z - xm.z);
    const float gyx = 0.5f * (yp.x - ym.x);
    const float gyy = 0.5f * (yp.y - ym.y);
    const float gyz = 0.5f * (yp.z - ym.z);
    const float m00 = gxx * gxx + gxy * gxy + gxz * gxz;
    const float m01 = gxx * gyx + gxy * gyy + gxz * gyz;
    const float m11 = gyx * gyx + gyy * gyy + gyz * gyz;
    const float a0 = m00 * m11 - m01 * m01;  // mat2det, src/overlay.wgsl:4-6
    const float a1 = -(m00 + m11);
    const float discrim = a1 * a1 - 4.0f * a0;
    const float root = sqrtf(fabsf(discrim));
    const bool real = discrim >= 0.0f;
    const float e0 = (real ? (-a1 - root) : -a1) / 2.0f;
    const float e1 = (real ? (-a1 + root) : root) / 2.0f;
    const float exponent = sqrtf(fmaxf(e0, e1));
    scaled = expf(scaling * exponent - 5.0f);
    if (scaled >= 0.8f) {
        const float alpha = fminf(1.0f, scaled) * 0.5f;
        const float inv = 1.0f - alpha;
        const unsigned int cx = c & 0xffu, cy = (c >> 8) & 0xffu, cz = (c >> 16) & 0xffu, cw = c >> 24;
        c = ovd_unorm8(alpha * 1.0f + inv * ((float)cx / 255.0f)) |
            (ovd_unorm8(alpha * 1.0f + inv * ((float)cy / 255.0f)) << 8) |
            (ovd_unorm8(alpha * 1.0f + inv * ((float)cz / 255.0f)) << 16) |
            (ovd_unorm8(alpha * 1.0f + inv * ((float)cw / 255.0f)) << 24);
    }
}

// one thread of a (32, 8) block: pixels (bx*32 + tx, by*16 + ty) and (bx*32 + tx, by*16 + ty + 8)
LINON_OV_HD void overlay_direct_thread(int tx, int ty, int bx, int by, int W, int H, int overlay_mode, int delta,
                                       float scaling, const unsigned int* __restrict__ raycast,
                                       const float4* __restrict__ endpoint, unsigned int* __restrict__ final_rgba,
                                       float* __restrict__ exponents) {
    const int x = bx * OVD_TX + tx;
    if (x >= W) return;
    const int padding = 2 * delta;
    const bool stencil = overlay_mode == 1;
    const bool inner_x = !(x < padding || x >= W - padding);
    const long long tap_x = (long long)delta * (long long)sizeof(float4);
    const long long tap_y = tap_x * W;
#pragma unroll
    for (int r = 0; r < OVD_TY / OVD_WARPS; ++r) {
        const int y = by * OVD_TY + ty + OVD_WARPS * r;
        if (y >= H) continue;
        const long long idx = (long long)y * W + x;
        unsigned int c = raycast[idx];
        if (stencil && inner_x && !(y < padding || y >= H - padding)) {
            // lyapunov_exponent, src/overlay.wgsl:99-112
            // one 64-bit address for the pixel, four warp-uniform byte offsets from it
            const char* e = (const char*)(endpoint + idx);
            const float4 xp = __ldg((const float4*)(e + tap_x)), xm = __ldg((const float4*)(e - tap_x));
            const float4 yp = __ldg((const float4*)(e + tap_y)), ym = __ldg((const float4*)(e - tap_y));
            float scaled;
            ovd_pixel(xp, xm, yp, ym, scaling, c, scaled);
            exponents[idx] = scaled;  // stride W (the reference uses size.y)
        }
        final_rgba[idx] = c;
    }
}

}  // namespace linon
