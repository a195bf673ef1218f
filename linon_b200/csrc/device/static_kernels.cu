// static_kernels.cu — kernels that do not depend on the field function (compiled ahead of time by nvcc
// for sm_100a and linked into liblinon_b200.so).
//
//   overlay_kernel       Lyapunov-exponent overlay, src/overlay.wgsl:99-146 (overlay_desktop).  HBM-bound
//                        stencil: 28 algorithmic bytes per pixel in mode 1 (4 colour in, 16 endpoint in, 4 colour
//                        out, 4 exponent out), 8 in mode 0.  Endpoints are staged per CTA into shared memory
//                        (tile + halo of `delta`, three float planes) so every texel is fetched from L2/HBM once
//                        per tile instead of four times.
//   overlay_tile_kernel  the default (LINON_OPT_OVERLAY_KERNEL = 2): tile + halo by one TMA load per CTA (overlay_tile.cuh)
//   overlay_direct_kernel  four global taps per pixel, no shared memory (overlay_direct.cuh; LINON_OPT_OVERLAY_KERNEL = 1)
//   outline_search       the eight extreme / first-hit searches over exponents >= 0.8 that pick the "Outline"
//                        sample rays (src/main_view.wgsl:381-468), restated as ONE parallel arg-reduction with
//                        the intended stride W instead of eight serial W*H scans.
//   click_positions      sample-ray positions around the clicked fragment, src/main_view.wgsl:469-472.
//   fp32_peak_kernel     FFMA-chain microbenchmark: the measured FP32 roofline denominator.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "static_kernels.h"

namespace linon {

// ------------------------------------------------------------------------------------------------------------
// overlay
// ------------------------------------------------------------------------------------------------------------
constexpr int OV_TX = 32;   // tile width  (pixels)
constexpr int OV_TY = 16;   // tile height (pixels)
constexpr int OV_MAX_DELTA = 10;  // GUI slider range, src/application.rs:297
constexpr int OV_SW = OV_TX + 2 * OV_MAX_DELTA;
constexpr int OV_SH = OV_TY + 2 * OV_MAX_DELTA;

__device__ __forceinline__ unsigned char ov_unorm8(float x) {
    x = fminf(fmaxf(x, 0.0f), 1.0f);
    return (unsigned char)__float2int_rn(x * 255.0f);
}

// src/overlay.wgsl:15-35 (vec2squareroots) on a = (det, -trace)
__device__ __forceinline__ float2 vec2squareroots(float a0, float a1) {
    const float discrim = a1 * a1 - 4.0f * a0;
    if (discrim >= 0.0f) {
        const float root = sqrtf(discrim);
        return make_float2((-a1 - root) / 2.0f, (-a1 + root) / 2.0f);
    }
    const float root = sqrtf(-discrim);
    return make_float2(-a1 / 2.0f, root / 2.0f);
}

__global__ void __launch_bounds__(OV_TX * 8)
overlay_kernel(int W, int H, int overlay_mode, int delta, float scaling, const uchar4* __restrict__ raycast,
               const float4* __restrict__ endpoint, uchar4* __restrict__ final_rgba, float* __restrict__ exponents) {
    __shared__ float s_x[OV_SH][OV_SW + 1];
    __shared__ float s_y[OV_SH][OV_SW + 1];
    __shared__ float s_z[OV_SH][OV_SW + 1];

    const int x0 = blockIdx.x * OV_TX, y0 = blockIdx.y * OV_TY;
    const bool stencil = overlay_mode == 1;
    if (stencil) {
        // tile + halo, row by row: 32 consecutive texels per warp (512 B, fully coalesced float4 loads), no index division
        const int sw = OV_TX + 2 * delta, sh = OV_TY + 2 * delta;
        for (int sy = threadIdx.y; sy < sh; sy += 8) {
            const int gy = y0 + sy - delta;
            if (gy < 0 || gy >= H) continue;
            const float4* __restrict__ row = endpoint + (size_t)gy * W;
            for (int sx = threadIdx.x; sx < sw; sx += OV_TX) {
                const int gx = x0 + sx - delta;
                if (gx >= 0 && gx < W) {
                    const float4 e = __ldg(row + gx);
                    s_x[sy][sx] = e.x; s_y[sy][sx] = e.y; s_z[sy][sx] = e.z;
                }
            }
        }
        __syncthreads();
    }
    const int padding = 2 * delta;
#pragma unroll
    for (int r = 0; r < OV_TY / 8; ++r) {
        const int ly = threadIdx.y + 8 * r;
        const int x = x0 + threadIdx.x, y = y0 + ly;
        if (x >= W || y >= H) continue;
        const size_t idx = (size_t)y * W + x;
        uchar4 c = raycast[idx];
        const bool inner = !(x < padding || y < padding || x >= W - padding || y >= H - padding);
        if (stencil && inner) {
            const int sx = threadIdx.x + delta, sy = ly + delta;
            // lyapunov_exponent, src/overlay.wgsl:99-112
            const float gxx = 0.5f * (s_x[sy][sx + delta] - s_x[sy][sx - delta]);
            const float gxy = 0.5f * (s_y[sy][sx + delta] - s_y[sy][sx - delta]);
            const float gxz = 0.5f * (s_z[sy][sx + delta] - s_z[sy][sx - delta]);
            const float gyx = 0.5f * (s_x[sy + delta][sx] - s_x[sy - delta][sx]);
            const float gyy = 0.5f * (s_y[sy + delta][sx] - s_y[sy - delta][sx]);
            const float gyz = 0.5f * (s_z[sy + delta][sx] - s_z[sy - delta][sx]);
            const float m00 = gxx * gxx + gxy * gxy + gxz * gxz;
            const float m01 = gxx * gyx + gxy * gyy + gxz * gyz;
            const float m11 = gyx * gyx + gyy * gyy + gyz * gyz;
            const float det = m00 * m11 - m01 * m01;       // mat2det, src/overlay.wgsl:4-6
            const float2 eig = vec2squareroots(det, -(m00 + m11));
            const float exponent = sqrtf(fmaxf(eig.x, eig.y));
            const float scaled = expf(scaling * exponent - 5.0f);
            exponents[idx] = scaled;                        // stride W (the reference uses size.y)
            if (scaled >= 0.8f) {
                const float alpha = fminf(1.0f, scaled) * 0.5f;
                const float inv = 1.0f - alpha;
                c = make_uchar4(ov_unorm8(alpha * 1.0f + inv * ((float)c.x / 255.0f)),
                                ov_unorm8(alpha * 1.0f + inv * ((float)c.y / 255.0f)),
                                ov_unorm8(alpha * 1.0f + inv * ((float)c.z / 255.0f)),
                                ov_unorm8(alpha * 1.0f + inv * ((float)c.w / 255.0f)));
            }
        }
        final_rgba[idx] = c;
    }
}

}  // namespace linon
#include "overlay_direct.cuh"
namespace linon {

// Direct variant (LINON_OPT_OVERLAY_KERNEL = 1): no shared-memory staging, see overlay_direct.cuh
__global__ void __launch_bounds__(OVD_TX * OVD_WARPS, 8)  // 32 registers: 64 resident warps per SM
overlay_direct_kernel(int W, int H, int overlay_mode, int delta, float scaling, const unsigned int* __restrict__ raycast,
                      const float4* __restrict__ endpoint, unsigned int* __restrict__ final_rgba,
                      float* __restrict__ exponents) {
    overlay_direct_thread((int)threadIdx.x, (int)threadIdx.y, (int)blockIdx.x, (int)blockIdx.y, W, H, overlay_mode, delta,
                          scaling, raycast, endpoint, final_rgba, exponents);
}

// ---- tile variant (LINON_OPT_OVERLAY_KERNEL = 2, default): the endpoint tile + halo arrives by one TMA load per CTA ----
__device__ __forceinline__ uint32_t ovt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ovt_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
}  // namespace linon
#include "overlay_tile.cuh"
namespace linon {

// One CTA = one 32 x TH tile.  Thread 0 arms the mbarrier with the byte count of the box and issues the TMA; every thread
// loads its colours, then waits on the barrier's phase and reads its taps from the tile (overlay_tile_thread).
template <int TH, int D>
__global__ void __launch_bounds__(OVT_TX * OVT_WARPS)
overlay_tile_kernel(const __grid_constant__ CUtensorMap endpoint_map, int W, int H, int overlay_mode, float scaling,
                    const unsigned int* __restrict__ raycast, unsigned int* __restrict__ final_rgba,
                    float* __restrict__ exponents) {
    constexpr int bw = OVT_TX + 2 * D, bh = TH + 2 * D;
    __shared__ __align__(128) float4 tile[bw * bh];
    __shared__ __align__(8) unsigned long long bar;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int x0 = blockIdx.x * OVT_TX, y0 = blockIdx.y * TH;
    const uint32_t bar_s = ovt_smem_u32(&bar);
    if (overlay_mode == 1) {
        if (tx == 0 && ty == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tx == 0 && ty == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"((uint32_t)(bw * bh * 16)) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(ovt_smem_u32(tile)), "l"(&endpoint_map), "r"(4 * (x0 - D)), "r"(y0 - D), "r"(bar_s) : "memory");
        }
    }
    overlay_tile_thread<TH, D>(tx, ty, x0, y0, W, H, overlay_mode, scaling, raycast, tile, final_rgba, exponents,
                               [bar_s] { ovt_mbar_wait(bar_s, 0); });
}

template <int D>
static void launch_overlay_tile(cudaStream_t stream, const CUtensorMap& map, int W, int H, int overlay_mode, float scaling,
                                const void* raycast, void* final_rgba, float* exponents, int tile_rows) {
    dim3 block(OVT_TX, OVT_WARPS), grid((W + OVT_TX - 1) / OVT_TX, (H + tile_rows - 1) / tile_rows);
    if constexpr (D <= OVT_TALL_MAX_DELTA) {
        if (tile_rows == OVT_TH_TALL) {
            overlay_tile_kernel<OVT_TH_TALL, D><<<grid, block, 0, stream>>>(map, W, H, overlay_mode, scaling, (const unsigned int*)raycast,
                                                                            (unsigned int*)final_rgba, exponents);
            return;
        }
    }
    overlay_tile_kernel<OVT_TH, D><<<grid, block, 0, stream>>>(map, W, H, overlay_mode, scaling, (const unsigned int*)raycast,
                                                                   (unsigned int*)final_rgba, exponents);
}

// Tile height of a frame: 32 x 64 tiles (8 pixels per thread, 27.5 M warp instructions at 3840 x 2160 against 29.9 M, 37.7 us
// against 40.2 us under ncu) once the grid still fills every SM several times over; 32 x 32 below that (a 1000 x 777 frame:
// 12.3 us against 13.8 us).  The tensor map's box must be made for the same height (overlay_tile_box).
int overlay_tile_rows(int W, int H, int delta, int n_sm) {
    const long long tall_ctas = (long long)((W + OVT_TX - 1) / OVT_TX) * ((H + OVT_TH_TALL - 1) / OVT_TH_TALL);
    const bool fits = delta <= OVT_TALL_MAX_DELTA;  // the tall tile + halo must fit 48 KB of static shared memory
    return (fits && tall_ctas >= 12LL * (n_sm > 0 ? n_sm : 148)) ? OVT_TH_TALL : OVT_TH;
}

void overlay_tile_box(int delta, int tile_rows, unsigned int* box_floats, unsigned int* box_rows) {
    *box_floats = 4u * (unsigned int)(OVT_TX + 2 * delta);
    *box_rows = (unsigned int)(tile_rows + 2 * delta);
}

void launch_overlay(cudaStream_t stream, int W, int H, int overlay_mode, int delta, float scaling, const void* raycast,
                    const void* endpoint, void* final_rgba, float* exponents, int variant, const void* endpoint_map,
                    int tile_rows) {
    if (variant == 2 && endpoint_map && delta >= 1 && delta <= OVT_MAX_DELTA && (tile_rows == OVT_TH || (tile_rows == OVT_TH_TALL && delta <= OVT_TALL_MAX_DELTA))) {
        const CUtensorMap& map = *(const CUtensorMap*)endpoint_map;
        switch (delta) {
#define LINON_OVT_CASE(D) case D: launch_overlay_tile<D>(stream, map, W, H, overlay_mode, scaling, raycast, final_rgba, exponents, tile_rows); return;
            LINON_OVT_CASE(1) LINON_OVT_CASE(2) LINON_OVT_CASE(3) LINON_OVT_CASE(4) LINON_OVT_CASE(5)
            LINON_OVT_CASE(6) LINON_OVT_CASE(7) LINON_OVT_CASE(8) LINON_OVT_CASE(9) LINON_OVT_CASE(10)
#undef LINON_OVT_CASE
        }
    }
    if (variant >= 1) {
        dim3 block(OVD_TX, OVD_WARPS), grid((W + OVD_TX - 1) / OVD_TX, (H + OVD_TY - 1) / OVD_TY);
        overlay_direct_kernel<<<grid, block, 0, stream>>>(W, H, overlay_mode, delta, scaling, (const unsigned int*)raycast,
                                                          (const float4*)endpoint, (unsigned int*)final_rgba, exponents);
        return;
    }
    dim3 block(OV_TX, 8), grid((W + OV_TX - 1) / OV_TX, (H + OV_TY - 1) / OV_TY);
    overlay_kernel<<<grid, block, 0, stream>>>(W, H, overlay_mode, delta, scaling, (const uchar4*)raycast,
                                               (const float4*)endpoint, (uchar4*)final_rgba, exponents);
}

// ------------------------------------------------------------------------------------------------------------
// outline search: 8 simultaneous arg-reductions.  Keys are 64-bit, larger wins, 0 = nothing found.
//   "max sum, last in scan order (x outer, y inner)"  -> (sum, x, y) lexicographic maximum
//   "first found in scan order (a outer, b inner)"    -> (a, b) lexicographic minimum, stored inverted
// ------------------------------------------------------------------------------------------------------------
}  // namespace linon
#include "outline_keys.cuh"
namespace linon {

__global__ void __launch_bounds__(256) outline_search_kernel(int W, int H, const float* __restrict__ exponents,
                                                             unsigned long long* __restrict__ keys) {
    unsigned long long best[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t n = (size_t)W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (!(exponents[i] >= 0.8f)) continue;
        const unsigned int y = (unsigned int)(i / W), x = (unsigned int)(i - (size_t)y * W);
        const unsigned int uw = (unsigned int)W, uh = (unsigned int)H;
        unsigned long long k[8];
        outline_keys(x, y, uw, uh, k);
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = best[j] > k[j] ? best[j] : k[j];
    }
    __shared__ unsigned long long s_best[8][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        unsigned long long v = best[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
            v = v > other ? v : other;
        }
        if (lane == 0) s_best[j][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        unsigned long long v = 0;
        for (int w = 0; w < 8; ++w) v = v > s_best[threadIdx.x][w] ? v : s_best[threadIdx.x][w];
        if (v) atomicMax(keys + threadIdx.x, v);
    }
}

// keys -> normalised positions, src/main_view.wgsl:467-468
__global__ void outline_finalize_kernel(int W, int H, const unsigned long long* __restrict__ keys, float* __restrict__ pos16) {
    const int j = threadIdx.x;
    if (j >= 8) return;
    const unsigned long long k = keys[j];
    float px = 0.0f, py = 0.0f;
    outline_decode(j, k, W, H, px, py);
    const float width = (float)W, height = (float)H;
    pos16[2 * j] = px / (width - 1.0f);
    pos16[2 * j + 1] = (height - 1.0f - py) / (height - 1.0f);
}

void launch_outline_search(cudaStream_t stream, int W, int H, const float* exponents, unsigned long long* keys,
                           float* pos16, int n_sm) {
    cudaMemsetAsync(keys, 0, 8 * sizeof(unsigned long long), stream);
    const size_t n = (size_t)W * H;
    int blocks = (int)((n + 255) / 256);
    const int cap = n_sm * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    outline_search_kernel<<<blocks, 256, 0, stream>>>(W, H, exponents, keys);
    outline_finalize_kernel<<<1, 32, 0, stream>>>(W, H, keys, pos16);
}

// src/main_view.wgsl:316-325, :470-471
__global__ void click_positions_kernel(float mouse_x, float mouse_y, float* __restrict__ pos16) {
    const float table[8][2] = {{-1.0f, -1.0f}, {0.0f, -1.0f}, {1.0f, -1.0f}, {1.0f, 0.0f},
                               {1.0f, 1.0f},   {0.0f, 1.0f},  {-1.0f, 1.0f}, {-1.0f, 0.0f}};
    const int j = threadIdx.x;
    if (j >= 8) return;
    pos16[2 * j] = mouse_x + 0.01f * table[j][0];
    pos16[2 * j + 1] = mouse_y + 0.01f * table[j][1];
}

void launch_click_positions(cudaStream_t stream, float mouse_x, float mouse_y, float* pos16) {
    click_positions_kernel<<<1, 32, 0, stream>>>(mouse_x, mouse_y, pos16);
}

// ------------------------------------------------------------------------------------------------------------
// FP32 peak microbenchmark: 8 independent FFMA chains per thread, no memory traffic in the loop.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp32_peak_kernel(int iters, float a, float b, float* out) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.0f, x2 = x0 + 2.0f, x3 = x0 + 3.0f;
    float x4 = x0 + 4.0f, x5 = x0 + 5.0f, x6 = x0 + 6.0f, x7 = x0 + 7.0f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const float s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123456.789f) out[0] = s;  // keeps the chains alive without a store on the hot path
}

double run_fp32_peak(int n_sm, int iters, float* d_scratch, float* ms_out) {
    const int blocks = n_sm * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    fp32_peak_kernel<<<blocks, 256>>>(iters / 8 + 1, 0.999f, 0.001f, d_scratch);  // warm-up
    cudaEventRecord(e0);
    fp32_peak_kernel<<<blocks, 256>>>(iters, 0.999f, 0.001f, d_scratch);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (ms_out) *ms_out = ms;
    const double flops = 2.0 * 8.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
    return flops / (ms * 1e-3) * 1e-12;  // TFLOP/s
}

}  // namespace linon
