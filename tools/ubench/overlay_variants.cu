// overlay_variants.cu — candidates for the overlay stencil (src/overlay.wgsl:99-146) timed side by side on the C3 frame
// (developer tool, run under gpurun; built by tools/r2_call2.sh).  Every variant computes the same arithmetic in the same
// order under --fmad=false as the shipped direct kernel (csrc/device/overlay_direct.cuh) and is compared with it bit for
// bit before it is timed; the L2 is flushed between launches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo -I linon_b200/csrc/device -o overlay_variants overlay_variants.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "overlay_direct.cuh"

using namespace linon;

// the per-pixel arithmetic of overlay_direct_thread from the four taps on (shared by the variants below)
__device__ __forceinline__ void ov_pixel(const float4 xp, const float4 xm, const float4 yp, const float4 ym, float scaling,
                                         unsigned int& c, float& scaled) {
    const float gxx = 0.5f * (xp.x - xm.x);
    const float gxy = 0.5f * (xp.y - xm.y);
    const float gxz = 0.5f * (xp.z - xm.z);
    const float gyx = 0.5f * (yp.x - ym.x);
    const float gyy = 0.5f * (yp.y - ym.y);
    const float gyz = 0.5f * (yp.z - ym.z);
    const float m00 = gxx * gxx + gxy * gxy + gxz * gxz;
    const float m01 = gxx * gyx + gxy * gyy + gxz * gyz;
    const float m11 = gyx * gyx + gyy * gyy + gyz * gyz;
    const float det = m00 * m11 - m01 * m01;
    float e0, e1;
    ovd_squareroots(det, -(m00 + m11), e0, e1);
    const float exponent = sqrtf(fmaxf(e0, e1));
    scaled = expf(scaling * exponent - 5.0f);
    if (scaled >= 0.8f) {
        const float alpha = fminf(1.0f, scaled) * 0.5f;
        const float inv = 1.0f - alpha;
        const unsigned int cx = c & 0xffu, cy = (c >> 8) & 0xffu, cz = (c >> 16) & 0xffu, cw = c >> 24;
        c = ovd_unorm8(alpha * 1.0f + inv * ((float)cx / 255.0f)) | (ovd_unorm8(alpha * 1.0f + inv * ((float)cy / 255.0f)) << 8) |
            (ovd_unorm8(alpha * 1.0f + inv * ((float)cz / 255.0f)) << 16) | (ovd_unorm8(alpha * 1.0f + inv * ((float)cw / 255.0f)) << 24);
    }
}

// V0: the shipped kernel
__global__ void __launch_bounds__(256, 8) v0_shipped(int W, int H, int mode, int delta, float scaling, const unsigned int* rc,
                                                     const float4* ep, unsigned int* fin, float* ex) {
    overlay_direct_thread((int)threadIdx.x, (int)threadIdx.y, (int)blockIdx.x, (int)blockIdx.y, W, H, mode, delta, scaling, rc, ep, fin, ex);
}
// V1: the same body with room for 64 registers (all eight taps of the thread's two rows in flight)
__global__ void __launch_bounds__(256, 4) v1_regs64(int W, int H, int mode, int delta, float scaling, const unsigned int* rc,
                                                    const float4* ep, unsigned int* fin, float* ex) {
    overlay_direct_thread((int)threadIdx.x, (int)threadIdx.y, (int)blockIdx.x, (int)blockIdx.y, W, H, mode, delta, scaling, rc, ep, fin, ex);
}

// V2/V3: ROWS consecutive rows per thread, all loads issued before any arithmetic; vertical taps shared between the rows of
// the strip when delta == 1 (row y's upper tap is row y+2's lower tap).  Block = (32, WARPS): a CTA covers 32 x (WARPS*ROWS).
template <int ROWS, int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) v_strip(int W, int H, int mode, int delta, float scaling,
                                                            const unsigned int* __restrict__ rc, const float4* __restrict__ ep,
                                                            unsigned int* __restrict__ fin, float* __restrict__ ex) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y0 = (blockIdx.y * WARPS + threadIdx.y) * ROWS;
    if (x >= W || y0 >= H) return;
    const int padding = 2 * delta;
    const bool inner_x = !(x < padding || x >= W - padding);
    const bool stencil = mode == 1;
    unsigned int c[ROWS];
    bool in[ROWS], st[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int y = y0 + r;
        in[r] = y < H;
        st[r] = in[r] && stencil && inner_x && !(y < padding || y >= H - padding);
        c[r] = in[r] ? rc[(long long)y * W + x] : 0u;
    }
    float4 xp[ROWS], xm[ROWS], yp[ROWS], ym[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (st[r]) {
            const float4* e = ep + (long long)(y0 + r) * W + x;
            xp[r] = __ldg(e + delta);
            xm[r] = __ldg(e - delta);
            yp[r] = __ldg(e + (long long)delta * W);
            ym[r] = __ldg(e - (long long)delta * W);
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        if (!in[r]) continue;
        const long long idx = (long long)(y0 + r) * W + x;
        if (st[r]) {
            float scaled;
            ov_pixel(xp[r], xm[r], yp[r], ym[r], scaling, c[r], scaled);
            ex[idx] = scaled;
        }
        fin[idx] = c[r];
    }
}

// V4: two horizontally adjacent pixels per thread (64-wide tiles): the inner x taps of the pair are each other's neighbours'
// columns, loads are 2 x float4 per tap row; stores are 8-byte.  delta is general.
template <int WARPS, int MINB>
__global__ void __launch_bounds__(32 * WARPS, MINB) v_pair(int W, int H, int mode, int delta, float scaling,
                                                           const unsigned int* __restrict__ rc, const float4* __restrict__ ep,
                                                           unsigned int* __restrict__ fin, float* __restrict__ ex) {
    const int x = (blockIdx.x * 32 + threadIdx.x) * 2;
    const int y = blockIdx.y * WARPS + threadIdx.y;
    if (x >= W || y >= H) return;
    const int padding = 2 * delta;
    const bool row_inner = mode == 1 && !(y < padding || y >= H - padding);
    const long long idx = (long long)y * W + x;
    const bool two = x + 1 < W;
    unsigned int c0 = rc[idx], c1 = two ? rc[idx + 1] : 0u;
    const bool s0 = row_inner && !(x < padding || x >= W - padding);
    const bool s1 = two && row_inner && !(x + 1 < padding || x + 1 >= W - padding);
    float4 xp0, xm0, yp0, ym0, xp1, xm1, yp1, ym1;
    const float4* e = ep + idx;
    if (s0) { xp0 = __ldg(e + delta); xm0 = __ldg(e - delta); yp0 = __ldg(e + (long long)delta * W); ym0 = __ldg(e - (long long)delta * W); }
    if (s1) { xp1 = __ldg(e + 1 + delta); xm1 = __ldg(e + 1 - delta); yp1 = __ldg(e + 1 + (long long)delta * W); ym1 = __ldg(e + 1 - (long long)delta * W); }
    float sc0 = 0.0f, sc1 = 0.0f;
    if (s0) { ov_pixel(xp0, xm0, yp0, ym0, scaling, c0, sc0); ex[idx] = sc0; }
    if (s1) { ov_pixel(xp1, xm1, yp1, ym1, scaling, c1, sc1); ex[idx + 1] = sc1; }
    fin[idx] = c0;
    if (two) fin[idx + 1] = c1;
}


// ---- TMA variants: the endpoint tile + halo arrives by ONE cp.async.bulk.tensor.2d per tile (tensor map over the endpoint
// image as a (4W, H) fp32 tensor, out-of-frame texels zero-filled - border pixels never use their taps), no per-thread
// address arithmetic for the taps: four LDS.128 from the dense tile.  Colours and outputs stay plain coalesced LDG/STG.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// T1: one 32 x TH tile per CTA, 256 threads, TH/8 rows per thread.  DT = compile-time delta (0: run-time, halo up to 10).
template <int TH, int MINB, int DT>
__global__ void __launch_bounds__(256, MINB) v_tma1(const __grid_constant__ CUtensorMap tm, int W, int H, int mode, int delta,
                                                    float scaling, const unsigned int* __restrict__ rc,
                                                    unsigned int* __restrict__ fin, float* __restrict__ ex) {
    constexpr int ROWS = TH / 8;
    constexpr int DMAX = DT ? DT : 10;
    __shared__ __align__(128) float4 tile[(32 + 2 * DMAX) * (TH + 2 * DMAX)];
    __shared__ __align__(8) unsigned long long bar;
    const int d = DT ? DT : delta;
    const int bw = 32 + 2 * d;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * TH;
    const bool stencil = mode == 1;
    const uint32_t bar_s = smem_u32(&bar);
    if (stencil) {
        if (tid == 0) {
            mbar_init(bar_s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar_s, (uint32_t)(bw * (TH + 2 * d) * 16));
            tma_load_2d(smem_u32(tile), &tm, 4 * (x0 - d), y0 - d, bar_s);
        }
    }
    const int x = x0 + tx;
    const int padding = 2 * d;
    const bool inner_x = x >= padding && x < W - padding;
    unsigned int c[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int y = y0 + ty + 8 * r;
        c[r] = (x < W && y < H) ? rc[y * W + x] : 0u;
    }
    if (stencil) mbar_wait(bar_s, 0);
    const float4* t0 = tile + (ty + d) * bw + (tx + d);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int y = y0 + ty + 8 * r;
        if (x >= W || y >= H) continue;
        const int idx = y * W + x;
        if (stencil && inner_x && y >= padding && y < H - padding) {
            const float4* t = t0 + 8 * r * bw;
            float scaled;
            ov_pixel(t[d], t[-d], t[d * bw], t[-d * bw], scaling, c[r], scaled);
            ex[idx] = scaled;
        }
        fin[idx] = c[r];
    }
}


// the same arithmetic with ONE square root for both signs of the discriminant: sqrt(|discrim|) is sqrt(discrim) for
// discrim >= 0 and sqrt(-discrim) otherwise, bit for bit (discrim = -0.0: sqrt gives -0.0 against +0.0 here, and the two sums
// it enters, -a1 - root and -a1 + root with -a1 >= +0, are +0 or -a1 either way)
__device__ __forceinline__ void ov_pixel1(const float4 xp, const float4 xm, const float4 yp, const float4 ym, float scaling,
                                          unsigned int& c, float& scaled) {
    const float gxx = 0.5f * (xp.x - xm.x);
    const float gxy = 0.5f * (xp.y - xm.y);
    const float gxz = 0.5f * (xp.z - xm.z);
    const float gyx = 0.5f * (yp.x - ym.x);
    const float gyy = 0.5f * (yp.y - ym.y);
    const float gyz = 0.5f * (yp.z - ym.z);
    const float m00 = gxx * gxx + gxy * gxy + gxz * gxz;
    const float m01 = gxx * gyx + gxy * gyy + gxz * gyz;
    const float m11 = gyx * gyx + gyy * gyy + gyz * gyz;
    const float a0 = m00 * m11 - m01 * m01;
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
        c = ovd_unorm8(alpha * 1.0f + inv * ((float)cx / 255.0f)) | (ovd_unorm8(alpha * 1.0f + inv * ((float)cy / 255.0f)) << 8) |
            (ovd_unorm8(alpha * 1.0f + inv * ((float)cz / 255.0f)) << 16) | (ovd_unorm8(alpha * 1.0f + inv * ((float)cw / 255.0f)) << 24);
    }
}

// T3: T1 with (i) tiles that touch neither the frame's edge nor its 2 delta border take a loop without any bounds test (the
// flag is uniform per CTA), (ii) ov_pixel1.  DT = compile-time delta only.
template <int TH, int MINB, int DT>
__global__ void __launch_bounds__(256, MINB) v_tma3(const __grid_constant__ CUtensorMap tm, int W, int H, int mode, int delta,
                                                    float scaling, const unsigned int* __restrict__ rc,
                                                    unsigned int* __restrict__ fin, float* __restrict__ ex) {
    constexpr int ROWS = TH / 8;
    constexpr int d = DT, bw = 32 + 2 * DT, padding = 2 * DT;
    __shared__ __align__(128) float4 tile[bw * (TH + 2 * DT)];
    __shared__ __align__(8) unsigned long long bar;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * TH;
    const bool stencil = mode == 1;
    const uint32_t bar_s = smem_u32(&bar);
    if (stencil) {
        if (tid == 0) {
            mbar_init(bar_s, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar_s, (uint32_t)(bw * (TH + 2 * d) * 16));
            tma_load_2d(smem_u32(tile), &tm, 4 * (x0 - d), y0 - d, bar_s);
        }
    }
    const int x = x0 + tx;
    const float4* t0 = tile + (ty + d) * bw + (tx + d);
    const bool interior = stencil && x0 >= padding && x0 + 32 <= W - padding && y0 >= padding && y0 + TH <= H - padding;
    if (interior) {
        const int idx0 = (y0 + ty) * W + x;
        unsigned int c[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) c[r] = rc[idx0 + 8 * r * W];
        mbar_wait(bar_s, 0);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float4* t = t0 + 8 * r * bw;
            float scaled;
            ov_pixel1(t[d], t[-d], t[d * bw], t[-d * bw], scaling, c[r], scaled);
            ex[idx0 + 8 * r * W] = scaled;
            fin[idx0 + 8 * r * W] = c[r];
        }
        return;
    }
    const bool inner_x = x >= padding && x < W - padding;
    unsigned int c[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int y = y0 + ty + 8 * r;
        c[r] = (x < W && y < H) ? rc[y * W + x] : 0u;
    }
    if (stencil) mbar_wait(bar_s, 0);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
        const int y = y0 + ty + 8 * r;
        if (x >= W || y >= H) continue;
        const int idx = y * W + x;
        if (stencil && inner_x && y >= padding && y < H - padding) {
            const float4* t = t0 + 8 * r * bw;
            float scaled;
            ov_pixel1(t[d], t[-d], t[d * bw], t[-d * bw], scaling, c[r], scaled);
            ex[idx] = scaled;
        }
        fin[idx] = c[r];
    }
}

// T2: persistent CTAs, S-stage ring of tiles: the TMA of tile k+S is issued as soon as every thread has finished reading
// the buffer of tile k; tiles are handed out in row-major order so CTAs that run together share their halos in L2.
template <int TH, int MINB, int S, int DT>
__global__ void __launch_bounds__(256, MINB) v_tma2(const __grid_constant__ CUtensorMap tm, int W, int H, int mode, int delta,
                                                    float scaling, const unsigned int* __restrict__ rc,
                                                    unsigned int* __restrict__ fin, float* __restrict__ ex) {
    constexpr int ROWS = TH / 8;
    constexpr int DMAX = DT ? DT : 10;
    constexpr int TILE = ((32 + 2 * DMAX) * (TH + 2 * DMAX) + 7) / 8 * 8;  // every stage 128-byte aligned (TMA destination)
    __shared__ __align__(128) float4 tile[S][TILE];
    __shared__ __align__(8) unsigned long long bar[S];
    const int d = DT ? DT : delta;
    const int bw = 32 + 2 * d;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int tiles_x = (W + 31) / 32, tiles_y = (H + TH - 1) / TH, n_tiles = tiles_x * tiles_y;
    const bool stencil = mode == 1;
    const uint32_t bytes = (uint32_t)(bw * (TH + 2 * d) * 16);
    if (stencil) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < S; ++s) mbar_init(smem_u32(&bar[s]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int t = (int)blockIdx.x + s * (int)gridDim.x;
                if (t < n_tiles) {
                    mbar_expect_tx(smem_u32(&bar[s]), bytes);
                    tma_load_2d(smem_u32(tile[s]), &tm, 4 * ((t % tiles_x) * 32 - d), (t / tiles_x) * TH - d, smem_u32(&bar[s]));
                }
            }
        }
    }
    const int padding = 2 * d;
    int k = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++k) {
        const int s = k % S;
        const uint32_t parity = (uint32_t)(k / S) & 1u;
        const int x0 = (t % tiles_x) * 32, y0 = (t / tiles_x) * TH;
        const int x = x0 + tx;
        const bool inner_x = x >= padding && x < W - padding;
        unsigned int c[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int y = y0 + ty + 8 * r;
            c[r] = (x < W && y < H) ? rc[y * W + x] : 0u;
        }
        if (stencil) mbar_wait(smem_u32(&bar[s]), parity);
        const float4* t0 = tile[s] + (ty + d) * bw + (tx + d);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const int y = y0 + ty + 8 * r;
            if (x >= W || y >= H) continue;
            const int idx = y * W + x;
            if (stencil && inner_x && y >= padding && y < H - padding) {
                const float4* tp = t0 + 8 * r * bw;
                float scaled;
                ov_pixel(tp[d], tp[-d], tp[d * bw], tp[-d * bw], scaling, c[r], scaled);
                ex[idx] = scaled;
            }
            fin[idx] = c[r];
        }
        if (stencil) {
            __syncthreads();  // every thread has read its taps: the buffer may be refilled
            const int tn = t + S * (int)gridDim.x;
            if (tid == 0 && tn < n_tiles) {
                mbar_expect_tx(smem_u32(&bar[s]), bytes);
                tma_load_2d(smem_u32(tile[s]), &tm, 4 * ((tn % tiles_x) * 32 - d), (tn / tiles_x) * TH - d, smem_u32(&bar[s]));
            }
        }
    }
}

// floor of the access pattern: 20 B/px read, 8 B/px written, no arithmetic to speak of (NOT a candidate: results differ)
__global__ void __launch_bounds__(256, 8) v_copy_floor(int W, int H, int mode, int delta, float scaling, const unsigned int* __restrict__ rc,
                                                       const float4* __restrict__ ep, unsigned int* __restrict__ fin, float* __restrict__ ex) {
    const int x = blockIdx.x * 32 + threadIdx.x;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int y = blockIdx.y * 16 + threadIdx.y + 8 * r;
        if (x >= W || y >= H) continue;
        const int idx = y * W + x;
        const float4 e = __ldg(ep + idx);
        ex[idx] = e.x + e.y * scaling + e.z;
        fin[idx] = rc[idx] + (unsigned int)delta;
    }
}

static CUtensorMap g_tm;
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void make_map(const float4* ep, int W, int H, int delta, int TH) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
    const cuuint64_t dims[2] = {(cuuint64_t)W * 4, (cuuint64_t)H};
    const cuuint64_t strides[1] = {(cuuint64_t)W * 16};
    const cuuint32_t box[2] = {(cuuint32_t)(4 * (32 + 2 * delta)), (cuuint32_t)(TH + 2 * delta)};
    const cuuint32_t es[2] = {1, 1};
    const CUresult r = ((EncodeTiled)fn)(&g_tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ep, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
}
static int g_sms = 148;
#define TMA1(TH, MINB, DT)                                                                                                        \
    [](int W, int H, int m, int d, float s, const unsigned int* rc, const float4* ep, unsigned int* fin, float* ex) {             \
        if (DT && d != DT) return;                                                                                                \
        make_map(ep, W, H, d, TH);                                                                                                \
        v_tma1<TH, MINB, DT><<<dim3((W + 31) / 32, (H + TH - 1) / TH), dim3(32, 8)>>>(g_tm, W, H, m, d, s, rc, fin, ex);           \
    }
#define TMA3(TH, MINB, DT)                                                                                                        \
    [](int W, int H, int m, int d, float s, const unsigned int* rc, const float4* ep, unsigned int* fin, float* ex) {             \
        if (d != DT) return;                                                                                                      \
        make_map(ep, W, H, d, TH);                                                                                                \
        v_tma3<TH, MINB, DT><<<dim3((W + 31) / 32, (H + TH - 1) / TH), dim3(32, 8)>>>(g_tm, W, H, m, d, s, rc, fin, ex);           \
    }
#define TMA2(TH, MINB, S, DT)                                                                                                     \
    [](int W, int H, int m, int d, float s, const unsigned int* rc, const float4* ep, unsigned int* fin, float* ex) {             \
        if (DT && d != DT) return;                                                                                                \
        make_map(ep, W, H, d, TH);                                                                                                \
        const int tiles = ((W + 31) / 32) * ((H + TH - 1) / TH);                                                                  \
        const int grid = tiles < g_sms * MINB ? tiles : g_sms * MINB;                                                             \
        v_tma2<TH, MINB, S, DT><<<grid, dim3(32, 8)>>>(g_tm, W, H, m, d, s, rc, fin, ex);                                          \
    }

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

struct Variant {
    const char* name;
    void (*launch)(int, int, int, int, float, const unsigned int*, const float4*, unsigned int*, float*);
};

template <class K> void launch_grid(K k, dim3 grid, dim3 block, int W, int H, int mode, int delta, float sc, const unsigned int* rc,
                                    const float4* ep, unsigned int* fin, float* ex) {
    k<<<grid, block>>>(W, H, mode, delta, sc, rc, ep, fin, ex);
}

#define STRIP(ROWS, WARPS, MINB)                                                                                                  \
    [](int W, int H, int m, int d, float s, const unsigned int* rc, const float4* ep, unsigned int* fin, float* ex) {             \
        launch_grid(v_strip<ROWS, WARPS, MINB>, dim3((W + 31) / 32, (H + ROWS * WARPS - 1) / (ROWS * WARPS)), dim3(32, WARPS), W, H, m, \
                    d, s, rc, ep, fin, ex);                                                                                       \
    }
#define PAIR(WARPS, MINB)                                                                                                         \
    [](int W, int H, int m, int d, float s, const unsigned int* rc, const float4* ep, unsigned int* fin, float* ex) {             \
        launch_grid(v_pair<WARPS, MINB>, dim3((W + 63) / 64, (H + WARPS - 1) / WARPS), dim3(32, WARPS), W, H, m, d, s, rc, ep, fin, ex); \
    }

int main(int argc, char** argv) {
    const int W = argc > 1 ? atoi(argv[1]) : 3840, H = argc > 2 ? atoi(argv[2]) : 2160, delta = argc > 3 ? atoi(argv[3]) : 1;
    const int reps = argc > 4 ? atoi(argv[4]) : 12;  // 0: one launch per variant (ncu captures), no timing
    const size_t n = (size_t)W * H;
    std::vector<float4> h_ep(n);
    std::vector<unsigned int> h_rc(n);
    // a smooth endpoint field with a few folds (so that some exponents pass the 0.8 threshold), like a mirage frame
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float u = (float)x / W, v = (float)y / H;
            const float bump = 0.3f * expf(-40.0f * ((u - 0.4f) * (u - 0.4f) + (v - 0.5f) * (v - 0.5f)));
            h_ep[(size_t)y * W + x] = make_float4(-2.0f + 3.0f * u + bump * sinf(900.0f * u), 1.5f * v + bump, -4.0f + 0.2f * u * v + bump * cosf(700.0f * v), 1.0f);
            h_rc[(size_t)y * W + x] = 0xff000000u | ((x * 7 + y * 3) & 0xff) | (((x ^ y) & 0xff) << 8) | (((x + y) & 0xff) << 16);
        }
    float4* ep; unsigned int *rc, *fin, *fin_ref; float *ex, *ex_ref; char* flush;
    CK(cudaMalloc(&ep, n * 16)); CK(cudaMalloc(&rc, n * 4)); CK(cudaMalloc(&fin, n * 4)); CK(cudaMalloc(&fin_ref, n * 4));
    CK(cudaMalloc(&ex, n * 4)); CK(cudaMalloc(&ex_ref, n * 4)); CK(cudaMalloc(&flush, 256 << 20));
    CK(cudaMemcpy(ep, h_ep.data(), n * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(rc, h_rc.data(), n * 4, cudaMemcpyHostToDevice));
    const float scaling = 50.0f;
    { cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); g_sms = prop.multiProcessorCount; }
    Variant variants[] = {
        {"v0 shipped direct (32x16 tile, 2 rows/thread, 32 regs)",
         [](int W, int H, int m, int d, float s, const unsigned int* rc, const float4* ep, unsigned int* fin, float* ex) {
             launch_grid(v0_shipped, dim3((W + 31) / 32, (H + 15) / 16), dim3(32, 8), W, H, m, d, s, rc, ep, fin, ex); }},
        {"v1 same body, 64 regs",
         [](int W, int H, int m, int d, float s, const unsigned int* rc, const float4* ep, unsigned int* fin, float* ex) {
             launch_grid(v1_regs64, dim3((W + 31) / 32, (H + 15) / 16), dim3(32, 8), W, H, m, d, s, rc, ep, fin, ex); }},
        {"strip 1 row  x 8 warps, min 8 CTAs", STRIP(1, 8, 8)},
        {"strip 1 row  x 4 warps, min 16 CTAs", STRIP(1, 4, 16)},
        {"strip 2 rows x 8 warps, min 4 CTAs", STRIP(2, 8, 4)},
        {"strip 2 rows x 8 warps, min 6 CTAs", STRIP(2, 8, 6)},
        {"strip 2 rows x 4 warps, min 8 CTAs", STRIP(2, 4, 8)},
        {"strip 4 rows x 8 warps, min 2 CTAs", STRIP(4, 8, 2)},
        {"strip 4 rows x 4 warps, min 4 CTAs", STRIP(4, 4, 4)},
        {"strip 4 rows x 4 warps, min 6 CTAs", STRIP(4, 4, 6)},
        {"pair 8 warps, min 4 CTAs", PAIR(8, 4)},
        {"pair 8 warps, min 6 CTAs", PAIR(8, 6)},
        {"pair 4 warps, min 8 CTAs", PAIR(4, 8)},
        {"tma1 32x16 tile/CTA, min 8 CTAs, delta=1 fixed", TMA1(16, 8, 1)},
        {"tma1 32x32 tile/CTA, min 6 CTAs, delta=1 fixed", TMA1(32, 6, 1)},
        {"tma1 32x32 tile/CTA, min 4 CTAs, delta=1 fixed", TMA1(32, 4, 1)},
        {"tma1 32x32 tile/CTA, min 8 CTAs, delta=1 fixed", TMA1(32, 8, 1)},
        {"tma1 32x64 tile/CTA, min 3 CTAs, delta=1 fixed", TMA1(64, 3, 1)},
        {"tma1 32x64 tile/CTA, min 4 CTAs, delta=1 fixed", TMA1(64, 4, 1)},
        {"tma1 32x16 tile/CTA, min 6 CTAs, run-time delta", TMA1(16, 6, 0)},
        {"tma3 32x32 interior fast path + one sqrt, min 6, delta=1", TMA3(32, 6, 1)},
        {"tma3 32x32 interior fast path + one sqrt, min 4, delta=1", TMA3(32, 4, 1)},
        {"tma3 32x64 interior fast path + one sqrt, min 3, delta=1", TMA3(64, 3, 1)},
        {"tma3 32x16 interior fast path + one sqrt, min 8, delta=1", TMA3(16, 8, 1)},
        {"tma3 32x32 interior fast path + one sqrt, min 4, delta=5", TMA3(32, 4, 5)},
        {"tma2 persistent 32x16, 8 CTAs/SM, 2 stages, delta=1", TMA2(16, 8, 2, 1)},
        {"tma2 persistent 32x16, 6 CTAs/SM, 3 stages, delta=1", TMA2(16, 6, 3, 1)},
        {"tma2 persistent 32x32, 4 CTAs/SM, 2 stages, delta=1", TMA2(32, 4, 2, 1)},
        {"tma2 persistent 32x32, 6 CTAs/SM, 2 stages, delta=1", TMA2(32, 6, 2, 1)},
        {"tma2 persistent 32x16, 6 CTAs/SM, 1 stage, run-time delta", TMA2(16, 6, 1, 0)},
        {"(floor) 20 B/px in, 8 B/px out, no arithmetic - NOT a candidate", 
         [](int W, int H, int m, int d, float s, const unsigned int* rc, const float4* ep, unsigned int* fin, float* ex) {
             launch_grid(v_copy_floor, dim3((W + 31) / 32, (H + 15) / 16), dim3(32, 8), W, H, m, d, s, rc, ep, fin, ex); }},
    };
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<unsigned int> a(n), b(n);
    std::vector<float> fa(n), fb(n);
    variants[0].launch(W, H, 1, delta, scaling, rc, ep, fin_ref, ex_ref);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(a.data(), fin_ref, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(fa.data(), ex_ref, n * 4, cudaMemcpyDeviceToHost));
    size_t blended = 0;
    for (size_t i = 0; i < n; ++i) blended += fa[i] >= 0.8f;
    printf("overlay variants, %dx%d delta %d, %zu pixels blended; algorithmic bytes %zu\n", W, H, delta, blended, n * 28);
    const double bytes = (double)n * 28.0;
    for (auto& v : variants) {
        CK(cudaMemset(fin, 0, n * 4)); CK(cudaMemset(ex, 0, n * 4));
        v.launch(W, H, 1, delta, scaling, rc, ep, fin, ex);
        CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(b.data(), fin, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(fb.data(), ex, n * 4, cudaMemcpyDeviceToHost));
        const bool same = memcmp(a.data(), b.data(), n * 4) == 0 && memcmp(fa.data(), fb.data(), n * 4) == 0;
        float best = 1e9f, sum = 0.0f;
        for (int i = 0; i < reps; ++i) {
            CK(cudaMemsetAsync(flush, i, 256 << 20));
            CK(cudaEventRecord(e0));
            v.launch(W, H, 1, delta, scaling, rc, ep, fin, ex);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (i >= 2) { sum += ms; best = fminf(best, ms); }
        }
        const float mean = sum / (reps - 2);
        if (reps <= 2) { printf("%-64s bit-equal %d\n", v.name, (int)same); continue; }
        printf("%-64s bit-equal %d  mean %.2f us  min %.2f us  %.0f GB/s (mean)\n", v.name, (int)same, mean * 1e3f, best * 1e3f, bytes / (mean * 1e-3) / 1e9);
    }
    return 0;
}
