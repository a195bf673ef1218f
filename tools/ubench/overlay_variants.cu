// overlay_variants.cu — candidates for the overlay stencil (src/overlay.wgsl:99-146) timed side by side on the C3 frame
// (developer tool, run under gpurun; built by tools/r2_call2.sh).  Every variant computes the same arithmetic in the same
// order under --fmad=false as the shipped direct kernel (csrc/device/overlay_direct.cuh) and is compared with it bit for
// bit before it is timed; the L2 is flushed between launches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -lineinfo -I linon_b200/csrc/device -o overlay_variants overlay_variants.cu
#include <cuda_runtime.h>

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
        const int reps = 12;
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
        printf("%-48s bit-equal %d  mean %.2f us  min %.2f us  %.0f GB/s (mean)\n", v.name, (int)same, mean * 1e3f, best * 1e3f, bytes / (mean * 1e-3) / 1e9);
    }
    return 0;
}
