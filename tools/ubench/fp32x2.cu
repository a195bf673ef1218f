// fp32x2.cu — pipe-throughput microbenchmark for the B200 march kernel design (developer tool, run under gpurun).
// Measures per-SM-cycle throughput of FFMA, FFMA2 (fma.rn.f32x2), FMUL2/FADD2, MUFU.{RCP,EX2,RSQ,SQRT} and mixes of
// FFMA2 + MUFU + ALU, to decide whether a two-rays-per-thread packed kernel can relieve the issue-slot bound.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2 fp32x2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define ITER 4096
#define CHAINS 8

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, float seed) {
    float2 x[CHAINS];
    float2 y = make_float2(seed * 0.999f, seed * 1.001f), z = make_float2(0.25f * seed, 0.5f * seed);
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = make_float2(threadIdx.x * 1e-3f + c, threadIdx.x * 2e-3f - c);
    float m[4] = {1.0f + threadIdx.x, 2.0f + threadIdx.x, 3.0f + threadIdx.x, 4.0f + threadIdx.x};
    for (int it = 0; it < ITER; ++it) {
        if (MODE == 0) {  // scalar FFMA, 16 independent chains
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) { x[c].x = fmaf(x[c].x, y.x, z.x); x[c].y = fmaf(x[c].y, y.y, z.y); }
        } else if (MODE == 1) {  // FFMA2, 8 independent chains
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = __ffma2_rn(x[c], y, z);
        } else if (MODE == 2) {  // FMUL2 + FADD2 kept apart
#pragma unroll
            for (int c = 0; c < CHAINS; c += 2) { x[c] = __fmul2_rn(x[c], y); x[c + 1] = __fadd2_rn(x[c + 1], z); }
#pragma unroll
            for (int c = 0; c < CHAINS; c += 2) { x[c] = __fmul2_rn(x[c], y); x[c + 1] = __fadd2_rn(x[c + 1], z); }
        } else if (MODE == 3) {  // MUFU.RCP only, 4 chains x 4
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) m[c] = __frcp_rn(m[c]);
        } else if (MODE == 4) {  // MUFU.EX2
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
        } else if (MODE == 5) {  // MUFU.RSQ
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
        } else if (MODE == 6) {  // was rcp; now ex2 again
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
        } else if (MODE == 7) {  // mix: 16 FFMA2 : 4 MUFU  (the packed march kernel's expected ratio is ~ 38 : 12)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) x[c] = __ffma2_rn(x[c], y, z);
#pragma unroll
            for (int c = 0; c < 4; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
        } else if (MODE == 8) {  // mix: 16 FFMA2 : 8 MUFU
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) x[c] = __ffma2_rn(x[c], y, z);
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
        } else if (MODE == 9) {  // mix: 32 scalar FFMA : 4 MUFU (today's kernel ratio, for comparison)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) { x[c].x = fmaf(x[c].x, y.x, z.x); x[c].y = fmaf(x[c].y, y.y, z.y); }
#pragma unroll
            for (int c = 0; c < 4; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
        } else if (MODE == 10) {  // MUFU.SQRT
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
        } else if (MODE == 11) {  // mix: 16 FFMA2 : 4 MUFU : 8 FMNMX (alu pipe)
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) x[c] = __ffma2_rn(x[c], y, z);
#pragma unroll
            for (int c = 0; c < 4; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) asm volatile("min.ftz.f32 %0, %0, %1;" : "+f"(x[c].x) : "f"(x[c].y));
        } else if (MODE == 12) {  // packed-kernel estimate per stage pair: 40 FFMA2 : 12 MUFU : 12 FMNMX
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) x[c] = __ffma2_rn(x[c], y, z);
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) asm volatile("min.ftz.f32 %0, %0, %1;" : "+f"(x[c].x) : "f"(x[c].y));
#pragma unroll
            for (int c = 0; c < 4; ++c) asm volatile("min.ftz.f32 %0, %0, %1;" : "+f"(x[c].y) : "f"(m[c]));
        } else if (MODE == 13) {  // the same work as scalar FFMA: 80 FFMA : 12 MUFU : 12 FMNMX
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) { x[c].x = fmaf(x[c].x, y.x, z.x); x[c].y = fmaf(x[c].y, y.y, z.y); }
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[c]));
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) asm volatile("min.ftz.f32 %0, %0, %1;" : "+f"(x[c].x) : "f"(x[c].y));
#pragma unroll
            for (int c = 0; c < 4; ++c) asm volatile("min.ftz.f32 %0, %0, %1;" : "+f"(x[c].y) : "f"(m[c]));
        }
    }
    float acc = m[0] + m[1] + m[2] + m[3];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc += x[c].x + x[c].y;
    if (acc == 12345.678f) out[0] = acc;
}

template <int MODE>
void run(const char* name, double fp_per_iter, double mufu_per_iter, int sms, float clock_ghz_hint) {
    float* d;
    cudaMalloc(&d, 64);
    const int grid = sms * 8, block = 256;
    bench<MODE><<<grid, block>>>(d, 1.0f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        bench<MODE><<<grid, block>>>(d, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double threads = (double)grid * block;
    const double flops = threads * ITER * fp_per_iter * 2.0;  // an FMA lane = 2 flop
    const double mufu = threads * ITER * mufu_per_iter;
    printf("{\"mode\": \"%s\", \"ms\": %.4f, \"fp32_tflops\": %.2f, \"mufu_per_clk_per_sm\": %.2f, \"fma_lanes_per_clk_per_sm\": %.2f}\n",
           name, best, flops / (best * 1e-3) / 1e12, mufu / (best * 1e-3) / (clock_ghz_hint * 1e9) / sms,
           threads * ITER * fp_per_iter / (best * 1e-3) / (clock_ghz_hint * 1e9) / sms);
    cudaFree(d);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const float ghz = p.clockRate * 1e-6f;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", p.name, sms, ghz);
    run<0>("ffma_scalar", 16, 0, sms, ghz);
    run<1>("ffma2", 16, 0, sms, ghz);
    run<2>("fmul2_fadd2", 8, 0, sms, ghz);   // counts a mul or add lane as half an fma (2 flop convention off by 2)
    run<3>("frcp_rn(IEEE)", 0, 16, sms, ghz);
    run<4>("mufu_ex2", 0, 16, sms, ghz);
    run<5>("mufu_rsq", 0, 16, sms, ghz);
    run<6>("mufu_rcp", 0, 16, sms, ghz);
    run<10>("mufu_sqrt", 0, 16, sms, ghz);
    run<7>("mix_16ffma2_4mufu", 32, 4, sms, ghz);
    run<8>("mix_16ffma2_8mufu", 32, 8, sms, ghz);
    run<9>("mix_32ffma_4mufu", 32, 4, sms, ghz);
    run<11>("mix_16ffma2_4mufu_8fmnmx", 32, 4, sms, ghz);
    run<12>("mix_40ffma2_12mufu_12fmnmx", 80, 12, sms, ghz);
    run<13>("mix_80ffma_12mufu_12fmnmx", 80, 12, sms, ghz);
    return 0;
}
