// MUFU throughput microbenchmark (run under gpurun): warp-instructions per clock per SM for each special-function op.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/mufu tools/ubench/mufu.cu && tools/ubench/mufu
// Each thread runs 8 independent dependent chains of one op, so the XU pipe is saturated and latency is hidden; the
// result is reported as lanes per clock per SM (16 = one warp instruction per 8 cycles per SM sub-partition x 4 ... ).
#include <cstdio>
#include <cuda_runtime.h>

#define OPS(X) X(rsqrt, "rsqrt.approx.ftz.f32") X(sqrt, "sqrt.approx.ftz.f32") X(ex2, "ex2.approx.ftz.f32") \
               X(rcp, "rcp.approx.ftz.f32") X(lg2, "lg2.approx.ftz.f32") X(sin, "sin.approx.ftz.f32") X(tanh, "tanh.approx.f32")

#define DEF(name, ptx)                                                                       \
    __global__ void k_##name(float* out, int iters, float seed) {                            \
        float x[8];                                                                          \
        for (int j = 0; j < 8; ++j) x[j] = seed + 0.01f * (threadIdx.x + j);                  \
        for (int i = 0; i < iters; ++i) {                                                    \
            _Pragma("unroll") for (int j = 0; j < 8; ++j) asm volatile(ptx " %0, %0;" : "+f"(x[j])); \
        }                                                                                    \
        float s = 0.f;                                                                       \
        for (int j = 0; j < 8; ++j) s += x[j];                                               \
        if (s == 123.456f) out[0] = s;                                                       \
    }
OPS(DEF)

// mixed: one MUFU per K FFMAs on independent registers (does MUFU issue interfere with FMA issue?)
template <int K> __global__ void k_mix(float* out, int iters, float seed) {
    float x[4], a[8];
    for (int j = 0; j < 4; ++j) x[j] = seed + 0.01f * (threadIdx.x + j);
    for (int j = 0; j < 8; ++j) a[j] = seed * j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
#pragma unroll
            for (int k = 0; k < K; ++k) asm volatile("fma.rn.ftz.f32 %0, %0, %1, %1;" : "+f"(a[(j * K + k) & 7]) : "f"(seed));
        }
    }
    float s = 0.f;
    for (int j = 0; j < 4; ++j) s += x[j];
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456f) out[0] = s;
}

template <class F> static float time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount, blocks = sms * 8, threads = 256, iters = 20000;
    float* out;
    cudaMalloc(&out, 4);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f", p.name, sms, khz / 1e3);
#define RUN(name, ptx)                                                                                   \
    {                                                                                                    \
        const float ms = time_ms([&] { k_##name<<<blocks, threads>>>(out, iters, 1.5f); });              \
        const double lanes = (double)blocks * threads * iters * 8;                                       \
        printf(", \"%s_lanes_per_clk_per_sm\": %.2f", #name, lanes / (ms * 1e-3) / (khz * 1e3) / sms);  \
    }
    OPS(RUN)
#define RUNMIX(K)                                                                                        \
    {                                                                                                    \
        const float ms = time_ms([&] { k_mix<K><<<blocks, threads>>>(out, iters, 1.5f); });              \
        const double warp_insts = (double)blocks * threads / 32 * iters * 4 * (1 + K);                   \
        printf(", \"mix_rsqrt_%dffma_warp_inst_per_clk_per_smsp\": %.3f", K, warp_insts / (ms * 1e-3) / (khz * 1e3) / sms / 4); \
    }
    RUNMIX(2) RUNMIX(4) RUNMIX(8) RUNMIX(12) RUNMIX(16)
    printf("}\n");
    return 0;
}
