// cuda_shim.hpp — TEST INFRASTRUCTURE: just enough of the CUDA device language for g++ to compile the march module's
// NVRTC source text (wgsl_prelude.cuh + field function + march_kernels.cuh, dumped by LINON_DUMP_SOURCE) and run it
// on the host, one warp at a time.  Never part of the product: linon_b200/ neither includes nor links this.
//
// Execution model: one CTA = one warp of 32 lanes (LINON_BLOCK = 32).  Every lane is a ucontext fiber on the calling OS
// thread; a lane runs until it reaches a warp collective (vote, shuffle, reduce, __syncthreads) or returns.  When every
// unfinished lane waits at a collective the operands are exchanged and the lanes resume - lockstep exactly where the
// device code asks for it, arbitrary interleaving elsewhere (what the SIMT model allows).  A full-mask collective that
// a finished lane never reached is reported as an error (on hardware it would hang).  Different warps run on
// different OS threads (OpenMP in the driver); __shared__ variables are thread_local statics, i.e. per CTA.
//
// Arithmetic: the special functions are libm's (correctly rounded sqrt / division), not the MUFU approximations, so a
// fast-math module emulated here differs from the device by the MUFU error (1-2 ulp per call) and nothing else;
// strict modules are bit-identical with the device (both IEEE).  Directed-rounding intrinsics are emulated through
// double precision (exact for sums and products of two floats).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static thread_local
#define __constant__ static const
#define __launch_bounds__(...)
#define __CUDACC_RTC__ 1  // the prelude's device branch (LINON_HD etc.)
#define __CUDA_ARCH__ 0   // no packed f32x2 instructions here (LINON_PACKED_OPS needs >= 1000)

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct uint3e { unsigned int x, y, z; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return {x, y, z, w}; }

namespace emu {
constexpr int kLanes = 32;
constexpr size_t kStack = 256 * 1024;
constexpr int kRing = 64;  // snapshots of completed collectives; at most one unread snapshot per lane
enum LaneState { RUNNABLE, WAITING, DONE };
struct Snapshot {
    unsigned long long operand[kLanes];
    unsigned int present;  // lanes that took part
};
struct Warp {
    ucontext_t main_ctx, lane_ctx[kLanes];
    char* stacks = nullptr;
    LaneState state[kLanes];
    bool wait_full[kLanes];              // the collective the lane waits at names all 32 lanes
    bool partial_hint[kLanes];           // the lane called __activemask() since its last collective
    unsigned long long operand[kLanes];  // what the lane contributed
    int snapshot_of[kLanes];             // ring slot the lane reads after it resumes
    Snapshot ring[kRing];
    int ring_next = 0;
    int cur = 0;
    void (*entry)(void*) = nullptr;
    void* arg = nullptr;
    bool failed = false;
};
inline thread_local Warp* g_warp = nullptr;
inline thread_local uint3e g_thread_idx = {0, 0, 0}, g_block_idx = {0, 0, 0}, g_block_dim = {kLanes, 1, 1}, g_grid_dim = {1, 1, 1};

inline void lane_trampoline() {
    Warp* w = g_warp;
    w->entry(w->arg);
    w->state[w->cur] = DONE;
    swapcontext(&w->lane_ctx[w->cur], &w->main_ctx);
}

inline void release(Warp* w, unsigned int group) {
    Snapshot& s = w->ring[w->ring_next];
    s.present = group;
    for (int l = 0; l < kLanes; ++l)
        if (group >> l & 1u) {
            s.operand[l] = w->operand[l];
            w->snapshot_of[l] = w->ring_next;
            w->state[l] = RUNNABLE;
        }
    w->ring_next = (w->ring_next + 1) % kRing;
}

// Runs `entry(arg)` on 32 lanes; returns false if a full-mask collective could not be completed by all 32 lanes.
// Scheduling: every runnable lane runs until it waits at a collective or returns.  Then (i) the lanes waiting at
// collectives with a PARTIAL mask (votes among "the lanes still in this loop", __activemask()) complete among
// themselves - which lanes share such a vote is the scheduler's choice on hardware too, and the device code must not
// depend on it; (ii) otherwise the lanes waiting at a FULL-mask collective complete once all 32 are there.
inline bool run_warp(void (*entry)(void*), void* arg, unsigned int block_idx) {
    static thread_local Warp warp;
    Warp* w = &warp;
    if (!w->stacks) w->stacks = (char*)malloc(kStack * kLanes);
    g_warp = w;
    w->entry = entry;
    w->arg = arg;
    w->failed = false;
    g_block_idx = {block_idx, 0, 0};
    for (int l = 0; l < kLanes; ++l) {
        w->state[l] = RUNNABLE;
        w->partial_hint[l] = false;
        getcontext(&w->lane_ctx[l]);
        w->lane_ctx[l].uc_stack.ss_sp = w->stacks + kStack * l;
        w->lane_ctx[l].uc_stack.ss_size = kStack;
        w->lane_ctx[l].uc_link = &w->main_ctx;
        makecontext(&w->lane_ctx[l], (void (*)())lane_trampoline, 0);
    }
    for (;;) {
        for (int l = 0; l < kLanes; ++l) {
            if (w->state[l] != RUNNABLE) continue;
            w->cur = l;
            g_thread_idx = {(unsigned)l, 0, 0};
            swapcontext(&w->main_ctx, &w->lane_ctx[l]);  // until the lane waits at a collective or returns
        }
        unsigned int partial = 0, full = 0, done = 0;
        for (int l = 0; l < kLanes; ++l) {
            if (w->state[l] == DONE) done |= 1u << l;
            else if (w->wait_full[l]) full |= 1u << l;
            else partial |= 1u << l;
        }
        if (done == 0xffffffffu) break;
        if (partial) {
            release(w, partial);
        } else if (full == 0xffffffffu) {
            release(w, full);
        } else {
            fprintf(stderr, "emu: full-mask collective reached by lanes %08x only, %08x have returned (block %u)\n", full, done, block_idx);
            w->failed = true;
            break;
        }
    }
    return !w->failed;
}

// The calling lane contributes `value` and waits; returns the snapshot of the completed collective.
inline const Snapshot& collective(unsigned long long value, bool full_mask) {
    Warp* w = g_warp;
    const int me = w->cur;
    w->operand[me] = value;
    w->wait_full[me] = full_mask;
    w->state[me] = WAITING;
    swapcontext(&w->lane_ctx[me], &w->main_ctx);
    return w->ring[w->snapshot_of[me]];
}
inline int lane_id() { return g_warp->cur; }
}  // namespace emu

#define threadIdx (emu::g_thread_idx)
#define blockIdx (emu::g_block_idx)
#define blockDim (emu::g_block_dim)
#define gridDim (emu::g_grid_dim)

// ---- warp collectives -------------------------------------------------------------------------------------------
inline void __syncthreads() { emu::collective(0, true); }
// The lanes that have not returned.  Which of them are converged with the caller is the scheduler's choice on hardware;
// here the lane's NEXT collective is treated as a partial-mask one whatever the value says (see run_warp).
inline unsigned int __activemask() {
    unsigned int m = 0;
    for (int l = 0; l < emu::kLanes; ++l) m |= emu::g_warp->state[l] == emu::DONE ? 0u : 1u << l;
    emu::g_warp->partial_hint[emu::g_warp->cur] = true;
    return m;
}
inline bool emu_full(unsigned int mask) {
    bool& hint = emu::g_warp->partial_hint[emu::g_warp->cur];
    const bool full = mask == 0xffffffffu && !hint;
    hint = false;
    return full;
}
inline void __syncwarp(unsigned int mask = 0xffffffffu) { emu::collective(0, emu_full(mask)); }
inline unsigned int __ballot_sync(unsigned int mask, int pred) {
    const emu::Snapshot& s = emu::collective(pred ? 1 : 0, emu_full(mask));
    unsigned int r = 0;
    for (int l = 0; l < emu::kLanes; ++l)
        if ((s.present >> l & 1u) && s.operand[l]) r |= 1u << l;
    return r & mask;
}
inline int __any_sync(unsigned int mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned int mask, int pred) {
    const emu::Snapshot& s = emu::collective(pred ? 1 : 0, emu_full(mask));
    for (int l = 0; l < emu::kLanes; ++l)
        if ((mask >> l & 1u) && (s.present >> l & 1u) && !s.operand[l]) return 0;
    return 1;
}
template <class T> inline unsigned long long emu_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle operand");
    unsigned long long b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T> inline T emu_unbits(unsigned long long b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}
template <class T> inline T emu_pick(const emu::Snapshot& s, int l, T own) {
    return (l >= 0 && l < emu::kLanes && (s.present >> l & 1u)) ? emu_unbits<T>(s.operand[l]) : own;
}
template <class T> inline T __shfl_sync(unsigned int mask, T v, int src, int width = 32) {
    const emu::Snapshot& s = emu::collective(emu_bits(v), emu_full(mask));
    return emu_pick(s, (emu::lane_id() & ~(width - 1)) | (src & (width - 1)), v);
}
template <class T> inline T __shfl_xor_sync(unsigned int mask, T v, int lane_mask, int = 32) {
    const emu::Snapshot& s = emu::collective(emu_bits(v), emu_full(mask));
    return emu_pick(s, emu::lane_id() ^ lane_mask, v);
}
template <class T> inline T __shfl_down_sync(unsigned int mask, T v, unsigned int delta, int = 32) {
    const emu::Snapshot& s = emu::collective(emu_bits(v), emu_full(mask));
    return emu_pick(s, emu::lane_id() + (int)delta, v);
}
template <class T> inline T __shfl_up_sync(unsigned int mask, T v, unsigned int delta, int = 32) {
    const emu::Snapshot& s = emu::collective(emu_bits(v), emu_full(mask));
    return emu_pick(s, emu::lane_id() - (int)delta, v);
}
inline int __reduce_min_sync(unsigned int mask, int v) {
    const emu::Snapshot& s = emu::collective((unsigned long long)(long long)v, emu_full(mask));
    int r = v;
    for (int l = 0; l < emu::kLanes; ++l)
        if ((mask >> l & 1u) && (s.present >> l & 1u)) r = std::min(r, (int)(long long)s.operand[l]);
    return r;
}
inline int __reduce_max_sync(unsigned int mask, int v) {
    const emu::Snapshot& s = emu::collective((unsigned long long)(long long)v, emu_full(mask));
    int r = v;
    for (int l = 0; l < emu::kLanes; ++l)
        if ((mask >> l & 1u) && (s.present >> l & 1u)) r = std::max(r, (int)(long long)s.operand[l]);
    return r;
}
inline unsigned int __reduce_min_sync(unsigned int mask, unsigned int v) {
    const emu::Snapshot& s = emu::collective(v, emu_full(mask));
    unsigned int r = v;
    for (int l = 0; l < emu::kLanes; ++l)
        if ((mask >> l & 1u) && (s.present >> l & 1u)) r = std::min(r, (unsigned int)s.operand[l]);
    return r;
}
inline unsigned int __reduce_add_sync(unsigned int mask, unsigned int v) {
    const emu::Snapshot& s = emu::collective(v, emu_full(mask));
    unsigned int r = 0;
    for (int l = 0; l < emu::kLanes; ++l)
        if ((mask >> l & 1u) && (s.present >> l & 1u)) r += (unsigned int)s.operand[l];
    return r;
}

// ---- atomics (warps of one CTA never overlap; CTAs run on different OS threads) -----------------------------------
inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
inline unsigned int atomicAdd(unsigned int* a, unsigned int v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v) {
    unsigned long long old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(a, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline long long clock64() { return 0; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned int x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned int min(unsigned int a, unsigned int b) { return a < b ? a : b; }
inline unsigned int max(unsigned int a, unsigned int b) { return a > b ? a : b; }

// ---- scalar intrinsics ------------------------------------------------------------------------------------------
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline unsigned int __float_as_uint(float f) { unsigned int i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(unsigned int i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float2int_rn(float x) { return (int)nearbyintf(x); }  // default rounding mode: nearest even
inline int __float2int_rd(float x) { return (int)floorf(x); }
inline int __float2int_rz(float x) { return (int)x; }
inline float __int2float_rn(int x) { return (float)x; }
inline float emu_round_dir(double exact, int dir) {  // dir > 0: towards +inf, dir < 0: towards -inf
    float r = (float)exact;
    if (dir > 0 && (double)r < exact) r = nextafterf(r, INFINITY);
    if (dir < 0 && (double)r > exact) r = nextafterf(r, -INFINITY);
    return r;
}
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __fmaf_ru(float a, float b, float c) { return emu_round_dir((double)a * b + c, +1); }
inline float __fmaf_rd(float a, float b, float c) { return emu_round_dir((double)a * b + c, -1); }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fadd_ru(float a, float b) { return emu_round_dir((double)a + b, +1); }
inline float __fadd_rd(float a, float b) { return emu_round_dir((double)a + b, -1); }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fsub_ru(float a, float b) { return emu_round_dir((double)a - b, +1); }
inline float __fsub_rd(float a, float b) { return emu_round_dir((double)a - b, -1); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fmul_ru(float a, float b) { return emu_round_dir((double)a * b, +1); }
inline float __fmul_rd(float a, float b) { return emu_round_dir((double)a * b, -1); }
#ifdef LINON_EMU_FDIVIDEF_ULPS
// The device's __fdividef is approximate (a MUFU.RCP and a multiply: up to 2 ulp off) where the host divides exactly, so an
// emulated frame could hide a dependence on the quotient's last bits.  This variant moves every quotient by a pseudo-random
// -N ... +N ulps: results that are exact by construction (tests/test_emulation_cpu.py) must not notice.
inline float __fdividef(float a, float b) {
    float q = a / b;
    uint32_t ua, ub, uq;
    memcpy(&ua, &a, 4); memcpy(&ub, &b, 4); memcpy(&uq, &q, 4);
    if ((uq & 0x7f800000u) == 0x7f800000u || (uq & 0x7fffffffu) < 0x00800000u + 16u) return q;  // inf / NaN / zero / subnormal
    uint32_t h = ua * 2654435761u ^ (ub + 0x9e3779b9u) * 2246822519u;
    h ^= h >> 15; h *= 3266489917u; h ^= h >> 13;
    const int n = (int)(h % (2u * LINON_EMU_FDIVIDEF_ULPS + 1u)) - (int)LINON_EMU_FDIVIDEF_ULPS;
    uq += (uint32_t)n;  // magnitude moves by n ulps (sign bit untouched: no borrow at these magnitudes)
    memcpy(&q, &uq, 4);
    return q;
}
#else
inline float __fdividef(float a, float b) { return a / b; }
#endif
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline float __frsqrt_rn(float a) { return 1.0f / sqrtf(a); }
inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
inline float __expf(float a) { return expf(a); }
inline float __exp2f(float a) { return exp2f(a); }
inline float __sinf(float a) { return sinf(a); }
inline float __cosf(float a) { return cosf(a); }
inline float __powf(float a, float b) { return powf(a, b); }
inline float __logf(float a) { return logf(a); }
inline float __log2f(float a) { return log2f(a); }
inline float __saturatef(float a) { return fminf(fmaxf(a, 0.0f), 1.0f); }
