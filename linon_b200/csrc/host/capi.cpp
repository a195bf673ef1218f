// capi.cpp — implementation of include/linon_b200.h: the device-facing half of the reference's MainView
// (src/main_view.rs:53-849) on CUDA.  wgpu textures/buffers become linear device buffers, the WGSL splice +
// pipeline rebuild (:649-690, :852-861) becomes WGSL->CUDA translation + NVRTC JIT for sm_100a with a variant
// cache, the two compute dispatches of MainView::render (:831-849) become the march kernel (+ sample rays on a
// second stream) and the overlay kernel.
//
// There is no CPU fallback: without libcuda.so.1 and a device every entry point that needs one fails with
// LINON_ERR_NO_DRIVER.

#include "../../../include/linon_b200.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../device/static_kernels.h"
#include "functions.hpp"
#include "mesh_tables.hpp"
#include "mirage_match.hpp"
#include "scene.hpp"
#include "wgsl_translate.h"

namespace {

#include "embedded_sources.inc"  // kPreludeSrc, kMarchSrc (generated from device/*.cuh by build.py)

std::string g_create_error;

std::string fmt(const char* f, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    return buf;
}

// ---- CUDA driver API, loaded at run time so the library itself loads on a machine without a driver ---------
struct Driver {
    void* lib = nullptr;
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream,
                             void**, void**) = nullptr;
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*OccupancyMaxActiveBlocksPerMultiprocessor)(int*, CUfunction, int, size_t) = nullptr;
    // virtual memory management: exportable allocations for the display hand-off (linon_export_buffer_fd)
    CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    // TMA descriptor of the endpoint image for the tile overlay kernel (optional: without it the direct kernel runs)
    CUresult (*TensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    bool has_vmm = false;
    std::string error;

    bool load() {
        if (lib) return true;
        lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) {
            error = "libcuda.so.1 not found: a CUDA driver and a B200 are required (no CPU fallback)";
            return false;
        }
        auto sym = [&](const char* n) { return dlsym(lib, n); };
        ModuleLoadData = (decltype(ModuleLoadData))sym("cuModuleLoadData");
        ModuleUnload = (decltype(ModuleUnload))sym("cuModuleUnload");
        ModuleGetFunction = (decltype(ModuleGetFunction))sym("cuModuleGetFunction");
        LaunchKernel = (decltype(LaunchKernel))sym("cuLaunchKernel");
        GetErrorString = (decltype(GetErrorString))sym("cuGetErrorString");
        FuncGetAttribute = (decltype(FuncGetAttribute))sym("cuFuncGetAttribute");
        OccupancyMaxActiveBlocksPerMultiprocessor =
            (decltype(OccupancyMaxActiveBlocksPerMultiprocessor))sym("cuOccupancyMaxActiveBlocksPerMultiprocessor");
        MemGetAllocationGranularity = (decltype(MemGetAllocationGranularity))sym("cuMemGetAllocationGranularity");
        MemCreate = (decltype(MemCreate))sym("cuMemCreate");
        MemRelease = (decltype(MemRelease))sym("cuMemRelease");
        MemAddressReserve = (decltype(MemAddressReserve))sym("cuMemAddressReserve");
        MemAddressFree = (decltype(MemAddressFree))sym("cuMemAddressFree");
        MemMap = (decltype(MemMap))sym("cuMemMap");
        MemUnmap = (decltype(MemUnmap))sym("cuMemUnmap");
        MemSetAccess = (decltype(MemSetAccess))sym("cuMemSetAccess");
        MemExportToShareableHandle = (decltype(MemExportToShareableHandle))sym("cuMemExportToShareableHandle");
        MemImportFromShareableHandle = (decltype(MemImportFromShareableHandle))sym("cuMemImportFromShareableHandle");
        TensorMapEncodeTiled = (decltype(TensorMapEncodeTiled))sym("cuTensorMapEncodeTiled");
        has_vmm = MemGetAllocationGranularity && MemCreate && MemRelease && MemAddressReserve && MemAddressFree && MemMap &&
                  MemUnmap && MemSetAccess && MemExportToShareableHandle && MemImportFromShareableHandle;
        if (!ModuleLoadData || !ModuleUnload || !ModuleGetFunction || !LaunchKernel || !GetErrorString ||
            !FuncGetAttribute || !OccupancyMaxActiveBlocksPerMultiprocessor) {
            error = "libcuda.so.1 lacks a required entry point";
            dlclose(lib);
            lib = nullptr;
            return false;
        }
        return true;
    }
    std::string err_str(CUresult r) {
        const char* s = nullptr;
        if (GetErrorString) GetErrorString(r, &s);
        return s ? s : "unknown driver error";
    }
};
Driver g_drv;
std::mutex g_drv_mutex;

// ---- NVRTC, loaded privately by absolute path --------------------------------------------------------------------------
// A process that imported torch first already holds torch's own libnvrtc.so.12 (12.8 in this image) under the same soname,
// and a library linked against "libnvrtc.so.12" would silently bind to that copy: the kernel a benchmark times would then be
// compiled by another compiler than the one the tests and the ncu captures saw (found in round 2: two cubin hashes for one
// variant).  The toolkit's NVRTC is therefore opened by the path recorded at build time (LINON_NVRTC_PATH; RTLD_LOCAL, its
// own copy whatever else is loaded), LINON_NVRTC_LIB overrides it, the bare soname is the last resort.
#ifndef LINON_NVRTC_PATH
#define LINON_NVRTC_PATH ""
#endif
struct Nvrtc {
    void* lib = nullptr;
    std::string path, error;
    int major = 0, minor = 0;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    nvrtcResult (*Version)(int*, int*) = nullptr;

    bool load() {
        if (lib) return true;
        const char* env = getenv("LINON_NVRTC_LIB");
        const char* candidates[] = {env ? env : "", LINON_NVRTC_PATH, "libnvrtc.so.12"};
        for (const char* c : candidates) {
            if (!c[0]) continue;
            lib = dlopen(c, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) {
            const char* why = dlerror();
            error = std::string("NVRTC not found (tried LINON_NVRTC_LIB, " LINON_NVRTC_PATH ", libnvrtc.so.12): ") + (why ? why : "");
            return false;
        }
        auto sym = [&](const char* n) { return dlsym(lib, n); };
        CreateProgram = (decltype(CreateProgram))sym("nvrtcCreateProgram");
        DestroyProgram = (decltype(DestroyProgram))sym("nvrtcDestroyProgram");
        CompileProgram = (decltype(CompileProgram))sym("nvrtcCompileProgram");
        GetProgramLogSize = (decltype(GetProgramLogSize))sym("nvrtcGetProgramLogSize");
        GetProgramLog = (decltype(GetProgramLog))sym("nvrtcGetProgramLog");
        GetCUBINSize = (decltype(GetCUBINSize))sym("nvrtcGetCUBINSize");
        GetCUBIN = (decltype(GetCUBIN))sym("nvrtcGetCUBIN");
        GetErrorString = (decltype(GetErrorString))sym("nvrtcGetErrorString");
        Version = (decltype(Version))sym("nvrtcVersion");
        if (!CreateProgram || !DestroyProgram || !CompileProgram || !GetProgramLogSize || !GetProgramLog || !GetCUBINSize ||
            !GetCUBIN || !GetErrorString || !Version) {
            error = "the NVRTC library lacks a required entry point";
            dlclose(lib);
            lib = nullptr;
            return false;
        }
        Version(&major, &minor);
        Dl_info info;
        if (dladdr((void*)CreateProgram, &info) && info.dli_fname) path = info.dli_fname;
        return true;
    }
};
Nvrtc g_nvrtc;
std::mutex g_nvrtc_mutex;

// ---- device-side parameter blocks (must match device/march_kernels.cuh) -------------------------------------
struct MarchParamsH {
    float cam_origin[4], cam_view[4], cam_up[4];
    linon_settings settings;
    int width, height, local_rows, shard_rank, shard_world, rows_per_block;
    const void* tris;
    const int* group_range;
    void* raycast;
    void* endpoint;
    int* hit_id;
    unsigned long long* counters;
    unsigned int* work;
    int* q_pixel;
    void* q_state;
    unsigned int q_capacity;
    unsigned int n_ray_slots;
    unsigned int n_pair_slots;
    float plane_w_max;
    const void* bvh;
    int n_fixed_steps;
};
struct SampleParamsH {
    MarchParamsH march;
    const float* pos16;
    void* ray_samples;
};
static_assert(sizeof(linon_camera) == 112, "CameraUniform is 112 bytes (src/main_view.rs:12-19)");
static_assert(sizeof(linon_settings) == 24, "Settings is 24 bytes (src/main_view.rs:43-51)");

struct Variant {
    float h_initial;
    int adaptive, n_faces, n_groups, cull, math, block, unroll, refill_min, max_regs, profile;
    int n_segs;           // runs of consecutive plane groups that share a bound (intersect_mesh skips a run whose box is out of reach)
    int n_p1, n_boxes;    // first pass of a culling check: planes of the enclosing clusters + one bounding box per other cluster
    int bvh;              // 1: large mesh, faces + bounding-volume hierarchy in global memory (n_faces / n_groups are 0 then)
    int fw_one;           // 1: settings.field_weight == 1 (the application's default): cur_dir = v without the blend
    int specialize;       // 1: fast math may splice the hand-fused restatement of a predefined mirage body
    int pairs;            // 1: also build linon_march_pairs (LinonField<f32x2>, two rays per thread)
    int pair_min_blocks;  // resident CTAs / SM requested for it (0 = kernel default)
    int quiet;            // 1: fixed-step loops take straight steps where a fused mirage body is the identity (LINON_QUIET_ZONE)
    int collapse;         // 1: quiet runs of at least LINON_COLLAPSE_MIN steps are taken binade by binade (advance_coord)
    int quiet_radius;     // sigmoid profiles: zone radius in permille of max_dist (0 = where the fp32 sigmoid saturates, 1.92:
                          // the default of mirage_fused.cuh; the library's default is 1500, LINON_OPT_QUIET_RADIUS)
};

struct Module {
    CUmodule mod = nullptr;
    CUfunction tiles = nullptr, pairs = nullptr, refill = nullptr, phase_a = nullptr, phase_b = nullptr, sample = nullptr,
               eval_field = nullptr, eval_helper = nullptr, eval_hit = nullptr, eval_rk4 = nullptr;
    int regs_tiles = 0, regs_pairs = 0, regs_refill = 0, regs_phase_a = 0, regs_phase_b = 0;
    uint64_t last_use = 0;  // linon_view::module_clock at the last get_module
    uint64_t cubin_hash = 0;  // FNV-1a of the cubin's code: ties a committed ncu capture to the kernel that is running
    uint64_t tiles_hash = 0;  // ... of the tile march kernel alone (what linon_march_cubin_hash reports when that kernel ran)
};

// The translated body becomes LinonField<S>::eval.  The member typedefs re-declare the type names the translator
// emits (vec3f, wg_f32, ...) for the scalar type of the instantiation: S = float is the reference's one ray per
// invocation, S = f32x2 carries two rays through the same text (linon_march_pairs).
const char kFieldHead[] =
    "template <class S> struct LinonField {\n"
    "    typedef S wg_f32;\n"
    "    typedef vec2t<S> vec2f; typedef vec3t<S> vec3f; typedef vec4t<S> vec4f;\n"
    "    typedef mat2x2t<S> mat2x2f; typedef mat3x3t<S> mat3x3f;\n"
    "    static __device__ __forceinline__ vec3t<S> eval(const vec3t<S> p_prev, const vec3t<S> p, const vec3t<S> v0,\n"
    "                                                    const vec3t<S> v, const S t, const WgslSettings& settings,\n"
    "                                                    const WgslCamera& camera) {\n"
    "#line 1 \"field_function.wgsl\"\n";
const char kFieldTail[] = "\n    }\n};\n#line 1 \"march_kernels.cuh\"\n";

// The mirage family, recognised structurally (mirage_match.hpp): kind 1..4 and the constants of the user's text, so that a
// body with edited constants (src/main_view.rs:852-861: the body is editable text) still runs the hand-fused restatement.
linon::MirageParams mirage_params(const std::string& wgsl_body) { return linon::match_mirage(wgsl_body); }

std::string build_source(const std::string& cuda_body, const linon::MirageParams& mirage, const linon::ModuleLayer& layer) {
    const int fused_kind = mirage.kind;
    std::string s;
    s.reserve(sizeof(kPreludeSrc) + sizeof(kMarchSrc) + sizeof(kMirageFusedSrc) + cuda_body.size() + layer.members.size() + 512);
    s += layer.march_defines;  // march-level constants of an edited main_view.wgsl (macros of march_kernels.cuh)
    s += "#line 1 \"wgsl_prelude.cuh\"\n";
    s += kPreludeSrc;
    s += "\n";
    if (fused_kind > 0) {  // hand-fused restatement of a predefined mirage body (fast math only)
        s += linon::mirage_defines(mirage) + "#line 1 \"mirage_fused.cuh\"\n";
        s += kMirageFusedSrc;
        s += "\n#line 1 \"march_kernels.cuh\"\n";
    } else {
        // helper layer of an edited main_view.wgsl (linon_reload_shader_source): member functions of LinonField<S>, found by
        // class-scope lookup before the prelude's helpers of the same name; module constants open every function body
        const size_t cut = std::string(kFieldHead).find("    static __device__ __forceinline__ vec3t<S> eval(");
        const std::string head(kFieldHead);
        const size_t line_directive = head.rfind("#line 1 \"field_function.wgsl\"");
        s += head.substr(0, cut);
        s += layer.members;
        s += head.substr(cut, line_directive - cut);
        s += layer.preamble;
        s += "\n";
        s += head.substr(line_directive);
        s += cuda_body;
        const std::string tail(kFieldTail);
        const size_t tail_line = tail.find("#line");
        s += tail.substr(0, tail_line);
        // every helper of the module is compiled, called or not, like naga validates the whole module (an error in an
        // unused helper is still an error of the reload)
        if (!layer.members.empty()) s += "template struct LinonField<float>;\n";
        s += tail.substr(tail_line);
    }
    s += kMarchSrc;
    return s;
}

uint64_t fnv1a64(const char* p, size_t n, uint64_t h = 1469598103934665603ull) {
    for (size_t i = 0; i < n; ++i) h = (h ^ (unsigned char)p[i]) * 1099511628211ull;
    return h;
}

// Identity of a module's MACHINE CODE: FNV-1a over the names and contents of the .text.* sections of the cubin (an ELF64
// image), in section order.  The -lineinfo tables are left out on purpose: an edit that only moves lines in a source file
// changes them, not the kernel - and a committed ncu capture is evidence about the kernel.
// only_kernel: hash the code of that one kernel (section ".text.<name>"), so that a capture of the tile kernel stays valid when
// another kernel of the module (sample rays, the opt-in march kernels) changes; nullptr: every .text section.
uint64_t cubin_code_hash(const std::vector<char>& cubin, const char* only_kernel = nullptr) {
    const size_t n = cubin.size();
    const unsigned char* b = (const unsigned char*)cubin.data();
    auto rd = [&](size_t off, int bytes) -> uint64_t {
        uint64_t v = 0;
        if (off + bytes > n) return 0;
        for (int i = 0; i < bytes; ++i) v |= (uint64_t)b[off + i] << (8 * i);
        return v;
    };
    if (n < 64 || memcmp(b, "\x7f" "ELF", 4) != 0 || b[4] != 2) return fnv1a64(cubin.data(), n);  // not ELF64: hash everything
    const uint64_t shoff = rd(0x28, 8);
    const uint64_t shentsize = rd(0x3a, 2), shnum = rd(0x3c, 2), shstrndx = rd(0x3e, 2);
    if (!shoff || shentsize < 64 || shstrndx >= shnum || shoff + shnum * shentsize > n) return fnv1a64(cubin.data(), n);
    const uint64_t str_off = rd(shoff + shstrndx * shentsize + 0x18, 8), str_size = rd(shoff + shstrndx * shentsize + 0x20, 8);
    uint64_t h = 1469598103934665603ull;
    for (uint64_t k = 0; k < shnum; ++k) {
        const uint64_t sh = shoff + k * shentsize;
        const uint64_t name = rd(sh, 4), off = rd(sh + 0x18, 8), size = rd(sh + 0x20, 8);
        if (name >= str_size || str_off + name >= n) continue;
        const char* nm = cubin.data() + str_off + name;
        const size_t nlen = strnlen(nm, n - (str_off + name));
        if (nlen < 5 || memcmp(nm, ".text", 5) != 0 || off + size > n) continue;
        if (only_kernel && (nlen != 6 + strlen(only_kernel) || memcmp(nm + 6, only_kernel, nlen - 6) != 0)) continue;
        h = fnv1a64(nm, nlen, h);
        h = fnv1a64(cubin.data() + off, size, h);
    }
    return h;
}

std::string float_literal(float x) {
    char b[64];
    snprintf(b, sizeof(b), "%.9gf", (double)x);
    std::string s = b;
    if (s.find('.') == std::string::npos && s.find('e') == std::string::npos) s.insert(s.size() - 1, ".0");
    return s;
}

// Translate + NVRTC compile to an sm_100a cubin.  No device needed.
bool compile_cubin(const std::string& wgsl_body, const Variant& var, const linon::ModuleLayer& layer, std::vector<char>& cubin,
                   std::string& log) {
    std::string cuda_body, terr;
    if (!linon::translate_wgsl_body(wgsl_body, layer.names, cuda_body, terr)) {
        log = terr;
        return false;
    }
    // an edited helper layer may redefine what the predefined bodies call: they then run the generic translation
    const linon::MirageParams mirage = (var.math == 2 && var.specialize && layer.empty()) ? mirage_params(wgsl_body) : linon::MirageParams();
    const std::string src = build_source(cuda_body, mirage, layer);
    {
        std::lock_guard<std::mutex> lock(g_nvrtc_mutex);
        if (!g_nvrtc.load()) {
            log = g_nvrtc.error;
            return false;
        }
    }
    nvrtcProgram prog = nullptr;
    if (g_nvrtc.CreateProgram(&prog, src.c_str(), "linon_march.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
        log = "nvrtcCreateProgram failed";
        return false;
    }
    std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo",
                                     "-DLINON_H_INITIAL=" + float_literal(var.h_initial),
                                     "-DLINON_ADAPTIVE=" + std::to_string(var.adaptive),
                                     "-DLINON_N_FACES=" + std::to_string(var.n_faces),
                                     "-DLINON_N_GROUPS=" + std::to_string(var.n_groups),
                                     "-DLINON_N_P1=" + std::to_string(var.n_p1),
                                     "-DLINON_N_BOXES=" + std::to_string(var.n_boxes),
                                     "-DLINON_N_SEGS=" + std::to_string(var.n_segs),
                                     "-DLINON_CULL=" + std::to_string(var.cull),
                                     "-DLINON_FACE_UNROLL=" + std::to_string(var.unroll),
                                     "-DLINON_BLOCK=" + std::to_string(var.block),
                                     "-DLINON_REFILL_MIN=" + std::to_string(var.refill_min),
                                     "-DLINON_PROFILE=" + std::to_string(var.profile),
                                     "-DLINON_PAIRS=" + std::to_string(var.pairs),
                                     "-DLINON_FW_ONE=" + std::to_string(var.fw_one),
                                     "-DLINON_BVH=" + std::to_string(var.bvh),
                                     "-DLINON_QUIET_ZONE=" + std::to_string(var.quiet),
                                     "-DLINON_COLLAPSE_RUNS=" + std::to_string(var.collapse),
                                     // strict: ptxas would contract mul.rn.f32x2 + add.rn.f32x2 whatever --fmad says
                                     "-DLINON_PACKED_OPS=" + std::to_string((var.math == 0 || getenv("LINON_DEV_UNPACKED")) ? 0 : 1),
                                     "-DLINON_STRICT_MATH=" + std::to_string(var.math == 0 ? 1 : 0)};
    if (var.pair_min_blocks > 0) opts.push_back("-DLINON_PAIR_MIN_BLOCKS=" + std::to_string(var.pair_min_blocks));
    if (var.quiet_radius > 0) opts.push_back("-DLINON_QUIET_RADIUS_SIGMOID=" + float_literal(0.001f * (float)var.quiet_radius));
    if (var.math == 0) {
        opts.insert(opts.end(), {"--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false"});
    } else if (var.math == 1) {
        opts.insert(opts.end(), {"--fmad=true", "--prec-div=true", "--prec-sqrt=true", "--ftz=false"});
    } else {
        // every option --use_fast_math implies is ALSO passed explicitly: NVRTC 12.9 keeps --prec-div / --prec-sqrt / --ftz of an
        // earlier compilation of the same process when a later one only says --use_fast_math, so a fast-math kernel compiled
        // after a strict or fused one came out as different machine code (found in round 2: a long test session compiled
        // production kernels that a fresh process never produces; tests/test_capi_cpu.py pins the order independence)
        opts.insert(opts.end(), {"--use_fast_math", "--fmad=true", "--ftz=true", "--prec-div=false", "--prec-sqrt=false"});
        opts.push_back("-DLINON_FAST_MATH=1");
    }
    if (const char* dev = getenv("LINON_DEV_DEFINES")) {  // developer aid (parameter sweeps): extra -D options, space separated;
        std::string d = dev;                               // NOT part of the variant key - one process per value
        size_t pos = 0;
        while (pos < d.size()) {
            size_t e = d.find(' ', pos);
            if (e == std::string::npos) e = d.size();
            if (e > pos && d.compare(pos, 2, "-D") == 0) opts.push_back(d.substr(pos, e - pos));
            pos = e + 1;
        }
    }
    if (var.max_regs > 0) {
        // __launch_bounds__ wins over --maxrregcount, so the cap is passed as resident CTAs per SM as well
        opts.push_back("--maxrregcount=" + std::to_string(var.max_regs));
        const int ctas = 65536 / (var.max_regs * var.block);
        if (ctas >= 1) opts.push_back("-DLINON_MIN_BLOCKS=" + std::to_string(ctas));
    }
    std::vector<const char*> copts;
    for (auto& o : opts) copts.push_back(o.c_str());
    if (const char* dump = getenv("LINON_DUMP_SOURCE")) {  // developer aid: the exact module text and options NVRTC gets
        if (FILE* f = fopen(dump, "wb")) {                  // (tests/emu compiles this text for the host)
            fputs("// nvrtc:", f);
            for (auto& o : opts) fprintf(f, " %s", o.c_str());
            fputs("\n", f);
            fwrite(src.data(), 1, src.size(), f);
            fclose(f);
        }
        if (getenv("LINON_DUMP_SOURCE_ONLY")) {  // the caller only wants the text (tests/emu): skip the NVRTC compile
            g_nvrtc.DestroyProgram(&prog);
            log = "LINON_DUMP_SOURCE_ONLY: source written, not compiled";
            return false;
        }
    }
    const nvrtcResult rc = g_nvrtc.CompileProgram(prog, (int)copts.size(), copts.data());
    size_t log_size = 0;
    g_nvrtc.GetProgramLogSize(prog, &log_size);
    std::string full_log(log_size, '\0');
    if (log_size > 1) g_nvrtc.GetProgramLog(prog, &full_log[0]);
    if (rc != NVRTC_SUCCESS) {
        // keep the diagnostics that point into the user's body first; fall back to the whole log
        std::string mine;
        size_t pos = 0;
        while (pos < full_log.size()) {
            size_t e = full_log.find('\n', pos);
            if (e == std::string::npos) e = full_log.size();
            const std::string line = full_log.substr(pos, e - pos);
            if (line.find("field_function.wgsl") != std::string::npos || line.find("main_view.wgsl") != std::string::npos) mine += line + "\n";
            pos = e + 1;
        }
        // Only diagnostics that point into the user's own text go back across the ABI.  Anything else (the translated body
        // broke a line of the kernel template) is reported by its message alone, without the module's paths and code.
        if (mine.empty()) {
            const size_t e = full_log.find("error");
            size_t eol = e == std::string::npos ? std::string::npos : full_log.find('\n', e);
            mine = "the field function does not fit the kernel template (" +
                   (e == std::string::npos ? std::string(g_nvrtc.GetErrorString(rc)) : full_log.substr(e, eol == std::string::npos ? std::string::npos : eol - e)) + ")";
        }
        log = mine;
        g_nvrtc.DestroyProgram(&prog);
        return false;
    }
    size_t n = 0;
    g_nvrtc.GetCUBINSize(prog, &n);
    cubin.resize(n);
    g_nvrtc.GetCUBIN(prog, cubin.data());
    g_nvrtc.DestroyProgram(&prog);
    log = full_log;
    return true;
}

std::string variant_key(const std::string& body, const Variant& v, const linon::ModuleLayer& layer) {
    return layer.members + "\x01" + layer.preamble + "\x01" + layer.march_defines + "\x01" + fmt("%a|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|", (double)v.h_initial, v.adaptive, v.n_faces, v.n_groups, v.n_p1, v.n_boxes, v.n_segs, v.cull, v.math,
               v.block, v.unroll, v.refill_min, v.max_regs, v.profile, v.pairs, v.pair_min_blocks, v.specialize, v.fw_one, v.bvh, v.quiet, v.quiet_radius, v.collapse) + body;
}

}  // namespace

// =============================================================================================================
struct linon_view {
    int device = 0;
    int sm_count = 0;
    uint32_t W = 0, H = 0;
    cudaStream_t stream = nullptr, aux = nullptr;
    cudaEvent_t ev[11] = {};  // 0,1 march  2,3 overlay  4,5 sample rays  6,7 outline  8 pos ready  9 aux done  10 phase A end
    bool ran_phases = false;
    bool ran_march = false, ran_overlay = false, ran_sample = false, ran_outline = false;

    linon_camera camera{};
    bool has_camera = false;
    linon_settings settings{};
    uint32_t n_faces = 0, n_groups = 0, n_p1 = 0, n_boxes = 0, n_segs = 0;
    float plane_w_max = 0.0f;  // max |offset| of the plane groups (BVH meshes: max |coordinate|)
    linon::ModuleLayer layer;  // helper layer of an edited main_view.wgsl (linon_reload_shader_source); empty = stock module
    bool use_bvh = false;      // mesh beyond the shared-memory stage: faces in hierarchy order + nodes in global memory
    void* d_bvh = nullptr;
    uint32_t bvh_nodes = 0, bvh_depth = 0;
    std::vector<float> host_verts;     // copy of the last mesh (re-staged when LINON_OPT_FORCE_BVH changes)
    std::vector<uint32_t> host_faces;
    void* d_tris = nullptr;   // n_faces x {v0, e1, e2} then n_groups x {unit normal, offset}
    int* d_group_range = nullptr;

    void *d_final = nullptr, *d_raycast = nullptr, *d_endpoint = nullptr;
    float* d_exponents = nullptr;
    int* d_hit = nullptr;
    void* d_ray_samples = nullptr;
    float* d_pos16 = nullptr;
    unsigned long long* d_counters = nullptr;
    unsigned long long* d_outline_keys = nullptr;
    unsigned int* d_work = nullptr;
    int* d_q_pixel = nullptr;
    void* d_q_state = nullptr;
    size_t q_capacity = 0;

    uint32_t shard_rank = 0, shard_world = 1, rows_per_block = 32;
    uint32_t local_rows = 0;

    std::string body;  // current WGSL body
    std::map<std::string, Module> modules;
    std::map<std::string, bool> pair_unfit;  // bodies whose LinonField<f32x2> instantiation failed to compile
    mutable std::map<std::string, bool> mirage_cache;  // body -> member of the mirage family (csrc/host/mirage_match.hpp)
    uint64_t module_clock = 0;  // use counter for the LRU eviction of `modules`

    int opt_math = 2, opt_kernel = 0, opt_pair_min_blocks = 0, opt_specialize = 1, opt_force_bvh = 0, opt_block = 256, opt_ctas_per_sm = 0, opt_refill_min = 1, opt_unroll = 4,
        opt_max_regs = 0, opt_profile = 0, opt_cull = 1, opt_quiet = 2, opt_quiet_radius = 1500, opt_overlay_kernel = 2, opt_collapse = 1;
    // tile overlay kernel: tensor map over d_endpoint as a (4 W, H) fp32 tensor, box = tile + halo of the current delta
    alignas(64) CUtensorMap endpoint_map;
    const void* endpoint_map_ptr = nullptr;
    uint32_t endpoint_map_w = 0, endpoint_map_h = 0;
    int endpoint_map_delta = 0, endpoint_map_rows = 0;  // rows: tile height the box was made for (overlay_tile_rows)
    bool endpoint_map_ok = false;

    linon_timings timings{};
    uint64_t last_rays = 0;
    uint64_t last_cubin_hash = 0;  // of the module whose march kernel ran last
    std::string error;

    // peer stores (linon_redirect_output): where the march stores colour / endpoint / hit id instead of this view's own
    // buffers - the frame buffers of the root shard, same process or mapped from another process (linon_ipc_open)
    uchar4* out_raycast = nullptr;
    float4* out_endpoint = nullptr;
    int* out_hit = nullptr;
    std::vector<void*> ipc_mappings;  // cudaIpcOpenMemHandle results, closed by linon_destroy
    // display hand-off (linon_export_buffer_fd): buffers that live in an exportable allocation instead of cudaMalloc memory
    struct VmmAlloc {
        CUmemGenericAllocationHandle handle = 0;
        CUdeviceptr ptr = 0;
        size_t size = 0;
    };
    std::map<int, VmmAlloc> exported;  // LINON_BUF_* -> the allocation that backs it now

    int fail(int code, const std::string& msg) {
        error = msg;
        return code;
    }
};

namespace {

#define CK(v, call)                                                                                     \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return (v)->fail(LINON_ERR_CUDA, fmt("%s failed: %s", #call, cudaGetErrorString(e__)));     \
    } while (0)

// Pixel indices are 32-bit ints in the kernels (texture_row * width + x, ray slots) and the outline search packs x and y into
// 21-bit key fields (csrc/device/outline_keys.cuh): larger frames are refused instead of overflowing (cudaMalloc would succeed
// on a 180 GB part).
bool frame_size_ok(uint32_t W, uint32_t H) {
    return W >= 2 && H >= 2 && W <= 65536u && H <= 65536u && (uint64_t)W * H < (1ull << 31);
}

uint32_t count_local_rows(uint32_t H, uint32_t rank, uint32_t world, uint32_t rpb) {
    uint32_t n = 0;
    for (uint32_t b = 0, y = 0; y < H; ++b, y += rpb)
        if (b % world == rank) n += (y + rpb <= H) ? rpb : (H - y);
    return n;
}

// frees a buffer whichever way it was allocated (cudaMalloc, or an exportable allocation mapped by linon_export_buffer_fd)
void release_buffer(linon_view* v, int buffer, void* ptr) {
    auto it = v->exported.find(buffer);
    if (it == v->exported.end()) {
        cudaFree(ptr);
        return;
    }
    g_drv.MemUnmap(it->second.ptr, it->second.size);
    g_drv.MemRelease(it->second.handle);
    g_drv.MemAddressFree(it->second.ptr, it->second.size);
    v->exported.erase(it);
}

void free_frame_buffers(linon_view* v) {
    release_buffer(v, LINON_BUF_FINAL_RGBA8, v->d_final);
    v->d_final = nullptr;
    cudaFree(v->d_final); cudaFree(v->d_raycast); cudaFree(v->d_endpoint); cudaFree(v->d_exponents);
    cudaFree(v->d_hit); cudaFree(v->d_q_pixel); cudaFree(v->d_q_state);
    v->d_final = v->d_raycast = v->d_endpoint = nullptr;
    v->d_exponents = nullptr; v->d_hit = nullptr; v->d_q_pixel = nullptr; v->d_q_state = nullptr;
    v->q_capacity = 0;
}

// The post-hit queue of the persistent kernel (52 B per ray of the shard) is allocated on first use: the default tile
// kernel never touches it, and at 8192 x 8192 it would be 3.5 GB.
int alloc_queue(linon_view* v, bool needed) {
    v->local_rows = count_local_rows(v->H, v->shard_rank, v->shard_world, v->rows_per_block);
    const size_t cap = needed ? (size_t)v->local_rows * v->W : 0;
    if (cap == v->q_capacity && (v->d_q_pixel || cap == 0)) return LINON_OK;
    cudaFree(v->d_q_pixel); cudaFree(v->d_q_state);
    v->d_q_pixel = nullptr; v->d_q_state = nullptr; v->q_capacity = 0;
    if (cap == 0) return LINON_OK;
    CK(v, cudaMalloc(&v->d_q_pixel, cap * sizeof(int)));
    CK(v, cudaMalloc(&v->d_q_state, cap * 3 * 16));
    v->q_capacity = cap;
    return LINON_OK;
}

int alloc_frame_buffers(linon_view* v, uint32_t W, uint32_t H) {
    free_frame_buffers(v);
    v->W = W; v->H = H;
    const size_t n = (size_t)W * H;
    CK(v, cudaMalloc(&v->d_final, n * 4));
    CK(v, cudaMalloc(&v->d_raycast, n * 4));
    CK(v, cudaMalloc(&v->d_endpoint, n * 16));
    CK(v, cudaMalloc(&v->d_exponents, n * 4));
    CK(v, cudaMalloc(&v->d_hit, n * 4));
    CK(v, cudaMemsetAsync(v->d_final, 0, n * 4, v->stream));
    CK(v, cudaMemsetAsync(v->d_raycast, 0, n * 4, v->stream));
    CK(v, cudaMemsetAsync(v->d_endpoint, 0, n * 16, v->stream));
    CK(v, cudaMemsetAsync(v->d_exponents, 0, n * 4, v->stream));  // wgpu zero-initialises buffers
    CK(v, cudaMemsetAsync(v->d_hit, 0xFF, n * 4, v->stream));
    return alloc_queue(v, false);
}

int get_module(linon_view* v, const std::string& body, const Variant& var, Module** out, float* jit_ms) {
    const std::string key = variant_key(body, var, v->layer);
    auto it = v->modules.find(key);
    if (it != v->modules.end()) {
        it->second.last_use = ++v->module_clock;
        *out = &it->second;
        return LINON_OK;
    }
    // A long editing session compiles a module per distinct body / variant / helper layer: keep the most recently used ones
    // (an evicted variant is rebuilt on demand, ~1 s).  Kernels of an evicted module may still be in flight: wait first.
    constexpr size_t kMaxModules = 48;
    if (v->modules.size() >= kMaxModules) {
        cudaStreamSynchronize(v->stream);
        cudaStreamSynchronize(v->aux);
        while (v->modules.size() >= kMaxModules) {
            auto oldest = v->modules.begin();
            for (auto m = v->modules.begin(); m != v->modules.end(); ++m)
                if (m->second.last_use < oldest->second.last_use) oldest = m;
            g_drv.ModuleUnload(oldest->second.mod);
            v->modules.erase(oldest);
        }
    }
    if (v->pair_unfit.size() > 256) v->pair_unfit.clear();
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<char> cubin;
    std::string log;
    if (!compile_cubin(body, var, v->layer, cubin, log)) return v->fail(LINON_ERR_COMPILE, log);
    Module m;
    m.cubin_hash = cubin_code_hash(cubin);
    m.tiles_hash = cubin_code_hash(cubin, "linon_march_tiles");
    CUresult r = g_drv.ModuleLoadData(&m.mod, cubin.data());
    if (r != CUDA_SUCCESS) return v->fail(LINON_ERR_CUDA, "cuModuleLoadData: " + g_drv.err_str(r));
    struct { const char* name; CUfunction* f; } fns[] = {
        {"linon_march_tiles", &m.tiles},       {"linon_march_phase_a", &m.phase_a},
        {"linon_march_pairs", var.pairs ? &m.pairs : nullptr}, {"linon_march_refill", &m.refill},
        {"linon_march_phase_b", &m.phase_b},   {"linon_sample_rays", &m.sample},
        {"linon_eval_field", &m.eval_field},   {"linon_eval_helper", &m.eval_helper},
        {"linon_eval_hit_triangle", &m.eval_hit}, {"linon_eval_rk4", &m.eval_rk4}};
    for (auto& f : fns) {
        if (!f.f) continue;
        r = g_drv.ModuleGetFunction(f.f, m.mod, f.name);
        if (r != CUDA_SUCCESS) {
            g_drv.ModuleUnload(m.mod);
            return v->fail(LINON_ERR_CUDA, fmt("cuModuleGetFunction(%s): %s", f.name, g_drv.err_str(r).c_str()));
        }
    }
    g_drv.FuncGetAttribute(&m.regs_tiles, CU_FUNC_ATTRIBUTE_NUM_REGS, m.tiles);
    if (m.pairs) g_drv.FuncGetAttribute(&m.regs_pairs, CU_FUNC_ATTRIBUTE_NUM_REGS, m.pairs);
    g_drv.FuncGetAttribute(&m.regs_refill, CU_FUNC_ATTRIBUTE_NUM_REGS, m.refill);
    g_drv.FuncGetAttribute(&m.regs_phase_a, CU_FUNC_ATTRIBUTE_NUM_REGS, m.phase_a);
    g_drv.FuncGetAttribute(&m.regs_phase_b, CU_FUNC_ATTRIBUTE_NUM_REGS, m.phase_b);
    const auto t1 = std::chrono::steady_clock::now();
    if (jit_ms) *jit_ms += std::chrono::duration<float, std::milli>(t1 - t0).count();
    m.last_use = ++v->module_clock;
    auto ins = v->modules.emplace(key, m);
    *out = &ins.first->second;
    return LINON_OK;
}

// Table sizes of the start-up mesh (src/application.rs:105-123), for builds that happen before a mesh is bound
void default_mesh_counts(Variant& var) {
    static const linon::MeshTables t = [] {
        linon::MeshTables m;
        const linon::Mesh mesh = linon::default_scene(nullptr);
        linon::build_mesh_tables(mesh.vertices.data(), (uint32_t)(mesh.vertices.size() / 3), mesh.indices.data(),
                                 (uint32_t)(mesh.indices.size() / 3), m);
        return m;
    }();
    var.n_faces = (int)(t.tris.size() / 12);
    var.n_groups = (int)t.n_groups();
    var.n_p1 = (int)t.n_p1();
    var.n_boxes = (int)t.n_boxes();
    var.n_segs = (int)t.n_segments();
}

bool body_is_mirage(const linon_view* v) {
    auto it = v->mirage_cache.find(v->body);
    if (it != v->mirage_cache.end()) return it->second;
    if (v->mirage_cache.size() > 256) v->mirage_cache.clear();
    const bool m = linon::match_mirage(v->body).kind > 0;
    v->mirage_cache[v->body] = m;
    return m;
}

Variant make_variant(const linon_view* v, uint32_t flags) {
    Variant var;
    const linon::ModuleLayer& L = v->layer;
    // Enhance replaces the TEXT `let h_initial = 0.1;` (src/main_view.rs:785-787): with an edited literal the replacement no
    // longer matches and the frame keeps the module's own step size, like in the reference
    const float h_module = L.h_initial > 0.0 ? (float)L.h_initial : 0.1f;
    var.h_initial = ((flags & LINON_RENDER_ENHANCE) && !(L.h_initial > 0.0)) ? 0.001f : h_module;
    // adaptive sampling can only act while h_initial > 0.002 (src/main_view.wgsl:252,267); `const adaptive_sampling` of an
    // edited module (:227) is the default the ADAPTIVE_OFF flag overrides
    const float min_h = L.reject_min_h >= 0.0 ? (float)L.reject_min_h : 0.002f;
    var.adaptive = (!(flags & LINON_RENDER_ADAPTIVE_OFF) && L.adaptive_sampling != 0 && var.h_initial > min_h) ? 1 : 0;
    var.bvh = v->use_bvh ? 1 : 0;
    var.n_faces = v->use_bvh ? 1 : (int)v->n_faces;   // the hierarchy kernels do not depend on the face count
    var.n_groups = v->use_bvh ? 1 : (int)v->n_groups;
    var.n_p1 = v->use_bvh ? 1 : (int)v->n_p1;
    var.n_boxes = v->use_bvh ? 0 : (int)v->n_boxes;
    var.n_segs = v->use_bvh ? 1 : (int)v->n_segs;
    var.cull = v->opt_cull;
    var.math = v->opt_math;
    var.block = v->opt_block;
    var.unroll = v->opt_unroll;
    var.refill_min = v->opt_refill_min;
    var.max_regs = v->opt_max_regs;
    var.profile = v->opt_profile;
    var.specialize = v->opt_specialize;
    // production math modes only: strict mode keeps the blend as written (0 * dir + 1 * v differs from v in the sign of a zero)
    // cur_dir = v instead of (1 - 1) * cur_dir + 1 * v: only where the velocity is known to be finite, i.e. for the fused mirage
    // restatements (bounded by construction).  A generic body may overflow, and 0 * inf + v is NaN in the reference while the
    // shortcut would keep v (round-1 advisor finding): those keep the blend as written in every math mode.
    var.fw_one = (v->opt_math == 2 && v->opt_specialize && v->layer.empty() && v->settings.field_weight == 1.0f &&
                  body_is_mirage(v)) ? 1 : 0;
    var.pairs = (v->opt_kernel == 2 && !v->use_bvh && !v->pair_unfit.count(v->body)) ? 1 : 0;
    var.pair_min_blocks = var.pairs ? v->opt_pair_min_blocks : 0;
    // only the tile kernel has the quiet-zone path (runs of straight steps in its fixed-step loops, a per-attempt test in
    // the adaptive loop), and only fused mirage bodies define the zone
    // (1: fixed-step variants only, 2: adaptive variants as well)
    var.quiet = (v->opt_quiet > (var.adaptive ? 1 : 0) && v->opt_math == 2 && v->opt_specialize) ? 1 : 0;
    var.quiet_radius = var.quiet ? v->opt_quiet_radius : 0;
    var.collapse = var.quiet ? v->opt_collapse : 0;
    return var;
}

// get_module for the current body; when the two-rays-per-thread instantiation of a user body does not compile
// (data-dependent control flow, integer casts of per-ray values ...) the body is remembered as unfit and the module is
// built without the pair kernel - the caller then launches the tile kernel.
int get_module_for_body(linon_view* v, const std::string& body, Variant& var, Module** out, float* jit_ms) {
    int rc = get_module(v, body, var, out, jit_ms);
    if (rc == LINON_ERR_COMPILE && var.pairs) {
        const std::string pair_error = v->error;
        var.pairs = 0;
        var.pair_min_blocks = 0;
        rc = get_module(v, body, var, out, jit_ms);
        if (rc == LINON_OK) v->pair_unfit[body] = true;  // else: a genuine error of the body, reported by the scalar build
        (void)pair_error;
    }
    return rc;
}

// Iterations of `loop { if (t > 5.0) { break; } ...; t = t + h; }` (src/main_view.wgsl:244-266) when h never changes: the
// same fp32 recurrence, evaluated once on the host (g++ -ffp-contract=off, IEEE binary32 additions), so fixed-step
// kernels count to an integer instead of comparing a per-lane float.  5001 for h = 0.001, 51 for h = 0.1.
int fixed_step_count(float h, float t_max = 5.0f) {
    if (!(h > 0.0f)) return 0;
    volatile float t = 0.0f;
    int n = 0;
    while (t <= t_max && n < (1 << 24)) {
        t = t + h;
        ++n;
    }
    return n;
}

void fill_march_params(const linon_view* v, MarchParamsH& P) {
    memset(&P, 0, sizeof(P));
    memcpy(P.cam_origin, v->camera.origin, 16);
    memcpy(P.cam_view, v->camera.view_direction, 16);
    memcpy(P.cam_up, v->camera.up, 16);
    P.settings = v->settings;
    P.width = (int)v->W; P.height = (int)v->H;
    P.local_rows = (int)v->local_rows;
    P.shard_rank = (int)v->shard_rank; P.shard_world = (int)v->shard_world; P.rows_per_block = (int)v->rows_per_block;
    P.tris = v->d_tris;
    P.group_range = v->d_group_range;
    P.raycast = v->out_raycast ? v->out_raycast : v->d_raycast;
    P.endpoint = v->out_endpoint ? v->out_endpoint : v->d_endpoint;
    P.hit_id = v->out_hit ? v->out_hit : v->d_hit;
    P.counters = v->d_counters;
    P.work = v->d_work; P.q_pixel = v->d_q_pixel; P.q_state = v->d_q_state;
    P.q_capacity = (unsigned int)v->q_capacity;
    const unsigned int tiles_x = (v->W + 7) / 8, tiles_y = (v->local_rows + 3) / 4;
    P.n_ray_slots = tiles_x * tiles_y * 32u;
    P.n_pair_slots = ((v->W + 15) / 16) * tiles_y * 32u;
    P.plane_w_max = v->plane_w_max;
    P.bvh = v->d_bvh;
}

int launch(linon_view* v, CUfunction f, unsigned grid, unsigned block, cudaStream_t s, void* params) {
    void* args[] = {params};
    const CUresult r = g_drv.LaunchKernel(f, grid, 1, 1, block, 1, 1, 0, (CUstream)s, args, nullptr);
    if (r != CUDA_SUCCESS) return v->fail(LINON_ERR_CUDA, "cuLaunchKernel: " + g_drv.err_str(r));
    return LINON_OK;
}

constexpr uint32_t kSharedMeshFaces = 1024;  // 48 B per face + plane groups fit the static shared memory of a CTA

// the bounding-volume hierarchy for meshes beyond the shared-memory stage is built by mesh_tables.hpp: build_bvh_tables
int set_mesh_bvh(linon_view* v, const float* verts, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces) {
    linon::BvhTables bt;
    uint32_t bad = 0;
    switch (linon::build_bvh_tables(verts, n_verts, faces, n_faces, bt, &bad)) {
        case linon::BVH_BAD_FACE: return v->fail(LINON_ERR_INVALID, fmt("face %u indexes past the vertex array", bad));
        case linon::BVH_NON_FINITE: return v->fail(LINON_ERR_INVALID, "linon_set_mesh: non-finite vertex coordinates");
        case linon::BVH_TOO_DEEP: return v->fail(LINON_ERR_INVALID, "linon_set_mesh: hierarchy deeper than the traversal stack");
        default: break;
    }
    CK(v, cudaSetDevice(v->device));
    CK(v, cudaStreamSynchronize(v->stream));
    cudaFree(v->d_tris); cudaFree(v->d_group_range); cudaFree(v->d_bvh);
    v->d_tris = nullptr; v->d_group_range = nullptr; v->d_bvh = nullptr;
    CK(v, cudaMalloc(&v->d_tris, bt.tris.size() * sizeof(float)));
    CK(v, cudaMemcpy(v->d_tris, bt.tris.data(), bt.tris.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(v, cudaMalloc(&v->d_bvh, bt.nodes.size() * sizeof(float)));
    CK(v, cudaMemcpy(v->d_bvh, bt.nodes.data(), bt.nodes.size() * sizeof(float), cudaMemcpyHostToDevice));
    v->n_faces = n_faces;
    v->n_groups = 0;
    v->use_bvh = true;
    v->bvh_nodes = (uint32_t)(bt.nodes.size() / 8);
    v->bvh_depth = bt.depth;
    v->plane_w_max = bt.extent;
    return LINON_OK;
}

}  // namespace

// =============================================================================================================
extern "C" {

const char* linon_last_error(const linon_view* v) { return v ? v->error.c_str() : g_create_error.c_str(); }

int linon_create(int device, uint32_t width, uint32_t height, linon_view** out) {
    if (!out || !frame_size_ok(width, height)) {
        g_create_error = "linon_create: need out != NULL, 2 <= width, height <= 65536 and width * height < 2^31";
        return LINON_ERR_INVALID;
    }
    *out = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_drv_mutex);
        if (!g_drv.load()) {
            g_create_error = g_drv.error;
            return LINON_ERR_NO_DRIVER;
        }
    }
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || device < 0 || device >= n_dev) {
        g_create_error = fmt("CUDA device %d not available (%d devices, %s): no CPU fallback", device, n_dev,
                             cudaGetErrorString(e));
        return LINON_ERR_NO_DRIVER;
    }
    std::unique_ptr<linon_view> v(new linon_view);
    v->device = device;
    auto bail = [&](const char* what, cudaError_t err) {
        g_create_error = fmt("%s: %s", what, cudaGetErrorString(err));
        return LINON_ERR_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaFree(nullptr)) != cudaSuccess) return bail("context init", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    if (prop.major != 10) {
        g_create_error = fmt("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major,
                             prop.minor);
        return LINON_ERR_NO_DRIVER;
    }
    v->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    if ((e = cudaStreamCreateWithFlags(&v->aux, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
    for (auto& ev : v->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("event", e);
    if ((e = cudaMalloc(&v->d_ray_samples, 800 * 32)) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemset(v->d_ray_samples, 0, 800 * 32);
    if ((e = cudaMalloc(&v->d_pos16, 64)) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemset(v->d_pos16, 0, 64);
    if ((e = cudaMalloc(&v->d_counters, 16 * 8)) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemset(v->d_counters, 0, 128);
    if ((e = cudaMalloc(&v->d_outline_keys, 8 * 8)) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc(&v->d_work, 64)) != cudaSuccess) return bail("cudaMalloc", e);
    // start-up state of MainView::new
    v->settings = linon_settings{1.0f, {0.5f, 0.6f}, 0, 1, 50.0f};                 // src/main_view.rs:133-139
    v->body = linon::to_code(linon::PredefinedFunction::MirageSphericalSigmoid);   // src/main_view.rs:100
    const int rc = alloc_frame_buffers(v.get(), width, height);
    if (rc != LINON_OK) {
        g_create_error = v->error;
        return rc;
    }
    *out = v.release();
    return LINON_OK;
}

void linon_destroy(linon_view* v) {
    if (!v) return;
    cudaSetDevice(v->device);
    cudaStreamSynchronize(v->stream);
    cudaStreamSynchronize(v->aux);
    for (auto& kv : v->modules) g_drv.ModuleUnload(kv.second.mod);
    for (void* m : v->ipc_mappings) cudaIpcCloseMemHandle(m);
    free_frame_buffers(v);
    cudaFree(v->d_bvh);
    release_buffer(v, LINON_BUF_RAY_SAMPLES, v->d_ray_samples);
    cudaFree(v->d_tris); cudaFree(v->d_group_range); cudaFree(v->d_pos16); cudaFree(v->d_counters);
    cudaFree(v->d_outline_keys); cudaFree(v->d_work);
    for (auto& ev : v->ev) cudaEventDestroy(ev);
    cudaStreamDestroy(v->stream);
    cudaStreamDestroy(v->aux);
    delete v;
}

int linon_set_mesh(linon_view* v, const float* verts, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces) {
    if (!v || !verts || !faces || n_faces == 0) return v ? v->fail(LINON_ERR_INVALID, "linon_set_mesh: null/empty mesh") : LINON_ERR_INVALID;
    if (n_faces > (1u << 24)) return v->fail(LINON_ERR_INVALID, "linon_set_mesh: more than 2^24 faces");
    // validate BEFORE anything of the current mesh is replaced: a rejected mesh must not be re-staged later by the
    // LINON_OPT_FORCE_BVH toggle (which works from the host copies)
    for (size_t k = 0; k < (size_t)3 * n_faces; ++k)
        if (faces[k] >= n_verts) return v->fail(LINON_ERR_INVALID, fmt("face %zu indexes past the vertex array", k / 3));
    for (size_t k = 0; k < (size_t)3 * n_verts; ++k)
        if (!std::isfinite(verts[k])) return v->fail(LINON_ERR_INVALID, "linon_set_mesh: non-finite vertex coordinates");
    if (verts != v->host_verts.data()) v->host_verts.assign(verts, verts + (size_t)3 * n_verts);
    if (faces != v->host_faces.data()) v->host_faces.assign(faces, faces + (size_t)3 * n_faces);
    verts = v->host_verts.data();
    faces = v->host_faces.data();
    if (n_faces > kSharedMeshFaces || v->opt_force_bvh) return set_mesh_bvh(v, verts, n_verts, faces, n_faces);
    v->use_bvh = false;
    // faces, plane groups and the first-pass bounds (mesh_tables.hpp)
    linon::MeshTables mt;
    uint32_t bad = 0;
    if (!linon::build_mesh_tables(verts, n_verts, faces, n_faces, mt, &bad))
        return v->fail(LINON_ERR_INVALID, fmt("face %u indexes past the vertex array", bad));
    std::vector<float> tris = mt.tris;
    tris.insert(tris.end(), mt.planes.begin(), mt.planes.end());
    tris.insert(tris.end(), mt.p1_planes.begin(), mt.p1_planes.end());
    tris.insert(tris.end(), mt.boxes.begin(), mt.boxes.end());
    std::vector<int> ranges = mt.ranges;  // per group, then the segments
    ranges.insert(ranges.end(), mt.segments.begin(), mt.segments.end());
    CK(v, cudaSetDevice(v->device));
    CK(v, cudaStreamSynchronize(v->stream));
    cudaFree(v->d_tris);
    cudaFree(v->d_group_range);
    v->d_tris = nullptr;
    v->d_group_range = nullptr;
    CK(v, cudaMalloc(&v->d_tris, tris.size() * sizeof(float)));
    CK(v, cudaMemcpy(v->d_tris, tris.data(), tris.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(v, cudaMalloc(&v->d_group_range, ranges.size() * sizeof(int)));
    CK(v, cudaMemcpy(v->d_group_range, ranges.data(), ranges.size() * sizeof(int), cudaMemcpyHostToDevice));
    v->n_groups = mt.n_groups();
    v->n_p1 = mt.n_p1();
    v->n_boxes = mt.n_boxes();
    v->n_segs = mt.n_segments();
    v->n_faces = n_faces;
    v->plane_w_max = mt.coord_max;
    return LINON_OK;
}

int linon_set_camera(linon_view* v, const linon_camera* camera) {
    if (!v || !camera) return LINON_ERR_INVALID;
    v->camera = *camera;
    v->has_camera = true;
    return LINON_OK;
}

int linon_set_settings(linon_view* v, const linon_settings* settings) {
    if (!v || !settings) return LINON_ERR_INVALID;
    if (settings->central_difference_delta < 1 || settings->central_difference_delta > 10)
        return v->fail(LINON_ERR_INVALID, "central_difference_delta must be in 1..10 (src/application.rs:297)");
    v->settings = *settings;
    return LINON_OK;
}

int linon_resize(linon_view* v, uint32_t width, uint32_t height) {
    if (!v) return LINON_ERR_INVALID;
    if (!frame_size_ok(width, height)) return v->fail(LINON_ERR_INVALID, "linon_resize: need 2 <= width, height <= 65536 and width * height < 2^31");
    CK(v, cudaSetDevice(v->device));
    CK(v, cudaStreamSynchronize(v->stream));
    v->out_raycast = nullptr; v->out_endpoint = nullptr; v->out_hit = nullptr;  // a redirection is for one frame size
    return alloc_frame_buffers(v, width, height);
}

int linon_set_shard(linon_view* v, uint32_t rank, uint32_t world, uint32_t rows_per_block) {
    if (!v || world == 0 || rank >= world || rows_per_block == 0) return v ? v->fail(LINON_ERR_INVALID, "linon_set_shard: bad arguments") : LINON_ERR_INVALID;
    CK(v, cudaSetDevice(v->device));
    CK(v, cudaStreamSynchronize(v->stream));
    v->shard_rank = rank; v->shard_world = world; v->rows_per_block = rows_per_block;
    return alloc_queue(v, false);
}

int linon_set_field_function(linon_view* v, const char* wgsl_body, char* err, size_t err_len) {
    if (!v || !wgsl_body) return LINON_ERR_INVALID;
    if (err && err_len) err[0] = '\0';
    CK(v, cudaSetDevice(v->device));
    // validate by building the default variant now; the previous body stays live on failure (src/main_view.rs:677-679)
    const std::string previous = v->body;
    v->body = wgsl_body;  // make_variant consults pair_unfit for the body being installed
    Variant var = make_variant(v, LINON_RENDER_DEFAULT);
    if (var.n_faces == 0) default_mesh_counts(var);
    Module* m = nullptr;
    float jit = 0.0f;
    const int rc = get_module_for_body(v, v->body, var, &m, &jit);
    if (rc != LINON_OK) {
        v->body = previous;
        if (err && err_len) snprintf(err, err_len, "%s", v->error.c_str());
        return rc;
    }
    return LINON_OK;
}

int linon_reload_shader_source(linon_view* v, const char* wgsl_module_source, char* err, size_t err_len) {
    if (!v) return LINON_ERR_INVALID;
    if (err && err_len) err[0] = '\0';
    CK(v, cudaSetDevice(v->device));
    linon::ModuleLayer layer;  // NULL / "" = the stock module
    if (wgsl_module_source && wgsl_module_source[0]) {
        std::string merr;
        bool unsupported = false;
        if (!linon::translate_wgsl_module(wgsl_module_source, layer, merr, unsupported)) {
            if (err && err_len) snprintf(err, err_len, "%s", merr.c_str());
            return v->fail(unsupported ? LINON_ERR_UNSUPPORTED : LINON_ERR_COMPILE, merr);
        }
    }
    // validate by building the default variant with the CURRENT field function; on failure the previous module text and
    // kernel stay live (src/main_view.rs:677-679 returns before `self.shader_src = new_src`)
    const linon::ModuleLayer previous = v->layer;
    v->layer = layer;
    Variant var = make_variant(v, LINON_RENDER_DEFAULT);
    if (var.n_faces == 0) default_mesh_counts(var);
    Module* m = nullptr;
    float jit = 0.0f;
    const int rc = get_module_for_body(v, v->body, var, &m, &jit);
    if (rc != LINON_OK) {
        v->layer = previous;
        if (err && err_len) snprintf(err, err_len, "%s", v->error.c_str());
        return rc;
    }
    return LINON_OK;
}

int linon_reload_shader(linon_view* v, const char* wgsl_module_source, const char* wgsl_body, char* err, size_t err_len) {
    if (!v || !wgsl_body) return LINON_ERR_INVALID;
    if (err && err_len) err[0] = '\0';
    CK(v, cudaSetDevice(v->device));
    linon::ModuleLayer layer;  // NULL / "" = the stock module
    if (wgsl_module_source && wgsl_module_source[0]) {
        std::string merr;
        bool unsupported = false;
        if (!linon::translate_wgsl_module(wgsl_module_source, layer, merr, unsupported)) {
            if (err && err_len) snprintf(err, err_len, "%s", merr.c_str());
            return v->fail(unsupported ? LINON_ERR_UNSUPPORTED : LINON_ERR_COMPILE, merr);
        }
    }
    // the PAIR is validated once, like the reference compiles new_src with the new field function spliced in
    // (src/main_view.rs:649-690); both are committed or neither
    const linon::ModuleLayer old_layer = v->layer;
    const std::string old_body = v->body;
    v->layer = layer;
    v->body = wgsl_body;
    Variant var = make_variant(v, LINON_RENDER_DEFAULT);
    if (var.n_faces == 0) default_mesh_counts(var);
    Module* m = nullptr;
    float jit = 0.0f;
    const int rc = get_module_for_body(v, v->body, var, &m, &jit);
    if (rc != LINON_OK) {
        v->layer = old_layer;
        v->body = old_body;
        if (err && err_len) snprintf(err, err_len, "%s", v->error.c_str());
        return rc;
    }
    return LINON_OK;
}

int linon_module_overrides(linon_view* v, char* names, size_t names_len) {
    if (!v || !names || !names_len) return LINON_ERR_INVALID;
    if (v->layer.names.size() + 1 > names_len) return LINON_ERR_SIZE;
    memcpy(names, v->layer.names.c_str(), v->layer.names.size() + 1);
    return LINON_OK;
}

namespace {
// (Re)encodes the TMA descriptor of the endpoint image when the buffer, the frame size or the delta changed.  Returns the
// descriptor, or nullptr when the tile kernel cannot run (option off, delta outside the slider's range, driver without
// cuTensorMapEncodeTiled, encode refused): launch_overlay then runs the direct kernel - same results, same device.
const void* endpoint_tensor_map(linon_view* v) {
    const int delta = v->settings.central_difference_delta;
    if (v->opt_overlay_kernel != 2 || delta < 1 || delta > 10 || !g_drv.TensorMapEncodeTiled || !v->d_endpoint) return nullptr;
    if (v->endpoint_map_ptr == v->d_endpoint && v->endpoint_map_w == v->W && v->endpoint_map_h == v->H && v->endpoint_map_delta == delta)
        return v->endpoint_map_ok ? &v->endpoint_map : nullptr;
    v->endpoint_map_rows = linon::overlay_tile_rows((int)v->W, (int)v->H, delta, v->sm_count);
    v->endpoint_map_ptr = v->d_endpoint;
    v->endpoint_map_w = v->W;
    v->endpoint_map_h = v->H;
    v->endpoint_map_delta = delta;
    unsigned int box_floats = 0, box_rows = 0;
    linon::overlay_tile_box(delta, v->endpoint_map_rows, &box_floats, &box_rows);
    const cuuint64_t dims[2] = {(cuuint64_t)v->W * 4, (cuuint64_t)v->H};
    const cuuint64_t strides[1] = {(cuuint64_t)v->W * 16};
    const cuuint32_t box[2] = {box_floats, box_rows};
    const cuuint32_t elem_strides[2] = {1, 1};
    const CUresult r = g_drv.TensorMapEncodeTiled(&v->endpoint_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)v->d_endpoint, dims, strides, box,
                                                  elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    v->endpoint_map_ok = r == CUDA_SUCCESS;
    return v->endpoint_map_ok ? &v->endpoint_map : nullptr;
}
}  // namespace

int linon_render(linon_view* v, uint32_t flags) {
    if (!v) return LINON_ERR_INVALID;
    const bool do_march = !(flags & LINON_RENDER_NO_MARCH);
    const bool do_overlay = !(flags & LINON_RENDER_NO_OVERLAY);
    const bool do_sample = !(flags & LINON_RENDER_NO_SAMPLE_RAYS);
    if ((do_march || do_sample) && (!v->d_tris || !v->has_camera))
        return v->fail(LINON_ERR_INVALID, "linon_render: set a mesh and a camera first");
    CK(v, cudaSetDevice(v->device));
    if (!(flags & LINON_RENDER_KEEP_TIMINGS)) {
        v->timings = linon_timings{};
        v->ran_march = v->ran_overlay = v->ran_sample = v->ran_outline = v->ran_phases = false;
    }

    if (do_march && v->opt_kernel == 1 && !v->use_bvh && !v->layer.linear_mode) {
        const int rc = alloc_queue(v, true);
        if (rc != LINON_OK) return rc;
    }
    Module* mod = nullptr;
    MarchParamsH P;
    fill_march_params(v, P);
    if (do_march || do_sample) {
        Variant var = make_variant(v, flags);
        const int rc = get_module_for_body(v, v->body, var, &mod, &v->timings.jit_ms);
        if (rc != LINON_OK) return rc;
        P.n_fixed_steps = var.adaptive ? 0 : fixed_step_count(var.h_initial, v->layer.t_max > 0.0 ? (float)v->layer.t_max : 5.0f);
    }

    if (do_sample) {
        // sample-ray start positions: outline search over the exponents of the PREVIOUS overlay pass (the march
        // dispatch runs before the overlay dispatch in the reference, src/main_view.rs:838-846) or the click ring
        if ((flags & LINON_RENDER_OUTLINE) || v->layer.sample_outline_rays == 1) {  // `const sample_outline_rays`, :337
            CK(v, cudaEventRecord(v->ev[6], v->stream));
            linon::launch_outline_search(v->stream, (int)v->W, (int)v->H, v->d_exponents, v->d_outline_keys, v->d_pos16,
                                         v->sm_count);
            CK(v, cudaEventRecord(v->ev[7], v->stream));
            v->ran_outline = true;
        } else {
            linon::launch_click_positions(v->stream, v->settings.mouse_pos[0], v->settings.mouse_pos[1], v->d_pos16);
        }
        CK(v, cudaGetLastError());
        CK(v, cudaEventRecord(v->ev[8], v->stream));
        CK(v, cudaStreamWaitEvent(v->aux, v->ev[8], 0));
        SampleParamsH S;
        S.march = P;
        S.pos16 = v->d_pos16;
        S.ray_samples = v->d_ray_samples;
        CK(v, cudaEventRecord(v->ev[4], v->aux));
        const int rc = launch(v, mod->sample, 8, 32, v->aux, &S);  // one ray per CTA (march_kernels.cuh: linon_sample_rays)
        if (rc != LINON_OK) return rc;
        CK(v, cudaEventRecord(v->ev[5], v->aux));
        v->ran_sample = true;
    }

    if (do_march && v->local_rows > 0) {
        CK(v, cudaMemsetAsync(v->d_counters, 0, 128, v->stream));
        const unsigned block = (unsigned)v->opt_block;
        const int kernel_opt = v->layer.linear_mode ? 0 : v->opt_kernel;  // linear_mode: one straight ray per pixel (tile kernel)
        const bool use_pairs = kernel_opt == 2 && mod->pairs != nullptr;
        const unsigned needed = ((use_pairs ? P.n_pair_slots : P.n_ray_slots) + block - 1) / block;
        const bool use_refill = kernel_opt == 3;
        v->timings.march_kernel = use_pairs ? 2 : ((kernel_opt == 1 && !v->use_bvh) ? 1 : (use_refill ? 3 : 0));
        v->timings.march_block = (int)block;
        CK(v, cudaEventRecord(v->ev[0], v->stream));
        if (kernel_opt == 1 && !v->use_bvh) {
            // persistent warps: phase A (pre-hit rays, lane refill from the fresh cursor) then phase B (queued post-hit
            // rays) as two launches on the same stream, so phase B sees a complete, immutable queue
            CK(v, cudaMemsetAsync(v->d_work, 0, 64, v->stream));
            unsigned grids[2];
            CUfunction fns[2] = {mod->phase_a, mod->phase_b};
            for (int k = 0; k < 2; ++k) {
                int per_sm = v->opt_ctas_per_sm;
                if (per_sm <= 0) {
                    const CUresult r = g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fns[k], (int)block, 0);
                    if (r != CUDA_SUCCESS || per_sm <= 0) return v->fail(LINON_ERR_CUDA, "occupancy query failed: " + g_drv.err_str(r));
                }
                grids[k] = (unsigned)(per_sm * v->sm_count);
                if (grids[k] > needed) grids[k] = needed;
            }
            int rc = launch(v, mod->phase_a, grids[0], block, v->stream, &P);
            if (rc != LINON_OK) return rc;
            CK(v, cudaEventRecord(v->ev[10], v->stream));
            rc = launch(v, mod->phase_b, grids[1], block, v->stream, &P);
            if (rc != LINON_OK) return rc;
            v->timings.march_regs = mod->regs_phase_a;
            v->timings.march_grid = (int)grids[0];
            v->ran_phases = true;
        } else if (use_refill) {
            CK(v, cudaMemsetAsync(v->d_work, 0, 64, v->stream));
            int per_sm = v->opt_ctas_per_sm;
            if (per_sm <= 0) {
                const CUresult r = g_drv.OccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mod->refill, (int)block, 0);
                if (r != CUDA_SUCCESS || per_sm <= 0) return v->fail(LINON_ERR_CUDA, "occupancy query failed: " + g_drv.err_str(r));
            }
            unsigned grid = (unsigned)(per_sm * v->sm_count);
            if (grid > needed) grid = needed;
            const int rc = launch(v, mod->refill, grid, block, v->stream, &P);
            if (rc != LINON_OK) return rc;
            v->timings.march_regs = mod->regs_refill;
            v->timings.march_grid = (int)grid;
        } else if (use_pairs) {
            const int rc = launch(v, mod->pairs, needed, block, v->stream, &P);
            if (rc != LINON_OK) return rc;
            v->timings.march_regs = mod->regs_pairs;
            v->timings.march_grid = (int)needed;
        } else {
            const int rc = launch(v, mod->tiles, needed, block, v->stream, &P);
            if (rc != LINON_OK) return rc;
            v->timings.march_regs = mod->regs_tiles;
            v->timings.march_grid = (int)needed;
        }
        CK(v, cudaEventRecord(v->ev[1], v->stream));
        v->ran_march = true;
        v->last_cubin_hash = v->timings.march_kernel == 0 ? mod->tiles_hash : mod->cubin_hash;
        v->last_rays = (uint64_t)v->local_rows * v->W;
    }

    if (do_overlay) {
        // mode 0 copies the colours and needs no tile either way; the tile kernel handles it without its TMA load
        const void* ep_map = endpoint_tensor_map(v);
        CK(v, cudaEventRecord(v->ev[2], v->stream));
        linon::launch_overlay(v->stream, (int)v->W, (int)v->H, v->settings.overlay_mode, v->settings.central_difference_delta,
                              v->settings.lyapunov_scaling, v->d_raycast, v->d_endpoint, v->d_final, v->d_exponents,
                              v->opt_overlay_kernel, ep_map, v->endpoint_map_rows);
        v->timings.overlay_kernel = (v->opt_overlay_kernel == 2 && !ep_map) ? 1 : v->opt_overlay_kernel;
        CK(v, cudaGetLastError());
        CK(v, cudaEventRecord(v->ev[3], v->stream));
        v->ran_overlay = true;
    }
    if (v->ran_sample) {  // join the sample-ray stream so one sync covers the frame
        CK(v, cudaEventRecord(v->ev[9], v->aux));
        CK(v, cudaStreamWaitEvent(v->stream, v->ev[9], 0));
    }
    return LINON_OK;
}

int linon_sync(linon_view* v) {
    if (!v) return LINON_ERR_INVALID;
    CK(v, cudaSetDevice(v->device));
    CK(v, cudaStreamSynchronize(v->stream));
    CK(v, cudaStreamSynchronize(v->aux));
    return LINON_OK;
}

size_t linon_buffer_bytes(const linon_view* v, int buffer) {
    if (!v) return 0;
    const size_t n = (size_t)v->W * v->H;
    switch (buffer) {
        case LINON_BUF_FINAL_RGBA8:
        case LINON_BUF_RAYCAST_RGBA8:
        case LINON_BUF_EXPONENTS_F32:
        case LINON_BUF_HIT_ID_I32: return n * 4;
        case LINON_BUF_ENDPOINT_F32X4: return n * 16;
        case LINON_BUF_RAY_SAMPLES: return 800 * 32;
        case LINON_BUF_SAMPLE_POS: return 64;
    }
    return 0;
}

void* linon_device_ptr(linon_view* v, int buffer) {
    if (!v) return nullptr;
    switch (buffer) {
        case LINON_BUF_FINAL_RGBA8: return v->d_final;
        case LINON_BUF_RAYCAST_RGBA8: return v->d_raycast;
        case LINON_BUF_ENDPOINT_F32X4: return v->d_endpoint;
        case LINON_BUF_EXPONENTS_F32: return v->d_exponents;
        case LINON_BUF_HIT_ID_I32: return v->d_hit;
        case LINON_BUF_RAY_SAMPLES: return v->d_ray_samples;
        case LINON_BUF_SAMPLE_POS: return v->d_pos16;
    }
    return nullptr;
}

void* linon_stream(linon_view* v) { return v ? (void*)v->stream : nullptr; }

// ---- peer stores: a shard's march writes straight into the root shard's frame buffers (no gather) ------------------
int linon_redirect_output(linon_view* v, int buffer, void* device_ptr) {
    if (!v) return LINON_ERR_INVALID;
    switch (buffer) {
        case LINON_BUF_RAYCAST_RGBA8: v->out_raycast = (uchar4*)device_ptr; return LINON_OK;
        case LINON_BUF_ENDPOINT_F32X4: v->out_endpoint = (float4*)device_ptr; return LINON_OK;
        case LINON_BUF_HIT_ID_I32: v->out_hit = (int*)device_ptr; return LINON_OK;
    }
    return v->fail(LINON_ERR_INVALID, "linon_redirect_output: only the march outputs (raycast, endpoint, hit id) can be redirected");
}

int linon_ipc_export(linon_view* v, int buffer, void* handle64) {
    if (!v || !handle64) return LINON_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    void* p = linon_device_ptr(v, buffer);
    if (!p) return v->fail(LINON_ERR_INVALID, "linon_ipc_export: unknown buffer");
    CK(v, cudaSetDevice(v->device));
    cudaIpcMemHandle_t h;
    CK(v, cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    return LINON_OK;
}

int linon_ipc_open(linon_view* v, const void* handle64, void** device_ptr) {
    if (!v || !handle64 || !device_ptr) return LINON_ERR_INVALID;
    CK(v, cudaSetDevice(v->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CK(v, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    v->ipc_mappings.push_back(p);
    *device_ptr = p;
    return LINON_OK;
}

// ---- display hand-off: the buffers the application SHOWS, shared with its graphics API without a copy -----------------------
namespace {
struct ImportedBuffer {
    CUmemGenericAllocationHandle handle = 0;
    CUdeviceptr ptr = 0;
    size_t size = 0;
    int device = 0;
};
CUmemAllocationProp exportable_prop(int device) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof(prop));
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}
}  // namespace

int linon_export_buffer_fd(linon_view* v, int buffer, int* fd, size_t* allocation_bytes) {
    if (!v || !fd) return LINON_ERR_INVALID;
    if (buffer != LINON_BUF_FINAL_RGBA8 && buffer != LINON_BUF_RAY_SAMPLES)
        return v->fail(LINON_ERR_INVALID, "linon_export_buffer_fd: only the buffers the application displays (final image, ray samples) are exported");
    if (!g_drv.has_vmm) return v->fail(LINON_ERR_NO_DRIVER, "this CUDA driver lacks the virtual memory management API (cuMemCreate ...)");
    CK(v, cudaSetDevice(v->device));
    CK(v, cudaStreamSynchronize(v->stream));
    CK(v, cudaStreamSynchronize(v->aux));
    const size_t bytes = linon_buffer_bytes(v, buffer);
    void** slot = buffer == LINON_BUF_FINAL_RGBA8 ? &v->d_final : &v->d_ray_samples;
    auto it = v->exported.find(buffer);
    if (it == v->exported.end()) {
        // move the buffer into an exportable allocation (once; linon_resize reallocates frame buffers plainly again)
        const CUmemAllocationProp prop = exportable_prop(v->device);
        size_t gran = 0;
        CUresult r = g_drv.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
        if (r != CUDA_SUCCESS || gran == 0) return v->fail(LINON_ERR_CUDA, "cuMemGetAllocationGranularity: " + g_drv.err_str(r));
        linon_view::VmmAlloc a;
        a.size = (bytes + gran - 1) / gran * gran;
        if ((r = g_drv.MemCreate(&a.handle, a.size, &prop, 0)) != CUDA_SUCCESS) return v->fail(LINON_ERR_CUDA, "cuMemCreate: " + g_drv.err_str(r));
        if ((r = g_drv.MemAddressReserve(&a.ptr, a.size, 0, 0, 0)) != CUDA_SUCCESS) {
            g_drv.MemRelease(a.handle);
            return v->fail(LINON_ERR_CUDA, "cuMemAddressReserve: " + g_drv.err_str(r));
        }
        CUmemAccessDesc access;
        memset(&access, 0, sizeof(access));
        access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        access.location.id = v->device;
        access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if ((r = g_drv.MemMap(a.ptr, a.size, 0, a.handle, 0)) != CUDA_SUCCESS || (r = g_drv.MemSetAccess(a.ptr, a.size, &access, 1)) != CUDA_SUCCESS) {
            g_drv.MemAddressFree(a.ptr, a.size);
            g_drv.MemRelease(a.handle);
            return v->fail(LINON_ERR_CUDA, "cuMemMap / cuMemSetAccess: " + g_drv.err_str(r));
        }
        cudaError_t e = cudaMemset((void*)a.ptr, 0, a.size);
        if (e == cudaSuccess) e = cudaMemcpy((void*)a.ptr, *slot, bytes, cudaMemcpyDeviceToDevice);  // keep what is on screen
        if (e != cudaSuccess) {
            g_drv.MemUnmap(a.ptr, a.size);
            g_drv.MemAddressFree(a.ptr, a.size);
            g_drv.MemRelease(a.handle);
            return v->fail(LINON_ERR_CUDA, std::string("copy into the exportable allocation: ") + cudaGetErrorString(e));
        }
        cudaFree(*slot);
        *slot = (void*)a.ptr;
        it = v->exported.emplace(buffer, a).first;
    }
    int out = -1;
    const CUresult r = g_drv.MemExportToShareableHandle(&out, it->second.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (r != CUDA_SUCCESS) return v->fail(LINON_ERR_CUDA, "cuMemExportToShareableHandle: " + g_drv.err_str(r));
    *fd = out;
    if (allocation_bytes) *allocation_bytes = it->second.size;
    return LINON_OK;
}

int linon_import_buffer_fd(int device, int fd, size_t allocation_bytes, void** device_ptr, void** import_handle) {
    if (fd < 0 || !allocation_bytes || !device_ptr || !import_handle) return LINON_ERR_INVALID;
    {
        std::lock_guard<std::mutex> lock(g_drv_mutex);
        if (!g_drv.load()) {
            g_create_error = g_drv.error;
            return LINON_ERR_NO_DRIVER;
        }
    }
    if (!g_drv.has_vmm) { g_create_error = "this CUDA driver lacks the virtual memory management API"; return LINON_ERR_NO_DRIVER; }
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) { g_create_error = "no CUDA device (no CPU fallback)"; return LINON_ERR_NO_DRIVER; }
    ImportedBuffer* b = new ImportedBuffer;
    b->size = allocation_bytes;
    b->device = device;
    CUresult r = g_drv.MemImportFromShareableHandle(&b->handle, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    if (r == CUDA_SUCCESS) r = g_drv.MemAddressReserve(&b->ptr, b->size, 0, 0, 0);
    if (r == CUDA_SUCCESS) r = g_drv.MemMap(b->ptr, b->size, 0, b->handle, 0);
    if (r == CUDA_SUCCESS) {
        CUmemAccessDesc access;
        memset(&access, 0, sizeof(access));
        access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        access.location.id = device;
        access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = g_drv.MemSetAccess(b->ptr, b->size, &access, 1);
    }
    if (r != CUDA_SUCCESS) {
        g_create_error = "importing the shared allocation: " + g_drv.err_str(r);
        if (b->ptr) g_drv.MemAddressFree(b->ptr, b->size);
        if (b->handle) g_drv.MemRelease(b->handle);
        delete b;
        return LINON_ERR_CUDA;
    }
    *device_ptr = (void*)b->ptr;
    *import_handle = b;
    return LINON_OK;
}

int linon_release_import(void* import_handle) {
    if (!import_handle) return LINON_ERR_INVALID;
    ImportedBuffer* b = (ImportedBuffer*)import_handle;
    cudaSetDevice(b->device);
    cudaDeviceSynchronize();
    g_drv.MemUnmap(b->ptr, b->size);
    g_drv.MemAddressFree(b->ptr, b->size);
    g_drv.MemRelease(b->handle);
    delete b;
    return LINON_OK;
}

int linon_copy_to_host(int device, const void* device_ptr, void* dst, size_t bytes) {
    if (!device_ptr || !dst) return LINON_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return LINON_ERR_NO_DRIVER;
    if (cudaDeviceSynchronize() != cudaSuccess) return LINON_ERR_CUDA;
    return cudaMemcpy(dst, device_ptr, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? LINON_OK : LINON_ERR_CUDA;
}

int linon_read(linon_view* v, int buffer, void* dst, size_t dst_bytes) {
    if (!v || !dst) return LINON_ERR_INVALID;
    const size_t n = linon_buffer_bytes(v, buffer);
    void* src = linon_device_ptr(v, buffer);
    if (!n || !src) return v->fail(LINON_ERR_INVALID, "linon_read: unknown buffer");
    if (dst_bytes < n) return v->fail(LINON_ERR_SIZE, fmt("linon_read: need %zu bytes, got %zu", n, dst_bytes));
    CK(v, cudaSetDevice(v->device));
    CK(v, cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, v->stream));
    CK(v, cudaStreamSynchronize(v->stream));
    return LINON_OK;
}

int linon_set_option(linon_view* v, int option, int value) {
    if (!v) return LINON_ERR_INVALID;
    switch (option) {
        case LINON_OPT_MATH: if (value < 0 || value > 2) break; v->opt_math = value; return LINON_OK;
        case LINON_OPT_KERNEL: if (value < 0 || value > 3) break; v->opt_kernel = value; return LINON_OK;
        case LINON_OPT_PAIR_MIN_BLOCKS: if (value < 0 || value > 32) break; v->opt_pair_min_blocks = value; return LINON_OK;
        case LINON_OPT_SPECIALIZE: if (value < 0 || value > 1) break; v->opt_specialize = value; return LINON_OK;
        case LINON_OPT_FORCE_BVH:
            if (value < 0 || value > 1) break;
            if (v->opt_force_bvh != value) {
                v->opt_force_bvh = value;
                if (!v->host_faces.empty())  // re-stage the current mesh for the other intersection path
                    return linon_set_mesh(v, v->host_verts.data(), (uint32_t)(v->host_verts.size() / 3), v->host_faces.data(),
                                          (uint32_t)(v->host_faces.size() / 3));
            }
            return LINON_OK;
        case LINON_OPT_BLOCK: if (value < 32 || value > 1024 || value % 32) break; v->opt_block = value; return LINON_OK;
        case LINON_OPT_CTAS_PER_SM: if (value < 0 || value > 32) break; v->opt_ctas_per_sm = value; return LINON_OK;
        case LINON_OPT_REFILL_MIN: if (value < 1 || value > 32) break; v->opt_refill_min = value; return LINON_OK;
        case LINON_OPT_FACE_UNROLL: if (value < 1 || value > 1024) break; v->opt_unroll = value; return LINON_OK;
        case LINON_OPT_MAX_REGS: if (value < 0 || value > 255) break; v->opt_max_regs = value; return LINON_OK;
        case LINON_OPT_PROFILE: if (value < 0 || value > 1) break; v->opt_profile = value; return LINON_OK;
        case LINON_OPT_CULL: if (value < 0 || value > 1) break; v->opt_cull = value; return LINON_OK;
        case LINON_OPT_QUIET_ZONE: if (value < 0 || value > 2) break; v->opt_quiet = value; return LINON_OK;
        case LINON_OPT_OVERLAY_KERNEL: if (value < 0 || value > 2) break; v->opt_overlay_kernel = value; return LINON_OK;
        case LINON_OPT_COLLAPSE_RUNS: if (value < 0 || value > 1) break; v->opt_collapse = value; return LINON_OK;
        case LINON_OPT_QUIET_RADIUS: if (value != 0 && (value < 1250 || value > 4000)) break; v->opt_quiet_radius = value; return LINON_OK;
    }
    return v->fail(LINON_ERR_INVALID, fmt("linon_set_option(%d, %d): invalid", option, value));
}

int linon_get_counters(linon_view* v, linon_counters* out) {
    if (!v || !out) return LINON_ERR_INVALID;
    unsigned long long c[5];
    CK(v, cudaSetDevice(v->device));
    CK(v, cudaMemcpyAsync(c, v->d_counters, 40, cudaMemcpyDeviceToHost, v->stream));
    CK(v, cudaStreamSynchronize(v->stream));
    out->rays = v->last_rays;
    out->accepted_steps = c[0];
    out->prehit_steps = c[1];
    out->rejected_steps = c[2];
    out->face_tests = c[3];
    out->culling_checks = c[4];
    out->plane_tests = v->opt_cull ? c[4] * v->n_groups : 0;
    return LINON_OK;
}

int linon_debug_profile(linon_view* v, uint64_t* out16) {
    if (!v || !out16) return LINON_ERR_INVALID;
    CK(v, cudaSetDevice(v->device));
    CK(v, cudaMemcpyAsync(out16, v->d_counters, 128, cudaMemcpyDeviceToHost, v->stream));
    CK(v, cudaStreamSynchronize(v->stream));
    return LINON_OK;
}

int linon_nvrtc_info(char* path, size_t path_len, int* major, int* minor) {
    std::lock_guard<std::mutex> lock(g_nvrtc_mutex);
    if (!g_nvrtc.load()) return LINON_ERR_COMPILE;
    if (path && path_len) snprintf(path, path_len, "%s", g_nvrtc.path.c_str());
    if (major) *major = g_nvrtc.major;
    if (minor) *minor = g_nvrtc.minor;
    return LINON_OK;
}

int linon_march_cubin_hash(linon_view* v, uint64_t* out) {
    if (!v || !out) return LINON_ERR_INVALID;
    *out = v->last_cubin_hash;
    return LINON_OK;
}

int linon_get_timings(linon_view* v, linon_timings* out) {
    if (!v || !out) return LINON_ERR_INVALID;
    const int rc = linon_sync(v);
    if (rc != LINON_OK) return rc;
    if (v->ran_march) cudaEventElapsedTime(&v->timings.march_ms, v->ev[0], v->ev[1]);
    if (v->ran_march && v->ran_phases) cudaEventElapsedTime(&v->timings.march_phase_a_ms, v->ev[0], v->ev[10]);
    if (v->ran_overlay) cudaEventElapsedTime(&v->timings.overlay_ms, v->ev[2], v->ev[3]);
    if (v->ran_sample) cudaEventElapsedTime(&v->timings.sample_rays_ms, v->ev[4], v->ev[5]);
    if (v->ran_outline) cudaEventElapsedTime(&v->timings.outline_search_ms, v->ev[6], v->ev[7]);
    *out = v->timings;
    return LINON_OK;
}

int linon_translate_wgsl(const char* wgsl_body, char* out, size_t out_len, char* err, size_t err_len) {
    if (!wgsl_body || !out) return LINON_ERR_INVALID;
    std::string cuda, e;
    if (!linon::translate_wgsl_body(wgsl_body, cuda, e)) {
        if (err && err_len) snprintf(err, err_len, "%s", e.c_str());
        return LINON_ERR_COMPILE;
    }
    if (cuda.size() + 1 > out_len) return LINON_ERR_SIZE;
    memcpy(out, cuda.c_str(), cuda.size() + 1);
    return LINON_OK;
}

namespace {
int compile_check_impl(const char* module_src, const char* wgsl_body, uint32_t flags, int math_mode, uint32_t n_faces,
                       size_t* cubin_bytes, char* err, size_t err_len) {
    if (!wgsl_body) return LINON_ERR_INVALID;
    linon_view tmp;
    tmp.n_faces = n_faces;
    tmp.n_groups = tmp.n_p1 = n_faces;  // an arbitrary mesh of that size: every face its own plane group, no boxes
    tmp.n_segs = 1;
    if (n_faces == 36) tmp.n_faces = 0;  // the start-up mesh: make_variant's caller fills the real table sizes
    if (n_faces > kSharedMeshFaces) tmp.use_bvh = true;
    tmp.opt_math = math_mode;
    tmp.opt_quiet = 0;  // the build check names the quiet-zone variants explicitly (LINON_CHECK_QUIET_ZONE)
    tmp.body = wgsl_body;
    tmp.settings.field_weight = 1.0f;  // the application's default (src/main_view.rs:133-139)
    if (module_src && module_src[0]) {
        std::string merr;
        bool unsupported = false;
        if (!linon::translate_wgsl_module(module_src, tmp.layer, merr, unsupported)) {
            if (err && err_len) snprintf(err, err_len, "%s", merr.c_str());
            return unsupported ? LINON_ERR_UNSUPPORTED : LINON_ERR_COMPILE;
        }
    }
    if (const char* k = getenv("LINON_CHECK_KERNEL")) tmp.opt_kernel = atoi(k);  // developer aid: 0 = without the pair kernel
    if (const char* k = getenv("LINON_CHECK_PAIR_MIN_BLOCKS")) tmp.opt_pair_min_blocks = atoi(k);
    if (const char* k = getenv("LINON_CHECK_BLOCK")) tmp.opt_block = atoi(k);
    if (const char* k = getenv("LINON_CHECK_MAX_REGS")) tmp.opt_max_regs = atoi(k);
    if (const char* k = getenv("LINON_CHECK_SPECIALIZE")) tmp.opt_specialize = atoi(k);
    if (const char* k = getenv("LINON_CHECK_BVH")) tmp.use_bvh = atoi(k) != 0;
    if (const char* k = getenv("LINON_CHECK_QUIET_ZONE")) tmp.opt_quiet = atoi(k);
    if (const char* k = getenv("LINON_CHECK_QUIET_RADIUS")) tmp.opt_quiet_radius = atoi(k);
    if (const char* k = getenv("LINON_CHECK_COLLAPSE")) tmp.opt_collapse = atoi(k);
    if (const char* k = getenv("LINON_CHECK_CULL")) tmp.opt_cull = atoi(k);
    Variant var = make_variant(&tmp, flags);
    if (var.n_faces == 0 && !tmp.use_bvh) default_mesh_counts(var);
    std::vector<char> cubin;
    std::string log;
    bool ok = compile_cubin(wgsl_body, var, tmp.layer, cubin, log);
    if (!ok && var.pairs && !getenv("LINON_DUMP_SOURCE_ONLY")) {  // same fallback as linon_render: the body may not instantiate for two rays per thread
        var.pairs = 0;
        std::string log2;
        ok = compile_cubin(wgsl_body, var, tmp.layer, cubin, log2);
        log = ok ? "pair kernel not built for this body:\n" + log : log2;
    }
    if (!ok) {
        if (err && err_len) snprintf(err, err_len, "%s", log.c_str());
        return LINON_ERR_COMPILE;
    }
    if (cubin_bytes) *cubin_bytes = cubin.size();
    if (const char* hf = getenv("LINON_DUMP_CUBIN_HASH")) {  // developer aid: the hash linon_march_cubin_hash reports for this variant
        if (FILE* f = fopen(hf, "w")) {
            fprintf(f, "%016llx\n", (unsigned long long)cubin_code_hash(cubin, "linon_march_tiles"));
            fclose(f);
        }
    }
    if (const char* dump = getenv("LINON_DUMP_CUBIN")) {  // developer aid: inspect with cuobjdump -sass / -res-usage
        if (FILE* f = fopen(dump, "wb")) {
            fwrite(cubin.data(), 1, cubin.size(), f);
            fclose(f);
        }
    }
    if (err && err_len) snprintf(err, err_len, "%s", log.c_str());  // warnings, if any
    return LINON_OK;
}
}  // namespace

int linon_compile_check(const char* wgsl_body, uint32_t flags, int math_mode, uint32_t n_faces, size_t* cubin_bytes,
                        char* err, size_t err_len) {
    return compile_check_impl(nullptr, wgsl_body, flags, math_mode, n_faces, cubin_bytes, err, err_len);
}

int linon_compile_check_module(const char* wgsl_module_source, const char* wgsl_body, uint32_t flags, int math_mode,
                               uint32_t n_faces, size_t* cubin_bytes, char* err, size_t err_len) {
    return compile_check_impl(wgsl_module_source, wgsl_body, flags, math_mode, n_faces, cubin_bytes, err, err_len);
}

int linon_translate_wgsl_module(const char* wgsl_module_source, char* out, size_t out_len, char* names, size_t names_len,
                                char* err, size_t err_len) {
    if (!wgsl_module_source) return LINON_ERR_INVALID;
    linon::ModuleLayer layer;
    std::string merr;
    bool unsupported = false;
    if (!linon::translate_wgsl_module(wgsl_module_source, layer, merr, unsupported)) {
        if (err && err_len) snprintf(err, err_len, "%s", merr.c_str());
        return unsupported ? LINON_ERR_UNSUPPORTED : LINON_ERR_COMPILE;
    }
    std::string text = layer.preamble.empty() ? layer.members : "// module constants: " + layer.preamble + "\n" + layer.members;
    if (!layer.march_defines.empty()) text = "// march-level constants (macros of the kernel template):\n" + layer.march_defines + text;
    if (layer.adaptive_sampling >= 0) text = fmt("// adaptive_sampling = %d\n", layer.adaptive_sampling) + text;
    if (layer.sample_outline_rays >= 0) text = fmt("// sample_outline_rays = %d\n", layer.sample_outline_rays) + text;
    if (layer.h_initial > 0.0) text = fmt("// h_initial = %.9g\n", layer.h_initial) + text;
    if (out && out_len) {
        if (text.size() + 1 > out_len) return LINON_ERR_SIZE;
        memcpy(out, text.c_str(), text.size() + 1);
    }
    if (names && names_len) snprintf(names, names_len, "%s", layer.names.c_str());
    return LINON_OK;
}

int linon_scan_wgsl_module(const char* wgsl_module_source, char* out, size_t out_len, char* err, size_t err_len) {
    if (!wgsl_module_source || !out || !out_len) return LINON_ERR_INVALID;
    std::vector<linon::ModuleItem> items;
    std::string merr;
    if (!linon::scan_wgsl_module(wgsl_module_source, items, merr)) {
        if (err && err_len) snprintf(err, err_len, "%s", merr.c_str());
        return LINON_ERR_COMPILE;
    }
    std::string text;
    for (const auto& it : items) {
        text += fmt("%s %s 0x%016llx %d 0x%016llx %zu", it.kind.c_str(), it.name.c_str(), (unsigned long long)it.hash, it.line,
                    (unsigned long long)it.masked_hash, it.literals.size());
        for (double x : it.literals) text += fmt(" %.17g", x);
        text += "\n";
    }
    if (text.size() + 1 > out_len) return LINON_ERR_SIZE;
    memcpy(out, text.c_str(), text.size() + 1);
    return LINON_OK;
}

// ---- parity / debug ------------------------------------------------------------------------------------------
namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t up(const void* src, size_t bytes) {
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 4);
        if (e != cudaSuccess) return e;
        return src ? cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) : cudaMemset(p, 0, bytes);
    }
};
int debug_module(linon_view* v, Module** m) {
    Variant var = make_variant(v, LINON_RENDER_DEFAULT);
    if (var.n_faces == 0) default_mesh_counts(var);
    return get_module_for_body(v, v->body, var, m, nullptr);
}
}  // namespace

int linon_debug_eval_field(linon_view* v, uint32_t n, const float* p_prev, const float* p, const float* v0,
                           const float* vel, const float* t, float* out) {
    if (!v || !n) return LINON_ERR_INVALID;
    CK(v, cudaSetDevice(v->device));
    Module* m = nullptr;
    int rc = debug_module(v, &m);
    if (rc != LINON_OK) return rc;
    DevBuf a, b, c, d, e, o;
    CK(v, a.up(p_prev, n * 12)); CK(v, b.up(p, n * 12)); CK(v, c.up(v0, n * 12)); CK(v, d.up(vel, n * 12));
    CK(v, e.up(t, n * 4)); CK(v, o.up(nullptr, n * 12));
    int ni = (int)n;
    linon_settings st = v->settings;
    void* args[] = {&ni, &a.p, &b.p, &c.p, &d.p, &e.p, &st, &o.p};
    const CUresult r = g_drv.LaunchKernel(m->eval_field, (n + 127) / 128, 1, 1, 128, 1, 1, 0, (CUstream)v->stream, args, nullptr);
    if (r != CUDA_SUCCESS) return v->fail(LINON_ERR_CUDA, "cuLaunchKernel(eval_field): " + g_drv.err_str(r));
    CK(v, cudaStreamSynchronize(v->stream));
    CK(v, cudaMemcpy(out, o.p, n * 12, cudaMemcpyDeviceToHost));
    return LINON_OK;
}

int linon_debug_eval_rk4(linon_view* v, uint32_t n, const float* p, const float* dir, const float* v0, const float* t,
                         const float* h, float* out) {
    if (!v || !n) return LINON_ERR_INVALID;
    CK(v, cudaSetDevice(v->device));
    Module* m = nullptr;
    int rc = debug_module(v, &m);
    if (rc != LINON_OK) return rc;
    DevBuf a, b, c, d, e, o;
    CK(v, a.up(p, n * 12)); CK(v, b.up(dir, n * 12)); CK(v, c.up(v0, n * 12)); CK(v, d.up(t, n * 4));
    CK(v, e.up(h, n * 4)); CK(v, o.up(nullptr, n * 12));
    int ni = (int)n;
    linon_settings st = v->settings;
    void* args[] = {&ni, &a.p, &b.p, &c.p, &d.p, &e.p, &st, &o.p};
    const CUresult r = g_drv.LaunchKernel(m->eval_rk4, (n + 127) / 128, 1, 1, 128, 1, 1, 0, (CUstream)v->stream, args, nullptr);
    if (r != CUDA_SUCCESS) return v->fail(LINON_ERR_CUDA, "cuLaunchKernel(eval_rk4): " + g_drv.err_str(r));
    CK(v, cudaStreamSynchronize(v->stream));
    CK(v, cudaMemcpy(out, o.p, n * 12, cudaMemcpyDeviceToHost));
    return LINON_OK;
}

int linon_debug_eval_helper(linon_view* v, int which, uint32_t n, const float* vec, const float* nrm, const float* p0,
                            const float* a_, const float* b_, const float* c_, float* out) {
    if (!v || !n) return LINON_ERR_INVALID;
    CK(v, cudaSetDevice(v->device));
    Module* m = nullptr;
    int rc = debug_module(v, &m);
    if (rc != LINON_OK) return rc;
    DevBuf dv, dn, dp, da, db, dc, o;
    CK(v, dv.up(vec, n * 12)); CK(v, dn.up(nrm, n * 12)); CK(v, dp.up(p0, n * 12));
    CK(v, da.up(a_, n * 4)); CK(v, db.up(b_, n * 4)); CK(v, dc.up(c_, n * 4)); CK(v, o.up(nullptr, n * 12));
    int ni = (int)n;
    void* args[] = {&which, &ni, &dv.p, &dn.p, &dp.p, &da.p, &db.p, &dc.p, &o.p};
    const CUresult r = g_drv.LaunchKernel(m->eval_helper, (n + 127) / 128, 1, 1, 128, 1, 1, 0, (CUstream)v->stream, args, nullptr);
    if (r != CUDA_SUCCESS) return v->fail(LINON_ERR_CUDA, "cuLaunchKernel(eval_helper): " + g_drv.err_str(r));
    CK(v, cudaStreamSynchronize(v->stream));
    CK(v, cudaMemcpy(out, o.p, n * 12, cudaMemcpyDeviceToHost));
    return LINON_OK;
}

int linon_debug_hit_triangle(linon_view* v, uint32_t n, const float* tri9, const float* origin, const float* dir,
                             float* t_out) {
    if (!v || !n) return LINON_ERR_INVALID;
    CK(v, cudaSetDevice(v->device));
    Module* m = nullptr;
    int rc = debug_module(v, &m);
    if (rc != LINON_OK) return rc;
    DevBuf dt, dorg, dd, o;
    CK(v, dt.up(tri9, n * 36)); CK(v, dorg.up(origin, n * 12)); CK(v, dd.up(dir, n * 12)); CK(v, o.up(nullptr, n * 4));
    int ni = (int)n;
    void* args[] = {&ni, &dt.p, &dorg.p, &dd.p, &o.p};
    const CUresult r = g_drv.LaunchKernel(m->eval_hit, (n + 127) / 128, 1, 1, 128, 1, 1, 0, (CUstream)v->stream, args, nullptr);
    if (r != CUDA_SUCCESS) return v->fail(LINON_ERR_CUDA, "cuLaunchKernel(eval_hit): " + g_drv.err_str(r));
    CK(v, cudaStreamSynchronize(v->stream));
    CK(v, cudaMemcpy(t_out, o.p, n * 4, cudaMemcpyDeviceToHost));
    return LINON_OK;
}

int linon_debug_match_mirage(const char* wgsl_body, int* kind, double* params8) {
    if (!wgsl_body || !kind) return LINON_ERR_INVALID;
    const linon::MirageParams m = linon::match_mirage(wgsl_body);
    *kind = m.kind;
    if (params8) {
        const double v[8] = {m.t_env, m.t_src, m.max_dist, m.c[0], m.c[1], m.c[2], m.sig_k, m.sig_c};
        memcpy(params8, v, sizeof(v));
    }
    return LINON_OK;
}

int linon_measure_fp32_peak(int device, double* tflops, int* sm_count) {
    if (!tflops) return LINON_ERR_INVALID;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) {
        g_create_error = "linon_measure_fp32_peak: no CUDA device (no CPU fallback)";
        return LINON_ERR_NO_DRIVER;
    }
    cudaSetDevice(device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    float* scratch = nullptr;
    if (cudaMalloc(&scratch, 256) != cudaSuccess) return LINON_ERR_CUDA;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        float ms = 0.0f;
        const double t = linon::run_fp32_peak(prop.multiProcessorCount, 20000, scratch, &ms);
        if (t > best) best = t;
    }
    cudaFree(scratch);
    if (cudaGetLastError() != cudaSuccess) return LINON_ERR_CUDA;
    *tflops = best;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    return LINON_OK;
}

}  // extern "C"
