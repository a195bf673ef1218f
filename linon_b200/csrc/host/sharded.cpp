// sharded.cpp — linon_sharded_*: ONE frame rendered by several B200s of a node, driven by one host thread through the C ABI
// (include/linon_b200.h).  SURVEY 8(b) "create_sharded(n_gpus) + render_sharded", 8(e); north star: "the image shards
// naturally by row tiles across the 8 B200s, with a single NCCL gather of the RGBA/endpoint tiles to GPU 0 over NVLink".
//
// The reference has no counterpart (one wgpu queue, one device: src/application.rs:67-89); what is mirrored is MainView's
// own surface (update_model / update_camera / update_settings / reload_shader / render, src/main_view.rs:460-849), applied
// to every shard.  Built on the public entry points only: one linon_view per device, rows dealt in blocks round-robin
// (linon_set_shard), and one of two ways to assemble the frame on the root device (shard 0):
//   LINON_SHARD_PEER_STORES (default)  every other shard's march stores its rows straight into the root's frame buffers
//       (linon_redirect_output; peer access enabled between the devices): the 24 bytes per ray cross NVLink while the march
//       runs, and the root's overlay is ordered behind the other shards' marches by CUDA events - no host synchronisation,
//       no copy, no collective.
//   LINON_SHARD_NCCL_GATHER  the north star's literal: every shard marches into its own buffers, then ONE grouped
//       ncclSend / ncclRecv exchange moves each non-root row block to the same rows of the root's buffers (libnccl.so.2 is
//       dlopen'ed; single-process communicators from ncclCommInitAll), then the overlay runs on the root.
// Both give the 1-GPU frame bit for bit (tests/test_gpu_sharded.py).  The overlay needs a halo of `delta` rows of endpoints,
// so it runs once on the root over the assembled frame; the sample rays run on the root only.
#include "../../../include/linon_b200.h"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

std::string g_sharded_create_error;

std::string sfmt(const char* f, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    return buf;
}

// ---- NCCL, loaded at run time (the library itself needs neither NCCL nor a GPU to load) -------------------------------
typedef struct ncclComm* ncclComm_t;
struct Nccl {
    void* lib = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
    bool load() {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (lib) break;
        }
        if (!lib) {
            error = "libnccl.so.2 not found (LINON_SHARD_NCCL_GATHER needs NCCL; LINON_SHARD_PEER_STORES does not)";
            return false;
        }
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        Send = (decltype(Send))dlsym(lib, "ncclSend");
        Recv = (decltype(Recv))dlsym(lib, "ncclRecv");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !Send || !Recv || !GetErrorString) {
            error = "libnccl.so.2 lacks a required entry point";
            dlclose(lib);
            lib = nullptr;
            return false;
        }
        return true;
    }
};
Nccl g_nccl;
constexpr int kNcclUint8 = 1;  // ncclDataType_t: ncclInt8 = 0, ncclUint8 = 1 (nccl.h)

const int kMarchBuffers[3] = {LINON_BUF_RAYCAST_RGBA8, LINON_BUF_ENDPOINT_F32X4, LINON_BUF_HIT_ID_I32};
const size_t kBytesPerPixel[3] = {4, 16, 4};

}  // namespace

struct linon_sharded {
    int mode = LINON_SHARD_PEER_STORES;
    uint32_t W = 0, H = 0, rows_per_block = 32;
    std::vector<int> devices;
    std::vector<linon_view*> views;
    std::vector<cudaEvent_t> march_done;  // per shard, recorded on the shard's stream after its march
    cudaEvent_t root_done = nullptr;      // recorded on the root's stream after it has consumed the assembled frame (overlay)
    bool root_done_valid = false;
    std::vector<ncclComm_t> comms;        // LINON_SHARD_NCCL_GATHER
    std::string error;

    int fail(int code, const std::string& msg) {
        error = msg;
        return code;
    }
    int fail_view(int rc, uint32_t rank) { return fail(rc, sfmt("shard %u: %s", rank, linon_last_error(views[rank]))); }
};

namespace {

#define SCK(s, call)                                                                                    \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return (s)->fail(LINON_ERR_CUDA, sfmt("%s failed: %s", #call, cudaGetErrorString(e__)));    \
    } while (0)

// (re)points every non-root shard's march outputs: at the root's buffers (peer stores) or at its own (gather)
int wire_outputs(linon_sharded* s) {
    for (uint32_t r = 1; r < s->views.size(); ++r)
        for (int b : kMarchBuffers) {
            void* target = s->mode == LINON_SHARD_PEER_STORES ? linon_device_ptr(s->views[0], b) : nullptr;
            const int rc = linon_redirect_output(s->views[r], b, target);
            if (rc != LINON_OK) return s->fail_view(rc, r);
        }
    return LINON_OK;
}

int enable_peer_access(linon_sharded* s) {
    const int root = s->devices[0];
    for (uint32_t r = 1; r < s->devices.size(); ++r) {
        const int dev = s->devices[r];
        if (dev == root) continue;  // shards on one device (tests on a single-GPU box): same address space
        int can = 0;
        SCK(s, cudaDeviceCanAccessPeer(&can, dev, root));
        if (!can) return s->fail(LINON_ERR_CUDA, sfmt("device %d cannot access device %d's memory (no NVLink / P2P): use LINON_SHARD_NCCL_GATHER", dev, root));
        SCK(s, cudaSetDevice(dev));
        const cudaError_t e = cudaDeviceEnablePeerAccess(root, 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) {
            cudaGetLastError();
        } else if (e != cudaSuccess) {
            return s->fail(LINON_ERR_CUDA, sfmt("cudaDeviceEnablePeerAccess(%d -> %d): %s", dev, root, cudaGetErrorString(e)));
        }
    }
    return LINON_OK;
}

// one grouped exchange: every non-root row block -> the same rows of the root's buffers
int nccl_gather(linon_sharded* s) {
    const uint32_t world = (uint32_t)s->views.size();
    int rc = g_nccl.GroupStart();
    if (rc != 0) return s->fail(LINON_ERR_CUDA, sfmt("ncclGroupStart: %s", g_nccl.GetErrorString(rc)));
    for (uint32_t b = 0, y = 0; y < s->H; ++b, y += s->rows_per_block) {
        const uint32_t src = b % world;
        if (src == 0) continue;
        const uint32_t rows = (y + s->rows_per_block <= s->H) ? s->rows_per_block : (s->H - y);
        for (int k = 0; k < 3; ++k) {
            const size_t off = (size_t)y * s->W * kBytesPerPixel[k], n = (size_t)rows * s->W * kBytesPerPixel[k];
            const char* from = (const char*)linon_device_ptr(s->views[src], kMarchBuffers[k]) + off;
            char* to = (char*)linon_device_ptr(s->views[0], kMarchBuffers[k]) + off;
            rc = g_nccl.Send(from, n, kNcclUint8, 0, s->comms[src], (cudaStream_t)linon_stream(s->views[src]));
            if (rc == 0) rc = g_nccl.Recv(to, n, kNcclUint8, (int)src, s->comms[0], (cudaStream_t)linon_stream(s->views[0]));
            if (rc != 0) {
                g_nccl.GroupEnd();
                return s->fail(LINON_ERR_CUDA, sfmt("ncclSend/ncclRecv: %s", g_nccl.GetErrorString(rc)));
            }
        }
    }
    rc = g_nccl.GroupEnd();
    if (rc != 0) return s->fail(LINON_ERR_CUDA, sfmt("ncclGroupEnd: %s", g_nccl.GetErrorString(rc)));
    return LINON_OK;
}

}  // namespace

extern "C" {

const char* linon_sharded_last_error(const linon_sharded* s) { return s ? s->error.c_str() : g_sharded_create_error.c_str(); }

int linon_sharded_create(const int* devices, uint32_t n_devices, uint32_t width, uint32_t height, uint32_t rows_per_block,
                         int mode, linon_sharded** out) {
    if (!out || !devices || n_devices == 0 || n_devices > 64 || rows_per_block == 0 ||
        (mode != LINON_SHARD_PEER_STORES && mode != LINON_SHARD_NCCL_GATHER)) {
        g_sharded_create_error = "linon_sharded_create: need 1..64 devices, rows_per_block > 0 and a valid mode";
        return LINON_ERR_INVALID;
    }
    *out = nullptr;
    linon_sharded* s = new linon_sharded;
    s->mode = mode;
    s->W = width; s->H = height; s->rows_per_block = rows_per_block;
    s->devices.assign(devices, devices + n_devices);
    auto bail = [&](int rc, const std::string& msg) {
        g_sharded_create_error = msg;
        linon_sharded_destroy(s);
        return rc;
    };
    for (uint32_t r = 0; r < n_devices; ++r) {
        linon_view* v = nullptr;
        const int rc = linon_create(devices[r], width, height, &v);
        if (rc != LINON_OK) return bail(rc, sfmt("shard %u (device %d): %s", r, devices[r], linon_last_error(nullptr)));
        s->views.push_back(v);
        const int rs = linon_set_shard(v, r, n_devices, rows_per_block);
        if (rs != LINON_OK) return bail(rs, sfmt("shard %u: %s", r, linon_last_error(v)));
        cudaEvent_t ev = nullptr;
        if (cudaSetDevice(devices[r]) != cudaSuccess || cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess)
            return bail(LINON_ERR_CUDA, "cudaEventCreate failed");
        s->march_done.push_back(ev);
    }
    if (cudaSetDevice(devices[0]) != cudaSuccess || cudaEventCreateWithFlags(&s->root_done, cudaEventDisableTiming) != cudaSuccess)
        return bail(LINON_ERR_CUDA, "cudaEventCreate failed");
    if (mode == LINON_SHARD_PEER_STORES) {
        int rc = enable_peer_access(s);
        if (rc == LINON_OK) rc = wire_outputs(s);
        if (rc != LINON_OK) return bail(rc, s->error);
    } else if (n_devices > 1) {
        if (!g_nccl.load()) return bail(LINON_ERR_NO_DRIVER, g_nccl.error);
        s->comms.assign(n_devices, nullptr);
        const int rc = g_nccl.CommInitAll(s->comms.data(), (int)n_devices, devices);
        if (rc != 0) {
            s->comms.clear();
            return bail(LINON_ERR_CUDA, sfmt("ncclCommInitAll: %s (one communicator per DISTINCT device)", g_nccl.GetErrorString(rc)));
        }
    }
    *out = s;
    return LINON_OK;
}

void linon_sharded_destroy(linon_sharded* s) {
    if (!s) return;
    for (linon_view* v : s->views) linon_sync(v);
    for (ncclComm_t c : s->comms)
        if (c) g_nccl.CommDestroy(c);
    for (size_t r = 0; r < s->march_done.size(); ++r) {
        cudaSetDevice(s->devices[r]);
        cudaEventDestroy(s->march_done[r]);
    }
    if (s->root_done) {
        cudaSetDevice(s->devices[0]);
        cudaEventDestroy(s->root_done);
    }
    // non-root shards first: they hold pointers into the root's buffers
    for (size_t r = s->views.size(); r-- > 0;) linon_destroy(s->views[r]);
    delete s;
}

uint32_t linon_sharded_world(const linon_sharded* s) { return s ? (uint32_t)s->views.size() : 0; }

linon_view* linon_sharded_view(linon_sharded* s, uint32_t rank) { return (s && rank < s->views.size()) ? s->views[rank] : nullptr; }

#define FOR_EACH_SHARD(s, call)                                    \
    do {                                                           \
        if (!(s)) return LINON_ERR_INVALID;                        \
        for (uint32_t r__ = 0; r__ < (s)->views.size(); ++r__) {   \
            linon_view* v = (s)->views[r__];                       \
            const int rc__ = (call);                               \
            if (rc__ != LINON_OK) return (s)->fail_view(rc__, r__);\
        }                                                          \
        return LINON_OK;                                           \
    } while (0)

int linon_sharded_set_mesh(linon_sharded* s, const float* verts, uint32_t n_verts, const uint32_t* faces, uint32_t n_faces) {
    FOR_EACH_SHARD(s, linon_set_mesh(v, verts, n_verts, faces, n_faces));
}
int linon_sharded_set_camera(linon_sharded* s, const linon_camera* camera) { FOR_EACH_SHARD(s, linon_set_camera(v, camera)); }
int linon_sharded_set_settings(linon_sharded* s, const linon_settings* settings) { FOR_EACH_SHARD(s, linon_set_settings(v, settings)); }
int linon_sharded_set_option(linon_sharded* s, int option, int value) { FOR_EACH_SHARD(s, linon_set_option(v, option, value)); }

int linon_sharded_set_field_function(linon_sharded* s, const char* wgsl_body, char* err, size_t err_len) {
    if (!s) return LINON_ERR_INVALID;
    // validate on the root first: a body that does not compile changes no shard (the previous kernel stays live everywhere,
    // src/main_view.rs:677-679)
    int rc = linon_set_field_function(s->views[0], wgsl_body, err, err_len);
    if (rc != LINON_OK) return s->fail_view(rc, 0);
    for (uint32_t r = 1; r < s->views.size(); ++r) {
        rc = linon_set_field_function(s->views[r], wgsl_body, err, err_len);
        if (rc != LINON_OK) return s->fail_view(rc, r);
    }
    return LINON_OK;
}

int linon_sharded_resize(linon_sharded* s, uint32_t width, uint32_t height) {
    if (!s) return LINON_ERR_INVALID;
    for (uint32_t r = 0; r < s->views.size(); ++r) {
        const int rc = linon_resize(s->views[r], width, height);  // drops every redirection of the view
        if (rc != LINON_OK) return s->fail_view(rc, r);
    }
    s->W = width; s->H = height;
    return s->mode == LINON_SHARD_PEER_STORES ? wire_outputs(s) : LINON_OK;
}

int linon_sharded_render(linon_sharded* s, uint32_t flags) {
    if (!s) return LINON_ERR_INVALID;
    const uint32_t world = (uint32_t)s->views.size();
    if (world == 1) {
        const int rc = linon_render(s->views[0], flags);
        return rc == LINON_OK ? rc : s->fail_view(rc, 0);
    }
    const bool do_march = !(flags & LINON_RENDER_NO_MARCH);
    if (do_march) {
        // every shard marches its rows (asynchronously, each on its own device and stream); the sample rays run on the root.
        // A peer store of THIS frame must not land while the root's overlay of the PREVIOUS frame still reads the buffers.
        for (uint32_t r = 1; r < world && s->root_done_valid; ++r) {
            SCK(s, cudaSetDevice(s->devices[r]));
            SCK(s, cudaStreamWaitEvent((cudaStream_t)linon_stream(s->views[r]), s->root_done, 0));
        }
        for (uint32_t r = 0; r < world; ++r) {
            uint32_t f = (flags | LINON_RENDER_NO_OVERLAY);
            if (r != 0) f |= LINON_RENDER_NO_SAMPLE_RAYS;
            const int rc = linon_render(s->views[r], f);
            if (rc != LINON_OK) return s->fail_view(rc, r);
        }
        if (s->mode == LINON_SHARD_NCCL_GATHER) {
            const int rc = nccl_gather(s);  // stream-ordered behind each shard's march; the root's stream receives
            if (rc != LINON_OK) return rc;
        } else {
            // the root's next work waits for the other shards' marches: their stores ARE the assembled frame
            cudaStream_t root_stream = (cudaStream_t)linon_stream(s->views[0]);
            for (uint32_t r = 1; r < world; ++r) {
                SCK(s, cudaSetDevice(s->devices[r]));
                SCK(s, cudaEventRecord(s->march_done[r], (cudaStream_t)linon_stream(s->views[r])));
                SCK(s, cudaSetDevice(s->devices[0]));
                SCK(s, cudaStreamWaitEvent(root_stream, s->march_done[r], 0));
            }
        }
    }
    if (!(flags & LINON_RENDER_NO_OVERLAY)) {
        // KEEP_TIMINGS: the root's linon_get_timings then reports its march AND the overlay of this frame
        const uint32_t f = (flags | LINON_RENDER_NO_MARCH | LINON_RENDER_NO_SAMPLE_RAYS | (do_march ? LINON_RENDER_KEEP_TIMINGS : 0u)) &
                           ~(uint32_t)LINON_RENDER_OUTLINE;
        const int rc = linon_render(s->views[0], f);
        if (rc != LINON_OK) return s->fail_view(rc, 0);
    }
    SCK(s, cudaSetDevice(s->devices[0]));
    SCK(s, cudaEventRecord(s->root_done, (cudaStream_t)linon_stream(s->views[0])));
    s->root_done_valid = true;
    return LINON_OK;
}

int linon_sharded_sync(linon_sharded* s) { FOR_EACH_SHARD(s, linon_sync(v)); }

}  // extern "C"
