// sqt_backend.cu -- context, scene upload and the C ABI (include/sqt.h) of the squigly-trace B200 backend.
//
// The kernels live in sqt_kernels.cuh (per-ray logic in sqt_core.cuh / sqt_paths.cuh), the NCCL shim in sqt_nccl.hpp,
// the host-side derivation of the device records in sqt_layout.hpp.  Every launch is bracketed by CUDA events on the
// context's stream.
//
// There is no CPU fallback in this file: every entry point needs a compute-capability-10.x device.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <atomic>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sqt.h"
#include "sqt_kernels.cuh"
#include "sqt_layout.hpp"
#include "sqt_nccl.hpp"

// ============================================================================== context
struct sqt_ctx {
    int device = 0, sm_count = 0, cc_major = 0, cc_minor = 0;
    char name[128] = {0};
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    std::string err;
    // scene
    bool has_scene = false;
    SceneView sc = {};
    float4 *d_slabs = nullptr, *d_tight = nullptr; uint32_t *d_levels = nullptr;          // subtree slabs (+ upload scratch)
    float4 *d_nodes = nullptr, *d_boxes = nullptr, *d_tris = nullptr, *d_mats = nullptr, *d_leaves = nullptr, *d_spheres = nullptr;
    float4 *d_sph_nodes = nullptr; uint32_t *d_sph_order = nullptr; int sphere_bvh = 1;      // extension: hierarchy over the spheres
    uint2 *d_ranges = nullptr;          // (first, count) per leaf, input of k_leaf_records
    uint32_t *d_flag = nullptr;         // material check result
    size_t cap_branches = 0, cap_leaves = 0, cap_tris = 0, cap_mats = 0;     // allocation sizes (records), reused across uploads
    uint64_t upload_bytes = 0; double upload_ms = 0, upload_layout_ms = 0;   // last sqt_upload_scene
    int leaf_cull = 1;
    int terminate_on_black_ok = 0;
    uint32_t tree_height = 0;
    // image buffers
    long long cap_pixels = 0;
    int2 *d_prim = nullptr;
    int *d_pixel_list = nullptr;
    float *d_sbuf = nullptr; long long cap_sbuf = 0;     // sample buffer of one round (bytes)
    long long sbuf_budget = 2ll << 30;
    float *d_accum = nullptr;
    uint8_t *d_rgb8 = nullptr;
    long long img_pixels = 0;
    int img_spp = 0;
    DeviceStats *d_stats = nullptr;
    DeviceStats *h_stats = nullptr;     // pinned
    // batch staging
    long long cap_rays = 0;
    float *d_org = nullptr, *d_dir = nullptr, *d_dist = nullptr, *d_point = nullptr;
    int *d_tri = nullptr;
    // pinned host staging for image I/O
    uint8_t *h_rgb8 = nullptr; float *h_accum = nullptr; long long cap_host_pixels = 0;
    Tune tune = {12, 0, 12};
    int pool_k = 2;                     // 0: one ray per lane (k_paths) ; K > 0: ray pools of 32*K rays per warp (k_paths_pool)
    PoolTune pool_tune = {4, 10, 16};
    int pool_blocks = 0;                // cap on resident CTAs per SM for k_paths_pool (0 = occupancy limit); fewer CTAs leave more L1
    float slab_ratio = kSlabRatioMax, slab_ratio_leaf = kSlabRatioMaxLeaf;   // subtree / leaf slabs: test a slab when it is at most this
                                                                             // fraction of the clipped box (SQT_SLAB_RATIO=branch[,leaf], 0 = never)
    int pool_carveout = 0;              // shared-memory carve-out of k_paths_pool in percent of 228 KB (0 = the driver's choice)
    float4 *d_gstack = nullptr; uint16_t *d_gpm = nullptr; uint4 *d_gpath = nullptr; long long cap_pool_slots = 0, cap_stack_entries = 0;
    bool comm_broken = false;           // the communicator was aborted after a rank failed
    // group
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
};

static thread_local std::string g_create_err;

static int fail(sqt_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_err = buf;
    return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, SQT_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

extern "C" int sqt_abi_version(void) { return SQT_ABI_VERSION; }

extern "C" const char *sqt_last_error(const sqt_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Environment knobs (development only; they change scheduling, never a result).  Malformed values are ignored and every
// value is clamped to a range in which the persistent kernels are guaranteed to make progress.
static void read_env_tuning(sqt_ctx *c) {
    if (const char *t = getenv("SQT_POOL")) { char *e = nullptr; long k = strtol(t, &e, 10); if (e != t && k >= 0 && k <= 4) c->pool_k = k == 3 ? 4 : (int)k; }      // rays per warp = 32 * K, K a power of two
    if (const char *t = getenv("SQT_POOL_BLOCKS")) { char *e = nullptr; long k = strtol(t, &e, 10); if (e != t && k >= 0 && k <= 32) c->pool_blocks = (int)k; }
    if (const char *t = getenv("SQT_SLAB_RATIO")) {
        char *e = nullptr; double v = strtod(t, &e);
        if (e != t && v >= 0.0 && v <= 2.0) {
            c->slab_ratio = (float)v;
            if (*e == ',') { const char *u = e + 1; double w = strtod(u, &e); if (e != u && w >= 0.0 && w <= 2.0) c->slab_ratio_leaf = (float)w; }
        }
    }
    if (const char *t = getenv("SQT_POOL_CARVEOUT")) { char *e = nullptr; long k = strtol(t, &e, 10); if (e != t && k >= 0 && k <= 100) c->pool_carveout = (int)k; }
    if (const char *t = getenv("SQT_POOL_TUNE")) {       // "burst_t,t_leave,c_min"
        int a, b, cm;
        if (sscanf(t, "%d,%d,%d", &a, &b, &cm) == 3) c->pool_tune = {clampi(a, 1, 64), clampi(b, 0, 31), clampi(cm, 1, 32)};
    }
    if (const char *t = getenv("SQT_SBUF_MB")) { char *e = nullptr; long long mb = strtoll(t, &e, 10); if (e != t && mb > 0 && mb < (1ll << 20)) c->sbuf_budget = mb << 20; }
    if (const char *t = getenv("SQT_TUNE")) {            // "a_leave,b_leave,c_min"
        int a, b, cm;
        if (sscanf(t, "%d,%d,%d", &a, &b, &cm) == 3) c->tune = {clampi(a, 0, 31), clampi(b, 0, 32), clampi(cm, 1, 32)};
    }
}

extern "C" int sqt_destroy(sqt_ctx *c);

extern "C" int sqt_create(int device, sqt_ctx **out) {
    if (!out) return fail(nullptr, SQT_E_INVALID, "sqt_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, SQT_E_NO_DEVICE, "no CUDA device (%s); this backend has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, SQT_E_INVALID, "device %d out of range (have %d)", device, n);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, SQT_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, SQT_E_NO_DEVICE, "device %d (%s) is compute capability %d.%d; this library is built for sm_100a only",
                    device, prop.name, prop.major, prop.minor);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, SQT_E_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    sqt_ctx *c = new sqt_ctx();
    c->device = device; c->sm_count = prop.multiProcessorCount; c->cc_major = prop.major; c->cc_minor = prop.minor;
    snprintf(c->name, sizeof c->name, "%s", prop.name);
    auto init = [&]() -> int {
        sqt_ctx *ctx = c;
        CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        for (auto &ev : c->ev) CU(cudaEventCreate(&ev));
        CU(cudaMalloc(&c->d_stats, sizeof(DeviceStats)));
        CU(cudaMalloc(&c->d_flag, 2 * sizeof(uint32_t)));
        CU(cudaMallocHost(&c->h_stats, sizeof(DeviceStats)));
        return SQT_OK;
    };
    const int rc = init();
    if (rc) { g_create_err = c->err; sqt_destroy(c); return rc; }        // no half-built context is leaked
    read_env_tuning(c);
    *out = c;
    return SQT_OK;
}

static void free_scene(sqt_ctx *c) {
    cudaFree(c->d_slabs); cudaFree(c->d_tight); cudaFree(c->d_levels); c->d_slabs = c->d_tight = nullptr; c->d_levels = nullptr;
    cudaFree(c->d_nodes); cudaFree(c->d_boxes); cudaFree(c->d_tris); cudaFree(c->d_mats); cudaFree(c->d_leaves); cudaFree(c->d_spheres); cudaFree(c->d_sph_nodes); cudaFree(c->d_sph_order); cudaFree(c->d_ranges);
    c->d_nodes = c->d_boxes = c->d_tris = c->d_mats = c->d_leaves = c->d_spheres = c->d_sph_nodes = nullptr; c->d_sph_order = nullptr; c->d_ranges = nullptr; c->has_scene = false;
    c->cap_branches = c->cap_leaves = c->cap_tris = c->cap_mats = 0;
}

extern "C" int sqt_destroy(sqt_ctx *c) {
    if (!c) return SQT_OK;
    cudaSetDevice(c->device);
    if (c->comm && nccl_api()->lib && !c->comm_broken) nccl_api()->CommDestroy(c->comm);
    free_scene(c);
    cudaFree(c->d_prim); cudaFree(c->d_accum); cudaFree(c->d_rgb8); cudaFree(c->d_stats); cudaFree(c->d_flag); cudaFree(c->d_pixel_list); cudaFree(c->d_sbuf); cudaFree(c->d_gstack); cudaFree(c->d_gpm); cudaFree(c->d_gpath);
    cudaFree(c->d_org); cudaFree(c->d_dir); cudaFree(c->d_dist); cudaFree(c->d_point); cudaFree(c->d_tri);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    if (c->h_rgb8) cudaFreeHost(c->h_rgb8);
    if (c->h_accum) cudaFreeHost(c->h_accum);
    for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SQT_OK;
}

extern "C" int sqt_device_info(sqt_ctx *c, int *sm_count, int *cc_major, int *cc_minor, char name_out[128]) {
    if (!c) return SQT_E_INVALID;
    if (sm_count) *sm_count = c->sm_count;
    if (cc_major) *cc_major = c->cc_major;
    if (cc_minor) *cc_minor = c->cc_minor;
    if (name_out) snprintf(name_out, 128, "%s", c->name);
    return SQT_OK;
}

// ------------------------------------------------------------------------------ scene upload
// Host: one iterative walk over the boundary tree (sqt_layout.hpp) validates it and derives the 16-byte device nodes and
// the clipped boxes (BIH.hs:130-141; plane values are copied, never computed).  Meanwhile a helper thread streams the
// triangle array -- the bulk of a large scene -- to every context of the call through two pinned staging buffers.
// Device: k_leaf_records derives the leaf records from the triangles, k_check_materials validates the material indices.
// Device allocations are kept and reused while the new scene fits them.
static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int ensure_scene_capacity(sqt_ctx *ctx, size_t n_br, size_t n_lf, size_t n_tris, size_t n_mats) {
    if (n_br > ctx->cap_branches) {
        cudaFree(ctx->d_nodes); cudaFree(ctx->d_boxes); cudaFree(ctx->d_slabs); cudaFree(ctx->d_tight); cudaFree(ctx->d_levels);
        ctx->d_nodes = ctx->d_boxes = ctx->d_slabs = ctx->d_tight = nullptr; ctx->d_levels = nullptr; ctx->cap_branches = 0;
        CU(cudaMalloc(&ctx->d_nodes, n_br * sizeof(float4))); CU(cudaMalloc(&ctx->d_boxes, 2 * n_br * sizeof(float4)));
        CU(cudaMalloc(&ctx->d_slabs, n_br * sizeof(float4))); CU(cudaMalloc(&ctx->d_tight, 2 * n_br * sizeof(float4)));
        CU(cudaMalloc(&ctx->d_levels, n_br * sizeof(uint32_t)));
        ctx->cap_branches = n_br;
    }
    if (n_lf > ctx->cap_leaves) {
        cudaFree(ctx->d_leaves); cudaFree(ctx->d_ranges); ctx->d_leaves = nullptr; ctx->d_ranges = nullptr; ctx->cap_leaves = 0;
        CU(cudaMalloc(&ctx->d_leaves, 2 * n_lf * sizeof(float4))); CU(cudaMalloc(&ctx->d_ranges, n_lf * sizeof(uint2)));
        ctx->cap_leaves = n_lf;
    }
    if (n_tris > ctx->cap_tris) {
        cudaFree(ctx->d_tris); ctx->d_tris = nullptr; ctx->cap_tris = 0;
        CU(cudaMalloc(&ctx->d_tris, n_tris * 48));
        ctx->cap_tris = n_tris;
    }
    if (n_mats > ctx->cap_mats) {
        cudaFree(ctx->d_mats); ctx->d_mats = nullptr; ctx->cap_mats = 0;
        CU(cudaMalloc(&ctx->d_mats, 3 * n_mats * sizeof(float4)));
        ctx->cap_mats = n_mats;
    }
    return SQT_OK;
}

// triangles -> every context, chunk by chunk through two pinned buffers (runs on a helper thread)
static void stream_tris(sqt_ctx **ctxs, int n, const sqt_tri *tris, size_t n_tris, std::atomic<int> *status, std::string *err) {
    const size_t CH = (size_t)32 << 20, bytes = n_tris * 48;
    void *pin[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> done(2 * (size_t)n);
    auto bail = [&](cudaError_t e, const char *what) { *err = std::string(what) + ": " + cudaGetErrorString(e); status->store(SQT_E_CUDA); };
    cudaError_t e = cudaSuccess;
    bool have_ev = false;
    do {
        if (bytes == 0) break;
        const size_t buf = bytes < CH ? bytes : CH;
        if ((e = cudaSetDevice(ctxs[0]->device)) != cudaSuccess) { bail(e, "cudaSetDevice"); break; }
        if ((e = cudaMallocHost(&pin[0], buf)) != cudaSuccess || (bytes > CH && (e = cudaMallocHost(&pin[1], buf)) != cudaSuccess)) { bail(e, "cudaMallocHost (staging)"); break; }
        for (int g = 0; g < n && e == cudaSuccess; ++g) { cudaSetDevice(ctxs[g]->device); for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaEventCreateWithFlags(&done[2 * g + b], cudaEventDisableTiming); }
        if (e != cudaSuccess) { bail(e, "cudaEventCreate"); break; }
        have_ev = true;
        size_t off = 0;
        for (int it = 0; off < bytes; ++it) {
            const int b = it & 1;
            const size_t len = bytes - off < CH ? bytes - off : CH;
            if (it >= 2) for (int g = 0; g < n; ++g) { cudaSetDevice(ctxs[g]->device); if ((e = cudaEventSynchronize(done[2 * g + b])) != cudaSuccess) break; }
            if (e != cudaSuccess) { bail(e, "cudaEventSynchronize"); break; }
            memcpy(pin[b], (const char *)tris + off, len);
            for (int g = 0; g < n; ++g) {
                cudaSetDevice(ctxs[g]->device);
                if ((e = cudaMemcpyAsync((char *)ctxs[g]->d_tris + off, pin[b], len, cudaMemcpyHostToDevice, ctxs[g]->stream)) != cudaSuccess) break;
                if ((e = cudaEventRecord(done[2 * g + b], ctxs[g]->stream)) != cudaSuccess) break;
            }
            if (e != cudaSuccess) { bail(e, "cudaMemcpyAsync (triangles)"); break; }
            off += len;
        }
        for (int g = 0; g < n; ++g) { cudaSetDevice(ctxs[g]->device); cudaStreamSynchronize(ctxs[g]->stream); }
    } while (false);
    if (have_ev) for (int g = 0; g < n; ++g) { cudaSetDevice(ctxs[g]->device); for (int b = 0; b < 2; ++b) cudaEventDestroy(done[2 * g + b]); }
    for (void *p : pin) if (p) cudaFreeHost(p);
}

static int upload_scene_to(sqt_ctx **ctxs, int n, const sqt_scene_desc *s) {
    sqt_ctx *ctx = ctxs[0];
    const double t0 = now_ms();
    if (!s->nodes || s->n_nodes == 0) return fail(ctx, SQT_E_INVALID, "scene has no BIH nodes");
    if (s->n_tris && !s->tris) return fail(ctx, SQT_E_INVALID, "tris is NULL");
    if (s->n_tris >= (1u << 27) || s->n_nodes >= (1u << 27)) return fail(ctx, SQT_E_UNSUPPORTED, "scene too large for the 27-bit indices");
    uint32_t n_br = 0, n_lf = 0;
    for (uint32_t i = 0; i < s->n_nodes; ++i) { if (s->nodes[i].b & SQT_NODE_LEAF) ++n_lf; else ++n_br; }
    for (int g = 0; g < n; ++g) {
        sqt_ctx *c = ctxs[g];
        cudaSetDevice(c->device);
        c->has_scene = false;
        const int rc = ensure_scene_capacity(c, n_br ? n_br : 1, n_lf ? n_lf : 1, s->n_tris ? s->n_tris : 1, s->n_mats ? s->n_mats : 1);
        if (rc) { if (g) ctx->err = c->err; return rc; }
    }
    // the triangles start to flow while the host walks the tree
    std::atomic<int> tri_status{0};
    std::string tri_err;
    std::thread streamer(stream_tris, ctxs, n, s->tris, (size_t)s->n_tris, &tri_status, &tri_err);
    DeviceLayout lay;
    std::string lerr;
    const int lrc = build_device_layout(*s, lay, lerr, /*check_materials=*/false);
    const double t1 = now_ms();
    streamer.join();
    if (lrc) return fail(ctx, lrc, "%s", lerr.c_str());
    if (tri_status.load()) return fail(ctx, tri_status.load(), "%s", tri_err.c_str());
    std::vector<uint2> ranges(lay.leaf_first.size());
    for (size_t k = 0; k < ranges.size(); ++k) ranges[k] = make_uint2(lay.leaf_first[k], lay.leaf_count[k]);
    // branches sorted by depth (counting sort): the bottom-up pass of the subtree slabs runs one launch per level
    std::vector<uint32_t> level_off(lay.height + 2, 0u), level_nodes(lay.n_branches ? lay.n_branches : 1);
    for (uint32_t b = 0; b < lay.n_branches; ++b) level_off[lay.branch_depth[b] + 1]++;
    for (size_t d = 1; d < level_off.size(); ++d) level_off[d] += level_off[d - 1];
    {
        std::vector<uint32_t> fill(level_off.begin(), level_off.end() - 1);
        for (uint32_t b = 0; b < lay.n_branches; ++b) level_nodes[fill[lay.branch_depth[b]]++] = b;
    }
    const uint32_t n_leaves = n_lf;
    for (int g = 0; g < n; ++g) {
        sqt_ctx *c = ctxs[g];
        sqt_ctx *ctx = c;       // CU() reports into this context
        CU(cudaSetDevice(c->device));
        cudaStream_t st = c->stream;
        const uint32_t init_flag[2] = {0xffffffffu, 0u};
        CU(cudaMemcpyAsync(c->d_flag, init_flag, sizeof init_flag, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(c->d_nodes, lay.nodes.data(), lay.nodes.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(c->d_boxes, lay.boxes.data(), lay.boxes.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(c->d_mats, lay.mats.data(), lay.mats.size() * sizeof(float4), cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(c->d_ranges, ranges.data(), ranges.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
        if (n_leaves) {
            k_leaf_records<<<(n_leaves + 255) / 256, 256, 0, st>>>(c->d_tris, c->d_ranges, n_leaves, c->d_leaves);
            CU(cudaGetLastError());
        }
        if (s->n_tris) {
            k_check_materials<<<(s->n_tris + 255) / 256, 256, 0, st>>>(c->d_tris, s->n_tris, s->n_mats, c->d_flag);
            CU(cudaGetLastError());
        }
        if (lay.n_branches) {
            CU(cudaMemcpyAsync(c->d_levels, level_nodes.data(), (size_t)lay.n_branches * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
            for (uint32_t d = lay.height; d >= 1; --d) {                      // deepest level first
                const uint32_t lo = level_off[d], cnt = level_off[d + 1] - lo;
                if (!cnt) continue;
                k_branch_tight<<<(cnt + 255) / 256, 256, 0, st>>>(c->d_levels + lo, cnt, c->d_nodes, c->d_leaves, c->d_tight);
                CU(cudaGetLastError());
            }
            k_child_slabs<<<(lay.n_branches + 255) / 256, 256, 0, st>>>(c->d_nodes, lay.n_branches, c->d_boxes, c->d_tight, c->d_leaves, lay.s_max, lay.c_max, c->slab_ratio, c->slab_ratio_leaf, c->d_slabs);
            k_flag_slabs<<<(lay.n_branches + 255) / 256, 256, 0, st>>>(c->d_nodes, lay.n_branches, c->d_slabs);
            CU(cudaGetLastError());
        }
    }
    for (int g = 0; g < n; ++g) {
        sqt_ctx *c = ctxs[g];
        sqt_ctx *ctx = c;
        CU(cudaSetDevice(c->device));
        uint32_t flag[2] = {0, 0};
        CU(cudaMemcpyAsync(flag, c->d_flag, sizeof flag, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        if (flag[0] != 0xffffffffu) return fail(ctxs[0], SQT_E_INVALID, "triangle %u: material %u out of range", flag[0], s->tris[flag[0]].material);
        SceneView v = {};
        v.nodes = c->d_nodes; v.boxes = c->d_boxes; v.tris = c->d_tris; v.mats = c->d_mats; v.leaves = c->d_leaves; v.slabs = c->d_slabs;
        for (int k = 0; k < 3; ++k) v.tame_c[k] = lay.tame_c[k];
        v.tame_r = lay.tame_r;
        v.leaf_cull = (uint32_t)c->leaf_cull;
        for (int k = 0; k < 3; ++k) { v.root_lo[k] = s->root_bounds[k]; v.root_hi[k] = s->root_bounds[3 + k]; }
        v.n_branches = lay.n_branches; v.n_tris = s->n_tris; v.n_mats = s->n_mats;
        v.root_is_leaf = (s->nodes[0].b & SQT_NODE_LEAF) ? 1u : 0u;
        v.planes_finite = (uint32_t)lay.planes_finite;
        c->sc = v; c->has_scene = true; c->terminate_on_black_ok = lay.terminate_on_black_ok; c->tree_height = lay.height;
        c->upload_bytes = (uint64_t)s->n_tris * 48 + lay.nodes.size() * 16 + lay.boxes.size() * 16 + lay.mats.size() * 16 + ranges.size() * 8 + (uint64_t)lay.n_branches * 4 + 8;
        c->upload_layout_ms = t1 - t0;
        c->upload_ms = now_ms() - t0;
    }
    return SQT_OK;
}

extern "C" int sqt_upload_scene(sqt_ctx *ctx, const sqt_scene_desc *s) {
    if (!ctx || !s) return SQT_E_INVALID;
    return upload_scene_to(&ctx, 1, s);
}

// the same scene onto every context of a single-process group: the tree is walked once, every staged chunk of
// triangles goes to all devices
extern "C" int sqt_upload_scene_group(sqt_ctx **ctxs, int n, const sqt_scene_desc *s) {
    if (!ctxs || n < 1 || !s) return SQT_E_INVALID;
    for (int i = 0; i < n; ++i) if (!ctxs[i]) return SQT_E_INVALID;
    return upload_scene_to(ctxs, n, s);
}

extern "C" int sqt_last_upload(const sqt_ctx *ctx, uint64_t *h2d_bytes, double *wall_ms, double *host_layout_ms) {
    if (!ctx) return SQT_E_INVALID;
    if (h2d_bytes) *h2d_bytes = ctx->upload_bytes;
    if (wall_ms) *wall_ms = ctx->upload_ms;
    if (host_layout_ms) *host_layout_ms = ctx->upload_layout_ms;
    return SQT_OK;
}

// extension: analytic spheres (include/sqt.h)
extern "C" int sqt_upload_spheres(sqt_ctx *ctx, const sqt_sphere *sp, uint32_t n) {
    if (!ctx) return SQT_E_INVALID;
    if (!ctx->has_scene) return fail(ctx, SQT_E_NO_SCENE, "sqt_upload_spheres before sqt_upload_scene");
    if (n && !sp) return fail(ctx, SQT_E_INVALID, "spheres is NULL");
    if ((uint64_t)ctx->sc.n_tris + n >= (1u << 27)) return fail(ctx, SQT_E_UNSUPPORTED, "too many surfaces");
    CU(cudaSetDevice(ctx->device));
    std::vector<float4> ds((size_t)2 * (n ? n : 1));
    for (uint32_t k = 0; k < n; ++k) {
        if (sp[k].material >= ctx->sc.n_mats) return fail(ctx, SQT_E_INVALID, "sphere %u: material %u out of range", k, sp[k].material);
        if (!(sp[k].radius > 0.0f)) return fail(ctx, SQT_E_INVALID, "sphere %u: radius must be positive", k);
        ds[2 * k] = make_float4(sp[k].center[0], sp[k].center[1], sp[k].center[2], sp[k].radius);
        ds[2 * k + 1] = make_float4(u2f(sp[k].material), 0.0f, 0.0f, 0.0f);
    }
    cudaFree(ctx->d_spheres); cudaFree(ctx->d_sph_nodes); cudaFree(ctx->d_sph_order);
    ctx->d_spheres = nullptr; ctx->d_sph_nodes = nullptr; ctx->d_sph_order = nullptr;
    ctx->sc.spheres = nullptr; ctx->sc.sph_nodes = nullptr; ctx->sc.sph_order = nullptr; ctx->sc.n_spheres = 0;
    if (n == 0) return SQT_OK;
    CU(cudaMalloc(&ctx->d_spheres, ds.size() * sizeof(float4)));
    CU(cudaMemcpy(ctx->d_spheres, ds.data(), ds.size() * sizeof(float4), cudaMemcpyHostToDevice));
    ctx->sc.spheres = ctx->d_spheres; ctx->sc.n_spheres = n;
    // a hierarchy over the spheres (sphere_step): only for finite, moderate coordinates -- the culling bound assumes no overflow
    bool tame = ctx->sphere_bvh != 0;
    for (uint32_t k = 0; k < n && tame; ++k)
        for (int q = 0; q < 4; ++q) { const float v = q < 3 ? sp[k].center[q] : sp[k].radius; if (!(fabsf(v) < 1.0e12f)) tame = false; }
    if (tame) {
        SphereBvh bvh;
        build_sphere_bvh(sp, n, bvh);
        CU(cudaMalloc(&ctx->d_sph_nodes, bvh.nodes.size() * sizeof(float4)));
        CU(cudaMalloc(&ctx->d_sph_order, bvh.order.size() * sizeof(uint32_t)));
        CU(cudaMemcpy(ctx->d_sph_nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_sph_order, bvh.order.data(), bvh.order.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        ctx->sc.sph_nodes = ctx->d_sph_nodes; ctx->sc.sph_order = ctx->d_sph_order;
    }
    return SQT_OK;
}

extern "C" int sqt_set_option(sqt_ctx *ctx, int option, int value) {
    if (!ctx) return SQT_E_INVALID;
    if (option == SQT_OPT_LEAF_CULL) { ctx->leaf_cull = value ? 1 : 0; ctx->sc.leaf_cull = (uint32_t)ctx->leaf_cull; return SQT_OK; }
    if (option == SQT_OPT_SPHERE_BVH) {          // takes effect at the next sqt_upload_spheres; 0 = test every sphere for every ray (the definition)
        ctx->sphere_bvh = value ? 1 : 0;
        if (!value) { ctx->sc.sph_nodes = nullptr; ctx->sc.sph_order = nullptr; }
        return SQT_OK;
    }
    return fail(ctx, SQT_E_INVALID, "unknown option %d", option);
}

// ------------------------------------------------------------------------------ helpers
static int ensure_image(sqt_ctx *ctx, long long npix) {
    if (npix <= ctx->cap_pixels) return SQT_OK;
    cudaFree(ctx->d_prim); cudaFree(ctx->d_accum); cudaFree(ctx->d_rgb8); cudaFree(ctx->d_pixel_list);
    ctx->d_prim = nullptr; ctx->d_accum = nullptr; ctx->d_rgb8 = nullptr; ctx->d_pixel_list = nullptr; ctx->cap_pixels = 0;
    CU(cudaMalloc(&ctx->d_prim, (size_t)npix * sizeof(int2)));
    CU(cudaMalloc(&ctx->d_pixel_list, (size_t)(npix + 32) * sizeof(int)));
    CU(cudaMalloc(&ctx->d_accum, (size_t)npix * 3 * sizeof(float)));
    CU(cudaMalloc(&ctx->d_rgb8, (size_t)npix * 3));
    ctx->cap_pixels = npix;
    return SQT_OK;
}
static int ensure_sbuf(sqt_ctx *ctx, long long bytes) {
    if (bytes <= ctx->cap_sbuf) return SQT_OK;
    cudaFree(ctx->d_sbuf); ctx->d_sbuf = nullptr; ctx->cap_sbuf = 0;
    CU(cudaMalloc(&ctx->d_sbuf, (size_t)bytes));
    ctx->cap_sbuf = bytes;
    return SQT_OK;
}
static int ensure_host_image(sqt_ctx *ctx, long long npix) {
    if (npix <= ctx->cap_host_pixels) return SQT_OK;
    cudaFreeHost(ctx->h_rgb8); cudaFreeHost(ctx->h_accum); ctx->h_rgb8 = nullptr; ctx->h_accum = nullptr; ctx->cap_host_pixels = 0;
    CU(cudaMallocHost(&ctx->h_rgb8, (size_t)npix * 3));
    CU(cudaMallocHost(&ctx->h_accum, (size_t)npix * 3 * sizeof(float)));
    ctx->cap_host_pixels = npix;
    return SQT_OK;
}
static int ensure_rays(sqt_ctx *ctx, long long n) {
    if (n <= ctx->cap_rays) return SQT_OK;
    cudaFree(ctx->d_org); cudaFree(ctx->d_dir); cudaFree(ctx->d_dist); cudaFree(ctx->d_point); cudaFree(ctx->d_tri);
    ctx->d_org = ctx->d_dir = ctx->d_dist = ctx->d_point = nullptr; ctx->d_tri = nullptr; ctx->cap_rays = 0;
    CU(cudaMalloc(&ctx->d_org, (size_t)n * 12)); CU(cudaMalloc(&ctx->d_dir, (size_t)n * 12));
    CU(cudaMalloc(&ctx->d_dist, (size_t)n * 4)); CU(cudaMalloc(&ctx->d_point, (size_t)n * 12));
    CU(cudaMalloc(&ctx->d_tri, (size_t)n * 4));
    ctx->cap_rays = n;
    return SQT_OK;
}
// persistent launch: as many 128-lane CTAs as stay resident on the 148 SMs (never more than the work needs)
template <class K>
static int persistent_grid(sqt_ctx *ctx, K kernel, long long nwork) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long grid = (long long)ctx->sm_count * per_sm, want = (nwork + 127) / 128;
    if (want < grid) grid = want ? want : 1;
    return (int)grid;
}
// k_paths_pool launch shape + the per-slot global state it needs.  `launch` false: only size the grid and make sure the
// allocations exist (the pre-pass of a render, so that nothing can fail between the ranks' agreement and the collective).
struct PoolPlan { int grid = 0; size_t smem = 0; int depth = 1, pm_stride = 8; };
template <bool COUNT, int K>
static int pool_step(sqt_ctx *ctx, const RenderParams &d, const RoundInfo &rd, int round, long long nitems, bool launch) {
    const size_t smem = (size_t)4 * (32 * K * PF_WORDS + 4 * (32 * K) / 4 + 64 + 8 + 8 + 32) * sizeof(uint32_t);
    auto kern = k_paths_pool<COUNT, K>;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    if (ctx->pool_blocks > 0 && ctx->pool_blocks < per_sm) {
        // keep the rest of the 228 KB for L1: the triangle / node / stack working set lives there
        per_sm = ctx->pool_blocks;
        int pct = (int)((smem + 1024) * per_sm * 100 / (228 * 1024)) + 1;
        if (pct > 100) pct = 100;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
    }
    if (ctx->pool_carveout > 0) CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->pool_carveout));
    long long grid = (long long)ctx->sm_count * per_sm, want = (nitems + 128 * K - 1) / (128 * K);
    if (want < grid) grid = want ? want : 1;
    const long long slots = grid * 4 * 32 * K;
    // a ray's stack holds at most one 16-byte entry per branch on a root-to-leaf path; the stacks of a warp's slots are
    // interleaved entry by entry, so the few entries in use of all resident slots share cache lines and stay in L2
    int depth = (int)ctx->tree_height;
    if (depth > kStackEntries) depth = kStackEntries;
    if (depth < 1) depth = 1;
    // the per-path material list of a slot: max_depth entries, packed (16-byte granules) for the same reason
    int pm_stride = (d.max_depth + 7) & ~7;
    if (pm_stride > SQT_MAX_DEPTH) pm_stride = SQT_MAX_DEPTH;
    if (slots > ctx->cap_pool_slots || slots * depth > ctx->cap_stack_entries) {
        cudaFree(ctx->d_gstack); cudaFree(ctx->d_gpm); cudaFree(ctx->d_gpath);
        ctx->d_gstack = nullptr; ctx->d_gpm = nullptr; ctx->d_gpath = nullptr; ctx->cap_pool_slots = 0; ctx->cap_stack_entries = 0;
        const long long cap = (long long)ctx->sm_count * 16 * 4 * 32 * K > slots ? (long long)ctx->sm_count * 16 * 4 * 32 * K : slots;
        CU(cudaMalloc(&ctx->d_gstack, (size_t)cap * depth * sizeof(float4)));
        CU(cudaMalloc(&ctx->d_gpm, (size_t)cap * SQT_MAX_DEPTH * sizeof(uint16_t)));
        CU(cudaMalloc(&ctx->d_gpath, (size_t)cap * 2 * sizeof(uint4)));
        ctx->cap_pool_slots = cap; ctx->cap_stack_entries = cap * depth;
    }
    if (!launch) return SQT_OK;
    kern<<<(int)grid, 128, smem, ctx->stream>>>(ctx->sc, d, rd, round, ctx->d_stats, ctx->pool_tune, ctx->d_gstack, ctx->d_gpm, ctx->d_gpath, depth, pm_stride);
    CU(cudaGetLastError());
    return SQT_OK;
}
template <bool COUNT>
static int pool_step_k(sqt_ctx *ctx, const RenderParams &d, const RoundInfo &rd, int round, long long nitems, bool launch) {
    switch (ctx->pool_k) {
    case 1: return pool_step<COUNT, 1>(ctx, d, rd, round, nitems, launch);
    case 2: return pool_step<COUNT, 2>(ctx, d, rd, round, nitems, launch);
    default: return pool_step<COUNT, 4>(ctx, d, rd, round, nitems, launch);
    }
}
static float ev_ms(cudaEvent_t a, cudaEvent_t b) { float ms = 0; cudaEventElapsedTime(&ms, a, b); return ms; }

// ------------------------------------------------------------------------------ intersect_batch
extern "C" int sqt_intersect_batch(sqt_ctx *ctx, const float *org, const float *dir, int64_t n, int32_t *tri_out,
                                   float *dist_out, float *point_out, sqt_stats *stats) {
    if (!ctx) return SQT_E_INVALID;
    if (!ctx->has_scene) return fail(ctx, SQT_E_NO_SCENE, "sqt_intersect_batch before sqt_upload_scene");
    if (n < 0 || (n > 0 && (!org || !dir || !tri_out))) return fail(ctx, SQT_E_INVALID, "bad batch arguments");
    if (stats) memset(stats, 0, sizeof *stats);
    if (n == 0) return SQT_OK;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_rays(ctx, n); if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CU(cudaEventRecord(ctx->ev[0], st));
    CU(cudaMemcpyAsync(ctx->d_org, org, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->d_dir, dir, (size_t)n * 12, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(ctx->d_stats, 0, sizeof(DeviceStats), st));
    CU(cudaEventRecord(ctx->ev[1], st));
    if (stats)      // instrumented variant: fills the visit/test counters
        k_intersect_batch<true><<<persistent_grid(ctx, k_intersect_batch<true>, n), 128, 0, st>>>(
            ctx->sc, ctx->d_org, ctx->d_dir, n, ctx->d_tri, dist_out ? ctx->d_dist : nullptr, point_out ? ctx->d_point : nullptr,
            ctx->d_stats, ctx->tune);
    else
        k_intersect_batch<false><<<persistent_grid(ctx, k_intersect_batch<false>, n), 128, 0, st>>>(
            ctx->sc, ctx->d_org, ctx->d_dir, n, ctx->d_tri, dist_out ? ctx->d_dist : nullptr, point_out ? ctx->d_point : nullptr,
            ctx->d_stats, ctx->tune);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->ev[2], st));
    CU(cudaMemcpyAsync(tri_out, ctx->d_tri, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (dist_out) CU(cudaMemcpyAsync(dist_out, ctx->d_dist, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (point_out) CU(cudaMemcpyAsync(point_out, ctx->d_point, (size_t)n * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->h_stats, ctx->d_stats, sizeof(DeviceStats), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(ctx->ev[3], st));
    CU(cudaStreamSynchronize(st));
    if (stats) {
        stats->h2d_ms = ev_ms(ctx->ev[0], ctx->ev[1]); stats->device_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
        stats->d2h_ms = ev_ms(ctx->ev[2], ctx->ev[3]);
        stats->rays_traced = ctx->h_stats->rays; stats->rays_reference = ctx->h_stats->rays;
        stats->branch_visits = ctx->h_stats->branch_visits; stats->child_box_tests = ctx->h_stats->child_box_tests;
        stats->tri_tests = ctx->h_stats->tri_tests; stats->leaves_culled = ctx->h_stats->leaves_culled;
        stats->mt_pass_a = ctx->h_stats->mt_pass_a; stats->mt_pass_u = ctx->h_stats->mt_pass_u;
        stats->mt_pass_v = ctx->h_stats->mt_pass_v; stats->mt_accept = ctx->h_stats->mt_accept;
        stats->h2d_bytes = (uint64_t)n * 24; stats->d2h_bytes = (uint64_t)n * (4 + (dist_out ? 4 : 0) + (point_out ? 12 : 0));
        stats->kernel_launches = 1;
    }
    return SQT_OK;
}

// ------------------------------------------------------------------------------ render
static int check_params(sqt_ctx *ctx, const sqt_camera *cam, const sqt_render_params *p) {
    if (!cam || !p) return fail(ctx, SQT_E_INVALID, "cam/params is NULL");
    if (!ctx->has_scene) return fail(ctx, SQT_E_NO_SCENE, "sqt_render before sqt_upload_scene");
    if (p->rows <= 0 || p->cols <= 0 || p->xdiv <= 0 || p->ydiv <= 0 || p->spp <= 0)
        return fail(ctx, SQT_E_INVALID, "rows/cols/xdiv/ydiv/spp must be positive");
    if (p->max_depth < 1 || p->max_depth > SQT_MAX_DEPTH) return fail(ctx, SQT_E_UNSUPPORTED, "max_depth must be in 1..%d", SQT_MAX_DEPTH);
    if (p->mode != 0 && p->mode != 1) return fail(ctx, SQT_E_INVALID, "mode must be 0 (trace) or 1 (cast)");
    if ((long long)p->rows * p->cols > (1ll << 31)) return fail(ctx, SQT_E_UNSUPPORTED, "image too large");
    return SQT_OK;
}

static RenderParams to_device_params(const sqt_ctx *ctx, const sqt_camera *cam, const sqt_render_params *p) {
    RenderParams d = {};
    d.rows = p->rows; d.cols = p->cols; d.xdiv = p->xdiv; d.ydiv = p->ydiv; d.seed_stride = p->seed_stride;
    d.spp = p->spp; d.max_depth = p->max_depth; d.mode = p->mode; d.seed = p->seed;
    d.rank = ctx->rank; d.world = ctx->world; d.split_samples = (p->flags & SQT_F_SPLIT_SAMPLES) ? 1 : 0;
    d.primary_reuse = (p->flags & SQT_F_NO_PRIMARY_REUSE) ? 0 : 1;
    for (int k = 0; k < 3; ++k) d.cam_pos[k] = cam->position[k];
    for (int k = 0; k < 9; ++k) d.cam_rot[k] = cam->rotation[k];
    d.terminate_on_black = (ctx->terminate_on_black_ok && !(p->flags & SQT_F_NO_EARLY_TERMINATION)) ? 1 : 0;
    return d;
}

// Everything of a render that can fail for lack of memory or because of its arguments, done BEFORE any kernel or
// collective is enqueued: image buffers, the sample buffer of a round, the pool slots' global state.
struct RenderPlan { long long nwork = 0; int k0 = 0, k1 = 0, log2_s = 0; long long sbytes = 0; };
static int prepare_render(sqt_ctx *ctx, const RenderParams &d, bool count, bool host_image, RenderPlan &pl) {
    const long long npix = (long long)d.rows * d.cols;
    int rc = ensure_image(ctx, npix); if (rc) return rc;
    if (host_image) { rc = ensure_host_image(ctx, npix); if (rc) return rc; }
    pl.nwork = work_items(d);
    if (d.mode == 1) return SQT_OK;
    sample_range(d, pl.k0, pl.k1);
    pl.log2_s = round_log2_s(pl.nwork, pl.k1 - pl.k0 > 0 ? pl.k1 - pl.k0 : 1, ctx->sbuf_budget);
    const int S = 1 << pl.log2_s;
    if ((pl.k1 - pl.k0 + S - 1) / S > 256) return fail(ctx, SQT_E_UNSUPPORTED, "more than 256 sample rounds (raise SQT_SBUF_MB)");
    pl.sbytes = pl.nwork * (long long)S * 12ll;
    rc = ensure_sbuf(ctx, pl.sbytes > 0 ? pl.sbytes : 12); if (rc) return rc;
    if (ctx->pool_k > 0) {
        RoundInfo none = {};
        rc = count ? pool_step_k<true>(ctx, d, none, 0, pl.nwork << pl.log2_s, false) : pool_step_k<false>(ctx, d, none, 0, pl.nwork << pl.log2_s, false);
        if (rc) return rc;
    }
    return SQT_OK;
}

// Multi-rank contexts: all ranks learn whether every rank got through check_params + prepare_render, with a one-word
// ncclAllReduce(max) -- a collective every rank reaches whatever happened to it locally.  Only if all agree is the
// frame (and its ncclReduce) enqueued; otherwise every rank returns an error and nobody waits in a collective.
static int group_agree(sqt_ctx *ctx, int rc_local) {
    if (ctx->world <= 1 || !ctx->comm) return rc_local;
    if (ctx->comm_broken) return fail(ctx, SQT_E_NCCL, "the communicator of this group was aborted after an earlier failure");
    NcclApi *na = nccl_api();
    const std::string local_err = ctx->err;
    cudaSetDevice(ctx->device);
    uint32_t *flag = ctx->d_flag + 1;
    const uint32_t v = rc_local ? 1u : 0u;
    uint32_t out = 1u;
    cudaError_t e = cudaMemcpyAsync(flag, &v, 4, cudaMemcpyHostToDevice, ctx->stream);
    ncclResult_t r = e == cudaSuccess ? na->AllReduce(flag, flag, 1, kNcclInt32, kNcclMax, ctx->comm, ctx->stream) : 1;
    if (r == 0) e = cudaMemcpyAsync(&out, flag, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (r == 0 && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (r != 0 || e != cudaSuccess) {
        na->CommAbort(ctx->comm); ctx->comm_broken = true;
        return fail(ctx, SQT_E_NCCL, "group agreement failed (%s); communicator aborted", r ? na->GetErrorString(r) : cudaGetErrorString(e));
    }
    if (rc_local) { ctx->err = local_err; return rc_local; }
    if (out) return fail(ctx, SQT_E_NCCL, "another rank of the group could not start this render; nothing was enqueued");
    return SQT_OK;
}
// a failure after the agreement (a launch error) would leave the other ranks waiting in ncclReduce: abort the communicator
static int group_abort(sqt_ctx *ctx, int rc) {
    if (rc && ctx->world > 1 && ctx->comm && !ctx->comm_broken) { nccl_api()->CommAbort(ctx->comm); ctx->comm_broken = true; }
    return rc;
}

// Enqueue all kernels of one render on ctx->stream (no host sync, no allocation).  Events: 0 start, 1 after primary,
// 2 after paths, 3 after reduce, 4 after tonemap.
static int enqueue_render(sqt_ctx *ctx, const RenderParams &d, bool count, const RenderPlan &pl, uint32_t *launches) {
    const long long npix = (long long)d.rows * d.cols;
    int rc = SQT_OK;
    cudaStream_t st = ctx->stream;
    CU(cudaMemsetAsync(ctx->d_stats, 0, sizeof(DeviceStats), st));
    CU(cudaMemsetAsync(ctx->d_accum, 0, (size_t)npix * 3 * sizeof(float), st));
    CU(cudaEventRecord(ctx->ev[0], st));
    const long long nwork = pl.nwork;
    uint32_t nl = 0;
    if (d.mode == 1) {
        CU(cudaEventRecord(ctx->ev[1], st));
        if (count) k_raycast<true><<<persistent_grid(ctx, k_raycast<true>, nwork), 128, 0, st>>>(ctx->sc, d, ctx->d_accum, ctx->d_stats, ctx->tune);
        else k_raycast<false><<<persistent_grid(ctx, k_raycast<false>, nwork), 128, 0, st>>>(ctx->sc, d, ctx->d_accum, ctx->d_stats, ctx->tune);
        CU(cudaGetLastError()); nl++;
    } else {
        if (d.primary_reuse) {
            if (count) k_primary<true><<<persistent_grid(ctx, k_primary<true>, nwork), 128, 0, st>>>(ctx->sc, d, ctx->d_prim, ctx->d_pixel_list, ctx->d_stats, ctx->tune);
            else k_primary<false><<<persistent_grid(ctx, k_primary<false>, nwork), 128, 0, st>>>(ctx->sc, d, ctx->d_prim, ctx->d_pixel_list, ctx->d_stats, ctx->tune);
            CU(cudaGetLastError()); nl++;
        }
        CU(cudaEventRecord(ctx->ev[1], st));
        // rounds of S samples per pixel: trace (persistent lanes, one counter per round), then add in sample order
        const int k0 = pl.k0, k1 = pl.k1;
        RoundInfo rd = {};
        rd.pixel_list = d.primary_reuse ? ctx->d_pixel_list : nullptr;
        rd.prim = d.primary_reuse ? ctx->d_prim : nullptr;
        rd.n_slots = nwork; rd.slot_stride = nwork;
        rd.log2_s = pl.log2_s;
        const int S = 1 << rd.log2_s;
        rd.sbuf = ctx->d_sbuf;
        const int agrid = (int)((nwork + 255) / 256 < (long long)ctx->sm_count * 8 ? (nwork + 255) / 256 : (long long)ctx->sm_count * 8);
        int round = 0;
        for (int kb = k0; kb < k1; kb += S, ++round) {
            rd.k0 = kb; rd.k1 = kb + S < k1 ? kb + S : k1;
            CU(cudaMemsetAsync(ctx->d_sbuf, 0, (size_t)(nwork * (long long)(rd.k1 - rd.k0) * 12ll), st));
            if (ctx->pool_k > 0) {
                rc = count ? pool_step_k<true>(ctx, d, rd, round, nwork << rd.log2_s, true) : pool_step_k<false>(ctx, d, rd, round, nwork << rd.log2_s, true);
                if (rc) return rc;
            } else if (count) k_paths<true><<<persistent_grid(ctx, k_paths<true>, nwork << rd.log2_s), 128, 0, st>>>(ctx->sc, d, rd, round, ctx->d_stats, ctx->tune);
            else k_paths<false><<<persistent_grid(ctx, k_paths<false>, nwork << rd.log2_s), 128, 0, st>>>(ctx->sc, d, rd, round, ctx->d_stats, ctx->tune);
            CU(cudaGetLastError()); nl++;
            k_accumulate<<<agrid > 0 ? agrid : 1, 256, 0, st>>>(d, rd, ctx->d_accum, ctx->d_stats);
            CU(cudaGetLastError()); nl++;
        }
    }
    CU(cudaEventRecord(ctx->ev[2], st));
    if (ctx->world > 1 && ctx->comm) {
        NcclApi *na = nccl_api();
        ncclResult_t r = na->Reduce(ctx->d_accum, ctx->d_accum, (size_t)npix * 3, kNcclFloat32, kNcclSum, 0, ctx->comm, st);
        if (r != 0) return fail(ctx, SQT_E_NCCL, "ncclReduce failed: %s", na->GetErrorString(r));
        nl++;
    }
    CU(cudaEventRecord(ctx->ev[3], st));
    if (ctx->rank == 0) {
        const float inv = 1.0f / (float)d.spp;
        k_tonemap<<<(int)((npix + 255) / 256), 256, 0, st>>>(ctx->d_accum, npix, inv, ctx->d_rgb8);
        CU(cudaGetLastError()); nl++;
    }
    CU(cudaEventRecord(ctx->ev[4], st));
    CU(cudaMemcpyAsync(ctx->h_stats, ctx->d_stats, sizeof(DeviceStats), cudaMemcpyDeviceToHost, st));
    ctx->img_pixels = npix; ctx->img_spp = d.spp;
    *launches = nl;
    return SQT_OK;
}

static void fill_stats(sqt_ctx *ctx, sqt_stats *s, uint32_t launches) {
    if (!s) return;
    s->primary_ms = ev_ms(ctx->ev[0], ctx->ev[1]); s->paths_ms = ev_ms(ctx->ev[1], ctx->ev[2]);
    s->reduce_ms = ev_ms(ctx->ev[2], ctx->ev[3]); s->tonemap_ms = ev_ms(ctx->ev[3], ctx->ev[4]);
    s->device_ms = ev_ms(ctx->ev[0], ctx->ev[4]);
    s->rays_traced = ctx->h_stats->rays; s->samples = ctx->h_stats->samples;
    s->rays_reference = ctx->h_stats->rays + ctx->h_stats->primary_reused;
    s->branch_visits = ctx->h_stats->branch_visits; s->child_box_tests = ctx->h_stats->child_box_tests;
    s->tri_tests = ctx->h_stats->tri_tests; s->leaves_culled = ctx->h_stats->leaves_culled;
    s->mt_pass_a = ctx->h_stats->mt_pass_a; s->mt_pass_u = ctx->h_stats->mt_pass_u;
    s->mt_pass_v = ctx->h_stats->mt_pass_v; s->mt_accept = ctx->h_stats->mt_accept;
    s->kernel_launches = launches;
    if (getenv("SQT_DEBUG_STATS") && ctx->h_stats->dbg[0]) {        // scheduler diagnostics of the instrumented pool kernel (development)
        const unsigned long long *d = ctx->h_stats->dbg;
        const double rays = (double)ctx->h_stats->rays;
        fprintf(stderr, "[sqt pool] per ray: T rounds %.3f (fill %.1f)  L rounds %.3f (fill %.1f)  R rounds %.3f (fill %.1f)\n", d[0] / rays, (double)d[3] / d[0],
                d[1] / rays, (double)d[4] / (d[1] ? d[1] : 1), d[2] / rays, (double)d[5] / (d[2] ? d[2] : 1));
        fprintf(stderr, "[sqt pool] T burst, lanes per step  ret:");
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %.1f", (double)d[16 + k] / d[0]);
        fprintf(stderr, "  desc:");
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %.1f", (double)d[8 + k] / d[0]);
        fprintf(stderr, "\n[sqt pool] L passes %.3f per L round, rays with triangles %.1f, tests %.1f, chunks %.2f per pass\n", (double)d[24] / (d[1] ? d[1] : 1),
                (double)d[25] / (d[24] ? d[24] : 1), (double)d[26] / (d[24] ? d[24] : 1), (double)d[27] / (d[24] ? d[24] : 1));
    }
}

extern "C" int sqt_render_resident(sqt_ctx *ctx, const sqt_camera *cam, const sqt_render_params *p, sqt_stats *stats) {
    if (!ctx) return SQT_E_INVALID;
    if (stats) memset(stats, 0, sizeof *stats);
    cudaSetDevice(ctx->device);
    RenderParams d = {};
    RenderPlan pl;
    int rc = check_params(ctx, cam, p);
    if (!rc) { d = to_device_params(ctx, cam, p); rc = prepare_render(ctx, d, (p->flags & SQT_F_COUNT_WORK) != 0, false, pl); }
    rc = group_agree(ctx, rc); if (rc) return rc;
    uint32_t nl = 0;
    rc = enqueue_render(ctx, d, (p->flags & SQT_F_COUNT_WORK) != 0, pl, &nl); if (rc) return group_abort(ctx, rc);
    CU(cudaStreamSynchronize(ctx->stream));
    fill_stats(ctx, stats, nl);
    return SQT_OK;
}

extern "C" int sqt_download_image(sqt_ctx *ctx, uint8_t *rgb8_out, float *accum_out) {
    if (!ctx) return SQT_E_INVALID;
    if (ctx->img_pixels <= 0) return fail(ctx, SQT_E_INVALID, "no rendered image to download");
    CU(cudaSetDevice(ctx->device));
    const long long npix = ctx->img_pixels;
    int rc = ensure_host_image(ctx, npix); if (rc) return rc;
    if (rgb8_out) CU(cudaMemcpyAsync(ctx->h_rgb8, ctx->d_rgb8, (size_t)npix * 3, cudaMemcpyDeviceToHost, ctx->stream));
    if (accum_out) CU(cudaMemcpyAsync(ctx->h_accum, ctx->d_accum, (size_t)npix * 12, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (rgb8_out) memcpy(rgb8_out, ctx->h_rgb8, (size_t)npix * 3);
    if (accum_out) memcpy(accum_out, ctx->h_accum, (size_t)npix * 12);
    return SQT_OK;
}

extern "C" int sqt_render(sqt_ctx *ctx, const sqt_camera *cam, const sqt_render_params *p, uint8_t *rgb8_out,
                          float *accum_out, sqt_stats *stats) {
    if (!ctx) return SQT_E_INVALID;
    if (stats) memset(stats, 0, sizeof *stats);
    cudaSetDevice(ctx->device);
    RenderParams d = {};
    RenderPlan pl;
    int rc = check_params(ctx, cam, p);
    if (!rc) { d = to_device_params(ctx, cam, p); rc = prepare_render(ctx, d, (p->flags & SQT_F_COUNT_WORK) != 0, true, pl); }
    rc = group_agree(ctx, rc); if (rc) return rc;
    const long long npix = (long long)d.rows * d.cols;
    uint32_t nl = 0;
    rc = enqueue_render(ctx, d, (p->flags & SQT_F_COUNT_WORK) != 0, pl, &nl); if (rc) return group_abort(ctx, rc);
    cudaStream_t st = ctx->stream;
    const bool root = ctx->rank == 0;
    CU(cudaEventRecord(ctx->ev[5], st));
    if (root && rgb8_out) CU(cudaMemcpyAsync(ctx->h_rgb8, ctx->d_rgb8, (size_t)npix * 3, cudaMemcpyDeviceToHost, st));
    if (root && accum_out) CU(cudaMemcpyAsync(ctx->h_accum, ctx->d_accum, (size_t)npix * 12, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(ctx->ev[6], st));
    CU(cudaStreamSynchronize(st));
    if (root && rgb8_out) memcpy(rgb8_out, ctx->h_rgb8, (size_t)npix * 3);
    if (root && accum_out) memcpy(accum_out, ctx->h_accum, (size_t)npix * 12);
    fill_stats(ctx, stats, nl);
    if (stats) {
        stats->d2h_ms = ev_ms(ctx->ev[5], ctx->ev[6]);
        stats->d2h_bytes = root ? (uint64_t)npix * ((rgb8_out ? 3 : 0) + (accum_out ? 12 : 0)) : 0;
        stats->h2d_bytes = sizeof(RenderParams);    // camera + parameters travel as kernel arguments
    }
    return SQT_OK;
}

extern "C" int sqt_tone_map(sqt_ctx *ctx, const float *mean_rgb, int64_t npix, uint8_t *rgb8_out) {
    if (!ctx) return SQT_E_INVALID;
    if (npix < 0 || (npix > 0 && (!mean_rgb || !rgb8_out))) return fail(ctx, SQT_E_INVALID, "bad tone_map arguments");
    if (npix == 0) return SQT_OK;
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_image(ctx, npix); if (rc) return rc;
    cudaStream_t st = ctx->stream;
    CU(cudaMemcpyAsync(ctx->d_accum, mean_rgb, (size_t)npix * 12, cudaMemcpyHostToDevice, st));
    k_tonemap<<<(int)((npix + 255) / 256), 256, 0, st>>>(ctx->d_accum, npix, 1.0f, ctx->d_rgb8);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(rgb8_out, ctx->d_rgb8, (size_t)npix * 3, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->img_pixels = 0;
    return SQT_OK;
}

// ------------------------------------------------------------------------------ group (NCCL)
extern "C" int sqt_comm_unique_id(uint8_t id_out[SQT_COMM_ID_BYTES]) {
    NcclApi *na = nccl_api();
    if (!na->lib) return fail(nullptr, SQT_E_NCCL, "%s", na->err.c_str());
    ncclUniqueId id;
    ncclResult_t r = na->GetUniqueId(&id);
    if (r != 0) return fail(nullptr, SQT_E_NCCL, "ncclGetUniqueId: %s", na->GetErrorString(r));
    memcpy(id_out, id.internal, SQT_COMM_ID_BYTES);
    return SQT_OK;
}

extern "C" int sqt_comm_init(sqt_ctx *ctx, int rank, int world, const uint8_t id_in[SQT_COMM_ID_BYTES]) {
    if (!ctx) return SQT_E_INVALID;
    if (world < 1 || rank < 0 || rank >= world) return fail(ctx, SQT_E_INVALID, "bad rank/world %d/%d", rank, world);
    if (world == 1) { ctx->rank = 0; ctx->world = 1; return SQT_OK; }
    NcclApi *na = nccl_api();
    if (!na->lib) return fail(ctx, SQT_E_NCCL, "%s", na->err.c_str());
    CU(cudaSetDevice(ctx->device));
    ncclUniqueId id; memcpy(id.internal, id_in, SQT_COMM_ID_BYTES);
    ncclResult_t r = na->CommInitRank(&ctx->comm, world, id, rank);
    if (r != 0) return fail(ctx, SQT_E_NCCL, "ncclCommInitRank: %s", na->GetErrorString(r));
    ctx->rank = rank; ctx->world = world;
    return SQT_OK;
}

extern "C" int sqt_comm_init_all(sqt_ctx **ctxs, int n) {
    if (!ctxs || n < 1) return SQT_E_INVALID;
    if (n == 1) { ctxs[0]->rank = 0; ctxs[0]->world = 1; return SQT_OK; }
    NcclApi *na = nccl_api();
    if (!na->lib) return fail(ctxs[0], SQT_E_NCCL, "%s", na->err.c_str());
    std::vector<int> devs(n); std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; ++i) devs[i] = ctxs[i]->device;
    ncclResult_t r = na->CommInitAll(comms.data(), n, devs.data());
    if (r != 0) return fail(ctxs[0], SQT_E_NCCL, "ncclCommInitAll: %s", na->GetErrorString(r));
    for (int i = 0; i < n; ++i) { ctxs[i]->comm = comms[i]; ctxs[i]->rank = i; ctxs[i]->world = n; }
    return SQT_OK;
}

extern "C" int sqt_render_group(sqt_ctx **ctxs, int n, const sqt_camera *cam, const sqt_render_params *p,
                                uint8_t *rgb8_out, float *accum_out, sqt_stats *stats) {
    if (!ctxs || n < 1) return SQT_E_INVALID;
    if (n == 1) return sqt_render(ctxs[0], cam, p, rgb8_out, accum_out, stats);
    // one host thread per device: the ncclReduce calls of a single-process group must be issued concurrently
    std::vector<int> rcs(n, 0);
    std::vector<std::thread> th;
    for (int i = 1; i < n; ++i)
        th.emplace_back([&, i]() { rcs[i] = sqt_render(ctxs[i], cam, p, nullptr, nullptr, nullptr); });
    rcs[0] = sqt_render(ctxs[0], cam, p, rgb8_out, accum_out, stats);
    for (auto &t : th) t.join();
    for (int i = 0; i < n; ++i) if (rcs[i]) { if (i) ctxs[0]->err = ctxs[i]->err; return rcs[i]; }
    return SQT_OK;
}

// ------------------------------------------------------------------------------ roofline microbenchmarks
extern "C" int sqt_measure_fp32_peak(sqt_ctx *ctx, double *gops) {
    if (!ctx || !gops) return SQT_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    float *d = nullptr; CU(cudaMalloc(&d, 16));
    const int iters = 4096, grid = ctx->sm_count * 8, block = 256;
    cudaStream_t st = ctx->stream;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(ctx->ev[0], st));
        k_fp32_peak<<<grid, block, 0, st>>>(d, iters, 1.0000001f, 1e-7f);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev[1], st));
        CU(cudaStreamSynchronize(st));
        const double ms = ev_ms(ctx->ev[0], ctx->ev[1]);
        const double ops = (double)grid * block * iters * 32.0;
        if (rep > 0 && ops / ms * 1e-6 > best) best = ops / ms * 1e-6;
    }
    cudaFree(d);
    *gops = best;
    return SQT_OK;
}

extern "C" int sqt_measure_l2_bandwidth(sqt_ctx *ctx, double *gbs) {
    if (!ctx || !gbs) return SQT_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    const long long bytes = 32ll << 20, n4 = bytes / 16;
    float4 *buf = nullptr; float *out = nullptr;
    CU(cudaMalloc(&buf, bytes)); CU(cudaMalloc(&out, 16));
    CU(cudaMemset(buf, 0, bytes));
    cudaStream_t st = ctx->stream;
    const int grid = ctx->sm_count * 8, block = 256, reps = 16;
    double best = 0;
    for (int rep = 0; rep < 5; ++rep) {
        CU(cudaEventRecord(ctx->ev[0], st));
        k_l2_read<<<grid, block, 0, st>>>(buf, n4, reps, out);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ctx->ev[1], st));
        CU(cudaStreamSynchronize(st));
        const double ms = ev_ms(ctx->ev[0], ctx->ev[1]);
        const double gb = (double)bytes * reps / ms * 1e-6;
        if (rep > 0 && gb > best) best = gb;
    }
    cudaFree(buf); cudaFree(out);
    *gbs = best;
    return SQT_OK;
}
