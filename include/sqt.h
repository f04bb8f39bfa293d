/*
 * sqt.h -- C ABI of the B200 (sm_100a) path-tracing backend for squigly-trace.
 *
 * This is the drop-in boundary for the reference's render hot path.  The reference has no FFI
 * today (no `foreign import` anywhere); the two function-typed seams it does have are
 *
 *   render    :: Scene a -> Camera -> Settings -> IO ()        src/Lib.hs:68-75   (caller app/Main.hs:43)
 *   intersect :: a -> Ray -> Maybe Intersection  (Scene field)  src/Geometry.hs:62-65 (chosen app/Main.hs:52-56,
 *                                                               called src/Lib.hs:131,143,150)
 *
 * Everything below `render` except `writeImage` (Lib.hs:75) moves behind sqt_render();
 * sqt_intersect_batch() is the batched form of `intersect` and carries the bit-exact contract.
 * The Haskell host keeps Obj.hs / BIH.hs / Main.hs, flattens its BIH into the records below and
 * binds these symbols with `foreign import ccall safe`.
 *
 * Conventions: plain C, no C++ types, no exceptions across the boundary.  Every function returns
 * 0 on success or a non-zero SQT_E_* code; sqt_last_error(ctx) gives the message.  All pointers
 * are caller-owned HOST memory, valid for the duration of the call only; the library owns all
 * device memory.  A context is bound to one CUDA device, is not re-entrant and blocks the
 * calling thread.  There is no CPU fallback: without a usable sm_100 device sqt_create fails.
 */
#ifndef SQT_H
#define SQT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQT_ABI_VERSION 1

enum {
    SQT_OK = 0,
    SQT_E_INVALID = 1,      /* bad argument / malformed scene */
    SQT_E_CUDA = 2,         /* CUDA runtime error (message has the cudaError string) */
    SQT_E_NO_DEVICE = 3,    /* no CUDA device, or not compute capability 10.x */
    SQT_E_NO_SCENE = 4,     /* render/intersect before sqt_upload_scene */
    SQT_E_UNSUPPORTED = 5,  /* e.g. BIH deeper than SQT_MAX_HEIGHT, max_depth > SQT_MAX_DEPTH */
    SQT_E_NCCL = 6          /* NCCL missing or failed */
};

#define SQT_MAX_HEIGHT 48   /* deepest supported BIH (reference scene: 13) */
#define SQT_MAX_DEPTH 64    /* largest max_depth (reference: 3, Lib.hs:129) */

/* ---- flattened BIH (BIH.hs:26,37-43), nodes in any order with node 0 = root ------------------
 * Branch (BIHN axis lmax rmin) l r : a = index(l) | axis<<30 ,  b = index(r)             (axis X=0,Y=1,Z=2)
 * Leaf tris                         : a = first triangle in tris[] , b = count | 0x80000000   (lmax,rmin ignored)
 * At most 2^27 - 1 nodes and 2^27 - 1 triangles.  Child boxes are NOT stored: the library derives them exactly as
 * BIH.hs:130-141 does, by clipping the parent's box (plain copies of lmax/rmin, no arithmetic).  The `pad` word of
 * a triangle is ignored on input (the device copy uses it). */
typedef struct sqt_node {
    float lmax, rmin;
    uint32_t a, b;
} sqt_node;                              /* 16 B */
#define SQT_NODE_LEAF 0x80000000u

/* Triangles in leaf order, i.e. `flatten` (BIH.hs:50-52): each Leaf's vector is a contiguous range.
 * e1 = v1 - v0, e2 = v2 - v0 computed in binary32 on the host: identical bits to edge1/edge2 of
 * Geometry.hs:130-131, and e1 x e2 is `normal` (Geometry.hs:79-80). */
typedef struct sqt_tri {
    float v0[3], e1[3], e2[3];
    uint32_t material;                   /* index into mats[] */
    uint32_t orig_index;                 /* position in the parsed triangle list (what tri_out reports) */
    uint32_t pad;
} sqt_tri;                               /* 48 B */

/* Color.hs:78-83 */
typedef struct sqt_material {
    float reflective, surf_color[3], emissive, emit_color[3];
} sqt_material;                          /* 32 B */

/* EXTENSION (named by the north star; the reference has no sphere primitive or syntax, so the semantics are this
 * library's own and are restated in oracle/oracle.c): analytic, double-sided spheres; a ray's result is the closest of
 * the BIH hit and all sphere hits, earlier candidates of the list [BIH hit, sphere 0, sphere 1, ...] win ties.  (The
 * library finds the candidate spheres through a hierarchy of its own, see SQT_OPT_SPHERE_BVH.)  In tri_out a
 * sphere k is reported as n_tris + k. */
typedef struct sqt_sphere {
    float center[3];
    float radius;
    uint32_t material;                   /* index into mats[] */
    uint32_t pad[3];
} sqt_sphere;                            /* 32 B */

typedef struct sqt_scene_desc {
    float root_bounds[6];                /* bounds of BIH (BIH.hs:41,62-65): lo.xyz, hi.xyz */
    const sqt_node *nodes;  uint32_t n_nodes;
    const sqt_tri *tris;    uint32_t n_tris;
    const sqt_material *mats; uint32_t n_mats;
} sqt_scene_desc;

/* Geometry.hs:41 : position + rotation matrix as produced by rotMatrixRads (row-major 3x3) */
typedef struct sqt_camera {
    float position[3];
    float rotation[9];
} sqt_camera;

/* Settings that reach the hot path (Lib.hs:54-63 `samples`, `dimensions`, `cast`) plus the
 * index convention of Lib.hs:69-85,107-114 made explicit:
 *   pixel (row y, col x):  xoffs = (x - xdiv/2)/xdiv ; yoffs = (ydiv/2 - y)/ydiv ;
 *   generator seed of sample k = spp*(x + y*seed_stride) + k.
 * Reference-literal `-d W,H`: rows=W cols=H xdiv=W ydiv=H seed_stride=W.
 * Corrected:                  rows=H cols=W xdiv=W ydiv=H seed_stride=W. */
typedef struct sqt_render_params {
    int32_t rows, cols;
    int32_t xdiv, ydiv;
    int32_t seed_stride;
    int32_t spp;
    int32_t max_depth;                   /* intersections per path; reference = 3 (Lib.hs:129) */
    int32_t mode;                        /* 0 = raytrace (Lib.hs:127), 1 = raycast / --cast (Lib.hs:141) */
    uint64_t seed;                       /* key of the counter-based RNG */
    uint32_t flags;                      /* SQT_F_* */
    uint32_t reserved;
} sqt_render_params;

#define SQT_F_COUNT_WORK      1u   /* run the instrumented kernels and fill the *_visits/_tests counters (slower) */
#define SQT_F_SPLIT_SAMPLES   2u   /* multi-GPU: split the spp range per rank instead of pixel groups */
#define SQT_F_NO_PRIMARY_REUSE 4u  /* re-trace the (identical) primary ray for every sample, as Lib.hs:81-87 does */
#define SQT_F_NO_EARLY_TERMINATION 8u /* keep tracing below a surface whose surfColor is (0,0,0), as Lib.hs:135 does
                                         (the product of such a level is exactly 0, so the image is the same) */

typedef struct sqt_stats {
    double device_ms;            /* CUDA-event time of all kernels of this call on the context's stream */
    double primary_ms, paths_ms, tonemap_ms, reduce_ms;
    double h2d_ms, d2h_ms;       /* host<->device copies of this call (0 for the *_resident entry points) */
    uint64_t rays_traced;        /* closest-hit queries actually executed on the device */
    uint64_t rays_reference;     /* queries the reference would execute for the same job (Lib.hs:131 per segment) */
    uint64_t samples;            /* paths (Lib.hs:84) */
    uint64_t branch_visits, child_box_tests, tri_tests;    /* only with SQT_F_COUNT_WORK */
    uint64_t leaves_culled;      /* leaf visits skipped by the conservative tight-box test (with SQT_F_COUNT_WORK) */
    uint64_t mt_pass_a, mt_pass_u, mt_pass_v, mt_accept;   /* triangle tests that got past the a / u / v / t guards
                                                               of Geometry.hs:118-122 (with SQT_F_COUNT_WORK) */
    uint64_t h2d_bytes, d2h_bytes;
    uint32_t kernel_launches;
    uint32_t reserved;
} sqt_stats;

typedef struct sqt_ctx sqt_ctx;

/* lifecycle ------------------------------------------------------------------------------- */
int sqt_abi_version(void);
int sqt_create(int device, sqt_ctx **out);
int sqt_destroy(sqt_ctx *ctx);
const char *sqt_last_error(const sqt_ctx *ctx);        /* ctx may be NULL: error of the last failed sqt_create */

/* replaces the in-memory `Scene BIH` value handed to render (Main.hs:39,55-56) ------------- */
int sqt_upload_scene(sqt_ctx *ctx, const sqt_scene_desc *scene);

/* The same scene onto every context of a single-process group (ctxs[0..n) on n distinct devices): the tree is walked
 * once and every staged chunk of triangles is sent to all devices.  What the Haskell host calls once per run. */
int sqt_upload_scene_group(sqt_ctx **ctxs, int n, const sqt_scene_desc *scene);

/* what the last sqt_upload_scene[_group] on this context moved and cost: bytes copied host -> device, wall time of the
 * call, and the part of it the host spent walking the tree (the rest is copies + two device kernels).  Any pointer may
 * be NULL. */
int sqt_last_upload(const sqt_ctx *ctx, uint64_t *h2d_bytes, double *wall_ms, double *host_layout_ms);

/* optional, after sqt_upload_scene: n = 0 removes the spheres again */
int sqt_upload_spheres(sqt_ctx *ctx, const sqt_sphere *spheres, uint32_t n);

/* batched Scene.intersect (Geometry.hs:64; intersectBIH BIH.hs:101-141) ---------------------
 * org/dir: n x 3 floats, xyz interleaved.  tri_out[i] = orig_index of the closest hit or -1.
 * dist_out / point_out (n x 3) may be NULL.  Bit-exact contract: same index, dist and point bits
 * as the reference algorithm on the same tree. */
int sqt_intersect_batch(sqt_ctx *ctx, const float *org, const float *dir, int64_t n,
                        int32_t *tri_out, float *dist_out, float *point_out, sqt_stats *stats_or_null);

/* replaces the body of render above writeImage (Lib.hs:70-74) --------------------------------
 * rgb8_out: rows*cols*3 bytes, row-major = massiv `Array S Ix2 (Pixel RGB Word8)`.
 * accum_out (optional): rows*cols*3 floats, per-pixel radiance SUM over samples (before the 1/spp of Lib.hs:88).
 * In a multi-rank group only rank 0 receives the image; other ranks may pass NULL. */
int sqt_render(sqt_ctx *ctx, const sqt_camera *cam, const sqt_render_params *params,
               uint8_t *rgb8_out, float *accum_out, sqt_stats *stats_or_null);

/* same job, result left in device memory (no host copies); fetch it with sqt_download_image */
int sqt_render_resident(sqt_ctx *ctx, const sqt_camera *cam, const sqt_render_params *params, sqt_stats *stats_or_null);
int sqt_download_image(sqt_ctx *ctx, uint8_t *rgb8_out, float *accum_out);

/* rgbFloatToPixelRGB (Lib.hs:93-104) on the device, for n_pixels mean colours */
int sqt_tone_map(sqt_ctx *ctx, const float *mean_rgb, int64_t n_pixels, uint8_t *rgb8_out);

/* multi-GPU: one context per GPU (one process per GPU or all in one process); the group sums the
 * per-rank accumulation buffers with ncclReduce(sum, f32) onto rank 0 ---------------------- */
#define SQT_COMM_ID_BYTES 128
int sqt_comm_unique_id(uint8_t id_out[SQT_COMM_ID_BYTES]);
int sqt_comm_init(sqt_ctx *ctx, int rank, int world, const uint8_t id[SQT_COMM_ID_BYTES]);
/* single-process form: ctxs[0..n) on n distinct devices, ncclCommInitAll underneath */
int sqt_comm_init_all(sqt_ctx **ctxs, int n);
/* render on every context of a single-process group concurrently (rank 0 gets the image) */
int sqt_render_group(sqt_ctx **ctxs, int n, const sqt_camera *cam, const sqt_render_params *params,
                     uint8_t *rgb8_out, float *accum_out, sqt_stats *stats_or_null);

/* measured roofline denominators on this device (microbenchmarks) ------------- */
int sqt_measure_fp32_peak(sqt_ctx *ctx, double *gops_per_s);     /* non-fused FADD/FMUL issue rate, Gop/s */
int sqt_measure_l2_bandwidth(sqt_ctx *ctx, double *gb_per_s);    /* L2-resident 128-bit read bandwidth */

/* options: SQT_OPT_LEAF_CULL (default 1) -- skip leaves whose conservatively enlarged tight box the ray
 * misses.  Exact; 0 reproduces the reference's triangle-test count one for one. */
#define SQT_OPT_LEAF_CULL 1
/* SQT_OPT_SPHERE_BVH (default 1) -- extension: find the spheres a ray can hit through a bounding-volume hierarchy built at
 * sqt_upload_spheres instead of testing every sphere for every ray.  Exact (same closest surface, same tie-break); 0
 * keeps the literal definition. */
#define SQT_OPT_SPHERE_BVH 2
int sqt_set_option(sqt_ctx *ctx, int option, int value);

/* introspection used by tests */
int sqt_device_info(sqt_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, char name_out[128]);

#ifdef __cplusplus
}
#endif
#endif /* SQT_H */
