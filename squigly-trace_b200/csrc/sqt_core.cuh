// sqt_core.cuh -- per-ray / per-path device logic of the squigly-trace B200 backend.
//
// Everything here is SQT_HD (__host__ __device__) so that tests/ can compile the same logic with
// g++ and check it against the oracle on the CPU box (tests/emu); the product only ever runs the
// __device__ instantiation from sqt_kernels.cu.
//
// Exactness rules (hard part 2): every FP32 operation on the intersection and shading
// path is an explicit round-to-nearest add/sub/mul/div/sqrt (X* wrappers = __f*_rn on the device,
// never contracted into FMA); min/max follow the Haskell class defaults whenever a NaN could be
// involved; comparisons are written exactly as in the reference.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SQT_HD __host__ __device__ __forceinline__
#define SQT_HD_NOINLINE __host__ __device__
#else
#define SQT_HD inline
#define SQT_HD_NOINLINE
#endif

#if defined(__CUDA_ARCH__)
#define XADD(a, b) __fadd_rn((a), (b))
#define XSUB(a, b) __fsub_rn((a), (b))
#define XMUL(a, b) __fmul_rn((a), (b))
#define XDIV(a, b) __fdiv_rn((a), (b))
#define XRCP(a) __frcp_rn((a))
#define XSQRT(a) __fsqrt_rn((a))
#define SQT_FMIN(a, b) fminf((a), (b))
#define SQT_FMAX(a, b) fmaxf((a), (b))
#define SQT_LDG4(p) __ldg((const float4 *)(p))
#else
// host build (tests/emu): compiled with -ffp-contract=off -fno-fast-math
#define XADD(a, b) ((float)((float)(a) + (float)(b)))
#define XSUB(a, b) ((float)((float)(a) - (float)(b)))
#define XMUL(a, b) ((float)((float)(a) * (float)(b)))
#define XDIV(a, b) ((float)((float)(a) / (float)(b)))
#define XRCP(a) ((float)(1.0f / (float)(a)))
#define XSQRT(a) sqrtf((a))
#define SQT_FMIN(a, b) fminf((a), (b))
#define SQT_FMAX(a, b) fmaxf((a), (b))
#define SQT_LDG4(p) (*(const float4 *)(p))
#if !defined(__CUDACC__)
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
#endif
#endif

namespace sqt {

// ---------------------------------------------------------------------------- device records
// Branch node, 64 B = 4 x 128-bit loads.  `lo`/`hi` are the node's own clipped box, i.e. the `bbox`
// argument intersectBIH' receives for this node (BIH.hs:111,130-141); Lhi = hi with hi[axis] := lmax is
// the upper corner of the left child's box, Rlo = lo with lo[axis] := rmin the lower corner of the right
// child's box.  All are static properties of the tree, derived at upload time by copying planes (no
// arithmetic).  Storing both corners makes the two child slab tests axis-free (no per-lane selects).
//   q0 = (lo.x, lo.y, lo.z, hi.x)   q1 = (hi.y, hi.z, Lhi.x, Lhi.y)   q2 = (Lhi.z, Rlo.x, Rlo.y, Rlo.z)
//   q3 = (left, right, lmeta, rmeta) as u32 bits:
//        child is Branch: ref = index into the branch array, meta = 0
//        child is Leaf  : ref = index into the leaf array,   meta = kLeaf | count
// Leaf record, 32 B = 2 x 128-bit loads: the TIGHT bounding box of the leaf's triangles and their longest
// edge, used only by the conservative leaf culling below (never by the reference algorithm):
//   b0 = (lo.x, lo.y, lo.z, hi.x)   b1 = (hi.y, hi.z, longest edge E, first triangle as u32 bits)
//        lmeta additionally carries the split axis in bits 27..28
// Traversal stack entries (the top word tells them apart):
//   phase A (near subtree in flight), 1 word : index of the branch (< 2^30, so kPhaseB is clear); plane, far child and
//                                              its meta are re-read from the node when the near subtree returns
//   phase B (near hit parked, far subtree in flight), 3 words : { t bits, dist bits, tri | kPhaseB }
constexpr uint32_t kLeaf = 0x80000000u;
constexpr uint32_t kPhaseB = 0x40000000u;
constexpr uint32_t kAxisShift = 27;
constexpr uint32_t kCountMask = 0x07ffffffu;
constexpr int kNodeQuads = 4;
constexpr int kStackWords = 3 * 48;            // at most one entry (1 or 3 words) per tree level
#ifndef SQT_MAX_DEPTH
#define SQT_MAX_DEPTH 64
#endif

struct SceneView {
    const float4 *nodes;     // kNodeQuads float4 per branch
    const float4 *leaves;    // 2 float4 per leaf
    const float4 *tris;      // 3 float4 per triangle: (v0.xyz,e1.x) (e1.yz,e2.xy) (e2.z, mat, orig, pad)
    const float4 *mats;      // 3 float4 per material: (refl, surf.rgb) (emissive, emit.rgb) (ec.rgb, flags)
    const float4 *spheres;   // extension: 2 float4 per sphere: (center.xyz, radius) (material bits, -, -, -)
    uint32_t n_spheres;
    float root_lo[3], root_hi[3];
    uint32_t n_branches, n_tris, n_mats;
    uint32_t root_is_leaf;   // tree = Leaf: no box test at all (BIH.hs:105)
    uint32_t leaf_cull;      // 1 = skip leaves whose enlarged tight box the ray provably misses (exact, see enter_leaf)
};

struct Ray { float ox, oy, oz, dx, dy, dz; };
struct Hit { int tri; float t; float dist; };      // tri = index in leaf order, -1 = Nothing
struct Counters {
    unsigned long long branch_visits, child_box_tests, tri_tests, rays, leaves_culled;
    unsigned long long mt_pass_a, mt_pass_u, mt_pass_v, mt_accept;      // triangle tests that got past each guard
};

SQT_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; __builtin_memcpy(&u, &f, 4); return u;
#endif
}
SQT_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; __builtin_memcpy(&f, &u, 4); return f;
#endif
}

// Haskell Ord Float class defaults: max x y = if x <= y then y else x ; min x y = if x <= y then x else y
SQT_HD float hs_max(float x, float y) { return (x <= y) ? y : x; }
SQT_HD float hs_min(float x, float y) { return (x <= y) ? x : y; }
// compare a b == GT  (anything against NaN is GT)
SQT_HD bool cmp_gt(float a, float b) { return !(a < b) && !(a == b); }
SQT_HD bool finite_f(float x) { return fabsf(x) < INFINITY; }   // false for inf and NaN

// V3.hs:25-26  dot = (a*d)+(b*e)+(c*f)
SQT_HD float dot3(float a, float b, float c, float d, float e, float f) {
    return XADD(XADD(XMUL(a, d), XMUL(b, e)), XMUL(c, f));
}

// ------------------------------------------------------------------------------- slab tests
// Geometry.hs:166-177, literal (used for the root box and for rays where a NaN can appear)
SQT_HD bool slab_exact(float lx, float ly, float lz, float hx, float hy, float hz, const Ray &r,
                       float dfx, float dfy, float dfz) {
    float t1 = XMUL(XSUB(lx, r.ox), dfx), t2 = XMUL(XSUB(hx, r.ox), dfx);
    float t3 = XMUL(XSUB(ly, r.oy), dfy), t4 = XMUL(XSUB(hy, r.oy), dfy);
    float t5 = XMUL(XSUB(lz, r.oz), dfz), t6 = XMUL(XSUB(hz, r.oz), dfz);
    float tmin = hs_max(hs_max(hs_min(t1, t2), hs_min(t3, t4)), hs_min(t5, t6));
    float tmax = hs_min(hs_min(hs_max(t1, t2), hs_max(t3, t4)), hs_max(t5, t6));
    return tmax > 0.0f && tmin < tmax;
}

// branch-free 3-way select (the compiler turns the ?: chain into divergent branches otherwise)
SQT_HD float sel3(int ax, float x, float y, float z) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("{\n\t.reg .pred p1, p2;\n\tsetp.eq.s32 p1, %4, 1;\n\tsetp.eq.s32 p2, %4, 2;\n\t"
        "selp.f32 %0, %2, %1, p1;\n\tselp.f32 %0, %3, %0, p2;\n\t}"
        : "=&f"(r) : "f"(x), "f"(y), "f"(z), "r"(ax));
    return r;
#else
    return ax == 0 ? x : (ax == 1 ? y : z);
#endif
}

// Both child boxes of a branch (BIH.hs:128-141): left = (lo, Lhi), right = (Rlo, hi).  `safe` rays (all
// 1/dir finite, origin finite) cannot produce a NaN slab value, and without NaNs min/max are exact,
// order-free operations whose zero sign never reaches the two comparisons -- so the hardware FMNMX path is
// used.  Unsafe rays take the literal Geometry.hs:166-177 path.
SQT_HD void slab_children(const float4 &q0, const float4 &q1, const float4 &q2, const Ray &r, float dfx, float dfy,
                          float dfz, bool safe, bool &hit_l, bool &hit_r) {
    if (safe) {
        const float lx = XMUL(XSUB(q0.x, r.ox), dfx), ly = XMUL(XSUB(q0.y, r.oy), dfy), lz = XMUL(XSUB(q0.z, r.oz), dfz);
        const float hx = XMUL(XSUB(q0.w, r.ox), dfx), hy = XMUL(XSUB(q1.x, r.oy), dfy), hz = XMUL(XSUB(q1.y, r.oz), dfz);
        const float ax_ = XMUL(XSUB(q1.z, r.ox), dfx), ay = XMUL(XSUB(q1.w, r.oy), dfy), az = XMUL(XSUB(q2.x, r.oz), dfz);
        const float bx = XMUL(XSUB(q2.y, r.ox), dfx), by = XMUL(XSUB(q2.z, r.oy), dfy), bz = XMUL(XSUB(q2.w, r.oz), dfz);
        const float tmin_l = SQT_FMAX(SQT_FMAX(SQT_FMIN(lx, ax_), SQT_FMIN(ly, ay)), SQT_FMIN(lz, az));
        const float tmax_l = SQT_FMIN(SQT_FMIN(SQT_FMAX(lx, ax_), SQT_FMAX(ly, ay)), SQT_FMAX(lz, az));
        const float tmin_r = SQT_FMAX(SQT_FMAX(SQT_FMIN(bx, hx), SQT_FMIN(by, hy)), SQT_FMIN(bz, hz));
        const float tmax_r = SQT_FMIN(SQT_FMIN(SQT_FMAX(bx, hx), SQT_FMAX(by, hy)), SQT_FMAX(bz, hz));
        hit_l = tmax_l > 0.0f && tmin_l < tmax_l;
        hit_r = tmax_r > 0.0f && tmin_r < tmax_r;
    } else {
        hit_l = slab_exact(q0.x, q0.y, q0.z, q1.z, q1.w, q2.x, r, dfx, dfy, dfz);
        hit_r = slab_exact(q2.y, q2.z, q2.w, q0.w, q1.x, q1.y, r, dfx, dfy, dfz);
    }
}

// ------------------------------------------------------------------------- Moller-Trumbore
// Geometry.hs:117-142 with edge1/edge2 precomputed.  Guard order a -> u -> v -> t.
SQT_HD bool moller_trumbore(const float4 &a0, const float4 &a1, const float4 &a2, const Ray &r, float &t_out,
                            float &dist_out, int &stage) {
    const float eps = 0.0001f;
    const float v0x = a0.x, v0y = a0.y, v0z = a0.z;
    const float e1x = a0.w, e1y = a1.x, e1z = a1.y;
    const float e2x = a1.z, e2y = a1.w, e2z = a2.x;
    // h = rayDir `cross` edge2
    float hx = XSUB(XMUL(r.dy, e2z), XMUL(r.dz, e2y));
    float hy = XSUB(XMUL(r.dz, e2x), XMUL(r.dx, e2z));
    float hz = XSUB(XMUL(r.dx, e2y), XMUL(r.dy, e2x));
    float a = dot3(e1x, e1y, e1z, hx, hy, hz);
    stage = 0;
    if (a > -eps && a < eps) return false;
    stage = 1;
    float f = XRCP(a);
    float sx = XSUB(r.ox, v0x), sy = XSUB(r.oy, v0y), sz = XSUB(r.oz, v0z);
    float u = XMUL(f, dot3(sx, sy, sz, hx, hy, hz));
    if (u < 0.0f || u > 1.0f) return false;
    stage = 2;
    // q = s `cross` edge1
    float qx = XSUB(XMUL(sy, e1z), XMUL(sz, e1y));
    float qy = XSUB(XMUL(sz, e1x), XMUL(sx, e1z));
    float qz = XSUB(XMUL(sx, e1y), XMUL(sy, e1x));
    float v = XMUL(f, dot3(r.dx, r.dy, r.dz, qx, qy, qz));
    if (v < 0.0f || XADD(u, v) > 1.0f) return false;
    stage = 3;
    float t = XMUL(f, dot3(e2x, e2y, e2z, qx, qy, qz));
    if (!(t > eps)) return false;
    stage = 4;
    // outInter = rayVert + t *^ rayDir ; rayDist = norm (outInter - rayVert)
    float px = XADD(r.ox, XMUL(t, r.dx)), py = XADD(r.oy, XMUL(t, r.dy)), pz = XADD(r.oz, XMUL(t, r.dz));
    float ex = XSUB(px, r.ox), ey = XSUB(py, r.oy), ez = XSUB(pz, r.oz);
    t_out = t;
    dist_out = XSQRT(dot3(ex, ey, ez, ex, ey, ez));
    return true;
}

// ------------------------------------------------------------------------------ traversal
// intersectBIH (BIH.hs:101-141) as an explicit-stack state machine that visits exactly the
// subtrees the recursion visits, in the same order, and combines results with the same rules:
//   * the own-box test of BIH.hs:112 is evaluated for the root only -- for every other branch it
//     repeats, with identical operands, the test its parent just passed at BIH.hs:128-129;
//   * a branch whose two children are both hit pushes one word (its index, "phase A") and enters
//     the near child; when that subtree returns, isClose (BIH.hs:121-123) is evaluated on the near
//     subtree's OWN result; if the far child must still be visited and near had a hit, that hit is
//     parked on the stack ("phase B", 3 words) and merged with min' when the far subtree returns.
//
// The machine is cut into three kinds of unit step so that a warp can run them in lock step
// (sqt_backend.cu: all lanes do traversal steps together, then all lanes do triangle steps together):
//   ST_DESC  enter subtree (child, meta), a Branch: test both child boxes, pick the next subtree
//   ST_ENTER enter subtree (child, meta), a Leaf: fetch its record, (conservatively) cull it or become ST_LEAF
//   ST_LEAF  test ONE triangle of the current leaf, walking from the last to the first
//   ST_RET   a subtree returned `cur`: pop ONE stack entry and act on it
//   ST_DONE  the ray is finished, result in `cur`
enum : int { ST_DONE = 0, ST_DESC = 1, ST_LEAF = 2, ST_RET = 3, ST_EXIT = 4, ST_ENTER = 5 };

struct TravLane {
    Ray r;
    float dfx, dfy, dfz;        // 1/dir (Geometry.hs:168), IEEE
    uint32_t child, meta;       // ST_DESC: subtree to enter ; ST_LEAF: child = first triangle of the leaf
    int i;                      // ST_LEAF: offset of the next triangle to test (counts down to 0)
    Hit cur;                    // result of the subtree that just returned / running best of the current leaf
    int sp;
    int state;
    uint32_t sgn;               // bit k set iff dir[k] > 0 (leftToRight on axis k, BIH.hs:127)
    float dfac;                 // |d|_1 * (1 + |d|_1), factor of the leaf-culling error bound
    bool safe;
    uint32_t *stack;            // kStackWords words of lane-private (local) memory, owned by the caller
};

// ---- extension: analytic spheres (no reference counterpart; semantics restated in oracle/oracle.c) ----------------
// oc = o - c ; A = d.d ; B = oc.d ; C = oc.oc - r*r ; disc = B*B - A*C ; disc < 0 -> Nothing
// t0 = (-B - sqrt disc)/A ; t1 = (-B + sqrt disc)/A ; t = first of t0, t1 that is > eps (1e-4, as mollerTrumbore)
// point = o + t *^ d ; dist = norm (point - o)     -- same conventions as Geometry.hs:134,141
SQT_HD bool ray_sphere(const float4 &s0, const Ray &r, float &t_out, float &dist_out) {
    const float eps = 0.0001f;
    const float ocx = XSUB(r.ox, s0.x), ocy = XSUB(r.oy, s0.y), ocz = XSUB(r.oz, s0.z);
    const float A = dot3(r.dx, r.dy, r.dz, r.dx, r.dy, r.dz);
    const float B = dot3(ocx, ocy, ocz, r.dx, r.dy, r.dz);
    const float C = XSUB(dot3(ocx, ocy, ocz, ocx, ocy, ocz), XMUL(s0.w, s0.w));
    const float disc = XSUB(XMUL(B, B), XMUL(A, C));
    if (!(disc >= 0.0f)) return false;
    const float sq = XSQRT(disc);
    const float t0 = XDIV(XSUB(-B, sq), A), t1 = XDIV(XADD(-B, sq), A);
    const float t = t0 > eps ? t0 : t1;
    if (!(t > eps)) return false;
    const float px = XADD(r.ox, XMUL(t, r.dx)), py = XADD(r.oy, XMUL(t, r.dy)), pz = XADD(r.oz, XMUL(t, r.dz));
    const float ex = XSUB(px, r.ox), ey = XSUB(py, r.oy), ez = XSUB(pz, r.oz);
    t_out = t;
    dist_out = XSQRT(dot3(ex, ey, ez, ex, ey, ez));
    return true;
}

// The BIH part of a ray is finished with `cur`: fold in the spheres (candidates in order [BIH hit, sphere 0, sphere 1, ..],
// minimumBy (comparing dist): an earlier candidate wins ties), then the ray is ST_DONE.  Surface n_tris + k = sphere k.
SQT_HD void finish_ray(const SceneView &sc, TravLane &L) {
    for (uint32_t k = 0; k < sc.n_spheres; ++k) {
        const float4 s0 = SQT_LDG4(sc.spheres + 2 * (size_t)k);
        float t, dist;
        if (ray_sphere(s0, L.r, t, dist)) {
            if (L.cur.tri < 0 || cmp_gt(L.cur.dist, dist)) { L.cur.tri = (int)(sc.n_tris + k); L.cur.t = t; L.cur.dist = dist; }
        }
    }
    L.state = ST_DONE;
}

// material index and (un-normalised) geometric normal of surface `idx` at the hit point (hx,hy,hz)
SQT_HD uint32_t surface_material(const SceneView &sc, int idx) {
    if ((uint32_t)idx >= sc.n_tris) return f2u(SQT_LDG4(sc.spheres + 2 * (size_t)((uint32_t)idx - sc.n_tris) + 1).x);
    return f2u(SQT_LDG4(sc.tris + 3 * (size_t)idx + 2).y);
}

// intersectBIH b = intersectBIH' (bounds b) (tree b)   (BIH.hs:101-102)
template <bool COUNT>
SQT_HD void start_ray(const SceneView &sc, TravLane &L, Counters *cn) {
    if (COUNT) cn->rays += 1;
    L.cur.tri = -1; L.cur.t = 0.0f; L.cur.dist = 0.0f;
    L.sp = 0;
    if (sc.root_is_leaf) {                       // tree = Leaf: no box test at all (BIH.hs:105)
        L.child = 0u; L.i = (int)sc.n_tris - 1;
        if (COUNT) cn->tri_tests += sc.n_tris;
        if (sc.n_tris) L.state = ST_LEAF; else finish_ray(sc, L);
        return;
    }
    L.dfx = XRCP(L.r.dx); L.dfy = XRCP(L.r.dy); L.dfz = XRCP(L.r.dz);
    L.safe = finite_f(L.dfx) && finite_f(L.dfy) && finite_f(L.dfz) && finite_f(L.r.dx) && finite_f(L.r.dy) &&
             finite_f(L.r.dz) && finite_f(L.r.ox) && finite_f(L.r.oy) && finite_f(L.r.oz);
    if (!slab_exact(sc.root_lo[0], sc.root_lo[1], sc.root_lo[2], sc.root_hi[0], sc.root_hi[1], sc.root_hi[2], L.r, L.dfx,
                    L.dfy, L.dfz)) { finish_ray(sc, L); return; }          // BIH.hs:112 at the root
    L.sgn = (L.r.dx > 0.0f ? 1u : 0u) | (L.r.dy > 0.0f ? 2u : 0u) | (L.r.dz > 0.0f ? 4u : 0u);
    { const float d1 = fabsf(L.r.dx) + fabsf(L.r.dy) + fabsf(L.r.dz); L.dfac = d1 * (1.0f + d1); }
    L.child = 0u; L.meta = 0u; L.state = ST_DESC;
}

// Entering Leaf `L.child` (index into the leaf array) with `count` triangles (BIH.hs:105-109).
//
// Conservative leaf culling (not in the reference; exact by a forward error bound, derivation in DESIGN.md section 5).
// A triangle test can only return Just if |a| >= 1e-4 and the computed u, v, u+v pass their guards and t > 1e-4
// (Geometry.hs:117-142).  With eps = 2^-24, E the longest edge in the leaf and s = origin - v0, every numerator is
// a 3-term dot of a 2-term cross and is off by at most 12*eps*|d|*E*(|s|_1 + E); after the division by |a| >= 1e-4
// an accepted hit lies, in exact arithmetic, within 0.0143*|d|*E^2*(|s|_1+E) of the triangle (origin at most that
// far behind it).  The leaf is skipped only if the ray misses the leaf's tight box enlarged by
// 0.03*|d|_1*(1+|d|_1)*E^2*(|s|_1+E)  (>= 2x the bound, 1-norm over-estimates)  +  1e-4 + 2^-20*(max|coord| + |s|_1):
// no skipped triangle could have been accepted, so the traversal result is bit-identical (tests: culling on/off
// agree on every ray; both agree with the oracle).  Rays with a zero/denormal/non-finite direction component
// (`safe` false) are never culled.
template <bool COUNT>
SQT_HD void enter_step(const SceneView &sc, TravLane &L, Counters *cn) {
    const uint32_t count = L.meta & kCountMask;
    L.cur.tri = -1;
    if (count == 0u) { L.state = ST_RET; return; }                        // empty leaf -> Nothing (BIH.hs:107)
    const float4 *lp = sc.leaves + 2 * (size_t)L.child;
    const float4 b0 = SQT_LDG4(lp), b1 = SQT_LDG4(lp + 1);
    if (sc.leaf_cull && L.safe) {
        const float E = b1.z;
        const float s1 = fabsf(L.r.ox - 0.5f * (b0.x + b0.w)) + fabsf(L.r.oy - 0.5f * (b0.y + b1.x)) +
                         fabsf(L.r.oz - 0.5f * (b0.z + b1.y)) + ((b0.w - b0.x) + (b1.x - b0.y) + (b1.y - b0.z));
        const float cmax = fmaxf(fmaxf(fmaxf(fabsf(b0.x), fabsf(b0.w)), fmaxf(fabsf(b0.y), fabsf(b1.x))), fmaxf(fabsf(b0.z), fabsf(b1.y)));
        const float m = 0.03f * (s1 + E) * L.dfac * (E * E) + (1.0e-4f + 9.5367431640625e-7f * (cmax + s1));
        const float lx = (b0.x - m - L.r.ox) * L.dfx, hx = (b0.w + m - L.r.ox) * L.dfx;
        const float ly = (b0.y - m - L.r.oy) * L.dfy, hy = (b1.x + m - L.r.oy) * L.dfy;
        const float lz = (b0.z - m - L.r.oz) * L.dfz, hz = (b1.y + m - L.r.oz) * L.dfz;
        const float tmin = SQT_FMAX(SQT_FMAX(SQT_FMIN(lx, hx), SQT_FMIN(ly, hy)), SQT_FMIN(lz, hz));
        const float tmax = SQT_FMIN(SQT_FMIN(SQT_FMAX(lx, hx), SQT_FMAX(ly, hy)), SQT_FMAX(lz, hz));
        // keep the leaf unless the ray clearly misses; a NaN (cannot happen for safe rays) keeps it too
        if (tmax < 0.0f || tmin > tmax) {
            if (COUNT) cn->leaves_culled += 1;
            L.state = ST_RET;
            return;
        }
    }
    if (COUNT) cn->tri_tests += count;
    L.child = f2u(b1.w);
    L.i = (int)count - 1;
    L.state = ST_LEAF;
}

// Word w of the lane's stack.  PLANE = 8: a plain array (lane-private local memory).  PLANE > 8: the stacks of a group of
// PLANE/8 rays are interleaved in 32-byte granules -- granule g of every ray of the group is contiguous -- so that the few
// granules in use (a stack is usually < 8 words deep) of all rays share cache lines instead of each ray owning lines of its
// own (k_paths_pool: the stacks of all resident pool slots then fit L2).
template <int PLANE>
SQT_HD uint32_t &stack_word(TravLane &L, int w) {
    return L.stack[PLANE == 8 ? w : (w >> 3) * PLANE + (w & 7)];
}

template <bool COUNT, int PLANE = 8>
SQT_HD void desc_step(const SceneView &sc, TravLane &L, Counters *cn) {
    const float4 *np = sc.nodes + kNodeQuads * (size_t)L.child;
    const float4 q0 = SQT_LDG4(np), q1 = SQT_LDG4(np + 1), q2 = SQT_LDG4(np + 2), q3 = SQT_LDG4(np + 3);
    const uint32_t left = f2u(q3.x), right = f2u(q3.y), lmeta = f2u(q3.z), rmeta = f2u(q3.w);
    const uint32_t ax = (lmeta >> kAxisShift) & 3u;
    bool hit_l, hit_r;
    slab_children(q0, q1, q2, L.r, L.dfx, L.dfy, L.dfz, L.safe, hit_l, hit_r);
    if (COUNT) { cn->branch_visits += 1; cn->child_box_tests += 2; }
    if (!(hit_l || hit_r)) { L.cur.tri = -1; L.state = ST_RET; return; }
    const bool ltr = ((L.sgn >> ax) & 1u) != 0u;                            // BIH.hs:127
    const bool both = hit_l && hit_r;
    const bool go_left = both ? ltr : hit_l;                                // near child first (BIH.hs:124-126)
    const uint32_t lm = lmeta & (kLeaf | kCountMask);
    if (both) {                                 // phase A frame: ONE word, this branch; ret_step re-reads plane and far child from the node
        stack_word<PLANE>(L, L.sp) = L.child;
        L.sp += 1;
    }
    L.child = go_left ? left : right;
    L.meta = go_left ? lm : rmeta;
    if (L.meta & kLeaf) L.state = ST_ENTER;
}

// One triangle of the leaf (BIH.hs:105-109): V.mapMaybe over the leaf's triangles, minimumBy (comparing dist).
// base-4.9 minimumBy = foldr1 min' with min' x y = GT -> y ; _ -> x : walk from the last triangle
// to the first, the earlier one wins unless it is strictly farther.
// the three 128-bit words of triangle `idx` (leaf order)
struct TriData { float4 a0, a1, a2; };
SQT_HD TriData tri_load(const SceneView &sc, uint32_t idx) {
    const float4 *p = sc.tris + 3 * (size_t)idx;
    TriData d; d.a0 = SQT_LDG4(p); d.a1 = SQT_LDG4(p + 1); d.a2 = SQT_LDG4(p + 2);
    return d;
}
// test the already loaded triangle L.child + L.i and advance (split from the load so that callers can prefetch)
template <bool COUNT>
SQT_HD void tri_apply(TravLane &L, const TriData &d, Counters *cn) {
    const uint32_t idx = L.child + (uint32_t)L.i;
    float t, dist;
    int stage;
    if (moller_trumbore(d.a0, d.a1, d.a2, L.r, t, dist, stage)) {
        if (L.cur.tri < 0 || !cmp_gt(dist, L.cur.dist)) { L.cur.tri = (int)idx; L.cur.t = t; L.cur.dist = dist; }
    }
    if (COUNT) { cn->mt_pass_a += stage >= 1; cn->mt_pass_u += stage >= 2; cn->mt_pass_v += stage >= 3; cn->mt_accept += stage >= 4; }
    if (--L.i < 0) L.state = ST_RET;
}
template <bool COUNT>
SQT_HD void tri_step(const SceneView &sc, TravLane &L, Counters *cn) {
    const TriData d = tri_load(sc, L.child + (uint32_t)L.i);
    tri_apply<COUNT>(L, d, cn);
}
// A subtree returned `cur`: pop entries until one of them sends the lane into a far subtree (-> ST_DESC / ST_ENTER) or
// the stack is empty (-> ST_DONE).  Two kinds of entry, told apart by kPhaseB in the top word:
//   phase A, 1 word  : index of a branch whose NEAR subtree just returned (both children were hit);
//   phase B, 3 words : (t, dist, tri | kPhaseB), the parked near hit of a branch whose FAR subtree just returned.
// Keeping phase A at one word (instead of caching plane / far child / meta in the entry) cuts the stack traffic to a
// third; the branch's node is re-read on the way back -- it was read on the way down and is normally still in L1/L2.
template <int PLANE = 8>
SQT_HD void ret_step(const SceneView &sc, TravLane &L) {
    for (;;) {
        if (L.sp == 0) { finish_ray(sc, L); return; }
        const uint32_t w = stack_word<PLANE>(L, L.sp - 1);
        if (w & kPhaseB) {                                                  // far subtree returned: min' near far
            const uint32_t w0 = stack_word<PLANE>(L, L.sp - 3), w1 = stack_word<PLANE>(L, L.sp - 2);
            L.sp -= 3;
            if (L.cur.tri < 0 || !cmp_gt(u2f(w1), L.cur.dist)) { L.cur.tri = (int)(w & ~kPhaseB); L.cur.dist = u2f(w1); L.cur.t = u2f(w0); }
            continue;
        }
        L.sp -= 1;
        // near subtree of branch w returned
        const float4 *np = sc.nodes + kNodeQuads * (size_t)w;
        const float4 q3 = SQT_LDG4(np + 3);
        const uint32_t lmeta = f2u(q3.z);
        const int ax = (int)((lmeta >> kAxisShift) & 3u);
        const bool ltr = ((L.sgn >> ax) & 1u) != 0u;                        // BIH.hs:127: the near child was left iff leftToRight
        if (L.cur.tri >= 0) {
            const float4 q1 = SQT_LDG4(np + 1), q2 = SQT_LDG4(np + 2);
            const float plane = ltr ? sel3(ax, q2.y, q2.z, q2.w) : sel3(ax, q1.z, q1.w, q2.x);      // rmin : lmax
            const float p = XADD(sel3(ax, L.r.ox, L.r.oy, L.r.oz), XMUL(L.cur.t, sel3(ax, L.r.dx, L.r.dy, L.r.dz)));   // intersectPoint on ax
            const bool close = ltr ? (p < plane) : (p > plane);             // BIH.hs:121-123
            if (close) continue;
            stack_word<PLANE>(L, L.sp) = f2u(L.cur.t); stack_word<PLANE>(L, L.sp + 1) = f2u(L.cur.dist);
            stack_word<PLANE>(L, L.sp + 2) = (uint32_t)L.cur.tri | kPhaseB;
            L.sp += 3;
        }
        L.child = ltr ? f2u(q3.y) : f2u(q3.x);
        L.meta = ltr ? f2u(q3.w) : (lmeta & (kLeaf | kCountMask));
        L.state = (L.meta & kLeaf) ? ST_ENTER : ST_DESC;
        return;
    }
}

// one lane alone, to completion (host emulation and the odd single ray)
template <bool COUNT>
SQT_HD Hit traverse(const SceneView &sc, const Ray &r, Counters *cn) {
    uint32_t stack[kStackWords];
    TravLane L;
    L.stack = stack;
    L.r = r;
    start_ray<COUNT>(sc, L, cn);
    while (L.state != ST_DONE) {
        if (L.state == ST_RET) ret_step(sc, L);
        if (L.state == ST_DESC) desc_step<COUNT>(sc, L, cn);
        if (L.state == ST_ENTER) enter_step<COUNT>(sc, L, cn);
        if (L.state == ST_LEAF) tri_step<COUNT>(sc, L, cn);
    }
    return L.cur;
}

// ------------------------------------------------------------------------------------- RNG
// Philox4x32-10; stream = the integer the reference hands to mkTFGen (Lib.hs:85-86), j = index of
// the Word32 in that generator's output.  counter = (stream lo, stream hi, j/4, "SQTR"), key = seed.
SQT_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#if defined(__CUDA_ARCH__)
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
#else
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t h0 = (uint32_t)(p0 >> 32), l0 = (uint32_t)p0, h1 = (uint32_t)(p1 >> 32), l1 = (uint32_t)p1;
#endif
        uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Lib.hs:183-188 at (0,1): Word32 -> Float (round to nearest) / 2^32 ; inclusive of 1.0
SQT_HD float random_r01(uint32_t n) {
#if defined(__CUDA_ARCH__)
    return XMUL(__uint2float_rn(n), 2.3283064365386963e-10f);
#else
    return XMUL((float)n, 2.3283064365386963e-10f);
#endif
}

// ------------------------------------------------------------------------------------ trig
// "sqt trig" (DESIGN.md section 6): plain binary32 polynomials, fixed operation order, no FMA, so the
// device and the oracle's restatement agree bit for bit.  Arguments: x in [0, 2*pi].
SQT_HD void sqt_sincos(float x, float &s_out, float &c_out) {
    const int k = (int)XADD(XMUL(x, 0.63661975f), 0.5f);
    const float kf = (float)k;
    const float r = XSUB(XSUB(XSUB(x, XMUL(kf, 1.5703125f)), XMUL(kf, 4.837512969970703125e-4f)),
                         XMUL(kf, 7.54978995489188216e-8f));
    const float z = XMUL(r, r);
    float s = XADD(XMUL(-1.9515295891e-4f, z), 8.3321608736e-3f);
    s = XSUB(XMUL(s, z), 1.6666654611e-1f);
    s = XADD(XMUL(XMUL(s, z), r), r);
    float c = XSUB(XMUL(2.443315711809948e-5f, z), 1.388731625493765e-3f);
    c = XADD(XMUL(c, z), 4.166664568298827e-2f);
    c = XMUL(XMUL(c, z), z);
    c = XSUB(c, XMUL(0.5f, z));
    c = XADD(c, 1.0f);
    const int q = k & 3;
    s_out = q == 0 ? s : (q == 1 ? c : (q == 2 ? -s : -c));
    c_out = q == 0 ? c : (q == 1 ? -s : (q == 2 ? -c : s));
}
SQT_HD float sqt_asin_core(float x, float z) {
    float p = XADD(XMUL(4.2163199048e-2f, z), 2.4181311049e-2f);
    p = XADD(XMUL(p, z), 4.5470025998e-2f);
    p = XADD(XMUL(p, z), 7.4953002686e-2f);
    p = XADD(XMUL(p, z), 1.6666752422e-1f);
    return XADD(XMUL(XMUL(p, z), x), x);
}
SQT_HD float sqt_acos(float x) {
    if (x < -0.5f) {
        const float z = XMUL(0.5f, XADD(1.0f, x)), y = XSQRT(z);
        return XSUB(3.14159265358979323846f, XMUL(2.0f, sqt_asin_core(y, z)));
    }
    if (x > 0.5f) {
        const float z = XMUL(0.5f, XSUB(1.0f, x)), y = XSQRT(z);
        return XMUL(2.0f, sqt_asin_core(y, z));
    }
    return XSUB(1.57079632679489661923f, sqt_asin_core(x, XMUL(x, x)));
}
SQT_HD float sqt_atan(float x) {   // x >= 0
    float y;
    if (x > 2.414213562373095f) { y = 1.57079632679489661923f; x = -XDIV(1.0f, x); }
    else if (x > 0.4142135623730950f) { y = 0.78539816339744830962f; x = XDIV(XSUB(x, 1.0f), XADD(x, 1.0f)); }
    else y = 0.0f;
    const float z = XMUL(x, x);
    float p = XSUB(XMUL(8.05374449538e-2f, z), 1.38776856032e-1f);
    p = XADD(XMUL(p, z), 1.99777106478e-1f);
    p = XSUB(XMUL(p, z), 3.33329491539e-1f);
    return XADD(y, XADD(XMUL(XMUL(p, z), x), x));
}

// -------------------------------------------------------------------------------- shading
struct RenderParams {
    int rows, cols, xdiv, ydiv, seed_stride, spp, max_depth, mode;
    unsigned long long seed;
    int rank, world, split_samples, primary_reuse;
    float cam_pos[3], cam_rot[9];
    int terminate_on_black;     // host proved C*L stays finite: a surface with surfColor == 0 ends the path exactly
};

// makeRay (Lib.hs:107-114) + rotVert (Geometry.hs:104-107: row vector x matrix, sum = foldl (+) 0)
SQT_HD Ray make_ray(const RenderParams &p, int y, int x) {
    const float ww = (float)p.xdiv, hh = (float)p.ydiv;
    const float xo = XDIV(XSUB((float)x, XDIV(ww, 2.0f)), ww);
    const float yo = XDIV(XSUB(XDIV(hh, 2.0f), (float)y), hh);
    const float *R = p.cam_rot;
    Ray r;
    r.ox = p.cam_pos[0]; r.oy = p.cam_pos[1]; r.oz = p.cam_pos[2];
    r.dx = XADD(XADD(XADD(0.0f, XMUL(1.0f, R[0])), XMUL(xo, R[3])), XMUL(yo, R[6]));
    r.dy = XADD(XADD(XADD(0.0f, XMUL(1.0f, R[1])), XMUL(xo, R[4])), XMUL(yo, R[7]));
    r.dz = XADD(XADD(XADD(0.0f, XMUL(1.0f, R[2])), XMUL(xo, R[5])), XMUL(yo, R[8]));
    return r;
}

SQT_HD float hs_signum(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : x); }

// bounceRay (Lib.hs:155-160) at a hit on triangle `tri` (leaf order) with hit parameter t.  The caller
// drew x = draw j (scatter iff reflective < x) and, for a scatter, v = draw j+1: scatter reuses x as the
// azimuth draw and takes v as the polar draw.  Returns the new ray; origin = intersectPoint.
SQT_HD Ray bounce_ray(const SceneView &sc, const Ray &in, int tri, float t, bool scatter, float x, float v) {
    Ray out;
    out.ox = XADD(in.ox, XMUL(t, in.dx)); out.oy = XADD(in.oy, XMUL(t, in.dy)); out.oz = XADD(in.oz, XMUL(t, in.dz));
    float nx, ny, nz;
    if ((uint32_t)tri >= sc.n_tris) {           // extension: sphere, normal = intersectPoint - center
        const float4 s0 = SQT_LDG4(sc.spheres + 2 * (size_t)((uint32_t)tri - sc.n_tris));
        nx = XSUB(out.ox, s0.x); ny = XSUB(out.oy, s0.y); nz = XSUB(out.oz, s0.z);
    } else {
        const float4 *p = sc.tris + 3 * (size_t)tri;
        const float4 a0 = SQT_LDG4(p), a1 = SQT_LDG4(p + 1), a2 = SQT_LDG4(p + 2);
        const float e1x = a0.w, e1y = a1.x, e1z = a1.y, e2x = a1.z, e2y = a1.w, e2z = a2.x;
        // normal = (b - a) `cross` (c - a)   Geometry.hs:79-80
        nx = XSUB(XMUL(e1y, e2z), XMUL(e1z, e2y));
        ny = XSUB(XMUL(e1z, e2x), XMUL(e1x, e2z));
        nz = XSUB(XMUL(e1x, e2y), XMUL(e1y, e2x));
    }
    if (scatter) {                              // scatterRay, Lib.hs:166-172 ; randomVector Lib.hs:192-198
        const float th = XMUL(6.28318530717958647692f, x);             // 2 * pi * u
        const float ph = sqt_acos(XSUB(XMUL(2.0f, v), 1.0f));
        float sth, cth, sph, cph;
        sqt_sincos(th, sth, cth);
        sqt_sincos(ph, sph, cph);
        const float ndx = XMUL(cth, sph), ndy = XMUL(sth, sph), ndz = cph;
        const float so = hs_signum(dot3(in.dx, in.dy, in.dz, nx, ny, nz));
        const float sn = hs_signum(dot3(ndx, ndy, ndz, nx, ny, nz));
        const bool flip = (so == sn);
        out.dx = flip ? -ndx : ndx; out.dy = flip ? -ndy : ndy; out.dz = flip ? -ndz : ndz;
    } else {                                    // reflectRay, Lib.hs:176-181
        const float nn = XSQRT(dot3(nx, ny, nz, nx, ny, nz));
        const float ux = XDIV(nx, nn), uy = XDIV(ny, nn), uz = XDIV(nz, nn);
        const float k = XMUL(2.0f, dot3(ux, uy, uz, in.dx, in.dy, in.dz));
        out.dx = XSUB(in.dx, XMUL(k, ux)); out.dy = XSUB(in.dy, XMUL(k, uy)); out.dz = XSUB(in.dz, XMUL(k, uz));
    }
    return out;
}

// rgbFloatToPixelRGB (Lib.hs:93-104); floor :: Float -> Word8 wraps through Integer, NaN -> 0
SQT_HD uint8_t to_w8(float s255) {
    if (s255 != s255) return 0;
    const float f = floorf(s255);
    long long w = (fabsf(f) < 9.0e18f) ? (long long)f : 0;
    uint32_t b = (uint32_t)((unsigned long long)w & 0xffull);
    return (uint8_t)(b < 255u ? b : 255u);
}
SQT_HD void tone_map(float r, float g, float b, uint8_t out[3]) {
    const float maxc = hs_max(hs_max(r, g), b), minc = hs_min(hs_min(r, g), b);
    const float lightness = XMUL(0.5f, XADD(maxc, minc));
    const float intensity = XDIV(sqt_atan(lightness), 1.57079632679489661923f);
    const float k = XDIV(intensity, maxc);
    out[0] = to_w8(XMUL(XMUL(k, r), 255.0f));
    out[1] = to_w8(XMUL(XMUL(k, g), 255.0f));
    out[2] = to_w8(XMUL(XMUL(k, b), 255.0f));
}

}  // namespace sqt
