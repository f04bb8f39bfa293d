// sqt_paths.cuh -- the per-lane integrator loop (renderPixel / raytrace / raycast, Lib.hs:79-151).
//
// One lane owns one pixel at a time and runs its samples k = k0..k1-1 in order, so the per-pixel
// radiance sum is the reference's sequential `sum = foldl (+) 0` (Lib.hs:88).  A lane whose path ends
// (miss, depth cut, black surface) immediately starts its next sample -- or fetches its next pixel --
// inside the inner "until I have a ray" loop, so that every lane enters the traversal with a live ray
// (path regeneration; keeps warps full although path lengths differ).
//
// SQT_HD like sqt_core.cuh: tests/emu compiles it for the host, the product only runs it on the device.
#pragma once
#include "sqt_core.cuh"

namespace sqt {

struct PathStats { unsigned long long rays, samples, primary_reused; };

// pixel owned by this rank for work item w (pixel-group partition: groups of 32 consecutive pixels,
// round-robin over ranks).  Returns -1 for the padding of the last group.
SQT_HD long long work_to_pixel(const RenderParams &p, long long w) {
    long long pix;
    if (p.world <= 1 || p.split_samples) pix = w;
    else pix = (((w >> 5) * (long long)p.world + (long long)p.rank) << 5) | (w & 31);
    return pix < (long long)p.rows * (long long)p.cols ? pix : -1;
}
SQT_HD long long work_items(const RenderParams &p) {
    const long long npix = (long long)p.rows * (long long)p.cols;
    if (p.world <= 1 || p.split_samples) return npix;
    const long long groups = (npix + 31) >> 5;
    const long long mine = groups > p.rank ? (groups - p.rank + p.world - 1) / p.world : 0;
    return mine << 5;
}
SQT_HD void sample_range(const RenderParams &p, int &k0, int &k1) {
    k0 = 0; k1 = p.spp;
    if (p.world > 1 && p.split_samples) {
        k0 = (int)((long long)p.spp * p.rank / p.world);
        k1 = (int)((long long)p.spp * (p.rank + 1) / p.world);
    }
}

// radiance of one finished path: L_j = surfColor_j * L_{j+1} + emissive_j *^ emitColor_j, evaluated from
// the deepest shaded bounce outwards exactly like the recursion of Lib.hs:135-137 (L beyond the end = black).
SQT_HD void fold_path(const SceneView &sc, const uint16_t *pm, int last, float &lr, float &lg, float &lb) {
    lr = 0.0f; lg = 0.0f; lb = 0.0f;
    for (int b = last; b >= 0; --b) {
        const float4 *m = sc.mats + 3 * (size_t)pm[b];
        const float4 m0 = SQT_LDG4(m), m2 = SQT_LDG4(m + 2);
        lr = XADD(XMUL(m0.y, lr), m2.x);
        lg = XADD(XMUL(m0.z, lg), m2.y);
        lb = XADD(XMUL(m0.w, lb), m2.z);
    }
}

constexpr uint32_t kMatEmits = 1u;       // emissive *^ emitColor != 0
constexpr uint32_t kMatBlack = 2u;       // surfColor == (0,0,0)

// Fetch: functor returning the next work item (>= 0) or -1 when the queue is empty.
// prim: per pixel (tri, t bits) of the primary hit, or nullptr when primary reuse is off.
template <bool COUNT, class Fetch>
SQT_HD_NOINLINE void render_lane(const SceneView &sc, const RenderParams &p, const int2 *prim, float *accum,
                                 Fetch &fetch, Counters *cn, PathStats &st) {
    int k0, k1;
    sample_range(p, k0, k1);
    uint16_t pm[64];                     // material of every shaded bounce of the current path
    DrawCache dc; dc.block = -1;
    dc.w[0] = dc.w[1] = dc.w[2] = dc.w[3] = 0u;
    long long pixel = -1;
    int k = 0, j = 0;                    // sample index, bounce index of the hit being shaded
    int ptri = -1; float pt = 0.0f;      // primary hit of the current pixel
    Ray pr;                              // primary ray of the current pixel
    pr.ox = pr.oy = pr.oz = pr.dx = pr.dy = pr.dz = 0.0f;
    Ray ray = pr, nr = pr;               // ray that produced the current hit ; ray to trace next
    int htri = -1; float ht = 0.0f;      // current hit
    unsigned long long rix = 0ull;
    float sr = 0.0f, sg = 0.0f, sb = 0.0f;
    bool any_emit = false, have_ray = false, path_over = false, need_pixel = true, start = false, done = false;

    for (;;) {
        while (!have_ray) {
            if (path_over) {
                // ---- finish sample k: add its radiance, in sample order (Lib.hs:87-88)
                if (any_emit) {
                    float lr, lg, lb;
                    fold_path(sc, pm, j, lr, lg, lb);
                    sr = XADD(sr, lr); sg = XADD(sg, lg); sb = XADD(sb, lb);
                }   // else the sample is exactly (+0,+0,+0) and sum + 0 == sum
                st.samples += 1;
                path_over = false;
                if (++k == k1) {
                    accum[3 * pixel] = sr; accum[3 * pixel + 1] = sg; accum[3 * pixel + 2] = sb;
                    need_pixel = true;
                } else start = true;
            }
            if (need_pixel) {
                const long long w = fetch();
                if (w < 0) { done = true; break; }
                pixel = work_to_pixel(p, w);
                if (pixel < 0) continue;
                const int y = (int)(pixel / p.cols), x = (int)(pixel % p.cols);
                pr = make_ray(p, y, x);
                rix = (unsigned long long)p.spp * ((unsigned long long)x + (unsigned long long)y * (unsigned long long)p.seed_stride);
                sr = sg = sb = 0.0f;
                k = k0;
                if (k0 >= k1) continue;
                if (prim) {
                    const int2 ph = prim[pixel];
                    ptri = ph.x; pt = u2f((uint32_t)ph.y);
                    if (ptri < 0) {                       // every sample of this pixel is black (accum pre-zeroed)
                        st.samples += (unsigned long long)(k1 - k0);
                        st.primary_reused += (unsigned long long)(k1 - k0);
                        continue;
                    }
                }
                need_pixel = false;
                start = true;
            }
            if (start) {
                // A sample begins at the cached primary hit (bounce 0 already intersected: the primary ray is
                // the same for every sample of a pixel, Lib.hs:81) or, with primary reuse off, by tracing the
                // primary ray again like Lib.hs:84 does.
                start = false; any_emit = false; dc.block = -1;
                if (prim) { ray = pr; htri = ptri; ht = pt; j = 0; st.primary_reused += 1; }
                else { nr = pr; j = -1; have_ray = true; continue; }
            }
            // ---- shade the hit of bounce j (raytrace, Lib.hs:127-137)
            const float4 *tp = sc.tris + 3 * (size_t)htri;
            const uint32_t mat = f2u(SQT_LDG4(tp + 2).y);
            pm[j] = (uint16_t)mat;
            const float4 *mp = sc.mats + 3 * (size_t)mat;
            const float4 m0 = SQT_LDG4(mp);
            const uint32_t mflags = f2u(SQT_LDG4(mp + 2).w);
            any_emit = any_emit || (mflags & kMatEmits);
            const bool terminal = (j + 1 > p.max_depth - 1) || (p.terminate_on_black && (mflags & kMatBlack));
            if (!terminal) {
                nr = bounce_ray(sc, ray, htri, ht, m0.x, dc, p.seed, rix + (unsigned long long)k, (uint32_t)j);
                have_ray = true;
            } else path_over = true;
        }
        if (done) break;
        // ---- trace one segment (Lib.hs:131)
        const Hit h = traverse<COUNT>(sc, nr, cn);
        st.rays += 1;
        have_ray = false;
        if (h.tri >= 0) { ray = nr; htri = h.tri; ht = h.t; j += 1; }
        else path_over = true;
    }
}

// --cast (Lib.hs:141-151): primary hit + one shadow ray to the hard-coded light; every sample identical.
template <bool COUNT>
SQT_HD void raycast_pixel(const SceneView &sc, const RenderParams &p, long long pixel, float *accum, Counters *cn,
                          PathStats &st) {
    const int y = (int)(pixel / p.cols), x = (int)(pixel % p.cols);
    const Ray r = make_ray(p, y, x);
    int k0, k1;
    sample_range(p, k0, k1);
    float cr = 0.0f, cg = 0.0f, cb = 0.0f;
    const Hit h = traverse<COUNT>(sc, r, cn);
    st.rays += 1;
    if (h.tri >= 0) {
        const float px = XADD(r.ox, XMUL(h.t, r.dx)), py = XADD(r.oy, XMUL(h.t, r.dy)), pz = XADD(r.oz, XMUL(h.t, r.dz));
        Ray s;                                                   // a `to` b = Ray a (b - a), light = V3 0 3 (-1)
        s.ox = px; s.oy = py; s.oz = pz;
        s.dx = XSUB(0.0f, px); s.dy = XSUB(3.0f, py); s.dz = XSUB(-1.0f, pz);
        const float ex = XSUB(px, 0.0f), ey = XSUB(py, 3.0f), ez = XSUB(pz, -1.0f);
        const float dl = XSQRT(dot3(ex, ey, ez, ex, ey, ez));
        const Hit sh = traverse<COUNT>(sc, s, cn);
        st.rays += 1;
        if (!(sh.tri >= 0 && !(sh.dist > dl))) {
            const uint32_t mat = f2u(SQT_LDG4(sc.tris + 3 * (size_t)h.tri + 2).y);
            const float4 m0 = SQT_LDG4(sc.mats + 3 * (size_t)mat);
            const float kk = XDIV(2.0f, dl);
            cr = XMUL(kk, m0.y); cg = XMUL(kk, m0.z); cb = XMUL(kk, m0.w);
        }
    }
    float sr = 0.0f, sg = 0.0f, sb = 0.0f;
    for (int k = k0; k < k1; ++k) { sr = XADD(sr, cr); sg = XADD(sg, cg); sb = XADD(sb, cb); }
    st.samples += (unsigned long long)(k1 > k0 ? k1 - k0 : 0);
    accum[3 * pixel] = sr; accum[3 * pixel + 1] = sg; accum[3 * pixel + 2] = sb;
}

}  // namespace sqt
