// emu.cpp -- TEST HELPER: host (g++) build of the kernel logic in csrc/sqt_core.cuh + sqt_paths.cuh.
//
// The GPU box is scarce, so the per-ray / per-path state machines are written __host__ __device__ and
// this file runs the very same code on the CPU, lane by lane, so that `-m "not gpu"` tests can compare it
// with the independent oracle (oracle/oracle.c) before any GPU time is spent.  It is NOT a fallback: the
// product library (libsqt_b200.so) neither links nor loads it, and it lives under tests/.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../squigly-trace_b200/csrc/sqt_layout.hpp"

using namespace sqt;

struct emu_scene {
    DeviceLayout lay;
    std::vector<float4> tris, spheres, leaves, slabs;
    SphereBvh bvh;
    int sphere_bvh = 1;
    SceneView view;
    std::string err;
};

static long long sbuf_budget = 1ll << 20;      // small on purpose: forces several sample rounds in the tests

struct VecAppend {
    std::vector<int> *v;
    void operator()(long long pixel) { v->push_back((int)pixel); }
};

struct SeqFetch {
    long long next, n;
    long long operator()() { return next < n ? next++ : -1; }
};

extern "C" {

emu_scene *emu_upload(const sqt_scene_desc *d) {
    emu_scene *s = new emu_scene();
    if (build_device_layout(*d, s->lay, s->err)) return s;
    s->tris.resize((size_t)3 * (d->n_tris ? d->n_tris : 1));
    if (d->n_tris) std::memcpy(s->tris.data(), d->tris, (size_t)d->n_tris * 48);
    SceneView v = {};
    s->leaves.resize(2 * s->lay.leaf_first.size());
    for (size_t k = 0; k < s->lay.leaf_first.size(); ++k) {
        make_leaf_record(s->tris.data(), s->lay.leaf_first[k], s->lay.leaf_count[k], s->leaves[2 * k], s->leaves[2 * k + 1]);
        if (s->lay.leaf_count[k] >= kLeafLong) s->tris[3 * (size_t)s->lay.leaf_first[k] + 2].w = u2f(s->lay.leaf_count[k]);
    }
    compute_slabs_host(s->lay, s->leaves, s->slabs);
    v.slabs = s->slabs.data();
    for (int k = 0; k < 3; ++k) v.tame_c[k] = s->lay.tame_c[k];
    v.tame_r = s->lay.tame_r;
    v.nodes = s->lay.nodes.data(); v.boxes = s->lay.boxes.data(); v.tris = s->tris.data(); v.mats = s->lay.mats.data(); v.leaves = s->leaves.data();
    v.leaf_cull = 1; v.planes_finite = (uint32_t)s->lay.planes_finite;
    for (int k = 0; k < 3; ++k) { v.root_lo[k] = d->root_bounds[k]; v.root_hi[k] = d->root_bounds[3 + k]; }
    v.n_branches = s->lay.n_branches; v.n_tris = d->n_tris; v.n_mats = d->n_mats;
    v.root_is_leaf = (d->nodes[0].b & SQT_NODE_LEAF) ? 1u : 0u;
    s->view = v;
    return s;
}
const char *emu_error(emu_scene *s) { return s->err.c_str(); }
void emu_free(emu_scene *s) { delete s; }
int emu_height(emu_scene *s) { return (int)s->lay.height; }
int emu_slow_nodes(emu_scene *s) { return (int)s->lay.n_slow; }
int emu_tight_children(emu_scene *s) {
    int n = 0;
    for (size_t b = 0; b < s->lay.n_branches; ++b)
        n += ((f2u(s->slabs[b].x) & 3u) != kSlabNone ? 1 : 0) + ((f2u(s->slabs[b].z) & 3u) != kSlabNone ? 1 : 0);
    return n;
}
void emu_set_leaf_cull(emu_scene *s, int on) { s->view.leaf_cull = on ? 1u : 0u; }
void emu_set_spheres(emu_scene *s, const sqt_sphere *sp, unsigned n) {
    s->spheres.assign((size_t)2 * (n ? n : 1), float4{0, 0, 0, 0});
    for (unsigned k = 0; k < n; ++k) {
        s->spheres[2 * k] = float4{sp[k].center[0], sp[k].center[1], sp[k].center[2], sp[k].radius};
        s->spheres[2 * k + 1] = float4{u2f(sp[k].material), 0, 0, 0};
    }
    s->view.spheres = s->spheres.data(); s->view.n_spheres = n;
    s->view.sph_nodes = nullptr; s->view.sph_order = nullptr;
    if (n && s->sphere_bvh) {
        build_sphere_bvh(sp, n, s->bvh);
        s->view.sph_nodes = s->bvh.nodes.data(); s->view.sph_order = s->bvh.order.data();
    }
}
void emu_set_sphere_bvh(emu_scene *s, int on) { s->sphere_bvh = on; if (!on) { s->view.sph_nodes = nullptr; s->view.sph_order = nullptr; } }
void emu_set_sbuf_budget(long long bytes) { sbuf_budget = bytes; }

void emu_intersect_batch(emu_scene *s, const float *org, const float *dir, long long n, int *tri_out, float *dist_out,
                         float *point_out, unsigned long long *counters4 /* 5 values */) {
    Counters cn = {};
    PathStats st = {0, 0, 0};
    const long long n_lanes = 5;
    for (long long l = 0; l < n_lanes; ++l) { BatchPolicy pol(org, dir, n, l, n_lanes, tri_out, dist_out, point_out, st); run_lane<true>(s->view, pol, &cn); }
    if (counters4) { counters4[0] = cn.branch_visits; counters4[1] = cn.child_box_tests; counters4[2] = cn.tri_tests; counters4[3] = cn.rays; counters4[4] = cn.leaves_culled; }
}

// The pool kernel's stack layout on the host: groups of 64 rays share ONE region in which their stacks are interleaved
// entry by entry (entry e of ray k at region[e * 64 + k], 16 bytes each; `height` entries per ray as launch_pool sizes it),
// and their rays live in a structure-of-arrays pool read through PoolRay, exactly as k_paths_pool does.  The 64 rays are
// stepped round-robin, one unit step each, so that an addressing error (overlap between slots, a region too small) would
// corrupt a neighbour's stack while it is live.  Canary entries behind the region catch overruns.  Returns 0, or 1 if a
// canary was overwritten.
int emu_intersect_batch_interleaved(emu_scene *s, const float *org, const float *dir, long long n, int *tri_out, float *dist_out) {
    constexpr int P = 64;
    const int depth = (int)std::max(1u, std::min<uint32_t>(s->lay.height, kStackEntries));
    const float4 canary = float4{-7.0f, -7.0f, -7.0f, -7.0f};
    std::vector<float4> region((size_t)P * depth + 64, canary);
    std::vector<uint32_t> pool((size_t)9 * P, 0u);
    Counters cn = {};
    int bad = 0;
    for (long long base = 0; base < n; base += P) {
        const int m = (int)std::min<long long>(P, n - base);
        TravLane L[P];
        for (int k = 0; k < m; ++k) {
            L[k].stack = region.data() + k;
            L[k].r = Ray{org[3 * (base + k)], org[3 * (base + k) + 1], org[3 * (base + k) + 2], dir[3 * (base + k)], dir[3 * (base + k) + 1], dir[3 * (base + k) + 2]};
            start_ray<false>(s->view, L[k], &cn);
            const float f[9] = {L[k].r.ox, L[k].r.oy, L[k].r.oz, L[k].r.dx, L[k].r.dy, L[k].r.dz, L[k].dfx, L[k].dfy, L[k].dfz};
            for (int w = 0; w < 9; ++w) pool[(size_t)w * P + k] = f2u(f[w]);
            // the steps below must read the ray from the pool only
            L[k].r = Ray{-1.0f, -1.0f, -1.0f, -1.0f, -1.0f, -1.0f}; L[k].dfx = L[k].dfy = L[k].dfz = -1.0f;
        }
        for (bool any = true; any;) {
            any = false;
            for (int k = 0; k < m; ++k) {
                TravLane &l = L[k];
                if (l.state == ST_DONE) continue;
                any = true;
                const PoolRay ra(pool.data() + k, P);
                if (l.state == ST_RET) ret_step<P>(s->view, l, ra);
                else if (l.state == ST_DESC) desc_step<false, P>(s->view, l, ra, &cn);
                else if (l.state == ST_ENTER) enter_step<false>(s->view, l, ra, &cn);
                else if (l.state == ST_LEAF) {
                    const TriData d = tri_load(s->view, l.child + (uint32_t)l.i);
                    tri_apply<false>(l, ra.ray(), d, &cn);
                }
                else if (l.state == ST_SPH) sphere_step(s->view, l, ra);
            }
        }
        for (int k = 0; k < m; ++k) {                        // report like BatchPolicy: position in the parsed list, -1 = Nothing
            const int t = L[k].cur.tri;
            tri_out[base + k] = t < 0 ? -1 : ((uint32_t)t >= s->view.n_tris ? t : (int)f2u(s->tris[3 * (size_t)t + 2].z));
            dist_out[base + k] = t < 0 ? 0.0f : L[k].cur.dist;
        }
        for (size_t c = (size_t)P * depth; c < region.size(); ++c) if (std::memcmp(&region[c], &canary, 16) != 0) bad = 1;
    }
    return bad;
}

// One render.  The device runs 32 lanes per warp in lock step; lanes are independent, so here `n_lanes` software
// lanes each run to completion, sharing one work queue like the device lanes share the atomic counter.
// stats5 = rays, samples, primary_reused, branch_visits, tri_tests
void emu_render(emu_scene *s, const sqt_camera *cam, const sqt_render_params *p, int rank, int world, float *accum,
                unsigned char *rgb8, unsigned long long *stats5) {
    RenderParams d = {};
    d.rows = p->rows; d.cols = p->cols; d.xdiv = p->xdiv; d.ydiv = p->ydiv; d.seed_stride = p->seed_stride;
    d.spp = p->spp; d.max_depth = p->max_depth; d.mode = p->mode; d.seed = p->seed;
    d.rank = rank; d.world = world; d.split_samples = (p->flags & SQT_F_SPLIT_SAMPLES) ? 1 : 0;
    d.primary_reuse = (p->flags & SQT_F_NO_PRIMARY_REUSE) ? 0 : 1;
    for (int k = 0; k < 3; ++k) d.cam_pos[k] = cam->position[k];
    for (int k = 0; k < 9; ++k) d.cam_rot[k] = cam->rotation[k];
    d.terminate_on_black = (s->lay.terminate_on_black_ok && !(p->flags & SQT_F_NO_EARLY_TERMINATION)) ? 1 : 0;
    const long long npix = (long long)d.rows * d.cols, nwork = work_items(d);
    std::memset(accum, 0, (size_t)npix * 12);
    Counters cn = {};
    PathStats st = {0, 0, 0};
    const long long n_lanes = 7;
    if (d.mode == 1) {
        for (long long l = 0; l < n_lanes; ++l) { CastPolicy pol(d, accum, nwork, l, n_lanes, st); run_lane<true>(s->view, pol, &cn); }
    } else {
        std::vector<int2> prim;
        std::vector<int> pixel_list;
        if (d.primary_reuse) {
            prim.resize((size_t)npix);
            VecAppend app{&pixel_list};
            // lanes interleave like the device's grid-stride lanes; the resulting list order is arbitrary on the device too
            for (long long l = n_lanes - 1; l >= 0; --l) { PrimaryPolicy<VecAppend> pol(d, prim.data(), app, nwork, l, n_lanes, st); run_lane<true>(s->view, pol, &cn); }
        }
        int k0, k1;
        sample_range(d, k0, k1);
        RoundInfo rd = {};
        rd.pixel_list = d.primary_reuse ? pixel_list.data() : nullptr;
        rd.prim = d.primary_reuse ? prim.data() : nullptr;
        rd.n_slots = d.primary_reuse ? (long long)pixel_list.size() : nwork;
        rd.slot_stride = nwork;
        rd.log2_s = round_log2_s(nwork, k1 - k0 > 0 ? k1 - k0 : 1, sbuf_budget);
        const int S = 1 << rd.log2_s;
        std::vector<float> sbuf((size_t)(nwork * S * 3 + 3));
        rd.sbuf = sbuf.data();
        uint16_t pm[SQT_MAX_DEPTH];
        for (int kb = k0; kb < k1; kb += S) {
            rd.k0 = kb; rd.k1 = kb + S < k1 ? kb + S : k1;
            std::fill(sbuf.begin(), sbuf.end(), 0.0f);
            const long long items = rd.n_slots << rd.log2_s;
            SeqFetch fetch{0, items};
            for (long long l = 0; l < n_lanes; ++l) {
                // every lane takes a contiguous share of the queue, like lanes racing on the device's counter
                SeqFetch part{fetch.next, l + 1 == n_lanes ? items : std::min(items, fetch.next + (items + n_lanes - 1) / n_lanes)};
                PathPolicy<SeqFetch> pol(d, rd, part, st, pm);
                run_lane<true>(s->view, pol, &cn);
                fetch.next = part.n;
            }
            for (long long slot = 0; slot < rd.n_slots; ++slot) accumulate_slot(d, rd, slot, accum);
        }
    }
    if (rgb8) {
        const float inv = 1.0f / (float)d.spp;
        for (long long i = 0; i < npix; ++i)
            tone_map(XMUL(inv, accum[3 * i]), XMUL(inv, accum[3 * i + 1]), XMUL(inv, accum[3 * i + 2]), rgb8 + 3 * i);
    }
    if (stats5) { stats5[0] = st.rays; stats5[1] = st.samples; stats5[2] = st.primary_reused; stats5[3] = cn.branch_visits; stats5[4] = cn.tri_tests; }
}

// Step trace of one path lane that renders the pixels [w0, w1) in order: one byte per unit step
// ('R' regeneration, 'T' traversal step, 'L' triangle step).  Used by tools/sched_sim.py to study warp
// schedulers offline.  Returns the number of bytes written (truncated at cap).
long long emu_trace_lane(emu_scene *s, const sqt_camera *cam, const sqt_render_params *p, long long w0, long long w1,
                         unsigned char *out, long long cap) {
    RenderParams d = {};
    d.rows = p->rows; d.cols = p->cols; d.xdiv = p->xdiv; d.ydiv = p->ydiv; d.seed_stride = p->seed_stride;
    d.spp = p->spp; d.max_depth = p->max_depth; d.mode = 0; d.seed = p->seed; d.rank = 0; d.world = 1;
    d.primary_reuse = 1;
    for (int k = 0; k < 3; ++k) d.cam_pos[k] = cam->position[k];
    for (int k = 0; k < 9; ++k) d.cam_rot[k] = cam->rotation[k];
    d.terminate_on_black = s->lay.terminate_on_black_ok;
    const long long npix = (long long)d.rows * d.cols;
    std::vector<int2> prim((size_t)npix);
    Counters cn = {};
    PathStats st = {0, 0, 0};
    for (long long w = w0; w < w1; ++w) {
        const Ray r = make_ray(d, (int)(w / d.cols), (int)(w % d.cols));
        const Hit h = traverse<false>(s->view, r, &cn);
        prim[(size_t)w].x = h.tri; prim[(size_t)w].y = (int)f2u(h.t);
    }
    // the lane traces all samples of the pixels [w0, w1) whose primary ray hits, one after the other
    std::vector<int> pixel_list;
    for (long long w = w0; w < w1; ++w) if (prim[(size_t)w].x >= 0) pixel_list.push_back((int)w);
    RoundInfo rd = {};
    rd.pixel_list = pixel_list.data(); rd.prim = prim.data(); rd.n_slots = (long long)pixel_list.size(); rd.slot_stride = rd.n_slots;
    rd.log2_s = 0; while ((1 << rd.log2_s) < d.spp) ++rd.log2_s;
    rd.k0 = 0; rd.k1 = d.spp;
    std::vector<float> sbuf((size_t)(rd.n_slots << rd.log2_s) * 3 + 3);
    rd.sbuf = sbuf.data();
    SeqFetch fetch{0, rd.n_slots << rd.log2_s};
    uint16_t pm[SQT_MAX_DEPTH];
    PathPolicy<SeqFetch> pol(d, rd, fetch, st, pm);
    float4 stack[kStackEntries];
    TravLane L;
    L.stack = stack; L.state = ST_DONE; L.sp = 0; L.cur.tri = -1;
    const LaneRay ra(L);
    long long n = 0;
    for (;;) {
        if (L.state == ST_DONE) { pol.regen<false>(s->view, L, &cn); if (n < cap) out[n] = 'R';
            n++; }
        if (L.state == ST_EXIT) break;
        if (L.state == ST_RET || L.state == ST_DESC) {
            if (L.state == ST_RET) ret_step(s->view, L, ra);
            if (L.state == ST_DESC) desc_step<false>(s->view, L, ra, &cn);
            if (n < cap) out[n] = 'T';
            n++;
        } else if (L.state == ST_ENTER) { enter_step<false>(s->view, L, ra, &cn); if (n < cap) out[n] = 'E'; n++; }
        else if (L.state == ST_LEAF) { tri_step<false>(s->view, L, &cn); if (n < cap) out[n] = 'L'; n++; }
        else if (L.state == ST_SPH) sphere_step(s->view, L, ra);
    }
    return n < cap ? n : cap;
}

// pack_slab_lo for n values and codes (tests: the packed bound is never above the input and carries the code)
void emu_pack_slab_lo(const float *lo, const unsigned *code, long long n, float *out) {
    for (long long i = 0; i < n; ++i) out[i] = pack_slab_lo(lo[i], code[i]);
}

// The pool kernel's division-free filter (moller_trumbore_au) against the full test, pair i = (triangle record i, ray i).
// tri = 12 floats per record (the device layout: v0.xyz e1.x | e1.yz e2.xy | e2.z - - -).
// out = { pairs, full test got past `u` (stage >= 2), filter passed, VIOLATIONS: full test got past `u` but the filter said no,
//         `a` guard disagreements }
void emu_filter_stats(const float *tri, const float *org, const float *dir, long long n, unsigned long long *out) {
    for (int k = 0; k < 5; ++k) out[k] = 0;
    for (long long i = 0; i < n; ++i) {
        const float *q = tri + 12 * i;
        const float4 a0 = mk4(q[0], q[1], q[2], q[3]), a1 = mk4(q[4], q[5], q[6], q[7]), a2 = mk4(q[8], q[9], q[10], q[11]);
        const Ray r{org[3 * i], org[3 * i + 1], org[3 * i + 2], dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
        float t, dist; int stage; bool pass_a;
        moller_trumbore(a0, a1, a2, r, t, dist, stage);
        const bool pass = moller_trumbore_au(a0, a1, a2, r, pass_a);
        out[0] += 1; out[1] += stage >= 2; out[2] += pass;
        if (stage >= 2 && !pass) out[3] += 1;
        if ((stage >= 1) != pass_a) out[4] += 1;
    }
}

// Diagnostic for DESIGN.md: how many leaf visits would a (conservatively enlarged) tight leaf bounding box reject?
// out = { leaf visits, triangle tests, visits whose ray misses the enlarged tight box, triangle tests in those,
//         visits with at least one accepted triangle, of those rejected by the box (must be 0) }
void emu_leaf_cull_stats(emu_scene *s, const float *org, const float *dir, long long n, float margin, unsigned long long *out) {
    for (int k = 0; k < 6; ++k) out[k] = 0;
    const uint32_t saved_cull = s->view.leaf_cull;
    s->view.leaf_cull = 0;
    Counters cn = {};
    float4 stack[kStackEntries];
    for (long long i = 0; i < n; ++i) {
        TravLane L; L.stack = stack;
        const LaneRay ra(L);
        L.r = Ray{org[3 * i], org[3 * i + 1], org[3 * i + 2], dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
        start_ray<false>(s->view, L, &cn);
        while (L.state != ST_DONE) {
            if (L.state == ST_RET) ret_step(s->view, L, ra);
            if (L.state == ST_DESC) desc_step<false>(s->view, L, ra, &cn);
            if (L.state == ST_SPH) sphere_step(s->view, L, ra);
            if (L.state == ST_ENTER) enter_step<false>(s->view, L, ra, &cn);
            if (L.state == ST_LEAF) {
                const uint32_t first = L.child, count = (uint32_t)L.i + 1;
                float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
                for (uint32_t t = 0; t < count; ++t) {
                    const float4 a0 = s->tris[3 * (size_t)(first + t)], a1 = s->tris[3 * (size_t)(first + t) + 1], a2 = s->tris[3 * (size_t)(first + t) + 2];
                    const float v[3][3] = {{a0.x, a0.y, a0.z}, {a0.x + a0.w, a0.y + a1.x, a0.z + a1.y}, {a0.x + a1.z, a0.y + a1.w, a0.z + a2.x}};
                    for (int q = 0; q < 3; ++q) for (int c = 0; c < 3; ++c) { lo[c] = std::min(lo[c], v[q][c]); hi[c] = std::max(hi[c], v[q][c]); }
                }
                float E2 = 0;
                for (uint32_t t = 0; t < count; ++t) {
                    const float4 a0 = s->tris[3 * (size_t)(first + t)], a1 = s->tris[3 * (size_t)(first + t) + 1], a2 = s->tris[3 * (size_t)(first + t) + 2];
                    const float e1[3] = {a0.w, a1.x, a1.y}, e2[3] = {a1.z, a1.w, a2.x};
                    float l1 = 0, l2 = 0, l3 = 0;
                    for (int c = 0; c < 3; ++c) { l1 += e1[c] * e1[c]; l2 += e2[c] * e2[c]; l3 += (e2[c] - e1[c]) * (e2[c] - e1[c]); }
                    E2 = std::max(E2, std::max(l1, std::max(l2, l3)));
                }
                const float E = std::sqrt(E2);
                const float s1 = std::fabs(L.r.ox - 0.5f * (lo[0] + hi[0])) + std::fabs(L.r.oy - 0.5f * (lo[1] + hi[1])) + std::fabs(L.r.oz - 0.5f * (lo[2] + hi[2])) + (hi[0] - lo[0]) + (hi[1] - lo[1]) + (hi[2] - lo[2]);
                const float d1 = std::fabs(L.r.dx) + std::fabs(L.r.dy) + std::fabs(L.r.dz);
                const float m = margin * (s1 + E) * d1 * (1.0f + d1) * E2 + 1e-4f;
                const bool box_hit = slab_exact(lo[0] - m, lo[1] - m, lo[2] - m, hi[0] + m, hi[1] + m, hi[2] + m, L.r, L.dfx, L.dfy, L.dfz);
                while (L.state == ST_LEAF) tri_step<false>(s->view, L, &cn);
                out[0] += 1; out[1] += count;
                if (!box_hit) { out[2] += 1; out[3] += count; }
                if (L.cur.tri >= 0) { out[4] += 1; if (!box_hit) out[5] += 1; }
            }
        }
    }
    s->view.leaf_cull = saved_cull;
}

// Diagnostic for DESIGN.md: with the production leaf culling ON, how many of the remaining triangle tests would a second,
// finer culling level remove -- the leaf's triangles cut into consecutive groups of `group`, each with its own tight box
// and the same conservative margin?  out = { leaf visits entered, triangle tests in them, tests left after group culling,
// group box tests, accepted hits lost (must be 0) }
void emu_group_cull_stats(emu_scene *s, const float *org, const float *dir, long long n, int group, unsigned long long *out) {
    for (int k = 0; k < 5; ++k) out[k] = 0;
    Counters cn = {};
    float4 stack[kStackEntries];
    for (long long i = 0; i < n; ++i) {
        TravLane L; L.stack = stack;
        const LaneRay ra(L);
        L.r = Ray{org[3 * i], org[3 * i + 1], org[3 * i + 2], dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
        start_ray<false>(s->view, L, &cn);
        while (L.state != ST_DONE) {
            if (L.state == ST_RET) ret_step(s->view, L, ra);
            if (L.state == ST_DESC) desc_step<false>(s->view, L, ra, &cn);
            if (L.state == ST_SPH) sphere_step(s->view, L, ra);
            if (L.state == ST_ENTER) enter_step<false>(s->view, L, ra, &cn);
            if (L.state == ST_LEAF) {
                const uint32_t first = L.child, count = (uint32_t)L.i + 1;
                out[0] += 1; out[1] += count;
                for (uint32_t g0 = 0; g0 < count; g0 += (uint32_t)group) {
                    const uint32_t gc = std::min<uint32_t>((uint32_t)group, count - g0);
                    float4 b0, b1;
                    make_leaf_record(s->tris.data(), first + g0, gc, b0, b1);
                    out[3] += 1;
                    bool keep = true;
                    if (!(L.rf & kRfUnsafe)) {
                        const Ray &r = L.r;
                        const float d1 = fabsf(r.dx) + fabsf(r.dy) + fabsf(r.dz), dfac = d1 * (1.0f + d1);
                        const float E = b1.z;
                        const float s1 = fabsf(r.ox - 0.5f * (b0.x + b0.w)) + fabsf(r.oy - 0.5f * (b0.y + b1.x)) + fabsf(r.oz - 0.5f * (b0.z + b1.y)) + ((b0.w - b0.x) + (b1.x - b0.y) + (b1.y - b0.z));
                        const float cmax = fmaxf(fmaxf(fmaxf(fabsf(b0.x), fabsf(b0.w)), fmaxf(fabsf(b0.y), fabsf(b1.x))), fmaxf(fabsf(b0.z), fabsf(b1.y)));
                        const float m = 0.03f * (s1 + E) * dfac * (E * E) + (1.0e-4f + 9.5367431640625e-7f * (cmax + s1));
                        const float lx = (b0.x - m - r.ox) * L.dfx, hx = (b0.w + m - r.ox) * L.dfx;
                        const float ly = (b0.y - m - r.oy) * L.dfy, hy = (b1.x + m - r.oy) * L.dfy;
                        const float lz = (b0.z - m - r.oz) * L.dfz, hz = (b1.y + m - r.oz) * L.dfz;
                        const float tmin = fmaxf(fmaxf(fminf(lx, hx), fminf(ly, hy)), fminf(lz, hz));
                        const float tmax = fminf(fminf(fmaxf(lx, hx), fmaxf(ly, hy)), fmaxf(lz, hz));
                        if (tmax < 0.0f || tmin > tmax) keep = false;
                    }
                    if (keep) out[2] += gc;
                    else {
                        for (uint32_t t = 0; t < gc; ++t) {
                            const TriData d = tri_load(s->view, first + g0 + t);
                            float tt, dd; int st;
                            if (moller_trumbore(d.a0, d.a1, d.a2, L.r, tt, dd, st)) out[4] += 1;
                        }
                    }
                }
                while (L.state == ST_LEAF) tri_step<false>(s->view, L, &cn);
            }
        }
    }
}

}  // extern "C"
