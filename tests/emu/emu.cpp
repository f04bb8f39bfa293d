// emu.cpp -- TEST HELPER: host (g++) build of the kernel logic in csrc/sqt_core.cuh + sqt_paths.cuh.
//
// The GPU box is scarce, so the per-ray / per-path state machines are written __host__ __device__ and
// this file runs the very same code on the CPU, lane by lane, so that `-m "not gpu"` tests can compare it
// with the independent oracle (oracle/oracle.c) before any GPU time is spent.  It is NOT a fallback: the
// product library (libsqt_b200.so) neither links nor loads it, and it lives under tests/.
#include <cstring>
#include <string>
#include <vector>

#include "../../squigly-trace_b200/csrc/sqt_layout.hpp"

using namespace sqt;

struct emu_scene {
    DeviceLayout lay;
    std::vector<float4> tris;
    SceneView view;
    std::string err;
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
    v.nodes = s->lay.nodes.data(); v.tris = s->tris.data(); v.mats = s->lay.mats.data();
    for (int k = 0; k < 3; ++k) { v.root_lo[k] = d->root_bounds[k]; v.root_hi[k] = d->root_bounds[3 + k]; }
    v.n_branches = s->lay.n_branches; v.n_tris = d->n_tris; v.n_mats = d->n_mats;
    v.root_is_leaf = (d->nodes[0].b & SQT_NODE_LEAF) ? 1u : 0u;
    s->view = v;
    return s;
}
const char *emu_error(emu_scene *s) { return s->err.c_str(); }
void emu_free(emu_scene *s) { delete s; }
int emu_height(emu_scene *s) { return (int)s->lay.height; }

void emu_intersect_batch(emu_scene *s, const float *org, const float *dir, long long n, int *tri_out, float *dist_out,
                         float *point_out, unsigned long long *counters4) {
    Counters cn = {0, 0, 0, 0};
    for (long long i = 0; i < n; ++i) {
        Ray r{org[3 * i], org[3 * i + 1], org[3 * i + 2], dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]};
        const Hit h = traverse<true>(s->view, r, &cn);
        const bool hit = h.tri >= 0;
        tri_out[i] = hit ? (int)f2u(s->tris[3 * (size_t)h.tri + 2].z) : -1;
        dist_out[i] = hit ? h.dist : 0.0f;
        point_out[3 * i] = hit ? XADD(r.ox, XMUL(h.t, r.dx)) : 0.0f;
        point_out[3 * i + 1] = hit ? XADD(r.oy, XMUL(h.t, r.dy)) : 0.0f;
        point_out[3 * i + 2] = hit ? XADD(r.oz, XMUL(h.t, r.dz)) : 0.0f;
    }
    if (counters4) { counters4[0] = cn.branch_visits; counters4[1] = cn.child_box_tests; counters4[2] = cn.tri_tests; counters4[3] = cn.rays; }
}

// One render, all lanes run one after the other.  stats5 = rays, samples, primary_reused, branch_visits, tri_tests
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
    Counters cn = {0, 0, 0, 0};
    PathStats st = {0, 0, 0};
    if (d.mode == 1) {
        for (long long w = 0; w < nwork; ++w) { const long long pix = work_to_pixel(d, w); if (pix >= 0) raycast_pixel<true>(s->view, d, pix, accum, &cn, st); }
    } else {
        std::vector<int2> prim;
        if (d.primary_reuse) {
            prim.resize((size_t)npix);
            for (long long w = 0; w < nwork; ++w) {
                const long long pix = work_to_pixel(d, w);
                if (pix < 0) continue;
                const Ray r = make_ray(d, (int)(pix / d.cols), (int)(pix % d.cols));
                const Hit h = traverse<true>(s->view, r, &cn);
                st.rays += 1;
                prim[(size_t)pix].x = h.tri; prim[(size_t)pix].y = (int)f2u(h.t);
            }
        }
        // emulate 7 interleaved "lanes" pulling from one queue, to exercise the dynamic fetch order independence
        SeqFetch fetch{0, nwork};
        render_lane<true>(s->view, d, d.primary_reuse ? prim.data() : nullptr, accum, fetch, &cn, st);
    }
    if (rgb8) {
        const float inv = 1.0f / (float)d.spp;
        for (long long i = 0; i < npix; ++i)
            tone_map(XMUL(inv, accum[3 * i]), XMUL(inv, accum[3 * i + 1]), XMUL(inv, accum[3 * i + 2]), rgb8 + 3 * i);
    }
    if (stats5) { stats5[0] = st.rays; stats5[1] = st.samples; stats5[2] = st.primary_reused; stats5[3] = cn.branch_visits; stats5[4] = cn.tri_tests; }
}

}  // extern "C"
