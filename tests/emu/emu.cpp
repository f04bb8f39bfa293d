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
    PathStats st = {0, 0, 0};
    const long long n_lanes = 5;
    for (long long l = 0; l < n_lanes; ++l) { BatchPolicy pol(org, dir, n, l, n_lanes, tri_out, dist_out, point_out, st); run_lane<true>(s->view, pol, &cn); }
    if (counters4) { counters4[0] = cn.branch_visits; counters4[1] = cn.child_box_tests; counters4[2] = cn.tri_tests; counters4[3] = cn.rays; }
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
    Counters cn = {0, 0, 0, 0};
    PathStats st = {0, 0, 0};
    const long long n_lanes = 7;
    if (d.mode == 1) {
        for (long long l = 0; l < n_lanes; ++l) { CastPolicy pol(d, accum, nwork, l, n_lanes, st); run_lane<true>(s->view, pol, &cn); }
    } else {
        std::vector<int2> prim;
        if (d.primary_reuse) {
            prim.resize((size_t)npix);
            for (long long l = 0; l < n_lanes; ++l) { PrimaryPolicy pol(d, prim.data(), nwork, l, n_lanes, st); run_lane<true>(s->view, pol, &cn); }
        }
        SeqFetch fetch{0, nwork};
        uint16_t pm[SQT_MAX_DEPTH];
        for (long long l = 0; l < n_lanes; ++l) {
            // each lane drains what is left of the queue after taking a few items, like lanes racing on the counter
            SeqFetch part{fetch.next, l + 1 == n_lanes ? nwork : std::min(nwork, fetch.next + (nwork + n_lanes - 1) / n_lanes)};
            PathPolicy<SeqFetch> pol(d, d.primary_reuse ? prim.data() : nullptr, accum, part, st, pm);
            run_lane<true>(s->view, pol, &cn);
            fetch.next = part.n;
        }
    }
    if (rgb8) {
        const float inv = 1.0f / (float)d.spp;
        for (long long i = 0; i < npix; ++i)
            tone_map(XMUL(inv, accum[3 * i]), XMUL(inv, accum[3 * i + 1]), XMUL(inv, accum[3 * i + 2]), rgb8 + 3 * i);
    }
    if (stats5) { stats5[0] = st.rays; stats5[1] = st.samples; stats5[2] = st.primary_reused; stats5[3] = cn.branch_visits; stats5[4] = cn.tri_tests; }
}

}  // extern "C"
