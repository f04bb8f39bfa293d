// sqt_layout.hpp -- host-side derivation of the device records from the C-ABI scene description.
//
// Walks the boundary tree once (iteratively), validating it and deriving for every Branch its 16-byte device node and
// the box intersectBIH' would receive for it: the root gets `bounds`, a left child gets its parent's box with
// hi[axis] := lmax, a right child the parent's box with lo[axis] := rmin (BIH.hs:130-141).  Plane values
// are copied, never computed.  The leaf records and the material-index check are per-triangle work and are left to the
// caller (device kernels in the product, make_leaf_record on the host in tests/emu).  Shared by sqt_backend.cu (the
// product) and tests/emu (host build of the kernel logic).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sqt.h"
#include "sqt_paths.cuh"

namespace sqt {

struct DeviceLayout {
    std::vector<float4> nodes;      // 1 per branch: (lmax, rmin, L, R)
    std::vector<float4> boxes;      // 2 per branch: the clipped box intersectBIH' receives for it (slow path only)
    std::vector<float4> mats;       // 3 per material
    std::vector<uint32_t> leaf_first, leaf_count;   // triangle range per leaf (leaf order = order of first appearance in the node
                                                    // array); the leaf RECORDS are derived from the triangles by make_leaf_record
    std::vector<uint32_t> branch_order;             // branches in the order of the walk (parents before children)
    std::vector<uint32_t> branch_depth;             // depth of every branch (root = 1)
    float s_max = 0.0f, c_max = 0.0f;               // constants of the subtree-slab margin (make_slab), tame_c / tame_r of SceneView
    float tame_c[3] = {0, 0, 0}, tame_r = 0.0f;
    uint32_t n_branches = 0, n_slow = 0, height = 0;
    int planes_finite = 1;
    int terminate_on_black_ok = 0;
};

inline int layout_fail(std::string &err, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    err = buf;
    return code;
}
inline float4 mk4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

inline int build_device_layout(const sqt_scene_desc &desc, DeviceLayout &out, std::string &err, bool check_materials = true) {
    const sqt_scene_desc *s = &desc;
    if (!s->nodes || s->n_nodes == 0) return layout_fail(err, SQT_E_INVALID, "scene has no BIH nodes");
    if (s->n_tris && !s->tris) return layout_fail(err, SQT_E_INVALID, "tris is NULL");
    if (!s->mats || s->n_mats == 0) return layout_fail(err, SQT_E_INVALID, "scene has no materials");
    if (s->n_mats > 65535) return layout_fail(err, SQT_E_UNSUPPORTED, "more than 65535 materials");
    if (s->n_tris >= (1u << 27) || s->n_nodes >= (1u << 27)) return layout_fail(err, SQT_E_UNSUPPORTED, "scene too large for the 27-bit indices");
    const uint32_t N = s->n_nodes;
    std::vector<int32_t> branch_id(N, -1), leaf_id(N, -1);
    uint32_t n_br = 0, n_lf = 0;
    for (uint32_t i = 0; i < N; ++i) {
        if (!(s->nodes[i].b & SQT_NODE_LEAF)) branch_id[i] = (int32_t)n_br++;
        else leaf_id[i] = (int32_t)n_lf++;
    }
    out.leaf_first.assign(n_lf ? n_lf : 1, 0u); out.leaf_count.assign(n_lf ? n_lf : 1, 0u);
    struct Box { float lo[3], hi[3]; };
    std::vector<float4> dn((size_t)(n_br ? n_br : 1)), db((size_t)2 * (n_br ? n_br : 1));
    out.branch_order.clear(); out.branch_order.reserve(n_br);
    out.branch_depth.assign(n_br ? n_br : 1, 0u);
    std::vector<uint8_t> seen(N, 0);
    struct Item { uint32_t node; Box box; uint32_t depth; };
    std::vector<Item> todo;
    Box root;
    int planes_finite = 1;
    for (int k = 0; k < 3; ++k) {
        root.lo[k] = s->root_bounds[k]; root.hi[k] = s->root_bounds[3 + k];
        if (!std::isfinite(root.lo[k]) || !std::isfinite(root.hi[k])) planes_finite = 0;
    }
    todo.push_back({0u, root, 1u});
    uint32_t height = 0, n_slow = 0;
    auto child_ref = [&](uint32_t child, uint32_t &ref) -> const char * {
        const sqt_node &c = s->nodes[child];
        if (c.b & SQT_NODE_LEAF) {
            const uint32_t cnt = c.b & ~SQT_NODE_LEAF, first = c.a;
            if ((uint64_t)first + cnt > s->n_tris) return "leaf triangle range out of bounds";
            ref = (uint32_t)leaf_id[child] | kLeaf;
        } else ref = (uint32_t)branch_id[child];
        return nullptr;
    };
    while (!todo.empty()) {
        Item it = todo.back(); todo.pop_back();
        if (it.node >= N) return layout_fail(err, SQT_E_INVALID, "node index %u out of range", it.node);
        if (seen[it.node]) return layout_fail(err, SQT_E_INVALID, "node %u reachable twice (not a tree)", it.node);
        seen[it.node] = 1;
        if (it.depth > height) height = it.depth;
        const sqt_node &nd = s->nodes[it.node];
        if (nd.b & SQT_NODE_LEAF) {
            const uint32_t cnt = nd.b & ~SQT_NODE_LEAF;
            if ((uint64_t)nd.a + cnt > s->n_tris) return layout_fail(err, SQT_E_INVALID, "leaf %u: triangle range out of bounds", it.node);
            out.leaf_first[(size_t)leaf_id[it.node]] = cnt ? nd.a : 0u; out.leaf_count[(size_t)leaf_id[it.node]] = cnt;
            continue;
        }
        const uint32_t ax = nd.a >> 30, l = nd.a & 0x3fffffffu, r = nd.b;
        if (ax > 2) return layout_fail(err, SQT_E_INVALID, "node %u: axis %u", it.node, ax);
        if (l >= N || r >= N) return layout_fail(err, SQT_E_INVALID, "node %u: child out of range", it.node);
        uint32_t lref = 0, rref = 0;
        const char *e1 = child_ref(l, lref), *e2 = child_ref(r, rref);
        if (e1 || e2) return layout_fail(err, SQT_E_INVALID, "node %u: %s", it.node, e1 ? e1 : e2);
        Box lb = it.box, rb = it.box;
        lb.hi[ax] = nd.lmax; rb.lo[ax] = nd.rmin;                           // BIH.hs:130-141
        if (!std::isfinite(nd.lmax) || !std::isfinite(nd.rmin)) planes_finite = 0;
        // interval stepping (desc_step) is exact iff both planes lie inside the node's own extent on the split axis
        const bool nested = it.box.lo[ax] <= nd.lmax && nd.lmax <= it.box.hi[ax] && it.box.lo[ax] <= nd.rmin && nd.rmin <= it.box.hi[ax];
        if (!nested) { lref |= kSlow; ++n_slow; }
        lref |= ax << kAxisShift;
        const size_t b = (size_t)branch_id[it.node];
        out.branch_order.push_back((uint32_t)b); out.branch_depth[b] = it.depth;
        dn[b] = mk4(nd.lmax, nd.rmin, u2f(lref), u2f(rref));
        db[2 * b] = mk4(it.box.lo[0], it.box.lo[1], it.box.lo[2], it.box.hi[0]);
        db[2 * b + 1] = mk4(it.box.hi[1], it.box.hi[2], 0.0f, 0.0f);
        todo.push_back({r, rb, it.depth + 1});
        todo.push_back({l, lb, it.depth + 1});
    }
    if (height > SQT_MAX_HEIGHT) return layout_fail(err, SQT_E_UNSUPPORTED, "BIH height %u exceeds SQT_MAX_HEIGHT=%d", height, SQT_MAX_HEIGHT);
    if (check_materials)
        for (uint32_t t = 0; t < s->n_tris; ++t)
            if (s->tris[t].material >= s->n_mats) return layout_fail(err, SQT_E_INVALID, "triangle %u: material %u out of range", t, s->tris[t].material);

    // materials: (reflective, surf) (emissive, emit) (emissive *^ emit, flags)
    std::vector<float4> dm((size_t)3 * s->n_mats);
    // terminate_on_black is exact iff no radiance can overflow: bound sum_j E_max * C_max^j over SQT_MAX_DEPTH levels
    double cmax = 0, emax = 0;
    for (uint32_t i = 0; i < s->n_mats; ++i) {
        const sqt_material &m = s->mats[i];
        const float ex = m.emissive * m.emit_color[0], ey = m.emissive * m.emit_color[1], ez = m.emissive * m.emit_color[2];
        uint32_t flags = 0;
        if (!(ex == 0.0f && ey == 0.0f && ez == 0.0f)) flags |= kMatEmits;
        if (m.surf_color[0] == 0.0f && m.surf_color[1] == 0.0f && m.surf_color[2] == 0.0f) flags |= kMatBlack;
        dm[3 * i] = mk4(m.reflective, m.surf_color[0], m.surf_color[1], m.surf_color[2]);
        dm[3 * i + 1] = mk4(m.emissive, m.emit_color[0], m.emit_color[1], m.emit_color[2]);
        dm[3 * i + 2] = mk4(ex, ey, ez, u2f(flags));
        for (int k = 0; k < 3; ++k) {
            const double c = fabs((double)m.surf_color[k]), ev = fabs((double)(k == 0 ? ex : (k == 1 ? ey : ez)));
            if (!(c <= cmax)) cmax = c;          // NaN-propagating on purpose
            if (!(ev <= emax)) emax = ev;
        }
    }
    double bound = 0, pw = 1;
    for (int j = 0; j < SQT_MAX_DEPTH; ++j) { bound += emax * pw; pw *= (cmax > 1 ? cmax : 1); }
    const int tob = (bound == bound) && bound < 1e30;

    if (s->nodes[0].b & SQT_NODE_LEAF) {
        // tree = Leaf geom: the leaf range must be the whole (leaf-ordered) triangle array
        if (s->nodes[0].a != 0 || (s->nodes[0].b & ~SQT_NODE_LEAF) != s->n_tris)
            return layout_fail(err, SQT_E_INVALID, "root leaf must cover tris[0..n_tris)");
    }
    {   // subtree-slab constants: tame rays start within 2 x the root's 1-norm half extent of its centre
        float h1 = 0.0f, cm = 0.0f;
        for (int k = 0; k < 3; ++k) {
            out.tame_c[k] = 0.5f * (root.lo[k] + root.hi[k]);
            h1 += 0.5f * (root.hi[k] - root.lo[k]);
            cm = std::fmax(cm, std::fmax(std::fabs(root.lo[k]), std::fabs(root.hi[k])));
        }
        const bool ok = planes_finite && h1 >= 0.0f && h1 < 1.0e15f;
        out.tame_r = ok ? 2.0f * h1 : -1.0f;                 // -1: no ray is tame, the slabs are never used
        out.s_max = 3.0f * h1 * 1.0001f; out.c_max = cm;
    }
    out.nodes.swap(dn); out.boxes.swap(db); out.mats.swap(dm);
    out.n_branches = n_br; out.n_slow = n_slow; out.height = height; out.terminate_on_black_ok = tob; out.planes_finite = planes_finite;
    return SQT_OK;
}

constexpr float kSlabRatioMax = 0.7f, kSlabRatioMaxLeaf = 0.7f;     // a subtree's / leaf's slab is tested when it is at most this fraction of the clipped box on its axis
// Host version of what k_branch_tight + k_child_slabs do on the device (tests/emu): tight records bottom-up, then every
// Branch decides for each child (Branch or Leaf) whether its slab is worth a test (code in the slab record), and references to Branches with such a slab are flagged kTight.
inline void compute_slabs_host(DeviceLayout &lay, const std::vector<float4> &leaves, std::vector<float4> &slabs, float ratio_max = kSlabRatioMax, float ratio_max_leaf = kSlabRatioMaxLeaf) {
    const size_t nb = lay.n_branches;
    std::vector<TightRec> tight(nb ? nb : 1);
    slabs.assign(nb ? nb : 1, mk4(u2f(kSlabNone), 0, u2f(kSlabNone), 0));
    auto child = [&](uint32_t ref) { return (ref & kLeaf) ? tight_of_leaf(leaves.data(), ref & kIdxMask) : tight[ref & kIdxMask]; };
    for (size_t i = lay.branch_order.size(); i-- > 0;) {
        const uint32_t b = lay.branch_order[i];
        tight[b] = tight_union(child(f2u(lay.nodes[b].z)), child(f2u(lay.nodes[b].w)));
    }
    for (size_t b = 0; b < nb; ++b) {
        const float4 q = lay.nodes[b];
        const uint32_t w[2] = {f2u(q.z), f2u(q.w)};
        const int ax = (int)((w[0] >> kAxisShift) & 3u);
        float4 sl[2];
        for (int c = 0; c < 2; ++c) {
            const uint32_t k = w[c] & kIdxMask;
            bool use;
            if (w[c] & kLeaf) {
                float4 c0, c1;
                clip_child_box(lay.boxes[2 * b], lay.boxes[2 * b + 1], ax, q.x, q.y, c, c0, c1);
                use = make_slab(tight_of_leaf(leaves.data(), k), c0, c1, lay.s_max, lay.c_max, ratio_max_leaf, sl[c]);
            } else {
                use = make_slab(tight[k], lay.boxes[2 * (size_t)k], lay.boxes[2 * (size_t)k + 1], lay.s_max, lay.c_max, ratio_max, sl[c]);
            }
            sl[c].x = pack_slab_lo(sl[c].x, use ? (f2u(sl[c].w) & 3u) : kSlabNone);
        }
        slabs[b] = mk4(sl[0].x, sl[0].y, sl[1].x, sl[1].y);
    }
    for (size_t b = 0; b < nb; ++b) {
        uint32_t w[2] = {f2u(lay.nodes[b].z), f2u(lay.nodes[b].w)};
        for (int c = 0; c < 2; ++c) {
            if (w[c] & kLeaf) continue;
            const float4 cs = slabs[w[c] & kIdxMask];
            if ((f2u(cs.x) & 3u) != kSlabNone || (f2u(cs.z) & 3u) != kSlabNone) w[c] |= kTight;
        }
        lay.nodes[b].z = u2f(w[0]); lay.nodes[b].w = u2f(w[1]);
    }
}

// ---- extension: bounding-volume hierarchy over the analytic spheres (sphere_step in sqt_core.cuh) -----------------
// Median split of the centres along the longest axis of their bounds, at most 4 spheres per leaf, pre-order numbering
// (root = 0): depth <= ceil(log2(n / 2)) + 1 <= 27 for the 2^27 surfaces the indices allow.  Node = 2 x float4:
// (lo.xyz, hi.x) (hi.yz, a, b); interior a = low child | split axis << 30, b = high child; leaf a = first entry of `order`,
// b = count | kLeaf.  Boxes bound centre +- radius * (1 + 2^-19), rounded outwards (the radius term of the culling bound).
struct SphereBvh { std::vector<float4> nodes; std::vector<uint32_t> order; };
inline void build_sphere_bvh(const sqt_sphere *sp, uint32_t n, SphereBvh &out) {
    out.nodes.clear(); out.order.resize(n);
    for (uint32_t i = 0; i < n; ++i) out.order[i] = i;
    if (n == 0) return;
    struct Job { uint32_t lo, hi, node; };
    std::vector<Job> todo;
    out.nodes.resize(2);
    todo.push_back({0u, n, 0u});
    while (!todo.empty()) {
        const Job j = todo.back(); todo.pop_back();
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
        for (uint32_t i = j.lo; i < j.hi; ++i) {
            const sqt_sphere &s = sp[out.order[i]];
            const double rr = (double)s.radius * (1.0 + 1.0 / 524288.0);
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::fmin(lo[k], (double)s.center[k] - rr); hi[k] = std::fmax(hi[k], (double)s.center[k] + rr);
                clo[k] = std::fmin(clo[k], (double)s.center[k]); chi[k] = std::fmax(chi[k], (double)s.center[k]);
            }
        }
        auto dn_ = [](double x) { return std::nextafterf((float)x, -INFINITY); };
        auto up_ = [](double x) { return std::nextafterf((float)x, INFINITY); };
        uint32_t a, b;
        if (j.hi - j.lo <= 4u) { a = j.lo; b = (j.hi - j.lo) | kLeaf; }
        else {
            int ax = 0;
            if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
            if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
            const uint32_t mid = j.lo + (j.hi - j.lo) / 2;
            std::nth_element(out.order.begin() + j.lo, out.order.begin() + mid, out.order.begin() + j.hi,
                             [&](uint32_t x, uint32_t y) { return sp[x].center[ax] < sp[y].center[ax] || (sp[x].center[ax] == sp[y].center[ax] && x < y); });
            a = (uint32_t)(out.nodes.size() / 2); b = a + 1;
            out.nodes.resize(out.nodes.size() + 4);
            todo.push_back({mid, j.hi, b});
            todo.push_back({j.lo, mid, a});
            a |= (uint32_t)ax << 30;
        }
        out.nodes[2 * (size_t)j.node] = mk4(dn_(lo[0]), dn_(lo[1]), dn_(lo[2]), up_(hi[0]));
        out.nodes[2 * (size_t)j.node + 1] = mk4(up_(hi[1]), up_(hi[2]), u2f(a), u2f(b));
    }
}

}  // namespace sqt
