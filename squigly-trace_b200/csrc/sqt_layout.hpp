// sqt_layout.hpp -- host-side derivation of the device records from the C-ABI scene description.
//
// Walks the boundary tree once (iteratively), validating it and deriving for every Branch the box
// intersectBIH' would receive for it: the root gets `bounds`, a left child gets its parent's box with
// hi[axis] := lmax, a right child the parent's box with lo[axis] := rmin (BIH.hs:130-141).  Plane values
// are copied, never computed.  Shared by sqt_backend.cu (the product) and tests/emu (host build of the
// kernel logic).
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sqt.h"
#include "sqt_paths.cuh"

namespace sqt {

struct DeviceLayout {
    std::vector<float4> nodes;      // kNodeQuads per branch
    std::vector<float4> mats;       // 3 per material
    std::vector<float4> leaves;     // 2 per leaf (leaf order = order of first appearance in the node array)
    uint32_t n_branches = 0, height = 0;
    int terminate_on_black_ok = 0;
};

inline int layout_fail(std::string &err, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    err = buf;
    return code;
}
inline float4 mk4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

inline int build_device_layout(const sqt_scene_desc &desc, DeviceLayout &out, std::string &err) {
    const sqt_scene_desc *s = &desc;
    if (!s->nodes || s->n_nodes == 0) return layout_fail(err, SQT_E_INVALID, "scene has no BIH nodes");
    if (s->n_tris && !s->tris) return layout_fail(err, SQT_E_INVALID, "tris is NULL");
    if (!s->mats || s->n_mats == 0) return layout_fail(err, SQT_E_INVALID, "scene has no materials");
    if (s->n_mats > 65535) return layout_fail(err, SQT_E_UNSUPPORTED, "more than 65535 materials");
    if (s->n_tris >= (1u << 27) || s->n_nodes >= (1u << 30)) return layout_fail(err, SQT_E_UNSUPPORTED, "scene too large for the 27/30-bit indices");
    const uint32_t N = s->n_nodes;
    std::vector<int32_t> branch_id(N, -1), leaf_id(N, -1);
    uint32_t n_br = 0, n_lf = 0;
    for (uint32_t i = 0; i < N; ++i) {
        if (!(s->nodes[i].b & SQT_NODE_LEAF)) branch_id[i] = (int32_t)n_br++;
        else leaf_id[i] = (int32_t)n_lf++;
    }
    std::vector<float4> dl((size_t)2 * (n_lf ? n_lf : 1));
    struct Box { float lo[3], hi[3]; };
    std::vector<float4> dn((size_t)kNodeQuads * (n_br ? n_br : 1));
    std::vector<uint8_t> seen(N, 0);
    std::vector<uint8_t> tri_cover(s->n_tris ? s->n_tris : 1, 0);
    struct Item { uint32_t node; Box box; uint32_t depth; };
    std::vector<Item> todo;
    Box root;
    for (int k = 0; k < 3; ++k) { root.lo[k] = s->root_bounds[k]; root.hi[k] = s->root_bounds[3 + k]; }
    todo.push_back({0u, root, 1u});
    uint32_t height = 0;
    auto leaf_meta = [&](uint32_t child, uint32_t &ref, uint32_t &meta) -> const char * {
        const sqt_node &c = s->nodes[child];
        if (c.b & SQT_NODE_LEAF) {
            const uint32_t cnt = c.b & ~SQT_NODE_LEAF, first = c.a;
            if ((uint64_t)first + cnt > s->n_tris) return "leaf triangle range out of bounds";
            ref = (uint32_t)leaf_id[child]; meta = kLeaf | cnt;
        } else { ref = (uint32_t)branch_id[child]; meta = 0; }
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
            // tight box of the triangles as Moller-Trumbore sees them (v0, v0+e1, v0+e2) and their longest edge
            double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, e2max = 0;
            for (uint32_t t = 0; t < cnt; ++t) {
                tri_cover[nd.a + t] = 1;
                const sqt_tri &tr = s->tris[nd.a + t];
                double l1 = 0, l2 = 0, l3 = 0;
                for (int k = 0; k < 3; ++k) {
                    const double v[3] = {(double)tr.v0[k], (double)tr.v0[k] + tr.e1[k], (double)tr.v0[k] + tr.e2[k]};
                    for (double x : v) { if (x < lo[k]) lo[k] = x; if (x > hi[k]) hi[k] = x; }
                    l1 += (double)tr.e1[k] * tr.e1[k]; l2 += (double)tr.e2[k] * tr.e2[k];
                    l3 += ((double)tr.e2[k] - tr.e1[k]) * ((double)tr.e2[k] - tr.e1[k]);
                }
                e2max = std::fmax(e2max, std::fmax(l1, std::fmax(l2, l3)));
            }
            if (cnt == 0) { for (int k = 0; k < 3; ++k) lo[k] = hi[k] = 0; }
            auto dn_ = [](double x) { return std::nextafterf((float)x, -INFINITY); };     // round outwards
            auto up_ = [](double x) { return std::nextafterf((float)x, INFINITY); };
            const size_t q = (size_t)2 * (size_t)leaf_id[it.node];
            dl[q] = mk4(dn_(lo[0]), dn_(lo[1]), dn_(lo[2]), up_(hi[0]));
            dl[q + 1] = mk4(up_(hi[1]), up_(hi[2]), up_(std::sqrt(e2max)), u2f(nd.a));
            continue;
        }
        const uint32_t ax = nd.a >> 30, l = nd.a & 0x3fffffffu, r = nd.b;
        if (ax > 2) return layout_fail(err, SQT_E_INVALID, "node %u: axis %u", it.node, ax);
        if (l >= N || r >= N) return layout_fail(err, SQT_E_INVALID, "node %u: child out of range", it.node);
        uint32_t lref = 0, lmeta = 0, rref = 0, rmeta = 0;
        const char *e1 = leaf_meta(l, lref, lmeta), *e2 = leaf_meta(r, rref, rmeta);
        if (e1 || e2) return layout_fail(err, SQT_E_INVALID, "node %u: %s", it.node, e1 ? e1 : e2);
        lmeta |= ax << kAxisShift;
        Box lb = it.box, rb = it.box;
        lb.hi[ax] = nd.lmax; rb.lo[ax] = nd.rmin;                           // BIH.hs:130-141
        const size_t b = (size_t)kNodeQuads * (size_t)branch_id[it.node];
        dn[b] = mk4(it.box.lo[0], it.box.lo[1], it.box.lo[2], it.box.hi[0]);
        dn[b + 1] = mk4(it.box.hi[1], it.box.hi[2], lb.hi[0], lb.hi[1]);
        dn[b + 2] = mk4(lb.hi[2], rb.lo[0], rb.lo[1], rb.lo[2]);
        dn[b + 3] = mk4(u2f(lref), u2f(rref), u2f(lmeta), u2f(rmeta));
        todo.push_back({r, rb, it.depth + 1});
        todo.push_back({l, lb, it.depth + 1});
    }
    if (height > SQT_MAX_HEIGHT) return layout_fail(err, SQT_E_UNSUPPORTED, "BIH height %u exceeds SQT_MAX_HEIGHT=%d", height, SQT_MAX_HEIGHT);
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
    out.nodes.swap(dn); out.mats.swap(dm); out.leaves.swap(dl);
    out.n_branches = n_br; out.height = height; out.terminate_on_black_ok = tob;
    return SQT_OK;
}

}  // namespace sqt
