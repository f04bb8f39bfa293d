// capi.cpp -- flat C view of the host library for the Python test/bench harness (ctypes).  Not part of the
// drop-in boundary (that is include/sqt.h); it only lets tests drive Obj/BIH/flatten without a C++ compiler.
#include <cstring>

#include "squigly.hpp"

using namespace squigly;

struct sqth_scene {
    BIH bih;
    FlatBIH flat;
    std::string err;
};
static thread_local std::string g_err;

extern "C" {

const char *sqth_last_error() { return g_err.c_str(); }

// trisFromObj + makeBIH + flatten
sqth_scene *sqth_load(const char *obj_path, const char *data_dir) {
    try {
        auto *s = new sqth_scene();
        s->bih = makeBIH(trisFromObj(false, readFile(obj_path), data_dir));
        s->flat = flattenForDevice(s->bih);
        return s;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
// synthetic scenes: n triangles as 9 floats, material index each, materials as 8 floats
sqth_scene *sqth_from_arrays(const float *v9, const int32_t *mat, int n, const float *m8, int n_mats) {
    try {
        ParsedScene ps;
        ps.triangles.resize(n);
        for (int i = 0; i < n; ++i) {
            const float *p = v9 + 9 * (size_t)i;
            ps.triangles[i] = Triangle{V3{p[0], p[1], p[2]}, V3{p[3], p[4], p[5]}, V3{p[6], p[7], p[8]}, (uint32_t)mat[i]};
        }
        ps.materials.resize(n_mats);
        for (int i = 0; i < n_mats; ++i) {
            const float *m = m8 + 8 * i;
            ps.materials[i] = Material{m[0], V3{m[1], m[2], m[3]}, m[4], V3{m[5], m[6], m[7]}};
        }
        auto *s = new sqth_scene();
        s->bih = makeBIH(std::move(ps));
        s->flat = flattenForDevice(s->bih);
        return s;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
void sqth_free(sqth_scene *s) { delete s; }
int sqth_n_tris(const sqth_scene *s) { return (int)s->bih.triangles.size(); }
int sqth_n_nodes(const sqth_scene *s) { return (int)s->bih.tree.size(); }
int sqth_n_mats(const sqth_scene *s) { return (int)s->bih.materials.size(); }
int sqth_height(const sqth_scene *s) { return height(s->bih); }
int sqth_num_leaves(const sqth_scene *s) { return numLeaves(s->bih); }
int sqth_longest_leaf(const sqth_scene *s) { return longestLeaf(s->bih); }
void sqth_get_tris(const sqth_scene *s, float *v9, int32_t *mat) {
    for (size_t i = 0; i < s->bih.triangles.size(); ++i) {
        const Triangle &t = s->bih.triangles[i]; float *p = v9 + 9 * i;
        p[0] = t.tFirst.x; p[1] = t.tFirst.y; p[2] = t.tFirst.z; p[3] = t.tSecond.x; p[4] = t.tSecond.y; p[5] = t.tSecond.z;
        p[6] = t.tThird.x; p[7] = t.tThird.y; p[8] = t.tThird.z; mat[i] = (int32_t)t.material;
    }
}
// the flattened records exactly as handed to sqt_upload_scene
void sqth_get_flat(const sqth_scene *s, float *root6, sqt_node *nodes, sqt_tri *tris, sqt_material *mats) {
    std::memcpy(root6, s->flat.root_bounds, 24);
    std::memcpy(nodes, s->flat.nodes.data(), s->flat.nodes.size() * sizeof(sqt_node));
    std::memcpy(tris, s->flat.tris.data(), s->flat.tris.size() * sizeof(sqt_tri));
    std::memcpy(mats, s->flat.mats.data(), s->flat.mats.size() * sizeof(sqt_material));
}
const sqt_scene_desc *sqth_desc(sqth_scene *s) {
    static thread_local sqt_scene_desc d;
    d = s->flat.desc();
    return &d;
}
int sqth_load_camera(const char *path, float *cam12) {
    try {
        Camera c = loadCamera(path);
        cam12[0] = c.position.x; cam12[1] = c.position.y; cam12[2] = c.position.z;
        std::memcpy(cam12 + 3, c.rotation, 36);
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return 1; }
}
void sqth_params(int w, int h, int spp, int bounces, int cast, int corrected, uint64_t seed, sqt_render_params *out) {
    Settings st; st.dimensions = {w, h}; st.samples = spp; st.bounces = bounces; st.cast = cast != 0; st.corrected = corrected != 0;
    st.seed = seed;
    *out = paramsFromSettings(st);
}
int sqth_write_png(const char *path, const uint8_t *rgb8, int rows, int cols) {
    try { writeImage(path, rgb8, rows, cols); return 0; } catch (const std::exception &e) { g_err = e.what(); return 1; }
}

}  // extern "C"
