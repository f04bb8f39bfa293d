// bih.cpp -- Bounding Interval Hierarchy build (BIH.hs:62-96) and the flattening into the C-ABI records.
//
// The build reproduces the reference tree bit for bit: same split rule, same +-0.001 padding, same
// tie-breaks, same sequential binary32 sums (SURVEY A.3 "Build").  It works on index ranges instead of
// Haskell lists, so it is O(n log n) and handles the 1M / 10M triangle configs the list-based original
// cannot reach.  Compile with -ffp-contract=off.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <thread>

#include "squigly.hpp"

namespace squigly {
namespace {

inline float proj(const V3 &v, Axis ax) { return ax == Axis::X ? v.x : (ax == Axis::Y ? v.y : v.z); }   // projectToAxis
// Ord Float class defaults
inline float hmax(float x, float y) { return x <= y ? y : x; }
inline float hmin(float x, float y) { return x <= y ? x : y; }
inline bool gt(float a, float b) { return !(a < b) && !(a == b); }     // compare == GT

// A subtree under construction: nodes with indices local to the subtree, triangle order local to the subtree.
struct Sub {
    std::vector<BIHTreeNode> nodes;
    std::vector<uint32_t> order;
};

class Builder {
  public:
    explicit Builder(BIH &out) : b_(out), tris_(out.triangles) {}

    // Large scenes: the top of the tree is split on this thread until there are a few subproblems per hardware
    // thread; the subtrees are then built concurrently (they work on disjoint index ranges) and spliced into
    // pre-order.  Every node is produced by the same arithmetic in the same order as the sequential build, so
    // the tree is bit-identical (tests/test_host.py compares both against the oracle).
    void run(unsigned threads) {
        const uint32_t n = (uint32_t)tris_.size();
        work_.resize(n); scratch_.resize(n);
        for (uint32_t i = 0; i < n; ++i) work_[i] = i;
        b_.bounds = boundingBox(0, n);
        if (threads < 2 || n < (1u << 17)) {
            Sub all;
            all.order.reserve(n);
            bih(all, b_.bounds, 0, n);
            b_.tree = std::move(all.nodes); b_.order = std::move(all.order);
            return;
        }
        grain_ = n / (4 * threads) + 1;
        const int root = top(b_.bounds, 0, n);
        std::vector<Sub> subs(tasks_.size());
        std::atomic<size_t> next{0};
        auto worker = [&]() {
            for (size_t t = next.fetch_add(1); t < tasks_.size(); t = next.fetch_add(1))
                bih(subs[t], tasks_[t].bbox, tasks_[t].lo, tasks_[t].hi);
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < threads; ++t) pool.emplace_back(worker);
        worker();
        for (auto &th : pool) th.join();
        b_.tree.clear(); b_.order.clear(); b_.order.reserve(n);
        emit(root, subs);
    }

  private:
    struct Task { Bounds bbox; uint32_t lo, hi; };
    struct TopNode { int task = -1; BIHTreeNode node; int left = -1, right = -1; };

    // boundingBox = getBounds . concatMap vertices  (Geometry.hs:155-163,196-197): minimum/maximum per axis
    Bounds boundingBox(uint32_t lo, uint32_t hi) const {
        Bounds bb;
        bool first = true;
        for (uint32_t i = lo; i < hi; ++i) {
            const Triangle &t = tris_[work_[i]];
            for (const V3 *v : {&t.tFirst, &t.tSecond, &t.tThird}) {
                if (first) { bb.lo = *v; bb.hi = *v; first = false; continue; }
                bb.lo.x = hmin(bb.lo.x, v->x); bb.hi.x = hmax(bb.hi.x, v->x);
                bb.lo.y = hmin(bb.lo.y, v->y); bb.hi.y = hmax(bb.hi.y, v->y);
                bb.lo.z = hmin(bb.lo.z, v->z); bb.hi.z = hmax(bb.hi.z, v->z);
            }
        }
        return bb;
    }
    // longestAxis (Geometry.hs:191-193): maximumBy keeps the LAST maximal element: Z over Y over X
    static Axis longestAxis(const Bounds &b) {
        const float dx = b.hi.x - b.lo.x, dy = b.hi.y - b.lo.y, dz = b.hi.z - b.lo.z;
        Axis best = Axis::Z; float bv = dz;
        if (gt(dy, bv)) { best = Axis::Y; bv = dy; }
        if (gt(dx, bv)) { best = Axis::X; }
        return best;
    }
    // averagePoints (vertices tri) projected (Geometry.hs:181-182): ((0 + a) + b) + c, then / 3
    float centroid(uint32_t tri, Axis ax) const {
        const Triangle &t = tris_[tri];
        float s = 0.0f;
        s = s + proj(t.tFirst, ax); s = s + proj(t.tSecond, ax); s = s + proj(t.tThird, ax);
        return s / 3.0f;
    }
    // split (BIH.hs:82-96) of work_[lo, hi): stable partition in place, returns the Branch node and the cut
    BIHTreeNode split(const Bounds &bbox, uint32_t lo, uint32_t hi, uint32_t &mid) {
        const uint32_t n = hi - lo;
        const Axis ax = longestAxis(bbox);
        float acc = 0.0f;
        for (uint32_t i = lo; i < hi; ++i) acc = acc + centroid(work_[i], ax);
        const float len = n <= 16777216u ? (float)n : 16777216.0f;         // genericLength at Float saturates at 2^24
        const float splitPlane = acc / len;
        uint32_t nl = 0, nr = 0;
        for (uint32_t i = lo; i < hi; ++i) {                               // filter underSplit / filter (not . underSplit)
            const uint32_t t = work_[i];
            if (centroid(t, ax) < splitPlane) work_[lo + nl++] = t; else scratch_[lo + nr++] = t;
        }
        std::copy(scratch_.begin() + lo, scratch_.begin() + lo + nr, work_.begin() + lo + nl);
        mid = lo + nl;
        float lbest = proj(bbox.lo, ax), rbest = proj(bbox.hi, ax);        // maximumDef leftSide / minimumDef rightSide
        bool any = false;
        for (uint32_t i = lo; i < mid; ++i) {
            const Triangle &t = tris_[work_[i]];
            for (const V3 *v : {&t.tFirst, &t.tSecond, &t.tThird}) { const float c = proj(*v, ax); lbest = any ? hmax(lbest, c) : c; any = true; }
        }
        any = false;
        for (uint32_t i = mid; i < hi; ++i) {
            const Triangle &t = tris_[work_[i]];
            for (const V3 *v : {&t.tFirst, &t.tSecond, &t.tThird}) { const float c = proj(*v, ax); rbest = any ? hmin(rbest, c) : c; any = true; }
        }
        BIHTreeNode nd;
        nd.leaf = false; nd.axis = ax;
        nd.lmax = 0.001f + lbest;
        nd.rmin = (-0.001f) + rbest;
        return nd;
    }
    uint32_t leaf(Sub &out, uint32_t lo, uint32_t hi) {
        BIHTreeNode n; n.leaf = true; n.first = (uint32_t)out.order.size(); n.count = hi - lo;
        out.order.insert(out.order.end(), work_.begin() + lo, work_.begin() + hi);
        out.nodes.push_back(n);
        return (uint32_t)out.nodes.size() - 1;
    }
    // bih (BIH.hs:67-80) over work_[lo, hi), appended to `out` in pre-order
    uint32_t bih(Sub &out, const Bounds &bbox, uint32_t lo, uint32_t hi) {
        if (hi - lo < 15) return leaf(out, lo, hi);                        // leafLimit = 15
        uint32_t mid;
        const BIHTreeNode nd = split(bbox, lo, hi, mid);
        const uint32_t id = (uint32_t)out.nodes.size();
        out.nodes.push_back(nd);
        uint32_t l, r;
        if (mid == lo || mid == hi) {      // one side empty: both children become leaves, recursion stops (BIH.hs:70-75)
            l = leaf(out, lo, mid); r = leaf(out, mid, hi);
        } else {
            const Bounds lb = boundingBox(lo, mid);
            l = bih(out, lb, lo, mid);
            const Bounds rb = boundingBox(mid, hi);
            r = bih(out, rb, mid, hi);
        }
        out.nodes[id].left = l; out.nodes[id].right = r;
        return id;
    }
    // the same recursion for the top of a large tree: ranges below the grain become tasks
    int top(const Bounds &bbox, uint32_t lo, uint32_t hi) {
        const int id = (int)top_.size();
        top_.emplace_back();
        if (hi - lo <= grain_) {
            top_[id].task = (int)tasks_.size();
            tasks_.push_back(Task{bbox, lo, hi});
            return id;
        }
        uint32_t mid;
        const BIHTreeNode nd = split(bbox, lo, hi, mid);
        top_[id].node = nd;
        int l, r;
        if (mid == lo || mid == hi) {      // degenerate split: two leaves, handled as (tiny or huge) tasks that are leaves
            l = (int)top_.size(); top_.emplace_back(); top_[l].task = -2; top_[l].node.first = lo; top_[l].node.count = mid - lo;
            r = (int)top_.size(); top_.emplace_back(); top_[r].task = -2; top_[r].node.first = mid; top_[r].node.count = hi - mid;
        } else {
            const Bounds lb = boundingBox(lo, mid);
            l = top(lb, lo, mid);
            const Bounds rb = boundingBox(mid, hi);
            r = top(rb, mid, hi);
        }
        top_[id].left = l; top_[id].right = r;
        return id;
    }
    // pre-order emission of the top part with the finished subtrees spliced in
    uint32_t emit(int t, const std::vector<Sub> &subs) {
        const TopNode &tn = top_[t];
        if (tn.task == -2) {               // forced leaf of a degenerate split
            BIHTreeNode n; n.leaf = true; n.first = (uint32_t)b_.order.size(); n.count = tn.node.count;
            b_.order.insert(b_.order.end(), work_.begin() + tn.node.first, work_.begin() + tn.node.first + tn.node.count);
            b_.tree.push_back(n);
            return (uint32_t)b_.tree.size() - 1;
        }
        if (tn.task >= 0) {
            const Sub &s = subs[(size_t)tn.task];
            const uint32_t node0 = (uint32_t)b_.tree.size(), tri0 = (uint32_t)b_.order.size();
            for (BIHTreeNode n : s.nodes) {
                if (n.leaf) n.first += tri0; else { n.left += node0; n.right += node0; }
                b_.tree.push_back(n);
            }
            b_.order.insert(b_.order.end(), s.order.begin(), s.order.end());
            return node0;
        }
        const uint32_t id = (uint32_t)b_.tree.size();
        b_.tree.push_back(tn.node);
        const uint32_t l = emit(tn.left, subs);
        const uint32_t r = emit(tn.right, subs);
        b_.tree[id].left = l; b_.tree[id].right = r;
        return id;
    }

    BIH &b_;
    const std::vector<Triangle> &tris_;
    std::vector<uint32_t> work_, scratch_;
    std::vector<TopNode> top_;
    std::vector<Task> tasks_;
    uint32_t grain_ = 0;
};

}  // namespace

BIH makeBIH(ParsedScene scene) {
    BIH b;
    b.triangles = std::move(scene.triangles);
    b.materials = std::move(scene.materials);
    unsigned threads = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("SQT_BIH_THREADS")) threads = (unsigned)std::atoi(e);
    Builder(b).run(threads ? threads : 1);
    return b;
}

namespace {
int heightAt(const BIH &b, uint32_t n) {
    const BIHTreeNode &t = b.tree[n];
    return t.leaf ? 1 : 1 + std::max(heightAt(b, t.left), heightAt(b, t.right));
}
}  // namespace
int height(const BIH &b) { return b.tree.empty() ? 0 : heightAt(b, 0); }
int numLeaves(const BIH &b) { int n = 0; for (const auto &t : b.tree) n += t.leaf; return n; }
int longestLeaf(const BIH &b) { int m = 0; for (const auto &t : b.tree) if (t.leaf) m = std::max(m, (int)t.count); return m; }

std::string showBIH(const BIH &b) {
    std::ostringstream o;
    o << "BIH {bounds = Bounds (V3 " << b.bounds.lo.x << ' ' << b.bounds.lo.y << ' ' << b.bounds.lo.z << ") (V3 " << b.bounds.hi.x
      << ' ' << b.bounds.hi.y << ' ' << b.bounds.hi.z << "), tree = ";
    struct F { const BIH &b; std::ostringstream &o;
        void go(uint32_t n) {
            const BIHTreeNode &t = b.tree[n];
            if (t.leaf) { o << "Leaf <" << t.count << " triangles>"; return; }
            o << "Branch (BIHN " << "XYZ"[(int)t.axis] << ' ' << t.lmax << ' ' << t.rmin << ") ("; go(t.left); o << ") ("; go(t.right); o << ")";
        } } f{b, o};
    if (!b.tree.empty()) f.go(0);
    o << "}";
    return o.str();
}

FlatBIH flattenForDevice(const BIH &b) {
    FlatBIH f;
    f.root_bounds[0] = b.bounds.lo.x; f.root_bounds[1] = b.bounds.lo.y; f.root_bounds[2] = b.bounds.lo.z;
    f.root_bounds[3] = b.bounds.hi.x; f.root_bounds[4] = b.bounds.hi.y; f.root_bounds[5] = b.bounds.hi.z;
    f.nodes.resize(b.tree.size());
    for (size_t i = 0; i < b.tree.size(); ++i) {
        const BIHTreeNode &t = b.tree[i];
        sqt_node &n = f.nodes[i];
        if (t.leaf) { n.lmax = 0; n.rmin = 0; n.a = t.first; n.b = t.count | SQT_NODE_LEAF; }
        else { n.lmax = t.lmax; n.rmin = t.rmin; n.a = t.left | ((uint32_t)t.axis << 30); n.b = t.right; }
    }
    f.tris.resize(b.order.size());
    for (size_t i = 0; i < b.order.size(); ++i) {
        const Triangle &t = b.triangles[b.order[i]];
        sqt_tri &d = f.tris[i];
        d.v0[0] = t.tFirst.x; d.v0[1] = t.tFirst.y; d.v0[2] = t.tFirst.z;
        // edge1 = vertex1 - vertex0 ; edge2 = vertex2 - vertex0   (Geometry.hs:130-131), binary32
        d.e1[0] = t.tSecond.x - t.tFirst.x; d.e1[1] = t.tSecond.y - t.tFirst.y; d.e1[2] = t.tSecond.z - t.tFirst.z;
        d.e2[0] = t.tThird.x - t.tFirst.x; d.e2[1] = t.tThird.y - t.tFirst.y; d.e2[2] = t.tThird.z - t.tFirst.z;
        d.material = t.material; d.orig_index = b.order[i]; d.pad = 0;
    }
    f.mats.resize(b.materials.size());
    for (size_t i = 0; i < b.materials.size(); ++i) {
        const Material &m = b.materials[i];
        sqt_material &d = f.mats[i];
        d.reflective = m.reflective; d.surf_color[0] = m.surfColor.x; d.surf_color[1] = m.surfColor.y; d.surf_color[2] = m.surfColor.z;
        d.emissive = m.emissive; d.emit_color[0] = m.emitColor.x; d.emit_color[1] = m.emitColor.y; d.emit_color[2] = m.emitColor.z;
    }
    return f;
}

sqt_scene_desc FlatBIH::desc() const {
    sqt_scene_desc d;
    std::memcpy(d.root_bounds, root_bounds, sizeof root_bounds);
    d.nodes = nodes.data(); d.n_nodes = (uint32_t)nodes.size();
    d.tris = tris.data(); d.n_tris = (uint32_t)tris.size();
    d.mats = mats.data(); d.n_mats = (uint32_t)mats.size();
    return d;
}

}  // namespace squigly
