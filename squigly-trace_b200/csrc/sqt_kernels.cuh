// sqt_kernels.cuh -- the sm_100a kernels of the squigly-trace B200 backend (included by sqt_backend.cu only).
//
//   k_intersect_batch  : one lane per ray, grid-stride; batched Scene.intersect (Geometry.hs:64 / BIH.hs:101)
//   k_primary          : one lane per pixel; makeRay (Lib.hs:107-114) + closest hit, cached per pixel
//   k_paths_pool       : the integrator: every warp owns a pool of 32*K rays in shared memory, keeps them in three
//                        queues by the kind of step they need and serves one queue per round with all lanes (default)
//   k_paths            : the integrator with one ray per lane and warp-synchronous phases (SQT_POOL=0)
//   k_accumulate       : adds a round's samples to the per-pixel sums in sample order (Lib.hs:88)
//   k_raycast          : --cast mode (Lib.hs:141-151)
//   k_tonemap          : mean + rgbFloatToPixelRGB (Lib.hs:88-104)
//   k_leaf_records, k_check_materials, k_branch_tight, k_child_slabs, k_flag_slabs : scene upload (leaf records from the triangles; material
//                        index validation; tight subtree slabs bottom-up)
//   k_fp32_peak, k_l2_read : roofline denominators measured on the device
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <type_traits>

#include "sqt_paths.cuh"

namespace cg = cooperative_groups;
using namespace sqt;

struct DeviceStats {
    unsigned long long rays, samples, primary_reused;
    unsigned long long branch_visits, child_box_tests, tri_tests, leaves_culled;
    unsigned long long mt_pass_a, mt_pass_u, mt_pass_v, mt_accept;
    unsigned long long n_hit;           // length of the pixel list k_primary builds
    unsigned long long work_next[256];  // dynamic work counters of the path kernels, one per round
    unsigned long long dbg[48];         // scheduler diagnostics of k_paths_pool (instrumented variant only), see SQT_DEBUG_STATS
};

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
// every lane of the block must call this (full warps)
__device__ __forceinline__ void flush_stats(DeviceStats *ds, const PathStats &st, const Counters &cn, bool count) {
    unsigned long long r = warp_sum(st.rays), s = warp_sum(st.samples), p = warp_sum(st.primary_reused);
    unsigned long long b = 0, c = 0, t = 0;
    unsigned long long lc = 0, ga = 0, gu = 0, gv = 0, gt = 0;
    if (count) {
        b = warp_sum(cn.branch_visits); c = warp_sum(cn.child_box_tests); t = warp_sum(cn.tri_tests); lc = warp_sum(cn.leaves_culled);
        ga = warp_sum(cn.mt_pass_a); gu = warp_sum(cn.mt_pass_u); gv = warp_sum(cn.mt_pass_v); gt = warp_sum(cn.mt_accept);
    }
    if ((threadIdx.x & 31) == 0) {
        if (r) atomicAdd(&ds->rays, r);
        if (s) atomicAdd(&ds->samples, s);
        if (p) atomicAdd(&ds->primary_reused, p);
        if (count) { atomicAdd(&ds->branch_visits, b); atomicAdd(&ds->child_box_tests, c); atomicAdd(&ds->tri_tests, t); atomicAdd(&ds->leaves_culled, lc);
            atomicAdd(&ds->mt_pass_a, ga); atomicAdd(&ds->mt_pass_u, gu); atomicAdd(&ds->mt_pass_v, gv); atomicAdd(&ds->mt_accept, gt); }
    }
}

// Scheduling knobs of the warp-synchronous loop (runtime so that they can be tuned without rebuilding; they
// change the order in which lanes get served, never a result):
//   a_leave : leave the traversal phase once at most this many lanes still want a traversal step
//   b_leave : leave the triangle phase once fewer than this many lanes still have triangles to test
//             (0 = the warp-cooperative triangle phase below, which always runs to completion)
//   c_min   : run the regeneration phase only when at least this many lanes are done (or nothing else can run)
struct Tune { int a_leave, b_leave, c_min; };

// Warp-cooperative triangle phase of the one-ray-per-lane kernels.  The lanes that wait in a leaf hold (first triangle,
// triangles left); their remaining (ray, triangle) tests are laid out consecutively by an exclusive scan and executed 32 at
// a time, one test per lane, whichever lane owns the ray: the owner of test p is found by a binary search over the scan
// (shuffles), the ray comes from the owner by shuffle.  Accepted hits (about one test in seventy) are handed back
// to the owner one after the other in test order, i.e. from the leaf's last triangle to its first, so every ray
// sees exactly the sequence of min' applications of BIH.hs:105-109 (base-4.9 minimumBy = foldr1 min').
// Moller-Trumbore is a pure function of (ray, triangle), so only the lane that evaluates it changes.
template <bool COUNT>
__device__ __forceinline__ void leaf_pairs(const SceneView &sc, TravLane &L, Counters *cn) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const bool in_leaf = L.state == ST_LEAF;
    const int left = in_leaf ? L.i + 1 : 0;
    const int cnt = left < 1024 ? left : 1024;            // a pathological leaf is worked off over several phases
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    const int start = incl - cnt;
    const int total = __shfl_sync(FULL, incl, 31);
    for (int base = 0; base < total; base += 32) {
        const int pr = base + lane;
        int own = 0;                                       // first lane whose inclusive scan exceeds pr
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const int v = __shfl_sync(FULL, incl, own + s - 1);
            if (v <= pr) own += s;
        }
        Ray r;
        r.ox = __shfl_sync(FULL, L.r.ox, own); r.oy = __shfl_sync(FULL, L.r.oy, own); r.oz = __shfl_sync(FULL, L.r.oz, own);
        r.dx = __shfl_sync(FULL, L.r.dx, own); r.dy = __shfl_sync(FULL, L.r.dy, own); r.dz = __shfl_sync(FULL, L.r.dz, own);
        const uint32_t first = __shfl_sync(FULL, L.child, own);
        const int oi = __shfl_sync(FULL, L.i, own), os = __shfl_sync(FULL, start, own);
        const uint32_t idx = first + (uint32_t)(oi - (pr - os));
        bool hit = false;
        float t = 0.0f, dist = 0.0f;
        if (pr < total) {
            const TriData d = tri_load(sc, idx);
            int stage;
            hit = moller_trumbore(d.a0, d.a1, d.a2, r, t, dist, stage);
            if (COUNT) { cn->mt_pass_a += stage >= 1; cn->mt_pass_u += stage >= 2; cn->mt_pass_v += stage >= 3; cn->mt_accept += stage >= 4; }
        }
        unsigned hm = __ballot_sync(FULL, hit);
        while (hm != 0u) {                                 // rare: hand each accepted hit to its owner, in test order
            const int src = __ffs(hm) - 1;
            hm &= hm - 1u;
            const int o_s = __shfl_sync(FULL, own, src);
            const float t_s = __shfl_sync(FULL, t, src), d_s = __shfl_sync(FULL, dist, src);
            const uint32_t i_s = __shfl_sync(FULL, idx, src);
            if (lane == o_s) fold_earlier(L.cur, (int)i_s, t_s, d_s);
        }
    }
    if (in_leaf) {
        L.i -= cnt;
        if (L.i < 0) L.state = ST_RET;
    }
}

// The persistent warp loop of the one-ray-per-lane kernels.  Every lane of the warp stays in it until all 32 have run
// out of work.  A round is three phases, each a tight loop whose trip count is decided by a warp vote: regeneration
// (consume the finished hit, shade, make the next ray), traversal steps (stack pops + one branch visit), triangle steps
// (one Moller-Trumbore test).  The votes force the 32 lanes back together at every phase boundary; an ordinary
// per-lane loop nest compiles to code where the lanes drift apart through the data-dependent traversal and
// never reconverge (measured: 2.4 of 32 lanes active, profiles/r01_k_paths_v0_divergent.txt).
template <class P, class = void> struct has_warp_regen : std::false_type {};
template <class P> struct has_warp_regen<P, std::void_t<decltype(P::kWarpRegen)>> : std::true_type {};

template <bool COUNT, class Policy>
__device__ __forceinline__ void warp_loop(const SceneView &sc, Policy &pol, Counters *cn, const Tune tn) {
    const unsigned FULL = 0xffffffffu;
    float4 stack[kStackEntries];
    TravLane L;
    L.stack = stack;
    L.state = ST_DONE; L.sp = 0; L.i = 0; L.child = 0u; L.tmin = 0.0f; L.tmax = 0.0f; L.rf = 0u;
    L.cur.tri = -1; L.cur.t = 0.0f; L.cur.dist = 0.0f;
    L.dfx = L.dfy = L.dfz = 0.0f;
    L.r.ox = L.r.oy = L.r.oz = L.r.dx = L.r.dy = L.r.dz = 0.0f;
    const LaneRay ra(L);
    for (;;) {
        // ---- (extension) rays whose BIH part is finished fold in the analytic spheres
        if (__any_sync(FULL, L.state == ST_SPH)) {
            if (L.state == ST_SPH) sphere_step(sc, L, ra);
            __syncwarp(FULL);
        }
        // ---- regeneration
        const unsigned m_done = __ballot_sync(FULL, L.state == ST_DONE);
        const unsigned m_busy = __ballot_sync(FULL, L.state == ST_DESC || L.state == ST_RET || L.state == ST_LEAF || L.state == ST_ENTER);
        if (m_done != 0u && (__popc(m_done) >= tn.c_min || m_busy == 0u)) {
            if constexpr (has_warp_regen<Policy>::value) pol.template regen_warp<COUNT>(sc, L, cn, L.state == ST_DONE);
            else { if (L.state == ST_DONE) pol.template regen<COUNT>(sc, L, cn); }
            __syncwarp(FULL);
        } else if (m_busy == 0u) break;                       // every lane is ST_EXIT
        // ---- traversal steps (stack pops + one branch visit) while more than a_leave lanes want one; then every
        //      lane that found a leaf enters it (record fetch + conservative culling; culled lanes traverse on).
        //      The few stragglers left over keep their state and continue next round.
        for (;;) {
            const unsigned mt = __ballot_sync(FULL, L.state == ST_DESC || L.state == ST_RET);
            if (__popc(mt) > tn.a_leave) {
                if (L.state == ST_RET) ret_step(sc, L, ra);
                if (L.state == ST_DESC) desc_step<COUNT>(sc, L, ra, cn);
                continue;
            }
            if (__any_sync(FULL, L.state == ST_ENTER)) {
                if (L.state == ST_ENTER) enter_step<COUNT>(sc, L, ra, cn);
                continue;
            }
            if (mt != 0u && !__any_sync(FULL, L.state == ST_LEAF)) {      // only stragglers are left and nobody has triangles
                if (L.state == ST_RET) ret_step(sc, L, ra);
                if (L.state == ST_DESC) desc_step<COUNT>(sc, L, ra, cn);
                continue;
            }
            break;
        }
        // ---- triangle tests: all lanes share the tests of the lanes that wait in a leaf, or one test per lane and step
        unsigned m = __ballot_sync(FULL, L.state == ST_LEAF);
        if (tn.b_leave == 0) {
            if (m != 0u) leaf_pairs<COUNT>(sc, L, cn);
            continue;
        }
        while (m != 0u) {
            if (L.state == ST_LEAF) tri_step<COUNT>(sc, L, cn);
            m = __ballot_sync(FULL, L.state == ST_LEAF);
            if (__popc(m) < tn.b_leave) break;
        }
    }
}

template <bool COUNT>
__global__ void __launch_bounds__(128) k_intersect_batch(SceneView sc, const float *__restrict__ org,
                                                         const float *__restrict__ dir, long long n,
                                                         int *__restrict__ tri_out, float *__restrict__ dist_out,
                                                         float *__restrict__ point_out, DeviceStats *ds, Tune tn) {
    Counters cn = {};
    PathStats st = {0, 0, 0};
    BatchPolicy pol(org, dir, n, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x, tri_out,
                    dist_out, point_out, st);
    warp_loop<COUNT>(sc, pol, &cn, tn);
    flush_stats(ds, st, cn, COUNT);
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }

struct DeviceAppend {
    unsigned long long *counter;
    int *list;
    __device__ __forceinline__ void operator()(long long pixel) {
        cg::coalesced_group g = cg::coalesced_threads();
        unsigned long long base = 0;
        if (g.thread_rank() == 0) base = atomicAdd(counter, (unsigned long long)g.size());
        base = g.shfl(base, 0);
        list[base + g.thread_rank()] = (int)pixel;
    }
};

template <bool COUNT>
__global__ void __launch_bounds__(128) k_primary(SceneView sc, RenderParams p, int2 *__restrict__ prim, int *__restrict__ pixel_list,
                                                 DeviceStats *ds, Tune tn) {
    Counters cn = {};
    PathStats st = {0, 0, 0};
    DeviceAppend app = {&ds->n_hit, pixel_list};
    PrimaryPolicy<DeviceAppend> pol(p, prim, app, work_items(p), (long long)blockIdx.x * blockDim.x + threadIdx.x,
                                    (long long)gridDim.x * blockDim.x, st);
    warp_loop<COUNT>(sc, pol, &cn, tn);
    flush_stats(ds, st, cn, COUNT);
}

struct DeviceFetch {
    unsigned long long *counter;
    long long n;
    __device__ __forceinline__ long long operator()() {
        // warp-aggregated claim: one atomic per group of lanes that need work at the same time
        cg::coalesced_group g = cg::coalesced_threads();
        unsigned long long base = 0;
        if (g.thread_rank() == 0) base = atomicAdd(counter, (unsigned long long)g.size());
        base = g.shfl(base, 0);
        const long long w = (long long)(base + g.thread_rank());
        return w < n ? w : -1;
    }
};

#ifndef SQT_WL_MIN_BLOCKS
#define SQT_WL_MIN_BLOCKS 6
#endif
// One round of samples (sqt_paths.cuh): persistent lanes pull (pixel, sample) items from the round's counter.
template <bool COUNT>
__global__ void __launch_bounds__(128, SQT_WL_MIN_BLOCKS) k_paths(SceneView sc, RenderParams p, RoundInfo rd, int round, DeviceStats *ds, Tune tn) {
    Counters cn = {};
    PathStats st = {0, 0, 0};
    if (rd.pixel_list) rd.n_slots = (long long)ds->n_hit;
    DeviceFetch fetch = {&ds->work_next[round], rd.n_slots << rd.log2_s};
    uint16_t pm[SQT_MAX_DEPTH];
    PathPolicy<DeviceFetch> pol(p, rd, fetch, st, pm);
    warp_loop<COUNT>(sc, pol, &cn, tn);
    flush_stats(ds, st, cn, COUNT);
}

// ------------------------------------------------------------------------------ ray pools
// k_paths_pool: the same per-ray logic as k_paths, scheduled differently.  Every warp owns a POOL of P = 32*K rays whose
// state lives in shared memory (structure of arrays, 16 words per ray; traversal stacks, per-path material lists and the
// integrator state of a slot in global memory, one region per pool slot).  A ray waits in one of four queues (byte rings of
// slot ids in shared memory, heads and fills packed in two warp-uniform registers; a served ray is appended with one MATCH + one
// packed REDUX at write-back, so there is no census and no gather):
//   T  traversal steps (ST_DESC): a short burst of branch visits; the ray's interval and child travel in registers, its
//      origin / direction components are read from the pool by split axis (PoolRayS)
//   L  leaf work (ST_ENTER / ST_LEAF): every gathered ray enters its leaf (record fetch + conservative culling), then the
//      (ray, triangle) tests of ALL gathered rays are laid out consecutively and executed 32 per step, one test per lane,
//      whichever lane gathered the ray -- any lane can read any ray from the pool -- and accepted hits are folded into
//      the owning ray's best hit in test order (from the leaf's last triangle to its first: BIH.hs:105-109, minimumBy =
//      foldr1 min'); a test first passes a division-free filter for the `a` / `u` guards, the survivors are re-run in full,
//      32 at a time; then the stack is popped right there with all gathered lanes (ret_step)
//   R  regeneration (ST_DONE): consume the hit, shade, bounce or fetch the next sample, start the ray (staged for all
//      gathered lanes at once, path_regen_warp)
//   S  (extension) fold the analytic spheres into the BIH result (ST_SPH)
// Each round the warp serves the longest queue with up to 32 lanes.  With one ray per lane at most ~10 of 32 lanes share a
// step kind at any time (tests/sched_sim.py); regrouping rays lifts that limit.  Scheduling never changes a result: every
// ray runs the same unit steps of sqt_core.cuh in the same order (test_every_path_kernel_scheduler_is_bit_exact).
#ifndef SQT_POOL_MIN_BLOCKS
#define SQT_POOL_MIN_BLOCKS 8
#endif
// The pool is read and written by 32-bit shared-memory address.  A generic pointer into shared memory makes the compiler rebuild
// the shared window base (S2UR SR_CgaCtaId, UMOV, ULEA, LEA) wherever it is short of registers -- inside the branch visit, in
// every chunk of the leaf phase -- and that sequence sits in front of the LDS it feeds.  The base is taken once, passed through a
// REDUX so that it lives in a uniform register, and every access is `[lane register + uniform base + immediate]`.
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
struct SharedWord {                       // PW(f, g) = x ; x = PW(f, g)
    uint32_t a;
    __device__ __forceinline__ operator uint32_t() const { return lds32(a); }
    __device__ __forceinline__ void operator=(uint32_t v) const { sts32(a, v); }
    SharedWord &operator=(const SharedWord &) = delete;
};
// PoolRay (sqt_core.cuh) by shared address: word k of the slot at a + k * STRIDE_B
template <int STRIDE_B>
struct PoolRayS {
    uint32_t a;
    __device__ __forceinline__ float f(int k) const { return __uint_as_float(lds32(a + (uint32_t)(k * STRIDE_B))); }
    __device__ __forceinline__ float o(int ax) const { return f(ax); }
    __device__ __forceinline__ float d(int ax) const { return f(3 + ax); }
    __device__ __forceinline__ float df(int ax) const { return f(6 + ax); }
    __device__ __forceinline__ Ray ray() const { Ray r; r.ox = f(0); r.oy = f(1); r.oz = f(2); r.dx = f(3); r.dy = f(4); r.dz = f(5); return r; }
    __device__ __forceinline__ void dfv(float &x, float &y, float &z) const { x = f(6); y = f(7); z = f(8); }
};

// burst_t: traversal steps per T round (at most); t_leave: end the burst early once at most this many lanes still traverse;
// c_min: serve the regeneration queue only when it holds at least this many rays (or nothing else can run)
struct PoolTune { int burst_t, t_leave, c_min; };

// 16 words = 64 B per ray in shared memory; the first nine are what PoolRayS reads.  PF_TMIN holds the interval's lower
// end while the ray descends and `i` (triangles left - 1) for a ray that starts inside a leaf (root leaf);
// PF_FLAGS = TravLane::rf (bits 0-2, 27, 30) | state << 4 | sp << 8.
enum { PF_OX = 0, PF_OY, PF_OZ, PF_DX, PF_DY, PF_DZ, PF_DFX, PF_DFY, PF_DFZ, PF_CHILD, PF_TMIN, PF_TMAX, PF_CTRI, PF_CT, PF_CDIST,
       PF_FLAGS, PF_WORDS };
enum { KT = 0, KL = 1, KR = 2, KS = 3, KNONE = 4 };

template <bool COUNT, int K>
__global__ void __launch_bounds__(128, SQT_POOL_MIN_BLOCKS) k_paths_pool(SceneView sc, RenderParams p, RoundInfo rd, int round, DeviceStats *ds, PoolTune tn,
                                                    float4 *__restrict__ gstack, uint16_t *__restrict__ gpm, uint4 *__restrict__ gpath, int stack_depth, int pm_stride) {
    extern __shared__ uint32_t pool_smem[];
    constexpr int P = 32 * K;                     // rays per warp (a power of two)
    constexpr int PT = 4 * P;                     // pool slots of the CTA: slot g = warp * P + s, and g is what the queues hold,
                                                  // so the hot accesses index CTA-wide arrays by a value the lane already has
                                                  // (one uniform base, see lds32)
    // per-warp scratch, kept small on purpose: 8 CTAs x (pool + scratch + 1 KB) must stay within the 164 KB shared-memory
    // carve-out, the next one (196 KB) would leave 32 KB instead of 64 KB of L1 for triangles, nodes and stacks
    constexpr int AUX_WORDS = 4 * P / 4 + 64 + 8 + 8 + 32;
    const unsigned FULL = 0xffffffffu;
    // (the warp index comes out of a REDUX, i.e. in a uniform register: the per-warp scratch addresses derived from it are then
    //  uniform too and are not rebuilt from SR_TID in every loop that is short of registers)
    const int warp = (int)__reduce_max_sync(FULL, threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int wbase = warp * P;
    const uint32_t sbase = __reduce_max_sync(FULL, (uint32_t)__cvta_generic_to_shared(pool_smem));      // word f of slot g at sbase + 4 (f PT + g)
    // per-warp scratch (shared addresses, warp-uniform):
    const uint32_t a_queue = sbase + 4u * (uint32_t)(PT * PF_WORDS + warp * AUX_WORDS);   // bytes: queue[k * P + i], k = KT, KL, KR, KS: rings of slot ids (g - wbase)
    const uint32_t a_surv = a_queue + 4u * P;                                   // words: ring of 64 (triangle | gathering lane << 27)
    const uint32_t a_lane_slot = a_surv + 256u;                                 // bytes: leaf round, the slot (g - wbase) each lane gathered
    const uint32_t a_own_lane = a_lane_slot + 32u;                              // bytes: leaf round, per owner rank: its lane ...
    const uint32_t a_own_tri = a_lane_slot + 64u;                               // words: ... and triangle of its test 0 + that test's position
    const int gbase = (int)(blockIdx.x * PT);                                   // global slot = gbase + g (< 2^31: a few hundred thousand exist)
    float4 *cstack = gstack + (size_t)gbase * (size_t)stack_depth;              // entry e of slot g at cstack[e * PT + g]
#define PW(f, g) SharedWord{sbase + 4u * (uint32_t)((f) * PT + (g))}
    typedef PoolRayS<4 * PT> SlotRay;
    Counters cn = {};
    PathStats st = {0, 0, 0};
    if (rd.pixel_list) rd.n_slots = (long long)ds->n_hit;
    DeviceFetch fetch = {&ds->work_next[round], rd.n_slots << rd.log2_s};
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int slot = wbase + lane + 32 * k;
        PW(PF_FLAGS, slot) = (uint32_t)ST_DONE << 4;
        PW(PF_CTRI, slot) = 0xffffffffu;
        sts8(a_queue + (uint32_t)(KR * P + lane + 32 * k), (uint32_t)(lane + 32 * k));
        gpath[2 * (gbase + slot)] = make_uint4(0u, 0u, 0u, 0u);
        gpath[2 * (gbase + slot) + 1] = make_uint4(0u, 0u, 0u, 0u);      // saved_j = -1 (stored +1), any_emit = in_flight = false
    }
    __syncwarp(FULL);
    unsigned lt_mask, le_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(le_mask));
    // queue state, warp-uniform, one byte per kind: ring head (< P) and fill (<= P <= 128)
    unsigned q_head = 0u, q_cnt = (unsigned)P << (8 * KR);     // (P = 128 fills its byte exactly: the top byte, KS, never holds more than P)
    unsigned long long dbg_rounds[3] = {0, 0, 0}, dbg_sel[3] = {0, 0, 0}, dbg_desc[8] = {}, dbg_ret[8] = {}, dbg_leaf[4] = {};
    for (;;) {
        // ---- pick the longest queue (regeneration only in batches, or when nothing else can run)
        const int n_t = (int)(q_cnt & 0xffu), n_l = (int)((q_cnt >> 8) & 0xffu), n_r = (int)((q_cnt >> 16) & 0xffu), n_s = (int)(q_cnt >> 24);
        const int c_r = (n_r >= tn.c_min || (q_cnt & 0xff00ffffu) == 0u) ? n_r : 0;
        int kind = KL, best = n_l;
        if (n_t > best) { kind = KT; best = n_t; }
        if (n_s > best) { kind = KS; best = n_s; }
        if (c_r > best) { kind = KR; best = c_r; }
        if (best == 0) break;                                               // every slot is ST_EXIT
        // ---- pop up to 32 rays of that kind onto the lanes
        const int n_sel = best < 32 ? best : 32;
        const bool act = lane < n_sel;
        int slot = wbase;                                                   // (inactive lanes: any valid slot of this warp)
        {
            const int sh = 8 * kind;
            const unsigned h = (q_head >> sh) & 0xffu;
            if (act) slot = wbase + (int)lds8(a_queue + (uint32_t)(kind * P) + ((h + (unsigned)lane) & (unsigned)(P - 1)));
            q_head = (q_head & ~(0xffu << sh)) | (((h + n_sel) & (P - 1)) << sh);
            q_cnt -= (unsigned)n_sel << sh;
        }
        TravLane L;
        L.stack = nullptr;
        L.state = ST_EXIT;
        const SlotRay ra{sbase + 4u * (uint32_t)slot};
        if (COUNT && kind < 3) { dbg_rounds[kind] += 1; dbg_sel[kind] += n_sel; }
        if (kind == KT) {
            // ---- traversal steps: branch visits only.  Stack pops happen at the end of the leaf round (stage 3), so the rare
            //      ray whose visit ends in ST_RET here (neither child box hit) simply queues for a leaf round with no leaf
            //      work; the best hit is not touched by a visit and stays in the pool.
            uint32_t fl = 0u;
            L.stack = cstack + slot;
            if (act) {
                L.child = PW(PF_CHILD, slot); L.tmin = u2f(PW(PF_TMIN, slot)); L.tmax = u2f(PW(PF_TMAX, slot));
                fl = PW(PF_FLAGS, slot);
                L.state = (int)((fl >> 4) & 15u); L.rf = fl; L.sp = (int)((fl >> 8) & 0xffu);
            }
            for (int b = 0; b < tn.burst_t; ++b) {
                if (COUNT) { dbg_desc[b < 7 ? b : 7] += __popc(__ballot_sync(FULL, L.state == ST_DESC)); }
                if (L.state == ST_DESC) desc_step<COUNT, PT>(sc, L, ra, &cn);
                if (__popc(__ballot_sync(FULL, L.state == ST_DESC)) <= tn.t_leave) break;
            }
            if (act) {
                PW(PF_CHILD, slot) = L.child; PW(PF_TMIN, slot) = f2u(L.tmin); PW(PF_TMAX, slot) = f2u(L.tmax);
                if (L.state == ST_RET) PW(PF_CTRI, slot) = 0xffffffffu;     // desc_step: the subtree returned Nothing
                PW(PF_FLAGS, slot) = (fl & 0xffff000fu) | ((uint32_t)L.state << 4) | ((uint32_t)L.sp << 8);
            }
        } else if (kind == KL) {
            // ---- leaf work.  Stage 1: enter the leaf (record fetch + conservative culling)
            uint32_t fl = 0u;
            uint32_t first = 0u;
            int rem = 0;                                                    // triangles of this lane's ray still to test
            sts8(a_lane_slot + (uint32_t)lane, (uint32_t)(slot & (P - 1)));
            if (act) {
                fl = PW(PF_FLAGS, slot);
                L.child = PW(PF_CHILD, slot);
                L.rf = fl;
                L.state = (int)((fl >> 4) & 15u);
                L.i = (int)PW(PF_TMIN, slot);                               // only meaningful for a ray that started inside a (root) leaf
                L.cur.tri = (int)PW(PF_CTRI, slot);
                if (L.state == ST_ENTER) {
                    enter_step<COUNT>(sc, L, ra, &cn);
                    PW(PF_CTRI, slot) = 0xffffffffu;                        // enter_step: cur = Nothing
                }
                if (L.state == ST_LEAF) { first = L.child; rem = L.i + 1; }
            }
            __syncwarp(FULL);
            // Stage 2: all (ray, triangle) tests of the gathered rays, 32 per step.  Test p of a pass belongs to the lane
            // whose inclusive scan first exceeds p; within a ray tests run from its last triangle to its first.  A test first
            // goes through a division-free filter for the `a` and `u` guards (moller_trumbore_au); the ~18 % that survive are appended -- in test order --
            // to a small ring in shared memory and re-run in full, 32 at a time, so that the expensive tail of
            // Moller-Trumbore (second cross product, v / t guards, point, IEEE sqrt) also executes with all lanes.
            int n_surv = 0, surv_head = 0;                                  // warp-uniform
            auto run_survivors = [&](int n) {
                const uint32_t e = lane < n ? lds32(a_surv + 4u * (uint32_t)((surv_head + lane) & 63)) : 0u;
                const int oslot = wbase + (int)lds8(a_lane_slot + (e >> 27));
                const uint32_t idx = e & kLeafFirstMask;
                bool hit = false;
                float t = 0.0f, dist = 0.0f;
                if (lane < n) {
                    const TriData d = tri_load(sc, idx);
                    const Ray r = SlotRay{sbase + 4u * (uint32_t)oslot}.ray();
                    int stage;
                    hit = moller_trumbore(d.a0, d.a1, d.a2, r, t, dist, stage);
                    if (COUNT) { cn.mt_pass_u += stage >= 2; cn.mt_pass_v += stage >= 3; cn.mt_accept += stage >= 4; }
                }
                unsigned hm = __ballot_sync(FULL, hit);
                while (hm != 0u) {                                          // fold each accepted hit into its ray, in test order
                    const int src = __ffs(hm) - 1;
                    hm &= hm - 1u;
                    if (lane == src) {                                      // fold_earlier on the ray's best hit in the pool
                        if ((int)PW(PF_CTRI, oslot) < 0 || !cmp_gt(dist, u2f(PW(PF_CDIST, oslot)))) {
                            PW(PF_CTRI, oslot) = idx; PW(PF_CT, oslot) = f2u(t); PW(PF_CDIST, oslot) = f2u(dist);
                        }
                    }
                    __syncwarp(FULL);
                }
                surv_head = (surv_head + n) & 63; n_surv -= n;
            };
            for (;;) {
                const int c = rem < 1024 ? rem : 1024;                     // a pathological leaf is worked off over several passes
                int incl = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(FULL, incl, 31);
                if (total == 0) break;
                if (COUNT) { dbg_leaf[0] += 1; dbg_leaf[1] += __popc(__ballot_sync(FULL, c > 0)); dbg_leaf[2] += total; dbg_leaf[3] += (total + 31) / 32; }
                // The rays that still have triangles ("owners"), in lane order, own consecutive ranges of tests.  Each owner
                // publishes (its lane, triangle of its test 0 + that test's position) in a small table by rank; per chunk of 32
                // tests the owners whose range starts inside the chunk set one bit each (one REDUX), and a test finds its
                // owner's rank by counting the bits at or below its position -- no search, two table reads.
                const unsigned om = __ballot_sync(FULL, c > 0);
                if (c > 0) {
                    const int rank = __popc(om & lt_mask);
                    sts8(a_own_lane + (uint32_t)rank, (uint32_t)lane);
                    sts32(a_own_tri + 4u * (uint32_t)rank, (uint32_t)((int)first + rem - 1 + (incl - c)));           // triangle of test p of this ray = that - p
                }
                __syncwarp(FULL);
                const int start = incl - c;
                int started = 0;                                            // owners whose range starts before `base` (warp-uniform)
                // the mapping of a chunk is computed one iteration ahead, and the chunk's triangles are prefetched into L1
                // while the previous chunk is tested
                auto map_chunk = [&](int base, int &own_o, uint32_t &idx_o, int &oslot_o) {
                    const unsigned rel = (unsigned)(start - base);
                    const unsigned sm = __reduce_or_sync(FULL, (c > 0 && rel < 32u) ? (1u << rel) : 0u);
                    int k = started + __popc(sm & le_mask) - 1;
                    started += __popc(sm);
                    if (base + lane >= total) k = 0;
                    own_o = (int)lds8(a_own_lane + (uint32_t)k);
                    idx_o = lds32(a_own_tri + 4u * (uint32_t)k) - (uint32_t)(base + lane);
                    oslot_o = wbase + (int)lds8(a_lane_slot + (uint32_t)own_o);
                    if (base + lane < total) prefetch_l1(sc.tris + 3 * (size_t)idx_o);
                };
                int own_n = 0, oslot_n = 0;
                uint32_t idx_n = 0u;
                map_chunk(0, own_n, idx_n, oslot_n);
                for (int base = 0; base < total; base += 32) {
                    const int pr = base + lane;
                    const int own = own_n, oslot = oslot_n;
                    const uint32_t idx = idx_n;
                    if (base + 32 < total) map_chunk(base + 32, own_n, idx_n, oslot_n);
                    bool pass = false;
                    if (pr < total) {
                        const TriData d = tri_load(sc, idx);
                        const Ray r = SlotRay{sbase + 4u * (uint32_t)oslot}.ray();
                        bool pass_a;
                        pass = moller_trumbore_au(d.a0, d.a1, d.a2, r, pass_a);
                        if (COUNT) cn.mt_pass_a += pass_a;
                    }
                    const unsigned pm = __ballot_sync(FULL, pass);
                    if (pass) sts32(a_surv + 4u * (uint32_t)((surv_head + n_surv + __popc(pm & lt_mask)) & 63), idx | ((uint32_t)own << 27));
                    n_surv += __popc(pm);
                    __syncwarp(FULL);
                    if (n_surv >= 32) run_survivors(32);
                }
                rem -= c;
            }
            if (n_surv > 0) run_survivors(n_surv);
            // Stage 3: every gathered ray now holds the result of its leaf (or Nothing, if the leaf was culled or empty):
            // pop its stack right here, with all gathered lanes, instead of in a traversal round -- a third of the rays go
            // straight on to another leaf (the far child) or are finished and never need a traversal round in between.
            L.stack = cstack + slot;
            if (act) {
                L.cur.tri = (int)PW(PF_CTRI, slot); L.cur.t = u2f(PW(PF_CT, slot)); L.cur.dist = u2f(PW(PF_CDIST, slot));
                L.rf = fl; L.sp = (int)((fl >> 8) & 0xffu);
                L.tmin = 0.0f; L.tmax = 0.0f; L.child = 0u;
                L.state = ST_RET;
                ret_step<PT>(sc, L, ra);
                PW(PF_CHILD, slot) = L.child; PW(PF_TMIN, slot) = f2u(L.tmin); PW(PF_TMAX, slot) = f2u(L.tmax);
                PW(PF_CTRI, slot) = (uint32_t)L.cur.tri; PW(PF_CT, slot) = f2u(L.cur.t); PW(PF_CDIST, slot) = f2u(L.cur.dist);
                PW(PF_FLAGS, slot) = (fl & 0xffff000fu) | ((uint32_t)L.state << 4) | ((uint32_t)L.sp << 8);
            }
        } else if (kind == KS) {
            // ---- (extension) fold the analytic spheres into the BIH result of every gathered ray (sphere_step)
            if (act) {
                const uint32_t fl = PW(PF_FLAGS, slot);
                L.cur.tri = (int)PW(PF_CTRI, slot); L.cur.t = u2f(PW(PF_CT, slot)); L.cur.dist = u2f(PW(PF_CDIST, slot));
                L.state = ST_SPH;
                sphere_step(sc, L, ra);
                PW(PF_CTRI, slot) = (uint32_t)L.cur.tri; PW(PF_CT, slot) = f2u(L.cur.t); PW(PF_CDIST, slot) = f2u(L.cur.dist);
                PW(PF_FLAGS, slot) = (fl & ~0xf0u) | ((uint32_t)L.state << 4);
            }
        } else {
            // ---- regeneration: consume the finished hit, shade, start the next ray (or the next sample); staged, all
            //      gathered lanes together (path_regen_warp)
            PathRay q;
            path_ray_init(q);
            uint4 *gp = gpath + 2 * (gbase + slot);
            L.dfx = L.dfy = L.dfz = 0.0f; L.child = 0u; L.tmin = 0.0f; L.tmax = 0.0f; L.i = 0; L.sp = 0; L.rf = 0u;
            L.r.ox = L.r.oy = L.r.oz = L.r.dx = L.r.dy = L.r.dz = 0.0f;
            L.cur.tri = -1; L.cur.t = 0.0f; L.cur.dist = 0.0f;
            if (act) {
                L.r.ox = u2f(PW(PF_OX, slot)); L.r.oy = u2f(PW(PF_OY, slot)); L.r.oz = u2f(PW(PF_OZ, slot));
                L.r.dx = u2f(PW(PF_DX, slot)); L.r.dy = u2f(PW(PF_DY, slot)); L.r.dz = u2f(PW(PF_DZ, slot));
                L.cur.tri = (int)PW(PF_CTRI, slot); L.cur.t = u2f(PW(PF_CT, slot)); L.cur.dist = u2f(PW(PF_CDIST, slot));
                L.state = ST_DONE;
                const uint4 g0 = gp[0], g1 = gp[1];
                const uint32_t pf = g1.y;
                q.sidx = g0.x; q.j = (int)g0.y;
                q.stream = (unsigned long long)g0.z | ((unsigned long long)g0.w << 32);
                q.saved_r = u2f(g1.x); q.saved_j = (int)(pf & 0xffffu) - 1;
                q.any_emit = ((pf >> 16) & 1u) != 0u; q.in_flight = ((pf >> 17) & 1u) != 0u;
            }
            path_regen_warp<COUNT>(sc, p, rd, fetch, st, q, gpm + (size_t)(gbase + slot) * (size_t)pm_stride, L, &cn, act);
            if (act) {
                PW(PF_OX, slot) = f2u(L.r.ox); PW(PF_OY, slot) = f2u(L.r.oy); PW(PF_OZ, slot) = f2u(L.r.oz);
                PW(PF_DX, slot) = f2u(L.r.dx); PW(PF_DY, slot) = f2u(L.r.dy); PW(PF_DZ, slot) = f2u(L.r.dz);
                PW(PF_DFX, slot) = f2u(L.dfx); PW(PF_DFY, slot) = f2u(L.dfy); PW(PF_DFZ, slot) = f2u(L.dfz);
                PW(PF_CHILD, slot) = L.child; PW(PF_TMIN, slot) = L.state == ST_LEAF ? (uint32_t)L.i : f2u(L.tmin); PW(PF_TMAX, slot) = f2u(L.tmax);
                PW(PF_CTRI, slot) = (uint32_t)L.cur.tri; PW(PF_CT, slot) = f2u(L.cur.t); PW(PF_CDIST, slot) = f2u(L.cur.dist);
                PW(PF_FLAGS, slot) = (L.rf & (7u | kRfTame | kRfUnsafe)) | ((uint32_t)L.state << 4) | ((uint32_t)L.sp << 8);
                gp[0] = make_uint4(q.sidx, (uint32_t)q.j, (uint32_t)q.stream, (uint32_t)(q.stream >> 32));
                gp[1] = make_uint4(f2u(q.saved_r), (uint32_t)(q.saved_j + 1) | (q.any_emit ? 0x10000u : 0u) | (q.in_flight ? 0x20000u : 0u), 0u, 0u);
            }
        }
        // ---- append every served ray to the queue of the step it needs next: lanes with the same destination find
        //      each other with one MATCH, the three fills grow by one packed REDUX
        {
            // ST_DONE 0 -> KR, ST_DESC 1 -> KT, ST_LEAF 2 -> KL, ST_RET 3 -> KL, ST_EXIT 4 -> none, ST_ENTER 5 -> KL, ST_SPH 6 -> KS
            const int nk = act ? (int)((0x3141102u >> (4 * L.state)) & 7u) : KNONE;
            const unsigned same = __match_any_sync(FULL, nk);
            const unsigned add = __reduce_add_sync(FULL, nk < KNONE ? (1u << (8 * nk)) : 0u);
            if (nk < KNONE) {
                const unsigned tail = ((q_head + q_cnt) >> (8 * nk)) & 0xffu;      // per byte: head < P, fill <= P, no carry
                sts8(a_queue + (uint32_t)(nk * P) + ((tail + (unsigned)__popc(same & lt_mask)) & (unsigned)(P - 1)), (uint32_t)(slot & (P - 1)));
            }
            q_cnt += add;
        }
        __syncwarp(FULL);
    }
#undef PW
    (void)gbase;
    if (COUNT && lane == 0) {
        for (int k = 0; k < 3; ++k) { atomicAdd(&ds->dbg[k], dbg_rounds[k]); atomicAdd(&ds->dbg[3 + k], dbg_sel[k]); }
        for (int k = 0; k < 8; ++k) { atomicAdd(&ds->dbg[8 + k], dbg_desc[k]); atomicAdd(&ds->dbg[16 + k], dbg_ret[k]); }
        for (int k = 0; k < 4; ++k) atomicAdd(&ds->dbg[24 + k], dbg_leaf[k]);
    }
    flush_stats(ds, st, cn, COUNT);
}

// sum the round's samples into the per-pixel running sums, in sample order (Lib.hs:88)
__global__ void __launch_bounds__(256) k_accumulate(RenderParams p, RoundInfo rd, float *__restrict__ accum, const DeviceStats *ds) {
    if (rd.pixel_list) rd.n_slots = (long long)ds->n_hit;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x; slot < rd.n_slots; slot += stride)
        accumulate_slot(p, rd, slot, accum);
}

template <bool COUNT>
__global__ void __launch_bounds__(128) k_raycast(SceneView sc, RenderParams p, float *__restrict__ accum, DeviceStats *ds, Tune tn) {
    Counters cn = {};
    PathStats st = {0, 0, 0};
    CastPolicy pol(p, accum, work_items(p), (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x, st);
    warp_loop<COUNT>(sc, pol, &cn, tn);
    flush_stats(ds, st, cn, COUNT);
}

// avg = (1 / fromIntegral sampleCount) *^ sum outcomes ; rgbFloatToPixelRGB avg   (Lib.hs:88-89)
__global__ void __launch_bounds__(256) k_tonemap(const float *__restrict__ accum, long long npix, float inv_spp,
                                                 uint8_t *__restrict__ rgb8) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    uint8_t o[3];
    tone_map(XMUL(inv_spp, accum[3 * i]), XMUL(inv_spp, accum[3 * i + 1]), XMUL(inv_spp, accum[3 * i + 2]), o);
    rgb8[3 * i] = o[0]; rgb8[3 * i + 1] = o[1]; rgb8[3 * i + 2] = o[2];
}

// scene upload: the leaf records (tight box, longest edge, range) from the triangles of every leaf, and the
// material-index check, on the device -- 10 M triangles take a millisecond here and half a second on one host core
__global__ void __launch_bounds__(256) k_leaf_records(float4 *tris, const uint2 *__restrict__ ranges, uint32_t n_leaves, float4 *__restrict__ leaves) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_leaves) return;
    const uint2 r = ranges[i];
    float4 b0, b1;
    make_leaf_record(tris, r.x, r.y, b0, b1);
    leaves[2 * (size_t)i] = b0; leaves[2 * (size_t)i + 1] = b1;
    if (r.y >= kLeafLong) tris[3 * (size_t)r.x + 2].w = __uint_as_float(r.y);
}
__global__ void __launch_bounds__(256) k_check_materials(const float4 *__restrict__ tris, uint32_t n_tris, uint32_t n_mats, uint32_t *bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tris && __float_as_uint(tris[3 * (size_t)i + 2].y) >= n_mats) atomicMin(bad, i);
}

// scene upload, subtree slabs (sqt_core.cuh "subtree slabs"): tight records bottom-up, one launch per tree level (the
// branches of a level only read records of deeper levels and of leaves), then every branch flags the children whose slab is worth a test
__global__ void __launch_bounds__(256) k_branch_tight(const uint32_t *__restrict__ level_nodes, uint32_t n, const float4 *__restrict__ nodes,
                                                      const float4 *__restrict__ leaves, float4 *tight) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = level_nodes[i];
    const float4 q = nodes[b];
    const uint32_t w[2] = {__float_as_uint(q.z), __float_as_uint(q.w)};
    TightRec c[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (w[k] & kLeaf) c[k] = tight_of_leaf(leaves, w[k] & kIdxMask);
        else { c[k].t0 = tight[2 * (size_t)(w[k] & kIdxMask)]; c[k].t1 = tight[2 * (size_t)(w[k] & kIdxMask) + 1]; }
    }
    const TightRec r = tight_union(c[0], c[1]);
    tight[2 * (size_t)b] = r.t0; tight[2 * (size_t)b + 1] = r.t1;
}
__global__ void __launch_bounds__(256) k_child_slabs(const float4 *__restrict__ nodes, uint32_t n_branches, const float4 *__restrict__ boxes,
                                                     const float4 *__restrict__ tight, const float4 *__restrict__ leaves, float s_max, float c_max,
                                                     float ratio_max, float ratio_max_leaf, float4 *__restrict__ slabs) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_branches) return;
    const float4 q = nodes[b];
    const uint32_t w[2] = {__float_as_uint(q.z), __float_as_uint(q.w)};
    const float4 p0 = boxes[2 * (size_t)b], p1 = boxes[2 * (size_t)b + 1];
    const int ax = (int)((w[0] >> kAxisShift) & 3u);
    float4 sl[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const uint32_t c = w[k] & kIdxMask;
        bool use;
        if (w[k] & kLeaf) {
            float4 c0, c1;
            clip_child_box(p0, p1, ax, q.x, q.y, k, c0, c1);
            use = make_slab(tight_of_leaf(leaves, c), c0, c1, s_max, c_max, ratio_max_leaf, sl[k]);
        } else {
            TightRec t; t.t0 = tight[2 * (size_t)c]; t.t1 = tight[2 * (size_t)c + 1];
            use = make_slab(t, boxes[2 * (size_t)c], boxes[2 * (size_t)c + 1], s_max, c_max, ratio_max, sl[k]);
        }
        sl[k].x = pack_slab_lo(sl[k].x, use ? (__float_as_uint(sl[k].w) & 3u) : kSlabNone);
    }
    slabs[b] = make_float4(sl[0].x, sl[0].y, sl[1].x, sl[1].y);
}
// ... and every reference to a Branch whose record holds a usable slab is flagged kTight
__global__ void __launch_bounds__(256) k_flag_slabs(float4 *nodes, uint32_t n_branches, const float4 *__restrict__ slabs) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_branches) return;
    float4 q = nodes[b];
    uint32_t w[2] = {__float_as_uint(q.z), __float_as_uint(q.w)};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (w[k] & kLeaf) continue;
        const float4 cs = slabs[w[k] & kIdxMask];
        if ((__float_as_uint(cs.x) & 3u) != kSlabNone || (__float_as_uint(cs.z) & 3u) != kSlabNone) w[k] |= kTight;
    }
    q.z = __uint_as_float(w[0]); q.w = __uint_as_float(w[1]);
    nodes[b] = q;
}

// Non-fused FP32 issue rate: 16 independent chains per lane, alternating FMUL / FADD (the op mix of the
// bit-exact intersection path, where FMA contraction is forbidden).
__global__ void __launch_bounds__(256) k_fp32_peak(float *out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) { x[i] = __fmul_rn(x[i], a); x[i + 1] = __fadd_rn(x[i + 1], b); }
#pragma unroll
        for (int i = 0; i < 16; i += 2) { x[i] = __fadd_rn(x[i], b); x[i + 1] = __fmul_rn(x[i + 1], a); }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456f) out[0] = s;      // keep the chains alive
}

// L2-resident 128-bit read bandwidth: every block sweeps the same `n4`-element buffer `reps` times.
__global__ void __launch_bounds__(256) k_l2_read(const float4 *__restrict__ buf, long long n4, int reps, float *out) {
    float s = 0.0f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r)
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(buf + i));
            s += v.x + v.y + v.z + v.w;
        }
    if (s == 123.456f) out[0] = s;
}
