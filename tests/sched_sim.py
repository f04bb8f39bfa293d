"""Offline study of warp schedulers for the persistent traversal loop.

Records, with the host build of the kernel logic (tests/emu), the exact sequence of unit steps every lane of a
warp would execute for data/scene.obj (R = regeneration, T = traversal step, L = triangle test), then replays 32
such lanes under different scheduling policies and reports issue slots per ray.  A policy only changes the ORDER
in which lanes are served, so any policy is valid; the cost model charges each executed phase iteration its
instruction count regardless of how many lanes are active (that is what SIMT does).
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "squigly-trace_b200"), os.path.join(ROOT, "tests")]
import pysqt
from common import emu_lib, Emu

COST = {"T": 85, "E": 50, "L": 62, "R": 650}     # instructions per step body (from SASS / ncu)
VOTE = 6                                   # ballot + popc + compare + branch per inner iteration
OUTER = 30                                 # one trip around the outer loop


def traces(n_lanes=32, pixels_per_lane=6, spp=8, depth=8, W=480, H=270):
    data = os.path.join(ROOT, "data")
    hs = pysqt.HostScene.load(os.path.join(data, "scene.obj"), data)
    cam = pysqt.camera_struct(pysqt.load_camera(os.path.join(data, "camera")))
    e = Emu(hs)
    E = emu_lib()
    E.emu_trace_lane.restype = C.c_longlong
    E.emu_trace_lane.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_longlong]
    p = pysqt.make_params(W, H, spp, max_depth=depth, seed=0)
    out = []
    base = (H // 2) * W + W // 3
    cap = 1 << 22
    buf = np.zeros(cap, np.uint8)
    for l in range(n_lanes):
        w0 = base + l * pixels_per_lane
        n = E.emu_trace_lane(e.h, C.byref(cam), C.byref(p), w0, w0 + pixels_per_lane, buf.ctypes.data_as(C.c_void_p), cap)
        out.append(bytes(buf[:n]).decode())
    return out


def simulate(tr, policy, **kw):
    """returns (cost, rays)"""
    pos = [0] * len(tr)
    n = len(tr)
    cost = 0
    rays = sum(t.count("R") for t in tr)

    def state(i):
        return tr[i][pos[i]] if pos[i] < len(tr[i]) else "X"

    def run(kind):
        nonlocal cost
        k = 0
        for i in range(n):
            if state(i) == kind:
                pos[i] += 1; k += 1
        cost += COST[kind] + VOTE
        return k

    def count(kind):
        return sum(1 for i in range(n) if state(i) == kind)

    if policy == "phased":
        a_leave, b_leave, c_min = kw["a_leave"], kw["b_leave"], kw["c_min"]
        while True:
            nd, nt, nl = count("R"), count("T"), count("L")
            if nd + nt + nl == 0: break
            cost += OUTER
            if nd and (nd >= c_min or nt + nl == 0): run("R")
            while count("T"):
                run("T")
                if count("T") <= a_leave: break
            while count("L"):
                run("L")
                if count("L") < b_leave: break
    elif policy == "majority":
        wt, wl, wd, c_min = kw["wt"], kw["wl"], kw["wd"], kw["c_min"]
        while True:
            nd, nt, nl = count("R"), count("T"), count("L")
            if nd + nt + nl == 0: break
            cost += 14
            sd = nd * wd if (nd >= c_min or nt + nl == 0) else 0
            st, sl = nt * wt, nl * wl
            if sl >= st and sl >= sd and nl: run("L")
            elif st >= sd and nt: run("T")
            else: run("R")
    elif policy == "thresh":
        # every trip: run each kind of step once if at least `thr` lanes want it (regen: c_min); if nothing
        # qualified, run the most wanted kind
        thr, c_min = kw["thr"], kw["c_min"]
        while True:
            cnt = {k: count(k) for k in "RTEL"}
            if sum(cnt.values()) == 0: break
            cost += 12
            ran = False
            for k in "TEL":
                if count(k) >= thr[k]: run(k); ran = True
            if count("R") >= c_min: run("R"); ran = True
            if not ran:
                cnt = {k: count(k) for k in "RTEL"}
                run(max(cnt, key=lambda k: cnt[k]))
    elif policy == "phased4":
        a_leave, b_leave, c_min = kw["a_leave"], kw["b_leave"], kw["c_min"]
        while True:
            nd, nt, ne, nl = count("R"), count("T"), count("E"), count("L")
            if nd + nt + ne + nl == 0: break
            cost += OUTER
            if nd and (nd >= c_min or nt + ne + nl == 0): run("R")
            while True:
                while count("T"):
                    run("T")
                    if count("T") <= a_leave: break
                if count("E"): run("E")
                if count("T") <= a_leave: break
            while count("L"):
                run("L")
                if count("L") < b_leave: break
    elif policy == "hyst":
        # stay in a phase while it has at least `stay` lanes, switch to the fullest one otherwise
        stay, c_min = kw["stay"], kw["c_min"]
        cur = "T"
        while True:
            nd, nt, nl = count("R"), count("T"), count("L")
            if nd + nt + nl == 0: break
            cnt = {"R": nd if (nd >= c_min or nt + nl == 0) else 0, "T": nt, "L": nl}
            if cnt[cur] < stay:
                cost += OUTER
                cur = max(cnt, key=lambda k: cnt[k])
            run(cur)
    return cost, rays


def study_phases():
    tr = traces()
    tot = sum(len(t) for t in tr)
    print("lanes", len(tr), "steps", tot, {k: sum(t.count(k) for t in tr) for k in "RTEL"})
    ideal = sum(COST[c] for t in tr for c in t) / 32
    rays = sum(t.count("R") for t in tr)
    print("ideal (perfect packing) slots/ray %.0f" % (ideal / rays))
    for a, b, c in [(0, 1, 4), (4, 1, 4), (8, 8, 4), (12, 12, 4), (16, 12, 4), (12, 16, 8)]:
        cst, r = simulate(tr, "phased4", a_leave=a, b_leave=b, c_min=c)
        print("phased4 a_leave=%2d b_leave=%2d c_min=%2d : %6.0f slots/ray  eff %.2f" % (a, b, c, cst / r, ideal / cst))
    for t, e, l, c in [(8, 8, 8, 4), (12, 8, 12, 4), (12, 12, 12, 8), (16, 8, 16, 8), (16, 16, 16, 8), (10, 6, 10, 6), (20, 10, 20, 8), (14, 10, 14, 10)]:
        cst, r = simulate(tr, "thresh", thr={"T": t, "E": e, "L": l}, c_min=c)
        print("thresh T>=%2d E>=%2d L>=%2d c_min=%2d : %6.0f slots/ray  eff %.2f" % (t, e, l, c, cst / r, ideal / cst))


def simulate_pool(tr_all, slots=64, burst_l=4, burst_t=2, overhead=86, pick="max"):
    """Warp-private pool: `slots` rays per warp; each round the warp takes up to 32 rays that are in the most
    populated state and runs up to burst_* steps of that kind for each before writing them back."""
    # tr_all: list of lane traces; we use `slots` of them as the warp's pool
    tr = tr_all[:slots]
    pos = [0] * len(tr)
    cost = 0
    rays = sum(t.count("R") for t in tr)

    def state(i):
        return tr[i][pos[i]] if pos[i] < len(tr[i]) else "X"
    util_num = util_den = 0
    while True:
        groups = {"R": [], "T": [], "E": [], "L": []}
        for i in range(len(tr)):
            s = state(i)
            if s != "X": groups[s].append(i)
        if not any(groups.values()): break
        kind = max(groups, key=lambda k: len(groups[k]) * (1 if k != "R" else 1))
        sel = groups[kind][:32]
        burst = {"L": burst_l, "T": burst_t, "R": 1, "E": 1}[kind]
        cost += overhead
        for b in range(burst):
            act = [i for i in sel if state(i) == kind]
            if not act: break
            for i in act: pos[i] += 1
            cost += COST[kind] + 4
            util_num += len(act); util_den += 32
    return cost, rays, util_num / max(util_den, 1)


def study_pool():
    tr128 = traces(n_lanes=128, pixels_per_lane=2, spp=6)
    ideal = sum(COST[c] for t in tr128 for c in t) / 32
    rays = sum(t.count("R") for t in tr128)
    print("pool study: ideal %.0f slots/ray" % (ideal / rays))
    for slots, bl, bt, ov in [(32, 1, 1, 0), (64, 4, 2, 86), (96, 4, 2, 86), (128, 4, 2, 86), (64, 2, 1, 86), (64, 8, 3, 86), (96, 8, 3, 86),
                              (64, 4, 2, 120), (96, 6, 2, 120), (128, 14, 4, 120)]:
        c, r, u = simulate_pool(tr128, slots, bl, bt, ov)
        print("pool slots=%3d burst L=%d T=%d overhead=%3d : %6.0f slots/ray  lane util %.2f" % (slots, bl, bt, ov, c / r, u))


def simulate_pool_refill(tr_all, slots=64, burst=8, overhead=86, refill_cost=30, c_min=16):
    """Pool with LANE-LEVEL REFILL: inside a burst a lane whose ray leaves the step kind writes it back and takes the next
    waiting ray of that kind (charged `refill_cost` instruction slots per iteration in which any lane refills)."""
    tr = tr_all[:slots]
    pos = [0] * len(tr)
    cost = 0
    rays = sum(t.count("R") for t in tr)

    def state(i):
        return tr[i][pos[i]] if pos[i] < len(tr[i]) else "X"
    num = den = 0
    while True:
        groups = {"R": [], "T": [], "E": [], "L": []}
        for i in range(len(tr)):
            s = state(i)
            if s != "X": groups[s].append(i)
        if not any(groups.values()): break
        cand = {k: len(v) for k, v in groups.items()}
        if cand["R"] < c_min and (cand["T"] or cand["E"] or cand["L"]): cand["R"] = 0
        kind = max(cand, key=lambda k: cand[k])
        waiting = list(groups[kind])
        lanes = [waiting.pop(0) for _ in range(min(32, len(waiting)))]
        cost += overhead
        for b in range(burst if kind in "TL" else 1):
            refilled = False
            for li, ray in enumerate(lanes):
                if ray is not None and state(ray) != kind:
                    lanes[li] = waiting.pop(0) if waiting else None
                    refilled = refilled or lanes[li] is not None
            act = [r for r in lanes if r is not None and state(r) == kind]
            if not act: break
            for r in act: pos[r] += 1
            cost += COST[kind] + 4 + (refill_cost if refilled else 0)
            num += len(act); den += 32
    return cost, rays, num / max(den, 1)


def study_refill():
    tr256 = traces(n_lanes=256, pixels_per_lane=1, spp=6)
    for slots in (64, 96, 128, 256):
        for burst in (4, 8, 16):
            c, r, u = simulate_pool_refill(tr256, slots, burst)
            print("pool+refill slots=%3d burst=%2d : %6.0f slots/ray  lane util %.2f" % (slots, burst, c / r, u))


def simulate_pool_pairs(tr_all, slots=64, burst_t=4, overhead=86, pair_overhead=60, pair_round=73, min_pairs=32, max_rounds=8, c_min=16):
    """Pool in which the triangle phase is expanded to (ray, triangle) PAIRS: the remaining triangles of all rays that
    wait in a leaf are laid out consecutively and tested 32 at a time, one pair per lane, whichever ray they belong to."""
    tr = tr_all[:slots]
    pos = [0] * len(tr)
    cost = 0
    rays = sum(t.count("R") for t in tr)

    def state(i):
        return tr[i][pos[i]] if pos[i] < len(tr[i]) else "X"

    def run_len(i):
        n = 0
        while pos[i] + n < len(tr[i]) and tr[i][pos[i] + n] == "L": n += 1
        return n
    num = den = 0
    while True:
        groups = {"R": [], "T": [], "E": [], "L": []}
        for i in range(len(tr)):
            s = state(i)
            if s != "X": groups[s].append(i)
        if not any(groups.values()): break
        pairs = sum(run_len(i) for i in groups["L"])
        cand = {k: len(groups[k]) for k in "RTE"}
        if cand["R"] < c_min and (cand["T"] or cand["E"] or pairs): cand["R"] = 0
        if pairs >= min_pairs or (pairs and not any(cand.values())):
            rounds = max(1, min(max_rounds, pairs // 32))
            budget = rounds * 32
            used = 0
            for i in groups["L"]:
                n = min(run_len(i), budget - used)
                pos[i] += n; used += n
                if used == budget: break
            cost += pair_overhead + pair_round * ((used + 31) // 32)
            num += used; den += 32 * ((used + 31) // 32)
            continue
        kind = max(cand, key=lambda k: cand[k])
        sel = groups[kind][:32]
        cost += overhead
        for b in range(burst_t if kind == "T" else 1):
            act = [i for i in sel if state(i) == kind]
            if not act: break
            for i in act: pos[i] += 1
            cost += COST[kind] + 4
            num += len(act); den += 32
    return cost, rays, num / max(den, 1)


def study_pairs():
    tr256 = traces(n_lanes=256, pixels_per_lane=1, spp=6)
    for slots in (64, 96):
        c, r, u = simulate_pool(tr256, slots, 8, 4, 86)
        print("pool        slots=%3d                : %6.0f slots/ray  lane util %.2f" % (slots, c / r, u))
        for mp in (32, 64, 96):
            for bt in (2, 4):
                c, r, u = simulate_pool_pairs(tr256, slots, bt, min_pairs=mp)
                print("pool+pairs  slots=%3d min_pairs=%3d bt=%d : %6.0f slots/ray  lane util %.2f" % (slots, mp, bt, c / r, u))


def simulate_warp_loop_pairs(tr_all, lanes=32, a_leave=12, c_min=12, pairs=True, pair_round=125, pair_overhead=30, vote=6, outer=30, b_leave=1):
    """One ray per lane (warp_loop) with the WARP-COOPERATIVE triangle phase: all triangle tests of the lanes that wait in a
    leaf are executed 32 at a time, whichever lane owns the ray (leaf_pairs in csrc/sqt_backend.cu); pairs=False is the old
    one-test-per-lane-and-step phase.  pair_round = instructions per 32 tests (measured: ~95 search/shuffle/hand-over + ~97
    Moller-Trumbore, but only ~60 % of the latter when few lanes pass the guards)."""
    tr = tr_all[:lanes]
    pos = [0] * len(tr)
    rays = sum(t.count("R") for t in tr)
    cost = 0

    def state(i):
        return tr[i][pos[i]] if pos[i] < len(tr[i]) else "X"

    def run_len(i):
        n = 0
        while pos[i] + n < len(tr[i]) and tr[i][pos[i] + n] == "L": n += 1
        return n

    def lanes_in(k):
        return [i for i in range(len(tr)) if state(i) == k]
    while any(state(i) != "X" for i in range(len(tr))):
        cost += outer
        done, busy = lanes_in("R"), [i for i in range(len(tr)) if state(i) in "TEL"]
        if done and (len(done) >= c_min or not busy):
            for i in done: pos[i] += 1
            cost += COST["R"]
        first = True
        while True:
            t = lanes_in("T")
            if not t or (not first and len(t) <= a_leave and (lanes_in("E") or lanes_in("L"))): break
            first = False
            for i in t: pos[i] += 1
            cost += COST["T"] + vote
        e = lanes_in("E")
        if e:
            for i in e: pos[i] += 1
            cost += COST["E"] + vote
        if pairs:
            ls = lanes_in("L")
            n = sum(run_len(i) for i in ls)
            if n:
                for i in ls: pos[i] += run_len(i)
                cost += pair_overhead + pair_round * ((n + 31) // 32)
        else:
            while True:
                ls = lanes_in("L")
                if not ls or (len(ls) <= b_leave and lanes_in("T")): break
                for i in ls: pos[i] += 1
                cost += COST["L"] + vote
    return cost, rays


def study_warp_loop_pairs():
    tr = traces(n_lanes=32, pixels_per_lane=1, spp=6)
    c, r = simulate_warp_loop_pairs(tr, pairs=False, a_leave=8, c_min=8)
    print("warp_loop, one test per lane and step       : %6.0f slots/ray" % (c / r))
    for a in (4, 8, 12, 16):
        for pr in (90, 125, 190):
            c, r = simulate_warp_loop_pairs(tr, a_leave=a, c_min=8, pair_round=pr)
            print("warp_loop + leaf_pairs a_leave=%2d pair_round=%3d : %6.0f slots/ray" % (a, pr, c / r))


STUDIES = {"phases": study_phases, "pool": study_pool, "refill": study_refill, "pairs": study_pairs, "warp_pairs": study_warp_loop_pairs}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(STUDIES)):
        print("==== " + name)
        STUDIES[name]()
