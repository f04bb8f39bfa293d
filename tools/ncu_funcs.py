"""Executed-instruction share per function of the traversal core from an ncu report (source page; needs -lineinfo).
usage: ncu_funcs.py rep"""
import csv, subprocess, sys, collections, re, os
rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# function line ranges from the sources (top-level SQT_HD / __device__ / __global__ definitions and structs)
ranges = {}
for f in ("sqt_core.cuh", "sqt_paths.cuh", "sqt_kernels.cuh", "sqt_backend.cu"):
    src = open(os.path.join(ROOT, "squigly-trace_b200", "csrc", f)).read().splitlines()
    starts = []
    for i, l in enumerate(src, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:SQT_HD|__device__|__global__|static|inline).*?\b([A-Za-z_0-9]+)\s*\(", l)
        if m and not l.startswith(" "): starts.append((i, m.group(1)))
        m = re.match(r"^struct\s+([A-Za-z_0-9]+)\s*\{", l)                # accessor structs (LaneRay, PoolRay): their methods are one-liners inside
        if m: starts.append((i, "struct " + m.group(1)))
    for (a, n), nxt in zip(starts, starts[1:] + [(len(src) + 1, "")]): ranges.setdefault(f, []).append((a, nxt[0] - 1, n))
def func(f, ln):
    for a, b, n in ranges.get(f, []):
        if a <= ln <= b: return n
    return f
cur_file = ""; hdr = None; agg = collections.Counter(); thr = collections.Counter(); tot = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; iex = hdr.index("Instructions Executed"); ith = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) <= ith: continue
    if r[0] and r[2] == "-":
        try: ex = int(r[iex]); th = int(r[ith]); ln = int(r[0])
        except ValueError: continue
        k = func(cur_file, ln); agg[k] += ex; thr[k] += th; tot += ex
print("total warp instructions %d" % tot)
for k, v in agg.most_common(25): print("%6.2f%%  lanes %5.1f  %s" % (100 * v / tot, thr[k] / max(v, 1), k))
