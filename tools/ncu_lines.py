"""Per-source-line executed-instruction share from an ncu report (needs -lineinfo). usage: ncu_lines.py rep [topN]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file = ""; hdr = None; lines = []; tot = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; iex = hdr.index("Instructions Executed"); ith = hdr.index("Thread Instructions Executed"); ismp = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= ith: continue
    if r[0] and r[2] == "-":
        try: ex = int(r[iex]); th = int(r[ith]); sm = int(r[ismp] or 0)
        except ValueError: continue
        lines.append((ex, th, sm, cur_file, r[0], r[1].strip()[:110])); tot += ex
lines.sort(reverse=True)
tsm = sum(l[2] for l in lines)
print("total warp instructions %d, avg threads %.2f" % (tot, sum(l[1] for l in lines) / max(tot, 1)))
for ex, th, sm, f, ln, src in lines[:top]:
    print("%5.2f%% smp %5.2f%% thr %5.1f  %s:%s  %s" % (100 * ex / tot, 100 * sm / max(tsm, 1), th / max(ex, 1), f, ln, src))
# per-file share
agg = collections.Counter()
for ex, th, sm, f, ln, src in lines: agg[f] += ex
print({k: "%.1f%%" % (100 * v / tot) for k, v in agg.items()})
