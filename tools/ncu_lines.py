"""Summarise an ncu source-page CSV: executed warp instructions and stall samples per source line of ONE file
(the kernel source: instructions inlined from headers are attributed to their call sites in that file's table; the
other files' tables repeat them), grouped into ranges of the kernel source.
Usage: ncu_lines.py src.csv file.cu [nframes]"""
import csv, sys, collections, os
rows = list(csv.reader(open(sys.argv[1])))
want = os.path.basename(sys.argv[2])
hdr = None; cur = None
per_line = collections.Counter(); samp = collections.Counter(); thr = collections.Counter()
for r in rows:
    if r and r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if r and r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); ns = hdr.index("# Samples"); te = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) < 10 or cur != want: continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    try:
        per_line[ln] += int(r[ie]); samp[ln] += int(r[ns]); thr[ln] += int(r[te])
    except ValueError:
        pass
src = open(sys.argv[2]).read().split("\n")
nfr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
tot = sum(per_line.values()); ts = sum(samp.values())
print("total warp-instr %.3g (%.0f per frame), samples %d" % (tot, tot / nfr, ts))
# group by marker comments
marks = []
for i, l in enumerate(src, 1):
    if "=================" in l or l.startswith("__device__") or l.startswith("__global__") or l.startswith("template <"):
        marks.append((i, l.strip()[:70]))
marks.append((len(src) + 1, "end"))
for (a, name), (b, _) in zip(marks, marks[1:]):
    n = sum(per_line[k] for k in range(a, b)); s = sum(samp[k] for k in range(a, b)); t = sum(thr[k] for k in range(a, b))
    if n:
        print("%5d-%5d %6.2f%% instr (%7.0f/frame, %4.1f thr/instr) %6.2f%% samples  %s" % (a, b - 1, 100.0 * n / tot, n / nfr, t / max(n, 1), 100.0 * s / ts, name))
