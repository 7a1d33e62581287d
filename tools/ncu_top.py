"""Top source lines of an ncu source-page CSV by executed warp instructions (uses the CSV's own source text).
Usage: ncu_top.py src.csv nframes [ntop]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
N = float(sys.argv[2]); ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 60
hdr = None
per = collections.Counter(); thr = collections.Counter(); smp = collections.Counter(); txt = {}
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); ns = hdr.index("# Samples"); continue
    if r and r[0] == "File Path":
        cur = r[1]; continue
    if hdr is None or len(r) < 10: continue
    try: ln = int(r[0])
    except ValueError: continue
    key = (cur.split("/")[-1], ln)
    try:
        per[key] += int(r[ie]); thr[key] += int(r[te]); smp[key] += int(r[ns]); txt[key] = r[1]
    except ValueError: pass
tot = sum(per.values()); ts = sum(smp.values())
print("total %.0f warp-instr/frame, %d samples" % (tot / N, ts))
byfile = collections.Counter()
for k, n in per.items(): byfile[k[0]] += n
for f, n in byfile.most_common(): print("  file %-28s %8.0f /frame" % (f, n / N))
for k, n in per.most_common(ntop):
    print("%-16s %5d %7.0f %5.1f thr %5.2f%%smp  %s" % (k[0][:16], k[1], n / N, thr[k] / max(n, 1), 100.0 * smp[k] / ts, txt[k].strip()[:100]))
