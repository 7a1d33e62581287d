"""Per source line: where the samples of one stall reason sit.  Usage: ncu_stalls.py src.csv reason [ntop] (reason e.g. stall_no_inst)"""
import csv, sys, collections
rows = csv.reader(open(sys.argv[1])); reason = sys.argv[2]; ntop = int(sys.argv[3]) if len(sys.argv) > 3 else 25
cur = None; idx = None; acc = collections.Counter(); txt = {}; tot = collections.Counter()
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": idx = r.index(reason); ns = r.index("# Samples"); continue
    if idx is None or len(r) <= idx or cur != "a52_decode.cu" or not r[0].isdigit(): continue
    try:
        acc[int(r[0])] += int(r[idx]); tot[int(r[0])] += int(r[ns]); txt[int(r[0])] = r[1].strip()[:100]
    except ValueError: pass
S = sum(acc.values()); T = sum(tot.values())
print("%s: %d of %d samples (%.1f%%)" % (reason, S, T, 100.0 * S / max(T, 1)))
for ln, n in acc.most_common(ntop): print("%5d %6.2f%%  %s" % (ln, 100.0 * n / max(S, 1), txt[ln]))
