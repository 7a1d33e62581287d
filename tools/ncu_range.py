"""Per-source-line executed warp instructions for a line range. Usage: ncu_range.py src.csv file.cu nframes lo hi"""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; per=collections.Counter(); thr=collections.Counter(); smp=collections.Counter()
for r in rows:
    if r and r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)<10: continue
    try: ln=int(r[0])
    except: continue
    ie=hdr.index("Instructions Executed"); te=hdr.index("Thread Instructions Executed"); ns=hdr.index("# Samples")
    try: per[ln]+=int(r[ie]); thr[ln]+=int(r[te]); smp[ln]+=int(r[ns])
    except: pass
src=open(sys.argv[2]).read().split('\n'); nf=float(sys.argv[3])
ts=sum(smp.values())
for ln in range(int(sys.argv[4]),int(sys.argv[5])+1):
    if per[ln]: print("%5d %7.0f %5.1f %5.2f%%  %s"%(ln, per[ln]/nf, thr[ln]/max(per[ln],1), 100.0*smp[ln]/ts, src[ln-1][:105]))
