"""Developer diagnostic (GPU box): per-block PCM error of the GPU decode vs the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
from refbind import *
from bitstream_writer import make_stream
from util import frame_offsets
eng = ge.load_engine()
ora = Oracle()
dec = eng.BatchDecoder(0)
cases = [(7, 1, A52_STEREO | A52_ADJUST_LEVEL, 0, 36), (7, 0, A52_MONO, 0, 30), (5, 0, A52_STEREO, 0, 33), (6, 0, A52_3F, 0, 30)]
for (acmod, lfe, flags, fscod, cod) in cases:
    es, fb = make_stream(2000 + acmod * 16 + (flags & 15), acmod, lfe, 4, ora.bit_allocate, fscod=fscod, frmsizecod=cod)
    nf, want = ora.decode_stream(es, flags, 1.0, 0.0)
    off = frame_offsets(es, ora)
    out = dec.decode_host(es, off, np.array([0, 4], np.uint32), flags, want_debug=True)
    nout = want.shape[1]
    got = out["pcm"][:, :6 * nout * 256].reshape(24, nout, 256)
    print("case", acmod, lfe, flags)
    for k in range(24):
        d = got[k].astype(np.float64) - want[k]
        e = np.sqrt((d * d).mean()) / max(np.sqrt((want[k].astype(np.float64) ** 2).mean()), 1e-30)
        x = out["info"][k // 6, k % 6, 12]
        print("  f%d b%d err=%.2e blksw=%s uniform=%d clev0=%d slev0=%d per-ch err=%s" % (
            k // 6, k % 6, e, bin(x & 31), (x >> 8) & 1, (x >> 9) & 1, (x >> 10) & 1,
            ["%.1e" % np.abs(d[c]).max() for c in range(nout)]))
