"""Developer check: statuses of one synthetic stream through the batch API for several requests/formats."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np
import __graft_entry__ as g
from refbind import Oracle
from bitstream_writer import make_stream
from util import frame_offsets
eng = g.load_engine()
o = Oracle()
dec = eng.BatchDecoder()
acmod, lfe = 3, 1
es, fb = make_stream(4100 + acmod * 2 + lfe, acmod, lfe, 3, o.bit_allocate, frmsizecod=30)
off = frame_offsets(es, o)
first = np.array([0, len(off)], np.uint32)
for flags in (2, 2 | 32, 10 | 32, 2 | 16, 0x100 | 16 | 32):
    for fmt in (0, 1, 2, 3):
        for bias in (0.0, 384.0):
            out = dec.decode_host(es, off, first, flags, 1.0, bias, out_fmt=fmt)
            dump = o.decode_dump(es, req_flags=flags & 0xff) if not flags & 0x100 else None
            print(flags, fmt, bias, out["status"], out["flags"], [f["status"] for f in dump] if dump else None)
