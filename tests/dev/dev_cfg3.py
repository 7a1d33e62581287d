"""Developer check: throughput on config-3-like streams (5.1 640 kb/s, block switching, coupling, dynrng, deltba
from tests/bitstream_writer.py) next to the stationary config-2 corpus, device-resident, stereo float out."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch
import __graft_entry__ as g
from refbind import Oracle
from bitstream_writer import make_stream
eng = g.load_engine(); o = Oracle(); dec = eng.BatchDecoder(0)
base = [make_stream(31337 + k, 7, 1, 24, o.bit_allocate, frmsizecod=36,
                    features=dict(blksw=0.5, cpl=0.9, dynrng=0.5, deltba=0.1))[0] for k in range(4)]
fb, nfr, ns = 2560, 24, 1776
es = np.concatenate([base[s % 4] for s in range(ns)])
d_es = torch.from_numpy(np.concatenate([es, np.zeros(64, np.uint8)])).cuda()
nframes = ns * nfr
off = torch.arange(nframes + 1, dtype=torch.int64, device="cuda") * fb
first = (torch.arange(ns + 1, dtype=torch.int32, device="cuda") * nfr)
for flags, nout, name in ((2 | 32, 2, "stereo"), (7 | 16, 6, "5.1")):
    pcm = torch.empty(nframes * 1536 * nout, dtype=torch.float32, device="cuda")
    status = torch.zeros(nframes, dtype=torch.int32, device="cuda")
    dec.set_max_frame_bytes(fb); dec.set_max_stream_frames(nfr)
    def step():
        dec.decode_device(d_es.data_ptr(), nframes * fb, off.data_ptr(), nframes, first.data_ptr(), ns, flags,
                          pcm.data_ptr(), status_ptr=status.data_ptr(), out_fmt=eng.PCM_F32_INTERLEAVED)
    for _ in range(3): step()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    assert int((status != 0).sum()) == 0
    print("config-3-like 640k transient corpus ->", name, "%.0f audio-s/s" % (nframes * 0.032 / dt), flush=True)
