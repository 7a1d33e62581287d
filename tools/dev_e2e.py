"""Developer check: end-to-end (host pointer) decode time against the chunk plan of the host pipeline.
usage: dev_e2e.py chunk_streams [chunk_streams ...]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
import __graft_entry__ as g
import bench
eng = g.load_engine()
S, F = 4096, 313
corpus = bench.build_corpus(S, F)
nframes = S * F
es_h = torch.from_numpy(corpus.reshape(-1).copy()).pin_memory()
pcm_h = torch.empty(nframes * 1536 * 2, dtype=torch.float32).pin_memory()
status_h = torch.zeros(nframes, dtype=torch.int32).pin_memory()
off_h = np.arange(nframes, dtype=np.uint64) * bench.FRAME_BYTES
first_h = (np.arange(S + 1, dtype=np.uint32) * F).astype(np.uint32)
for arg in sys.argv[1:]:
    cs, cc = arg.split(":")
    os.environ["A52_B200_HOST_CHUNK_STREAMS"] = cs
    os.environ["A52_B200_HOST_CONCURRENCY"] = cc
    dec = eng.BatchDecoder(0)
    dec.set_max_frame_bytes(1792)
    def step():
        dec.decode_host_into(es_h.data_ptr(), es_h.numel(), off_h, first_h, bench.REQ_FLAGS, pcm_h.data_ptr(),
                             status_h.data_ptr(), out_fmt=eng.PCM_F32_INTERLEAVED)
    step()
    dec.kernel_ms()
    t0 = time.perf_counter()
    for _ in range(2):
        step()
    wall = (time.perf_counter() - t0) / 2
    ms, n = dec.kernel_ms()
    print("chunk_streams", cs, "concurrency", cc, "ms_per_step %.1f" % (1e3 * wall), "audio-s/s %.0f" % (nframes * 0.032 / wall),
          "kernel ms avg %.2f over %d launches" % (ms, n), flush=True)
    dec.close()
