"""Throughput of the GPU frame indexer (a52_batch_index_device, row f2): the config-2 batch as 4096 elementary streams
of 313 frames, clean and with every 50th stream damaged (a run of garbage bytes the walk has to slide over)."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np, torch
import __graft_entry__ as g
import bench
eng = g.load_engine(); dev = torch.device("cuda", 0)
uniq = bench.unique_frames_gpu(eng, dev)
S, F = 4096, 313
es = torch.zeros(S * F * 1792 + 64, dtype=torch.uint8, device=dev)
es[:S * F * 1792].copy_(bench.tile_corpus(uniq, S, F).reshape(-1))
soff = (np.arange(S + 1, dtype=np.uint64) * (F * 1792))
dec = eng.BatchDecoder(0)
res = {}
for name in ("clean", "damaged"):
    if name == "damaged":
        v = es[:S * F * 1792].view(S, F * 1792)
        v[::50, 100 * 1792 + 5: 100 * 1792 + 5 + 3000] = 0x0b                 # ~1.7 frames of junk in every 50th stream
    off = torch.zeros(S * F + 1, dtype=torch.int64, device=dev); first = torch.zeros(S + 1, dtype=torch.int32, device=dev)
    n = dec.index_device(es.data_ptr(), soff, off.data_ptr(), S * F, first.data_ptr())
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10):
        n = dec.index_device(es.data_ptr(), soff, off.data_ptr(), S * F, first.data_ptr())
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
    res[name] = {"frames_found": int(n), "ms": dt * 1e3, "GB_per_s_of_bitstream": S * F * 1792 / dt / 1e9,
                 "frames_per_s": n / dt}
    print(name, res[name], flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "index_bench.json"), "w"), indent=1)
