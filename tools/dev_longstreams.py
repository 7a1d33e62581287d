"""Few long streams: {1, 16, 256} streams x 20 000 frames (10.7 min of audio each) of the config-2 corpus, slices
chained by the carry record against frame-independent slices (scan pass + prefix sum + one frame of look-back).
Device resident, stereo float out.  Prints one JSON line per case; the two modes must produce the same bits."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import __graft_entry__ as g
import bench
eng = g.load_engine()
dev = torch.device("cuda", 0)
uniq = bench.unique_frames_gpu(eng, dev)                       # [256][313][1792]
F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
res = []
for S in (1, 16, 256):
    idx = (torch.arange(F, device=dev)[None, :] + 11 * torch.arange(S, device=dev)[:, None]) % 313
    es = torch.zeros(S * F * 1792 + 64, dtype=torch.uint8, device=dev)
    es[:S * F * 1792].copy_(uniq[torch.arange(S, device=dev)[:, None] % 256, idx].reshape(-1))
    n = S * F
    off = torch.arange(n + 1, dtype=torch.int64, device=dev) * 1792
    first = (torch.arange(S + 1, dtype=torch.int64, device=dev) * F).to(torch.int32)
    pcm = {}
    row = {"streams": S, "frames_per_stream": F}
    for mode, name in ((eng.SLICES_CHAINED, "chained"), (eng.SLICES_INDEPENDENT, "independent")):
        dec = eng.BatchDecoder(0); dec.set_slice_mode(mode)
        dec.set_max_frame_bytes(1792); dec.set_max_stream_frames(F)
        out = torch.empty(n * 3072, dtype=torch.float32, device=dev); st = torch.zeros(n, dtype=torch.int32, device=dev)
        def step():
            dec.decode_device(es.data_ptr(), n * 1792, off.data_ptr(), n, first.data_ptr(), S, 2 | 32, out.data_ptr(),
                              status_ptr=st.data_ptr(), out_fmt=eng.PCM_F32_INTERLEAVED)
        step(); torch.cuda.synchronize()
        reps = 1 if (name == "chained" and S == 1) else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        assert int((st != 0).sum()) == 0
        row[name + "_audio_s_per_s"] = n * 0.032 / (ms / 1e3); row[name + "_ms"] = ms
        pcm[name] = out
        dec.close()
    row["bit_identical"] = bool(torch.equal(pcm["chained"].view(torch.int32), pcm["independent"].view(torch.int32)))
    row["speedup"] = row["independent_audio_s_per_s"] / row["chained_audio_s_per_s"]
    print(json.dumps(row), flush=True); res.append(row)
    del es, pcm, out
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "longstreams.json"), "w"), indent=1)
