"""BASELINE.json configs[0] timed: one stereo 48 kHz stream encoded at 192 kb/s, decoded by the a52dec command line
(-o wav).  Three decoders, wall clock of the whole process (CUDA start-up included), output to a file on /tmp:
  a52dec_ref       the unmodified reference CLI on one host core                         (oracle/_ref/a52dec_ref)
  a52dec (drop-in) the SAME unmodified CLI linked against liba52_b200.so: one frame per a52_frame / a52_block
                   round trip, as the liba52 API hands them over                          (oracle/_ref/a52dec_b200)
  a52dec_b200      this repository's batch CLI: the whole file in one a52_batch_decode     (ac-3-acm-codec_b200/a52dec_b200)
for a 60 s stream (the config) and a 2 h one (a film's worth: 225 000 frames)."""
import json, os, subprocess, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from refbind import RefAc3Enc
from synth import synth_pcm

enc = RefAc3Enc()
pcm = synth_pcm(1, 0, 2, 48000 * 60)                       # C1: 60 s stereo, 3 sines + noise per channel
fb, es = enc.encode_stream(pcm, 48000, 192000)
es = np.asarray(es, np.uint8)
files = {"60s": es, "2h": np.tile(es, 120)}
exes = {"a52dec_ref (1 host core)": os.path.join(ROOT, "oracle", "_ref", "a52dec_ref"),
        "a52dec drop-in (liba52_b200.so, frame by frame)": os.path.join(ROOT, "oracle", "_ref", "a52dec_b200"),
        "a52dec_b200 (batch CLI)": os.path.join(ROOT, "ac-3-acm-codec_b200", "a52dec_b200")}
res = {"frame_bytes": int(fb), "what": "wall seconds of the whole process, -o wav to a file; audio seconds / wall in brackets"}
for tag, data in files.items():
    path = "/tmp/c1_%s.ac3" % tag
    data.tofile(path)
    audio = len(data) / fb * 0.032
    for name, exe in exes.items():
        if tag == "2h" and "drop-in" in name:
            continue                                       # minutes of single-frame round trips: the 60 s figure scales
        best = None
        for rep in range(2):
            t = time.perf_counter()
            with open("/tmp/c1_out.wav", "wb") as f:
                r = subprocess.run([exe, "-o", "wav", path], stdout=f, stderr=subprocess.PIPE)
            dt = time.perf_counter() - t
            assert r.returncode == 0, (name, r.stderr[-300:])
            best = dt if best is None else min(best, dt)
        size = os.path.getsize("/tmp/c1_out.wav")
        res["%s | %s" % (tag, name)] = {"wall_s": best, "audio_s_per_s": audio / best, "wav_bytes": size}
        print(tag, name, "%.3f s (%.0f audio-s/s)" % (best, audio / best), flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "config1_cli.json"), "w"), indent=1)
