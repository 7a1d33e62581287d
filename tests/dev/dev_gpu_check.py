"""Developer smoke check run on the GPU box: GPU decode vs the reference build (oracle/_ref)."""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
spec = importlib.util.spec_from_file_location("ac3_b200", os.path.join(ROOT, "ac-3-acm-codec_b200", "__init__.py"))
eng = importlib.util.module_from_spec(spec)
spec.loader.exec_module(eng)
from refbind import RefA52, RefAc3Enc, A52_STEREO, A52_3F2R, A52_LFE, A52_ADJUST_LEVEL, A52_MONO, A52_DOLBY, A52_3F, A52_2F2R  # noqa
from synth import synth_pcm

LIBA52 = [0, -1, -2, 3, -3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 14, 16]


def relrms(a, b):
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.sqrt((d * d).mean()) / max(np.sqrt((b.astype(np.float64) ** 2).mean()), 1e-30))


def main():
    ref, enc = RefA52(), RefAc3Enc()
    dec = eng.BatchDecoder(0)
    ok = True
    for (nch, br, rate, flags) in [(6, 448000, 48000, A52_3F2R | A52_LFE), (6, 448000, 48000, A52_STEREO | A52_ADJUST_LEVEL),
                                   (2, 192000, 48000, A52_STEREO | A52_ADJUST_LEVEL), (1, 96000, 48000, A52_STEREO),
                                   (5, 384000, 44100, A52_DOLBY | A52_ADJUST_LEVEL), (3, 192000, 32000, A52_MONO),
                                   (4, 256000, 48000, A52_3F), (6, 640000, 48000, A52_2F2R | A52_LFE)]:
        nfr = 8
        pcm = synth_pcm(2, nch, nch, 1536 * nfr, rate)
        fb, es = enc.encode_stream(pcm, rate, br)
        off = eng.index_frames(es)
        assert len(off) == nfr, (len(off), nfr)
        first = np.array([0, nfr], np.uint32)
        out = dec.decode_host(es, off, first, flags, 1.0, 0.0, want_debug=True)
        nf, rpcm = ref.decode_stream(es, flags, 1.0, 0.0)
        nout = rpcm.shape[1]
        g = out["pcm"][:, : 6 * nout * 256].reshape(nfr * 6, nout, 256)
        e_pcm = relrms(g, rpcm)
        msg = "nch=%d br=%d rate=%d flags=%d status=%s relrms=%.3g maxabs=%.3g" % (
            nch, br, rate, flags, out["status"].tolist(), e_pcm, np.abs(g - rpcm).max())
        # exp / bap / coef against the reference dump (decoded to 5.1 so that planes are per channel)
        fr = ref.decode_dump(es, req_flags=flags)
        bad_e = bad_b = 0
        for f in range(nfr):
            for b in range(6):
                blk = fr[f]["blocks"][b]
                info = blk["info"]
                nf_ch = [2, 1, 2, 3, 3, 4, 4, 5][info[9]]
                for ch in range(nf_ch):
                    end = info[ch]
                    bad_e += int((out["exp"][f, b, ch, :end] != blk["exp"][ch, :end]).sum())
                    gb = np.array(LIBA52)[out["bap"][f, b, ch, :end]]
                    bad_b += int((gb != blk["bap"][ch, :end]).sum())
                if info[10]:
                    bad_e += int((out["exp"][f, b, 5, :7] != blk["exp"][5, :7]).sum())
                    bad_b += int((np.array(LIBA52)[out["bap"][f, b, 5, :7]] != blk["bap"][5, :7]).sum())
        lf = fr[-1]["blocks"][-1]["info"][8]
        msg += " exp_bad=%d bap_bad=%d lfsr %d/%d" % (bad_e, bad_b, out["info"][-1, -1, 8], lf)
        print(msg, flush=True)
        if e_pcm > 1e-5 or bad_e or bad_b or out["info"][-1, -1, 8] != lf:
            ok = False
    # coefficient parity in the 5.1 -> 5.1 case (no mixing in either decoder)
    pcm = synth_pcm(2, 7, 6, 1536 * 4)
    fb, es = enc.encode_stream(pcm, 48000, 448000)
    off = eng.index_frames(es)
    out = dec.decode_host(es, off, np.array([0, 4], np.uint32), A52_3F2R | A52_LFE, 1.0, 0.0, want_debug=True)
    fr = ref.decode_dump(es, req_flags=A52_3F2R | A52_LFE)
    bad = 0
    for f in range(4):
        for b in range(6):
            for (plane, kind, co) in fr[f]["blocks"][b]["coeffs"]:
                mine = out["coef"][f, b, 5 if plane == 0 else plane - 1]
                bad += int((mine.view(np.uint32) != co.view(np.uint32)).sum())
    print("coef bit mismatches (5.1):", bad, flush=True)
    ok = ok and bad == 0

    # first throughput probe: many copies of one stream
    nfr, ns = 32, 2048
    pcm = synth_pcm(2, 1, 6, 1536 * nfr)
    fb, es1 = enc.encode_stream(pcm, 48000, 448000)
    es = np.tile(es1, ns)
    off = np.arange(ns * nfr, dtype=np.uint64) * fb
    first = (np.arange(ns + 1) * nfr).astype(np.uint32)
    for fmt in (eng.PCM_F32_INTERLEAVED, eng.PCM_F32_INTERLEAVED, eng.PCM_F32_PLANAR):
        dec.kernel_ms()
        t = time.time()
        out = dec.decode_host(es, off, first, A52_STEREO | A52_ADJUST_LEVEL, 1.0, 0.0, out_fmt=fmt)
        dt = time.time() - t
        ms, n = dec.kernel_ms()
        secs = ns * nfr * 1536 / 48000.0
        print("batch %d streams x %d frames: host call %.3f s, kernel %.3f ms (%d launches) -> %.0f audio-s/s kernel-only; status ok=%s"
              % (ns, nfr, dt, ms, n, secs / (ms / 1e3), bool((out["status"] == 0).all())), flush=True)
    print("DEV CHECK", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
