#!/usr/bin/env python3
"""Generate the committed golden vectors from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference to have been compiled by
`make -C oracle ref`):   python tests/golden/make_golden.py
Outputs (committed):
  tests/golden/decode_vectors.npz  - small streams + reference PCM / exponents / baps / lfsr
  tests/golden/c2_fixture.npz      - 4 reference-encoded 5.1 448 kb/s streams of 64 frames
                                     (the corpus bench.py tiles) + reference stereo PCM digests
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refbind import (RefA52, RefAc3Enc, Oracle, A52_STEREO, A52_3F2R, A52_LFE, A52_ADJUST_LEVEL, A52_MONO,  # noqa
                     A52_DOLBY, A52_CHANNEL, A52_2F2R, A52_3F, A52_2F1R, A52_CHANNEL1)
from synth import synth_pcm  # noqa
from bitstream_writer import make_stream  # noqa


def main():
    ref, enc, ora = RefA52(), RefAc3Enc(), Oracle()
    out = {}
    cases = []
    # reference-encoded streams (configs 1, 2 and friends)
    for name, nch, br, rate, flags, bias in [
            ("enc51_stereo", 6, 448000, 48000, A52_STEREO | A52_ADJUST_LEVEL, 0.0),
            ("enc51_51", 6, 448000, 48000, A52_3F2R | A52_LFE, 0.0),
            ("enc20_stereo_bias", 2, 192000, 48000, A52_STEREO | A52_ADJUST_LEVEL, 384.0),
            ("enc10_stereo", 1, 96000, 48000, A52_STEREO, 0.0),
            ("enc50_dolby_441", 5, 320000, 44100, A52_DOLBY | A52_ADJUST_LEVEL, 0.0),
            ("enc30_mono_32k", 3, 192000, 32000, A52_MONO | A52_ADJUST_LEVEL, 0.0),
            ("enc20_halfrate", 2, 64000, 24000, A52_STEREO, 0.0)]:
        pcm = synth_pcm(1, len(cases), nch, 1536 * 4, rate)
        fb, es = enc.encode_stream(pcm, rate, br)
        cases.append((name, es, flags, bias))
    # feature-rich synthetic streams (config 3): block switching, coupling, rematrix, dynrng, deltba, skip
    for name, acmod, lfe, flags, fscod, cod in [
            ("syn51_stereo", 7, 1, A52_STEREO | A52_ADJUST_LEVEL, 0, 36),
            ("syn51_51", 7, 1, A52_3F2R | A52_LFE, 0, 36),
            ("syn20_remat", 2, 0, A52_STEREO, 0, 24),
            ("syn11_dualmono", 0, 0, A52_CHANNEL, 0, 24),
            ("syn31_2f2r_441", 5, 1, A52_2F2R | A52_LFE, 1, 33),
            ("syn22_dolby_32k", 6, 0, A52_DOLBY | A52_ADJUST_LEVEL, 2, 30)]:
        es, fb = make_stream(4242 + len(cases), acmod, lfe, 4, ora.bit_allocate, fscod=fscod, frmsizecod=cod)
        cases.append((name, es, flags, bias if False else 0.0))
    names = []
    for name, es, flags, bias in cases:
        nf, pcm = ref.decode_stream(es, flags, 1.0, bias)
        assert nf == 4, (name, nf)
        fr = ref.decode_dump(es, req_flags=flags, bias=bias)
        b0 = fr[0]["blocks"][0]
        out[name + ".es"] = es
        out[name + ".req"] = np.array([flags, int(bias)], np.int32)
        out[name + ".pcm"] = pcm.astype(np.float32)
        out[name + ".lfsr"] = np.array([fr[-1]["blocks"][-1]["info"][8]], np.int32)
        out[name + ".info"] = np.stack([blk["info"] for f in fr for blk in f["blocks"]]).astype(np.int32)
        out[name + ".exp"] = np.stack([blk["exp"] for blk in fr[0]["blocks"]])
        out[name + ".bap"] = np.stack([blk["bap"] for blk in fr[0]["blocks"]])
        names.append(name)
        print(name, len(es), pcm.shape, float(np.abs(pcm).max()))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "decode_vectors.npz"), **out)

    # config-2 corpus base
    streams = []
    for s in range(4):
        pcm = synth_pcm(2, s, 6, 1536 * 64)
        fb, es = enc.encode_stream(pcm, 48000, 448000)
        assert fb == 1792
        streams.append(es.reshape(64, 1792))
    streams = np.stack(streams)
    dig = []
    for s in range(4):
        nf, pcm = ref.decode_stream(streams[s].reshape(-1), A52_STEREO | A52_ADJUST_LEVEL, 1.0, 0.0)
        assert nf == 64
        p = pcm.astype(np.float64)
        dig.append([p.sum(), (p * p).sum(), np.abs(p).max()])
    np.savez_compressed(os.path.join(HERE, "c2_fixture.npz"), frames=streams, digest=np.array(dig))
    print("c2 fixture", streams.shape, dig)


if __name__ == "__main__":
    main()
