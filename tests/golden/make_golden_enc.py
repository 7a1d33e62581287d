#!/usr/bin/env python3
"""Generate tests/golden/encode_vectors.npz from the UNMODIFIED reference encoder (oracle/_ref/ac3enc_ref.so).

Run in the build container after `make -C oracle ref`:   python tests/golden/make_golden_enc.py
Per case: the synthetic int16 PCM (seeded, tests/synth.py), the reference's frames, and per frame the
exponent strategies, snr offsets and a digest of baps / mdct coefficients.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from refbind import RefAc3Enc  # noqa
from synth import synth_pcm  # noqa

CASES = [("e51_448", 6, 448000, 48000, 4, None), ("e20_192", 2, 192000, 48000, 6, None),
         ("e10_96", 1, 96000, 48000, 3, None), ("e50_384_441", 5, 384000, 44100, 3, None),
         ("e30_192_32k", 3, 192000, 32000, 3, None), ("e20_64_half", 2, 64000, 24000, 3, None),
         ("e51_640_map", 6, 640000, 48000, 3, [0, 2, 1, 4, 5, 3]), ("e40_256_bursts", 4, 256000, 48000, 4, None)]


def main():
    enc = RefAc3Enc()
    out = {"names": np.array([c[0] for c in CASES])}
    for name, nch, br, rate, nfr, chmap in CASES:
        pcm = synth_pcm(4, len(out), nch, 1536 * nfr, rate, bursts="bursts" in name)
        fb = enc.lib.ref_ac3enc_init(rate, br, nch)
        frames, strat, snr, bapsum, coefsum = [], [], [], [], []
        cm = np.array(chmap if chmap else list(range(nch)), np.uint8)
        for f in range(nfr):
            dst = np.zeros(3840 + 64, np.uint8)
            import ctypes as C
            x = np.ascontiguousarray(pcm[f * 1536:(f + 1) * 1536])
            n = enc.lib.ref_ac3enc_frame(dst.ctypes.data_as(C.POINTER(C.c_uint8)),
                                         x.ctypes.data_as(C.POINTER(C.c_short)),
                                         cm.ctypes.data_as(C.POINTER(C.c_uint8)))
            assert n == fb
            frames.append(dst[:fb].copy())
            strat.append(enc.get(2))
            snr.append(enc.get(6)[:2])
            bapsum.append(sum(int(enc.get(4)[:, c, :(7 if (nch == 6 and c == 5) else 223)].astype(np.int64).sum()) for c in range(nch)))  # valid bins only: beyond them the reference holds stack garbage (bap1, ac3enc.cpp:858)
            coefsum.append(int(np.abs(enc.get(0)[:, :nch].astype(np.int64)).sum()))
        out[name + ".cfg"] = np.array([nch, br, rate, nfr], np.int32)
        out[name + ".chmap"] = cm
        out[name + ".pcm"] = pcm
        out[name + ".frames"] = np.stack(frames)
        out[name + ".strategy"] = np.stack(strat)
        out[name + ".snr"] = np.stack(snr).astype(np.int32)
        out[name + ".bapsum"] = np.array(bapsum, np.int64)
        out[name + ".coefsum"] = np.array(coefsum, np.int64)
        print(name, fb, out[name + ".snr"].tolist())
    np.savez_compressed(os.path.join(HERE, "encode_vectors.npz"), **out)


if __name__ == "__main__":
    main()
