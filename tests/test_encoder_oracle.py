"""CPU tests of the encoder oracle (oracle/ac3enc_oracle.c) against the committed golden frames
(tests/golden/encode_vectors.npz, generated from the unmodified reference encoder) and, where the
reference build exists, against the reference itself on more configurations: frames byte for byte,
MDCT coefficients, exponents, strategies, baps and snr offsets (ac3enc.cpp:1640-1763)."""
import ctypes as C
import os

import numpy as np
import pytest

from refbind import OracleEnc
from synth import synth_pcm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def egold():
    return np.load(os.path.join(ROOT, "tests", "golden", "encode_vectors.npz"))


def test_encoder_oracle_matches_golden_frames(engine, egold):
    ora = OracleEnc()
    for name in egold["names"]:
        name = str(name)
        nch, br, rate, nfr = [int(x) for x in egold[name + ".cfg"]]
        pcm = egold[name + ".pcm"]
        fb = ora.init(rate, br, nch)
        assert fb == egold[name + ".frames"].shape[1]
        for f in range(nfr):
            fr = ora.frame(pcm[f * 1536:(f + 1) * 1536], egold[name + ".chmap"])
            assert (fr == egold[name + ".frames"][f]).all(), (name, f)
            assert (ora.get(2)[:, :nch] == egold[name + ".strategy"][f][:, :nch]).all(), (name, f)
            assert (ora.get(6)[:2] == egold[name + ".snr"][f]).all(), (name, f)
            assert sum(int(ora.get(4)[:, c, :(7 if (nch == 6 and c == 5) else 223)].astype(np.int64).sum()) for c in range(nch)) == int(egold[name + ".bapsum"][f])
            assert int(np.abs(ora.get(0)[:, :nch].astype(np.int64)).sum()) == int(egold[name + ".coefsum"][f])


def test_encoder_init_rejections(engine):
    ora = OracleEnc()
    assert ora.init(48000, 448000, 7) == 0 and ora.init(48000, 448000, 0) == 0
    assert ora.init(47000, 448000, 2) == 0 and ora.init(48000, 450000, 2) == 0
    assert ora.init(48000, 448000, 6) == 1792 and ora.init(44100, 320000, 2) == 2 * (320000 * 1536 // (44100 * 16))
    assert ora.init(24000, 64000, 2) == 512 and ora.init(12000, 32000, 1) == 512      # half / quarter rates


def test_encoder_tables_vs_reference(engine, refenc):
    ora = OracleEnc()
    refenc.lib.ref_ac3enc_init(48000, 448000, 6)
    w = np.zeros(256, np.int16)
    refenc.lib.ref_ac3enc_get(7, w.ctypes.data_as(C.c_void_p))
    assert (w == ora.window()).all()


@pytest.mark.parametrize("nch,br,rate", [(6, 448000, 48000), (6, 384000, 44100), (6, 256000, 32000), (5, 320000, 48000),
                                         (4, 192000, 44100), (3, 128000, 48000), (2, 96000, 48000), (2, 448000, 48000),
                                         (1, 64000, 32000), (2, 128000, 22050), (1, 32000, 12000)])
def test_encoder_oracle_vs_reference(engine, refenc, nch, br, rate):
    ora = OracleEnc()
    for seed, noise, bursts in ((1, 0.02, False), (2, 0.3, True), (3, 0.0005, False)):
        pcm = synth_pcm(6, seed * 8 + nch, nch, 1536 * 8, rate, noise=noise, bursts=bursts)
        if seed == 3:
            pcm[1536 * 3:1536 * 5] = 0                      # digital silence: all-zero coefficients
        fb, a = refenc.encode_stream(pcm, rate, br)
        fb2, b = ora.encode_stream(pcm, rate, br)
        assert fb == fb2 and (a == b).all(), (seed, int((a != b).sum()))
    # intermediates frame by frame, with a channel map
    chmap = list(range(nch))[::-1]
    pcm = synth_pcm(6, 99 + nch, nch, 1536 * 4, rate)
    assert refenc.lib.ref_ac3enc_init(rate, br, nch) == ora.init(rate, br, nch)
    cm = np.array(chmap, np.uint8)
    for f in range(4):
        x = np.ascontiguousarray(pcm[f * 1536:(f + 1) * 1536])
        dst = np.zeros(3840 + 64, np.uint8)
        n = refenc.lib.ref_ac3enc_frame(dst.ctypes.data_as(C.POINTER(C.c_uint8)), x.ctypes.data_as(C.POINTER(C.c_short)),
                                        cm.ctypes.data_as(C.POINTER(C.c_uint8)))
        fr = ora.frame(x, chmap)
        assert (fr == dst[:n]).all()
        assert (refenc.get(0)[:, :nch] == ora.get(0)[:, :nch]).all()          # mdct coefficients
        assert (refenc.get(5)[:, :nch] == ora.get(5)[:, :nch]).all()          # block exponent shifts
        assert (refenc.get(2)[:, :nch] == ora.get(2)[:, :nch]).all()          # strategies
        assert (refenc.get(3)[:, :nch, :7] == ora.get(3)[:, :nch, :7]).all()  # encoded exponents (lfe range)
        for ch in range(nch):
            ncoef = 7 if (nch == 6 and ch == 5) else 223
            assert (refenc.get(3)[:, ch, :ncoef] == ora.get(3)[:, ch, :ncoef]).all()
            assert (refenc.get(4)[:, ch, :ncoef] == ora.get(4)[:, ch, :ncoef]).all()
        assert (refenc.get(6)[:2] == ora.get(6)[:2]).all()


def test_round_trip_snr(engine, oracle):
    """PCM -> encoder oracle -> decoder oracle reproduces the input within codec noise."""
    from refbind import A52_3F2R, A52_LFE
    ora = OracleEnc()
    pcm = synth_pcm(6, 5, 6, 1536 * 10, 48000)
    fb, es = ora.encode_stream(pcm, 48000, 448000)
    nf, dec = oracle.decode_stream(es, A52_3F2R | A52_LFE, 1.0, 0.0)
    assert nf == 10
    # decoder planes: LFE, L, C, R, LS, RS ; encoder coded order: L, C, R, LS, RS, LFE ; 256-sample latency
    dec = dec.reshape(10, 6, 6, 256).transpose(0, 1, 3, 2).reshape(-1, 6)
    x = pcm[:, [5, 0, 1, 2, 3, 4]].astype(np.float64) / 32768.0
    n = len(x) - 256
    err = dec[256:256 + n] - x[:n]
    snr = 10 * np.log10((x[:n] ** 2).sum() / (err ** 2).sum())
    assert snr > 20, snr
