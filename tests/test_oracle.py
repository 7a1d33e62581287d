"""CPU tests of the oracle (oracle/liboracle.so, our restatement of liba52's decode path).

Pins the oracle against
  * the committed golden vectors (tests/golden/decode_vectors.npz, generated from the
    UNMODIFIED reference by tests/golden/make_golden.py), and
  * the reference build itself (oracle/_ref) where it exists: fuzzed bit allocation
    (bit_allocate.c:124-265), both transforms (imdct.c:258-345), whole streams incl. the
    feature-rich synthetic ones, error returns of corrupted frames.
"""
import numpy as np
import pytest

from refbind import (A52_3F2R, A52_LFE, A52_STEREO, A52_ADJUST_LEVEL, A52_MONO, A52_DOLBY, A52_2F2R, A52_3F,
                     A52_CHANNEL, A52_CHANNEL1, A52_CHANNEL2, A52_2F1R, A52_3F1R)
from bitstream_writer import make_stream, frame_bytes


def test_oracle_matches_golden_pcm_bit_exact(oracle, golden):
    for name in golden["names"]:
        name = str(name)
        es = golden[name + ".es"]
        flags, bias = [int(x) for x in golden[name + ".req"]]
        nf, pcm = oracle.decode_stream(es, flags, 1.0, float(bias))
        assert nf == 4, name
        ref = golden[name + ".pcm"]
        assert pcm.shape == ref.shape, name
        assert (pcm.view(np.uint32) == ref.view(np.uint32)).all(), name


def test_oracle_matches_golden_exp_bap_lfsr(oracle, golden):
    for name in golden["names"]:
        name = str(name)
        es = golden[name + ".es"]
        flags, bias = [int(x) for x in golden[name + ".req"]]
        fr = oracle.decode_dump(es, req_flags=flags, bias=float(bias))
        assert len(fr) == 4 and all(f["status"] == 0 for f in fr), name
        info = np.stack([blk["info"] for f in fr for blk in f["blocks"]])
        ginfo = golden[name + ".info"]
        # column 12 (`downmixed`) is an internal flag of liba52 the restatement does not keep;
        # the coupling fields (5, 6, 13) and rematflg (14) are uninitialised malloc memory in the
        # reference until a stream uses them (parse.c:59-67 zeroes nothing but the samples)
        cols = [7, 8, 9, 10, 11, 15]
        assert (info[:, cols] == ginfo[:, cols]).all(), name
        nfch0 = [2, 1, 2, 3, 3, 4, 4, 5][ginfo[0, 9]]
        assert (info[:, :nfch0] == ginfo[:, :nfch0]).all(), name      # endmant of the coded channels
        cpl = ginfo[:, 7] != 0
        assert (info[cpl][:, [5, 6, 13]] == ginfo[cpl][:, [5, 6, 13]]).all(), name
        st = ginfo[:, 9] == 2
        assert (info[st][:, 14] == ginfo[st][:, 14]).all(), name
        assert fr[-1]["blocks"][-1]["info"][8] == int(golden[name + ".lfsr"][0]), name
        for b in range(6):
            blk = fr[0]["blocks"][b]
            gi = ginfo[b]
            nfch = [2, 1, 2, 3, 3, 4, 4, 5][gi[9]]
            for ch in range(nfch):
                end = gi[ch]
                assert (blk["exp"][ch, :end] == golden[name + ".exp"][b, ch, :end]).all(), (name, b, ch)
                assert (blk["bap"][ch, :end] == golden[name + ".bap"][b, ch, :end]).all(), (name, b, ch)
            if gi[10]:
                assert (blk["exp"][5, :7] == golden[name + ".exp"][b, 5, :7]).all()
                assert (blk["bap"][5, :7] == golden[name + ".bap"][b, 5, :7]).all()
            if gi[7]:
                s, e = gi[5], gi[6]
                assert (blk["exp"][6, s:e] == golden[name + ".exp"][b, 6, s:e]).all()
                assert (blk["bap"][6, s:e] == golden[name + ".bap"][b, 6, s:e]).all()


def test_c2_fixture_digest(oracle, c2):
    """The committed config-2 corpus decodes to the digests recorded from the reference."""
    for s in range(c2["frames"].shape[0]):
        nf, pcm = oracle.decode_stream(c2["frames"][s].reshape(-1), A52_STEREO | A52_ADJUST_LEVEL, 1.0, 0.0)
        assert nf == 64
        p = pcm.astype(np.float64)
        d = np.array([p.sum(), (p * p).sum(), np.abs(p).max()])
        assert np.allclose(d, c2["digest"][s], rtol=0, atol=0), (s, d, c2["digest"][s])


def test_syncinfo_all_headers(oracle):
    """a52_syncinfo over every (fscod, frmsizecod, bsid, acmod) header (parse.c:86-129)."""
    rates = [48000, 44100, 32000]
    kb = [32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384, 448, 512, 576, 640]
    for fscod in range(4):
        for cod in range(64):
            for bsid in (0, 6, 8, 9, 10, 11, 12, 16):
                hdr = bytes([0x0B, 0x77, 0, 0, (fscod << 6) | cod, bsid << 3, 7 << 5 | 1])
                n, fl, sr, br = oracle.syncinfo(np.frombuffer(hdr, np.uint8))
                if fscod == 3 or cod >= 38 or bsid >= 12:
                    assert n == 0
                    continue
                half = max(bsid - 8, 0)
                k = kb[cod >> 1]
                want = 4 * k if fscod == 0 else 6 * k if fscod == 2 else 2 * (320 * k // 147 + (cod & 1))
                assert n == want and n == frame_bytes(fscod, cod)
                assert sr == rates[fscod] >> half and br == (k * 1000) >> half
    assert oracle.syncinfo(np.frombuffer(bytes([0x0B, 0x76, 0, 0, 0, 0x40, 0]), np.uint8))[0] == 0


# ---------------------------------------------------------------------------
# against the reference build
# ---------------------------------------------------------------------------
def test_syncinfo_vs_reference(oracle, ref):
    rng = np.random.RandomState(1)
    for _ in range(3000):
        h = rng.randint(0, 256, 7).astype(np.uint8)
        h[0], h[1] = 0x0B, 0x77
        assert oracle.syncinfo(h) == ref.syncinfo(h)


def test_bit_allocate_fuzz_vs_reference(oracle, ref):
    rng = np.random.RandomState(7)
    for it in range(600):
        # legal ends only: 7 (lfe) or >= 37 (fbw: 73 + 3*chbwcod, or cplstrtmant = 37 + 12*cplbegf);
        # for other values the reference reads exponents past `end` (bit_allocate.c:203-215)
        end = 7 if it % 10 == 0 else int(rng.randint(37, 254))
        exp = np.clip(np.cumsum(rng.randint(-2, 3, 256)) + 8, 0, 24).astype(np.uint8)
        fscod, half = int(rng.randint(3)), int(rng.randint(3))
        bai, csnr, chbai = int(rng.randint(2048)), int(rng.randint(64)), int(rng.randint(128))
        deltba = None
        if it % 3 == 0:
            deltba = rng.randint(-4, 5, 50).astype(np.int8)
        a = oracle.bit_allocate(fscod, half, bai, csnr, chbai, exp, end, deltba)
        b = ref.bit_allocate(fscod, half, bai, csnr, chbai, exp, end, deltba)
        assert (a[:end] == b[:end]).all(), it
    # coupling-channel form (start != 0, leak initialisers)
    cpl_band = [31, 35, 37, 39, 41, 42, 43, 44, 45, 45, 46, 46, 47, 47, 48, 48]
    for it in range(300):
        begf = int(rng.randint(0, 12))
        endf = int(rng.randint(max(begf - 2, 0), 16))
        start, end = 37 + 12 * begf, 73 + 12 * endf
        exp = np.clip(np.cumsum(rng.randint(-2, 3, 256)) + 8, 0, 24).astype(np.uint8)
        args = (int(rng.randint(3)), int(rng.randint(3)), int(rng.randint(2048)), int(rng.randint(64)),
                int(rng.randint(128)), exp, end, None, cpl_band[begf], start,
                (9 - int(rng.randint(8))) << 8, (9 - int(rng.randint(8))) << 8)
        a, b = oracle.bit_allocate(*args), ref.bit_allocate(*args)
        assert (a[start:end] == b[start:end]).all(), it


def test_imdct_bit_exact_vs_reference(oracle, ref):
    rng = np.random.RandomState(3)
    for kind in (512, 256):
        for bias in (0.0, 384.0):
            x = (rng.rand(256).astype(np.float32) - 0.5)
            d = (rng.rand(256).astype(np.float32) - 0.5)
            a, ad = oracle.imdct(kind, x, d, bias)
            b, bd = ref.imdct(kind, x, d, bias)
            assert (a.view(np.uint32) == b.view(np.uint32)).all()
            assert (ad.view(np.uint32) == bd.view(np.uint32)).all()


@pytest.mark.parametrize("acmod,lfe,flags,fscod,cod", [
    (7, 1, A52_STEREO | A52_ADJUST_LEVEL, 0, 36), (7, 1, A52_3F2R | A52_LFE, 0, 36),
    (7, 0, A52_MONO, 0, 30), (7, 1, A52_DOLBY, 0, 34), (7, 1, A52_2F1R | A52_LFE | A52_ADJUST_LEVEL, 0, 36),
    (2, 0, A52_STEREO, 0, 24), (2, 0, A52_MONO | A52_ADJUST_LEVEL, 1, 22),
    (0, 0, A52_CHANNEL, 0, 24), (0, 0, A52_CHANNEL1, 0, 24), (0, 1, A52_CHANNEL2, 0, 24), (0, 0, A52_MONO, 0, 24),
    (1, 0, A52_STEREO, 2, 14), (3, 0, A52_STEREO | A52_ADJUST_LEVEL, 0, 28), (4, 1, A52_3F1R | A52_LFE, 0, 28),
    (5, 1, A52_2F2R | A52_LFE, 1, 33), (6, 0, A52_DOLBY | A52_ADJUST_LEVEL, 2, 30), (6, 0, A52_3F, 0, 30),
])
def test_feature_streams_oracle_vs_reference(oracle, ref, acmod, lfe, flags, fscod, cod):
    """Config-3 style streams from the test-only writer: block switching, coupling, rematrixing,
    dynrng, delta bit allocation, skip fields, every acmod - PCM and state bit-exact."""
    es, fb = make_stream(1000 + acmod * 16 + (flags & 15), acmod, lfe, 3, oracle.bit_allocate, fscod=fscod,
                         frmsizecod=cod)
    nf_r, pcm_r = ref.decode_stream(es, flags, 1.0, 0.0)
    nf_o, pcm_o = oracle.decode_stream(es, flags, 1.0, 0.0)
    assert nf_r == 3 and nf_o == 3          # the unmodified reference accepts every frame
    assert (pcm_r.view(np.uint32) == pcm_o.view(np.uint32)).all()
    fr_r = ref.decode_dump(es, req_flags=flags)
    fr_o = oracle.decode_dump(es, req_flags=flags)
    assert fr_r[-1]["blocks"][-1]["info"][8] == fr_o[-1]["blocks"][-1]["info"][8]     # lfsr_state
    # drc off
    nf_r, pcm_r = ref.decode_stream(es, flags, 1.0, 0.0, dynrng_off=True)
    nf_o, pcm_o = oracle.decode_stream(es, flags, 1.0, 0.0, dynrng_off=True)
    assert (pcm_r.view(np.uint32) == pcm_o.view(np.uint32)).all()


def test_reference_encoded_streams_oracle_vs_reference(oracle, ref, refenc):
    from synth import synth_pcm
    for nch, br, rate, flags, bias in [(6, 448000, 48000, A52_STEREO | A52_ADJUST_LEVEL, 384.0),
                                       (2, 192000, 48000, A52_STEREO | A52_ADJUST_LEVEL, 384.0),
                                       (4, 256000, 44100, A52_3F, 0.0), (1, 64000, 32000, A52_MONO, 0.0)]:
        pcm = synth_pcm(5, nch, nch, 1536 * 6, rate)
        fb, es = refenc.encode_stream(pcm, rate, br)
        nf_r, pcm_r = ref.decode_stream(es, flags, 1.0, bias)
        nf_o, pcm_o = oracle.decode_stream(es, flags, 1.0, bias)
        assert nf_r == nf_o == 6
        assert (pcm_r.view(np.uint32) == pcm_o.view(np.uint32)).all()


def test_corrupted_frames_same_errors(oracle, ref):
    """Bit flips in the middle frame of a 3-frame stream: same per-frame error returns as the
    reference (parse.c:163-164, 227-256, 287-288, 612-621, 697-698) and the same PCM for every
    block the reference produces.  The first frame is left intact so that the decoder state the
    damaged frame may fall back on (exponent / bit-allocation reuse) is initialised - on a fresh
    state the reference would read uninitialised malloc memory there (parse.c:59-67)."""
    rng = np.random.RandomState(11)
    es0, fb = make_stream(77, 7, 1, 3, oracle.bit_allocate)
    nerr = 0
    for it in range(200):
        es = es0.copy()
        for _ in range(int(rng.randint(1, 4))):
            p = fb + int(rng.randint(6, 500))
            es[p] ^= 1 << int(rng.randint(8))
        a = ref.decode_dump(es, req_flags=A52_STEREO)
        b = oracle.decode_dump(es, req_flags=A52_STEREO)
        sa, sb = [f["status"] for f in a], [f["status"] for f in b]
        assert sa == sb, (it, sa, sb)
        nerr += any(sa)
        for fa, fb_ in zip(a, b):
            # (never-sent fields read as zero on both sides: the checker build of the reference
            # callocs its state, see oracle/refbuild/a52_ref_wrap.c)
            for ba, bb in zip(fa["blocks"], fb_["blocks"]):
                assert np.array_equal(ba["pcm"].view(np.uint32), bb["pcm"].view(np.uint32)), it
    assert nerr > 10        # the fuzz does reach the error returns


@pytest.mark.parametrize("acmod,flags", [(4, 2), (6, 2), (5, 3), (7, 3 | 32), (7, 2), (6, 4), (3, 10), (0, 1)])
def test_bias_placement_oracle_vs_reference(oracle, ref, acmod, flags):
    """Where liba52 adds the bias depends on the transform path: the IMDCT of pass-through channels and the
    time-domain mixers add it, and with slev == 0 no mixer runs for 2/1, 2/2 -> stereo and 3/1, 3/2 -> 3F
    (downmix.c:526-573, parse.c:893-918).  The CUDA path is checked against the oracle for this
    (test_bias_follows_the_reference_path_by_path); here the oracle is pinned to the reference, bit for bit."""
    from bitstream_writer import make_stream
    es, fb = make_stream(900 + acmod, acmod, 0, 6, oracle.bit_allocate, frmsizecod=30, features=dict(blksw=0.5))
    for bias in (384.0, 1.0, 0.0):
        n1, a = oracle.decode_stream(es, flags, 1.0, bias)
        n2, b = ref.decode_stream(es, flags, 1.0, bias)
        assert n1 == n2 == 6 and (a.view(np.uint32) == b.view(np.uint32)).all(), (acmod, flags, bias)


def test_oracle_error_returns_match_reference_on_crafted_fields():
    """Every `return 1` site of a52_block (parse.c:218-294, 600-701), hit by a deliberately invalid field: the oracle
    port stops in the same block as the unmodified reference."""
    import pytest
    from refbind import RefA52, Oracle, have_ref, A52_STEREO
    from bitstream_writer import make_stream
    if not have_ref():
        pytest.skip("reference not built")
    ref, ora = RefA52(), Oracle()
    for site, acmod, lfe, blk in [("chbwcod", 7, 1, 0), ("exp_code", 2, 0, 0), ("exp_range", 7, 1, 0), ("cpl_range", 5, 0, 3),
                                  ("cpl_mono", 0, 1, 2), ("cpl_mono", 1, 0, 4), ("deltba_len", 7, 1, 5), ("deltba_len", 2, 0, 0)]:
        es, fb = make_stream(900 + 7 * acmod + blk, acmod, lfe, 3, ora.bit_allocate, frmsizecod=30, inject=(1, blk, site))
        a = [f["status"] for f in ref.decode_dump(es, req_flags=A52_STEREO)]
        b = [f["status"] for f in ora.decode_dump(es, req_flags=A52_STEREO)]
        assert a == b == [0, 2 + blk, 0], (site, acmod, blk, a, b)
