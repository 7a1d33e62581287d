"""Round-2 features on the GPU: frame-independent slices, the scan pass, dynamic range control by table and the
a52_dynrng callback of the drop-in API (against the UNMODIFIED reference, oracle/_ref, where it is present)."""
import ctypes as C
import os

import numpy as np
import pytest

from refbind import A52_STEREO, A52_3F2R, A52_LFE, A52_ADJUST_LEVEL, have_ref
from bitstream_writer import make_stream

pytestmark = pytest.mark.gpu


def _long_streams(c2, oracle, n_c2=3, frames=150):
    """A few long streams: reference-encoded material (tiled fixture) and a feature-rich synthetic one."""
    chunks, off, first, pos = [], [], [0], 0
    for s in range(n_c2):
        idx = (np.arange(frames + 13 * s) + 5 * s) % c2["frames"].shape[1]
        fr = c2["frames"][s % 4, idx].reshape(-1)
        chunks.append(fr)
        off += [pos + 1792 * k for k in range(len(idx))]
        pos += len(fr)
        first.append(first[-1] + len(idx))
    es_s, fb = make_stream(4711, 7, 1, 70, oracle.bit_allocate, frmsizecod=36,
                           features=dict(blksw=0.5, cpl=0.9, dynrng=0.5, deltba=0.1))
    chunks.append(np.asarray(es_s, np.uint8))
    off += [pos + fb * k for k in range(70)]
    pos += len(es_s)
    first.append(first[-1] + 70)
    return np.concatenate(chunks), np.array(off, np.uint64), np.array(first, np.uint32)


@pytest.mark.parametrize("flags", [A52_STEREO | A52_ADJUST_LEVEL, A52_3F2R | A52_LFE])
def test_frame_independent_slices_are_bit_identical(engine, c2, oracle, flags):
    """All slices of a stream decoded side by side (scan pass + prefix sum of the dither draws + one frame of
    look-back) give the same bits, the same status and the same final carry as the chained walk."""
    es, off, first = _long_streams(c2, oracle)
    res = []
    for mode in (engine.SLICES_CHAINED, engine.SLICES_INDEPENDENT):
        dec = engine.BatchDecoder(0)
        dec.set_slice_mode(mode)
        carry = [engine.CarryStruct() for _ in range(len(first) - 1)]
        carry[1].dither_index = 12345                       # a caller-supplied starting state travels too
        out = dec.decode_host(es, off, first, flags, carry=carry)
        res.append(out)
        dec.close()
    a, b = res
    assert (a["status"] == 0).all() and (b["status"] == 0).all()
    assert (a["pcm"].view(np.uint32) == b["pcm"].view(np.uint32)).all()
    assert (a["flags"] == b["flags"]).all()
    for ca, cb in zip(a["carry"], b["carry"]):
        assert ca.dither_index == cb.dither_index and ca.per_channel == cb.per_channel
        assert bytes(ca.delay) == bytes(cb.delay)


def test_scan_counts_dither_draws_and_finds_dynrng_words(decoder, engine, c2, oracle):
    es, off, first = _long_streams(c2, oracle, n_c2=1, frames=40)
    flags = A52_STEREO | A52_ADJUST_LEVEL
    sc = decoder.scan_host(es, off, first, flags)
    assert (sc["status"] == 0).all()
    carry = [engine.CarryStruct() for _ in range(len(first) - 1)]
    out = decoder.decode_host(es, off, first, flags, carry=carry)
    for s in range(len(first) - 1):
        draws = int(sc["dither_draws"][first[s]:first[s + 1]].astype(np.int64).sum())
        assert draws % 65535 == out["carry"][s].dither_index
    # the reference encoder writes no dynrng words, the synthetic stream has them in about half of its blocks
    w0 = sc["dynrng"][first[0]:first[1]]
    w1 = sc["dynrng"][first[1]:first[2]]
    assert (w0 == -1).all()
    present = (w1[:, :, 0] >= 0).mean()
    assert 0.3 < present < 0.7 and (w1[:, :, 1] == -1).all()
    # a table that repeats liba52's own mapping of every word reproduces the stream mode bit for bit
    d = w1.astype(np.int32)
    d8 = np.where(d >= 128, d - 256, d)
    rng_tab = (((d8 & 0x1f) | 0x20) << 13).astype(np.float32) * np.ldexp(np.float32(1.0), -(18 - (d8 >> 5))).astype(np.float32)
    table = np.ones((len(off), 6, 2), np.float32)
    table[first[1]:first[2]] = np.where(d >= 0, rng_tab, 1.0)
    decoder.set_drc_table(table)
    out_t = decoder.decode_host(es, off, first, flags, drc=engine.DRC_TABLE)
    decoder.set_drc_table(None)
    assert (out_t["pcm"].view(np.uint32) == out["pcm"].view(np.uint32)).all()
    # and a table of ones is compression off
    decoder.set_drc_table(np.ones((len(off), 6, 2), np.float32))
    out_1 = decoder.decode_host(es, off, first, flags, drc=engine.DRC_TABLE)
    decoder.set_drc_table(None)
    out_off = decoder.decode_host(es, off, first, flags, drc=engine.DRC_OFF)
    assert (out_1["pcm"].view(np.uint32) == out_off["pcm"].view(np.uint32)).all()


@pytest.mark.skipif(not have_ref(), reason="reference not built (make -C oracle ref)")
def test_dynrng_callback_like_liba52(engine, oracle):
    """a52_dynrng(state, call, data): the callback sees every dynrng word's range in block order and what it
    returns is applied (parse.c:207-216, 586-595) - drop-in API against the unmodified reference."""
    from refbind import RefA52
    ref = RefA52()
    L = engine.load_library()
    es, fb = make_stream(99, 7, 1, 6, oracle.bit_allocate, frmsizecod=36, features=dict(blksw=0.3, cpl=0.5, dynrng=0.7))
    es = np.ascontiguousarray(np.concatenate([np.asarray(es, np.uint8), np.zeros(16, np.uint8)]))
    CB = C.CFUNCTYPE(C.c_float, C.c_float, C.c_void_p)
    seen = {"ref": [], "gpu": []}

    def make_cb(key):
        def cb(rng, data):
            seen[key].append(rng)
            return C.c_float(0.5 * rng + 0.25).value          # any mapping: halve the compression and shift it
        return CB(cb)

    outs = {}
    for key, lib, pre in (("ref", ref.lib, "ref_"), ("gpu", L, "")):
        fn = lambda name: getattr(lib, pre + name)
        init = fn("a52_init"); init.restype = C.c_void_p
        samples = fn("a52_samples"); samples.restype = C.POINTER(C.c_float); samples.argtypes = [C.c_void_p]
        frame = fn("a52_frame"); frame.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_float]
        dyn = fn("a52_dynrng"); dyn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        block = fn("a52_block"); block.argtypes = [C.c_void_p]
        free = fn("a52_free"); free.argtypes = [C.c_void_p]
        st = init(0)
        assert st
        cb = make_cb(key)
        pcm = []
        for f in range(6):
            flags, level = C.c_int(A52_STEREO | A52_ADJUST_LEVEL), C.c_float(1.0)
            assert frame(st, es.ctypes.data + f * fb, C.byref(flags), C.byref(level), 0.0) == 0
            dyn(st, C.cast(cb, C.c_void_p), None)
            for b in range(6):
                assert block(st) == 0
                pcm.append(np.ctypeslib.as_array(samples(st), (2 * 256,)).copy())
        free(st)
        outs[key] = np.stack(pcm)
    assert len(seen["ref"]) > 10 and seen["gpu"] == seen["ref"]
    d = outs["gpu"].astype(np.float64) - outs["ref"]
    assert np.sqrt((d * d).mean()) / np.sqrt((outs["ref"].astype(np.float64) ** 2).mean()) < 1e-5


# ---------------------------------------------------------------------------
# every `return 1` of a52_block, triggered by a crafted field (parse.c:218-294, 600-701)
# ---------------------------------------------------------------------------
ERROR_SITES = [
    ("chbwcod", 7, 1, 0), ("chbwcod", 3, 0, 0), ("exp_code", 7, 1, 0), ("exp_code", 2, 0, 0), ("exp_range", 7, 1, 0),
    ("exp_range", 1, 0, 0), ("cpl_range", 7, 1, 1), ("cpl_range", 2, 0, 0), ("cpl_range", 5, 0, 3), ("cpl_mono", 1, 0, 4),
    ("cpl_mono", 0, 0, 0), ("cpl_mono", 0, 1, 2), ("deltba_len", 7, 1, 5), ("deltba_len", 3, 0, 2), ("deltba_len", 2, 0, 0),
]


@pytest.mark.parametrize("site,acmod,lfe,blk", ERROR_SITES)
def test_every_error_return_of_a52_block(decoder, engine, oracle, site, acmod, lfe, blk):
    """A frame whose block `blk` carries one invalid field fails exactly there (status 16 + blk, what a52_block
    returning 1 in that block means), its first `blk` blocks and its neighbours decode as in the reference."""
    from refbind import RefA52
    chk = RefA52() if have_ref() else oracle
    es, fb = make_stream(900 + 7 * acmod + blk, acmod, lfe, 3, oracle.bit_allocate, frmsizecod=30, inject=(1, blk, site))
    flags = A52_STEREO
    dump = chk.decode_dump(es, req_flags=flags)
    want = [0 if f["status"] == 0 else (2 if f["status"] == 1 else 16 + f["status"] - 2) for f in dump]
    assert want == [0, 16 + blk, 0], (site, want)               # the crafted field is reached and refused
    off = np.arange(3, dtype=np.uint64) * fb
    out = decoder.decode_host(es, off, np.array([0, 3], np.uint32), flags)
    assert out["status"].tolist() == want
    # frame 0 and the good blocks of frame 1 carry the reference's samples, the rest of frame 1 is silence
    ref0 = np.stack([b["pcm"] for b in dump[0]["blocks"]])
    d = out["pcm"][0].reshape(6, 2, 256).astype(np.float64) - ref0
    assert np.sqrt((d * d).mean()) / max(np.sqrt((ref0.astype(np.float64) ** 2).mean()), 1e-30) < 1e-5
    got1 = out["pcm"][1].reshape(6, 2, 256)
    for b in range(blk):
        r = dump[1]["blocks"][b]["pcm"]
        dd = got1[b].astype(np.float64) - r
        assert np.sqrt((dd * dd).mean()) <= 1e-5 * max(np.sqrt((r.astype(np.float64) ** 2).mean()), 1e-30) + 1e-9
    assert not got1[blk:].any()


# ---------------------------------------------------------------------------
# coefficient bit patterns on feature-rich streams (coupling, rematrixing, block switching, delta bit allocation)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("acmod,lfe,flags,cod", [(7, 1, A52_3F2R | A52_LFE, 36), (2, 0, A52_STEREO, 24), (5, 1, 5 | A52_LFE, 33)])
def test_feature_stream_coefficients_are_bit_exact(decoder, engine, oracle, acmod, lfe, flags, cod):
    """No downmix, compression off: every dequantised coefficient (coupled bins and rematrixed bands included) has the
    reference's bit pattern - the mantissa integers, the dither draws and the coupling arithmetic are the same."""
    es, fb = make_stream(3100 + acmod, acmod, lfe, 4, oracle.bit_allocate, frmsizecod=cod,
                         features=dict(blksw=0.4, cpl=0.9, deltba=0.2, dynrng=0.5))
    off = np.arange(4, dtype=np.uint64) * fb
    out = decoder.decode_host(es, off, np.array([0, 4], np.uint32), flags, drc=engine.DRC_OFF, want_debug=True)
    dump = oracle.decode_dump(es, req_flags=flags, dynrng_off=True)
    assert (out["status"] == 0).all()
    nfch = [2, 1, 2, 3, 3, 4, 4, 5][acmod]
    bad = 0
    for f in range(4):
        for b in range(6):
            blk = dump[f]["blocks"][b]
            g, r = out["coef"][f, b], blk["coef"]
            for ch in list(range(nfch)) + ([5] if lfe else []):
                # (value equality: a rematrixed or phase-flipped zero may carry either sign)
                bad += int((g[ch] != r[ch]).sum())
    assert bad == 0, bad
