"""GPU parity tests of the encoder: CUDA path through the C ABI (include/ac3enc_batch.h, include/ac3enc.h)
against the encoder oracle (oracle/ac3enc_oracle.c, itself pinned byte for byte to the reference encoder).
Bar: frames byte-identical (=> 0 dB SNR delta), every integer intermediate identical."""
import ctypes as C
import os

import numpy as np
import pytest

from refbind import OracleEnc
from synth import synth_pcm

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def encoder(engine):
    enc = engine.BatchEncoder(0)
    yield enc
    enc.close()


@pytest.fixture(scope="module")
def egold():
    return np.load(os.path.join(ROOT, "tests", "golden", "encode_vectors.npz"))


def test_golden_frames(encoder, egold):
    for name in egold["names"]:
        name = str(name)
        nch, br, rate, nfr = [int(x) for x in egold[name + ".cfg"]]
        out = encoder.encode_host(egold[name + ".pcm"][None], rate, br, chmap=egold[name + ".chmap"], want_debug=True)
        assert (out["status"] == 0).all()
        assert (out["strategy"][0][:, :, :nch] == egold[name + ".strategy"][:, :, :nch]).all(), name
        assert (out["snr"][0] == egold[name + ".snr"]).all(), name
        assert (out["frames"][0] == egold[name + ".frames"]).all(), name


@pytest.mark.parametrize("nch,br,rate", [(6, 448000, 48000), (6, 384000, 44100), (5, 320000, 48000), (4, 192000, 44100),
                                         (3, 128000, 48000), (2, 192000, 48000), (2, 96000, 48000), (1, 64000, 32000),
                                         (2, 128000, 22050), (6, 640000, 48000)])
def test_encoder_all_stages_vs_oracle(encoder, nch, br, rate):
    ora = OracleEnc()
    nfr = 6
    pcm = np.stack([synth_pcm(7, 10 * nch + s, nch, 1536 * nfr, rate, noise=[0.02, 0.2, 0.001][s], bursts=(s == 1))
                    for s in range(3)])
    pcm[2, 1536 * 2:1536 * 3] = 0                                   # a silent frame
    chmap = list(range(nch))[::-1] if nch > 2 else None
    out = encoder.encode_host(pcm, rate, br, chmap=chmap, want_debug=True)
    assert (out["status"] == 0).all()
    for s in range(3):
        fb = ora.init(rate, br, nch)
        assert fb == out["frame_bytes"]
        for f in range(nfr):
            fr = ora.frame(pcm[s, f * 1536:(f + 1) * 1536], chmap)
            assert (out["coef"][s, f][:, :nch] == ora.get(0)[:, :nch]).all(), (s, f, "mdct")
            assert (out["exp_shift"][s, f][:, :nch] == ora.get(5)[:, :nch]).all(), (s, f, "shift")
            assert (out["strategy"][s, f][:, :nch] == ora.get(2)[:, :nch]).all(), (s, f, "strategy")
            for ch in range(nch):
                n = 7 if (nch == 6 and ch == 5) else 223
                assert (out["encoded_exp"][s, f][:, ch, :n] == ora.get(3)[:, ch, :n]).all(), (s, f, ch, "exp")
                assert (out["bap"][s, f][:, ch, :n] == ora.get(4)[:, ch, :n]).all(), (s, f, ch, "bap")
            assert (out["snr"][s, f] == ora.get(6)[:2]).all(), (s, f, "snr")
            assert (out["frames"][s, f] == fr).all(), (s, f, int((out["frames"][s, f] != fr).sum()))


def test_carry_split_equals_one_call(encoder, engine):
    pcm = synth_pcm(7, 77, 6, 1536 * 8, 48000)[None]
    whole = encoder.encode_host(pcm, 48000, 448000)
    a = encoder.encode_host(pcm[:, : 1536 * 3], 48000, 448000, carry=[engine.EncCarryStruct()])
    b = encoder.encode_host(pcm[:, 1536 * 3:], 48000, 448000, carry=[a["carry"][0]])
    got = np.concatenate([a["frames"], b["frames"]], axis=1)
    assert (got == whole["frames"]).all()


def test_dropin_api(engine):
    """AC3_encode_init / AC3_encode_frame used the way AC3ACM.cpp:1940, 1762 uses them."""
    L = engine.load_library()
    ora = OracleEnc()
    pcm = synth_pcm(7, 5, 2, 1536 * 4, 48000)
    fb = L.AC3_encode_init(48000, 192000, 2)
    assert fb == 768 == ora.init(48000, 192000, 2)
    cm = np.array([0, 1], np.uint8)
    for f in range(4):
        x = np.ascontiguousarray(pcm[f * 1536:(f + 1) * 1536])
        dst = np.zeros(3840 + 8, np.uint8)
        n = L.AC3_encode_frame(dst.ctypes.data, x.ctypes.data, cm.ctypes.data)
        assert n == fb
        assert (dst[:fb] == ora.frame(x)).all(), f
    assert L.AC3_encode_init(48000, 448000, 9) == 0


def test_infeasible_bitrate_reports_no_fit(encoder):
    """5.1 at 32 kb/s cannot fit (the reference prints "Yack" and emits garbage, ac3enc.cpp:930-933): the
    batch call flags the frames and still emits decodable frames, identical to the oracle's."""
    ora = OracleEnc()
    pcm = synth_pcm(7, 3, 6, 1536 * 3, 48000, noise=0.3)
    out = encoder.encode_host(pcm[None], 48000, 32000)
    fb, want = ora.encode_stream(pcm, 48000, 32000)
    assert (out["status"] == 1).any()
    assert (out["frames"][0].reshape(-1) == want).all()


def test_large_batch_properties_and_round_trip(encoder, engine, decoder, oracle):
    """Config 4 shape: many 5.1 streams at 448 kb/s in one launch through the device-pointer entry; every
    copy of a base stream gives identical bytes, the bytes match the oracle, and the GPU decoder plays the
    GPU encoder's frames back within codec noise."""
    import torch
    from refbind import A52_3F2R, A52_LFE
    ns, nfr, nb = 1024, 8, 4
    base = np.stack([synth_pcm(8, s, 6, 1536 * nfr, 48000) for s in range(nb)])
    pick = np.random.RandomState(0).randint(0, nb, ns)
    pcm = torch.from_numpy(np.ascontiguousarray(base[pick])).cuda()
    out = torch.zeros((ns, nfr, 1792), dtype=torch.uint8, device="cuda")
    status = torch.full((ns, nfr), -1, dtype=torch.int32, device="cuda")
    encoder.encode_device(pcm.data_ptr(), ns, nfr, 48000, 448000, 6, out.data_ptr(), status_ptr=status.data_ptr())
    torch.cuda.synchronize()
    assert int((status != 0).sum()) == 0
    ora = OracleEnc()
    pick_t = torch.from_numpy(pick).cuda()
    for b in range(nb):
        sel = out[pick_t == b]
        assert bool((sel == sel[0:1]).all())
        fb, want = ora.encode_stream(base[b], 48000, 448000)
        assert (sel[0].cpu().numpy().reshape(-1) == want).all()
    # round trip on the GPU
    es = out[0].cpu().numpy().reshape(-1)
    off = np.arange(nfr, dtype=np.uint64) * 1792
    dec = decoder.decode_host(es, off, np.array([0, nfr], np.uint32), A52_3F2R | A52_LFE)
    y = dec["pcm"].reshape(nfr, 6, 6, 256).transpose(0, 1, 3, 2).reshape(-1, 6)
    x = base[pick[0]][:, [5, 0, 1, 2, 3, 4]].astype(np.float64) / 32768.0
    n = len(x) - 256
    err = y[256:256 + n] - x[:n]
    assert 10 * np.log10((x[:n] ** 2).sum() / (err ** 2).sum()) > 20


def test_mixed_corpus_round_trip_sharded(encoder, engine, decoder, oracle):
    """Config 5 in small: a grid of (sample rate, channels, bitrate) cells inside the reference encoder's
    feasible region is encoded on the GPU (frames = the oracle's), then all streams - different frame
    sizes, channel modes and rates, 44.1 kHz frames unaligned - are decoded in ONE batch and again split
    into 2 and 4 logical shards (LPT partition + 16-byte repack): PCM matches the oracle decoder and is
    bit-identical whatever the sharding."""
    import importlib
    from refbind import A52_STEREO, A52_ADJUST_LEVEL
    from util import relrms, frame_offsets
    shard = importlib.import_module("ac3_acm_codec_b200.shard")
    ora = OracleEnc()
    cells = [(48000, 1, 64000), (48000, 2, 128000), (44100, 2, 96000), (32000, 3, 160000), (44100, 4, 224000),
             (48000, 5, 384000), (44100, 6, 448000), (32000, 6, 256000), (48000, 6, 640000), (24000, 2, 64000)]
    nfr = 5
    streams = []
    for k, (rate, nch, br) in enumerate(cells):
        pcm = synth_pcm(9, k, nch, 1536 * nfr, rate)
        out = encoder.encode_host(pcm[None], rate, br)
        fb, want = ora.encode_stream(pcm, rate, br)
        assert (out["status"] == 0).all() and (out["frames"][0].reshape(-1) == want).all(), (rate, nch, br)
        streams.append(out["frames"][0].reshape(-1))
    es = np.concatenate(streams)
    off, first, flen = [], [0], []
    pos = 0
    for sbytes in streams:
        o = frame_offsets(sbytes, oracle)
        n = oracle.syncinfo(sbytes[:7])[0]
        off += [pos + int(x) for x in o]
        flen += [n] * len(o)
        first.append(first[-1] + len(o))
        pos += len(sbytes)
    off, first, flen = np.array(off, np.uint64), np.array(first, np.uint32), np.array(flen)
    flags = A52_STEREO | A52_ADJUST_LEVEL
    whole = decoder.decode_host(es, off, first, flags)
    assert (whole["status"] == 0).all()
    for k, sbytes in enumerate(streams):
        nf, want = oracle.decode_stream(sbytes, flags, 1.0, 0.0)
        got = whole["pcm"][first[k]:first[k + 1]].reshape(-1, 2, 256)
        assert nf == nfr and relrms(got, want) < 1e-5, cells[k]
    costs = [len(sb) for sb in streams]
    for world in (2, 4):
        parts = shard.partition_streams(costs, world)
        assert sorted(np.concatenate(parts).tolist()) == list(range(len(cells)))
        for r in range(world):
            if len(parts[r]) == 0:
                continue
            sub, soff, sfirst = shard.shard_batch(es, off, flen, first, parts[r])
            res = decoder.decode_host(sub, soff, sfirst, flags)
            for j, sidx in enumerate(parts[r]):
                a = res["pcm"][sfirst[j]:sfirst[j + 1]]
                b = whole["pcm"][first[sidx]:first[sidx + 1]]
                assert (a.view(np.uint32) == b.view(np.uint32)).all(), (world, r, sidx)


def test_pipelined_host_batch_equals_single_launch(encoder):
    """Host-pointer calls of 1024 streams and more are cut into chunks that flow through H2D | kernel | D2H;
    the frames, statuses and carry records must not depend on that."""
    rng = np.random.RandomState(4)
    base = np.stack([synth_pcm(11, s, 2, 1536 * 3, 48000, noise=0.03) for s in range(8)])
    pick = rng.randint(0, 8, 1300)
    pcm = np.ascontiguousarray(base[pick])
    big = encoder.encode_host(pcm, 48000, 192000, carry=[None] * 1300)
    small = encoder.encode_host(base, 48000, 192000, carry=[None] * 8)
    assert (big["status"] == 0).all()
    assert (big["frames"] == small["frames"][pick]).all()
    for k in (0, 511, 650, 1299):
        a, b = big["carry"][k], small["carry"][pick[k]]
        assert bytes(a) == bytes(b)


def test_time_slices_are_invisible(engine, monkeypatch):
    """Streams longer than a slice (32 frames) are encoded as several work units linked by the carry record
    (last 256 samples per channel, SNR search warm start): byte-identical to whole-stream units, to the
    oracle, and to a caller who cuts the stream himself."""
    nch, rate, br, nfr = 2, 48000, 192000, 75
    pcm = np.stack([synth_pcm(12, s, nch, 1536 * nfr, rate, noise=0.03, bursts=(s == 1)) for s in range(5)])
    outs = []
    for sf in ("32", "7", "1000"):
        monkeypatch.setenv("AC3_B200_SLICE_FRAMES", sf)
        enc = engine.BatchEncoder(0)
        outs.append(enc.encode_host(pcm, rate, br, carry=[None] * 5))
        enc.close()
    for o in outs[1:]:
        assert (o["frames"] == outs[0]["frames"]).all() and (o["status"] == 0).all()
        assert all(bytes(a) == bytes(b) for a, b in zip(o["carry"], outs[0]["carry"]))
    ora = OracleEnc()
    fb, want = ora.encode_stream(pcm[1], rate, br)
    assert outs[0]["frames"][1].reshape(-1).tobytes() == want.tobytes()
    # the caller's own cut: 40 + 35 frames with the carry record
    monkeypatch.setenv("AC3_B200_SLICE_FRAMES", "32")
    enc = engine.BatchEncoder(0)
    a = enc.encode_host(pcm[:, :1536 * 40], rate, br, carry=[None] * 5)
    b = enc.encode_host(pcm[:, 1536 * 40:], rate, br, carry=list(a["carry"]))
    enc.close()
    assert (np.concatenate([a["frames"], b["frames"]], axis=1) == outs[0]["frames"]).all()
