"""the encoder front end the reference keeps in its ACM wrapper - WAVE -> coded channel map
(AC3ACM.cpp:1631-1662), the format validation of the stream open (:128-149, :1890-1936), the PCM gather /
carry-over buffering of stream_convert_pcm (:1665-1798) - and the `ac3enc_b200` WAV front end built on them.
CPU tests cover the pure host logic; GPU tests compare the streaming conversion with a model of the wrapper's
loop driving the encoder oracle (itself byte-identical to the reference encoder, tests/test_encoder_oracle.py)."""
import os
import struct
import subprocess

import numpy as np
import pytest

from refbind import OracleEnc
from synth import synth_pcm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENC_CLI = os.path.join(ROOT, "ac-3-acm-codec_b200", "ac3enc_b200")
DEC_CLI = os.path.join(ROOT, "ac-3-acm-codec_b200", "a52dec_b200")

# golden data: the wrapper's frame-size table (AC3ACM.cpp:128-149): 16-bit words at 32 k, 44.1 k, 48 k; kb/s
FRAMESIZES = [(96, 69, 64, 32), (120, 87, 80, 40), (144, 104, 96, 48), (168, 121, 112, 56), (192, 139, 128, 64),
              (240, 174, 160, 80), (288, 208, 192, 96), (336, 243, 224, 112), (384, 278, 256, 128),
              (480, 348, 320, 160), (576, 417, 384, 192), (672, 487, 448, 224), (768, 557, 512, 256),
              (960, 696, 640, 320), (1152, 835, 768, 384), (1344, 975, 896, 448), (1536, 1114, 1024, 512),
              (1728, 1253, 1152, 576), (1920, 1393, 1280, 640)]


def test_wav_channel_map(engine):
    # FC | FL FR | FL FR FC -> L C R | FL FR BL BR | FL FR FC BL BR -> L C R SL SR | FL FR FC LF BL BR -> L C R SL SR LFE
    want = {1: [0], 2: [0, 1], 3: [0, 2, 1], 4: [0, 1, 2, 3], 5: [0, 2, 1, 3, 4], 6: [0, 2, 1, 4, 5, 3]}
    for n, m in want.items():
        assert engine.wav_channel_map(n).tolist() == m
    for n in (0, 7, -1):
        with pytest.raises(ValueError):
            engine.wav_channel_map(n)


def test_acm_format_validation(engine):
    L = engine.load_library()
    rates = (32000, 44100, 48000)
    for row in FRAMESIZES:
        kbps = row[3]
        for ri, rate in enumerate(rates):
            assert L.ac3_acm_block_align(rate, kbps) == 2 * row[ri]
            assert L.ac3_acm_bitrate(rate, 125 * kbps) == kbps
            # integer division by 125 forgives up to 124 bytes/s above the nominal rate, nothing below
            assert L.ac3_acm_bitrate(rate, 125 * kbps + 124) == kbps
            if rate != 44100:
                assert L.ac3_acm_bitrate(rate, 125 * kbps - 1) == 0
        special = (row[1] * 2 * 44100 + 768) // 1536        # the 44.1 kHz byte rate of the unpadded frame
        assert L.ac3_acm_bitrate(44100, special) == kbps
        if special // 125 not in [r[3] for r in FRAMESIZES]:
            assert L.ac3_acm_bitrate(48000, special) == 0
        # the encoder emits exactly the wrapper's block size (44.1 kHz frames are never padded)
        for ri, rate in enumerate(rates):
            fb = L.ac3_batch_frame_bytes(rate, kbps * 1000, 2)
            assert fb in (0, 2 * row[ri])
    for bad in (0, 1000, 125 * 33, 125 * 700):
        assert L.ac3_acm_bitrate(48000, bad) == 0
    assert L.ac3_acm_block_align(22050, 192) == 0 and L.ac3_acm_block_align(48000, 100) == 0
    assert not L.ac3_stream_open(None, 48000, 24000, 2)       # no context, no stream (and no CPU fallback)


def test_ac3enc_cli_usage(engine, tmp_path):
    assert os.path.exists(ENC_CLI), "run sh ac-3-acm-codec_b200/build.sh"
    assert subprocess.run([ENC_CLI], capture_output=True, timeout=60).returncode == 1
    bad = tmp_path / "x.wav"
    bad.write_bytes(b"not a wave file at all")
    r = subprocess.run([ENC_CLI, str(bad)], capture_output=True, timeout=60)
    assert r.returncode == 1 and b"RIFF" in r.stderr
    lo = tmp_path / "lo.wav"
    lo.write_bytes(wav_bytes(np.zeros((3072, 2), np.int16), 22050))
    r = subprocess.run([ENC_CLI, str(lo)], capture_output=True, timeout=60)
    assert r.returncode == 1 and b"32000" in r.stderr          # AC3ACM.cpp:1890-1891
    ok = tmp_path / "ok.wav"
    ok.write_bytes(wav_bytes(np.zeros((3072, 2), np.int16), 48000))
    r = subprocess.run([ENC_CLI, "-b", "100", str(ok)], capture_output=True, timeout=60)
    assert r.returncode == 1 and b"bitrate" in r.stderr


def wav_bytes(pcm, rate, extensible=None, open_ended=False):
    """RIFF/WAVE of int16 [nsamples, nch]; extensible header above two channels like libao's wav6."""
    nch = pcm.shape[1]
    data = np.ascontiguousarray(pcm, "<i2").tobytes()
    ext = nch > 2 if extensible is None else extensible
    fmt = struct.pack("<HHIIHH", 0xFFFE if ext else 1, nch, rate, rate * 2 * nch, 2 * nch, 16)
    if ext:
        fmt += struct.pack("<HHI", 22, 16, {3: 7, 4: 0x33, 5: 0x37, 6: 0x3f}.get(nch, 3))
        fmt += bytes([1, 0, 0, 0, 0, 0, 0x10, 0, 0x80, 0, 0, 0xaa, 0, 0x38, 0x9b, 0x71])
    dlen = 0xFFFFFFD8 if open_ended else len(data)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 4) + b"abcd" \
        + b"data" + struct.pack("<I", dlen) + data
    return b"RIFF" + struct.pack("<I", 0xFFFFFFFC if open_ended else len(body)) + body


@pytest.fixture(scope="module")
def encoder(engine):
    enc = engine.BatchEncoder(0)
    yield enc
    enc.close()


class AcmModel:
    """stream_convert_pcm (AC3ACM.cpp:1665-1798) restated around a per-frame encoder: gather `needed` bytes,
    encode, hand out what the destination takes, carry the rest of the frame over to the next call."""

    def __init__(self, enc, rate, kbps, nch, chmap):
        self.enc, self.chmap, self.nch = enc, chmap, nch
        self.fb = enc.init(rate, kbps * 1000, nch)
        self.needed = 1536 * nch * 2
        self.buf = b""
        self.carry = b""

    def convert(self, src, dst_len, start=False):
        out = b""
        used = 0
        if start:
            self.buf, self.carry = b"", b""
        elif self.carry:
            out, self.carry = self.carry[:dst_len], self.carry[dst_len:]
            dst_len -= len(out)
            if dst_len <= 0:
                return 0, out
        while used < len(src):
            tc = min(self.needed - len(self.buf), len(src) - used)
            self.buf += src[used:used + tc]
            used += tc
            if len(self.buf) >= self.needed:
                fr = bytes(self.enc.frame(np.frombuffer(self.buf, np.int16).reshape(1536, self.nch), self.chmap))
                self.buf = b""
                out += fr[:dst_len]
                self.carry = fr[dst_len:]
                dst_len -= min(len(fr), dst_len)
                if dst_len <= 0:
                    break
        return used, out


@pytest.mark.gpu
@pytest.mark.parametrize("nch,rate,kbps", [(2, 48000, 192), (6, 48000, 448), (5, 44100, 320), (1, 32000, 64), (3, 44100, 160)])
def test_stream_convert_matches_the_acm_loop(engine, encoder, nch, rate, kbps):
    rng = np.random.RandomState(100 + nch)
    pcm = synth_pcm(8, nch, nch, 1536 * 23 + 700, rate, noise=0.05, bursts=True)
    raw = np.ascontiguousarray(pcm, "<i2").tobytes()
    chmap = engine.wav_channel_map(nch)
    avg = 125 * kbps if rate != 44100 else (dict((r[3], r[1]) for r in FRAMESIZES)[kbps] * 2 * 44100 + 768) // 1536
    st = engine.PcmToAc3Stream(encoder, rate, avg, nch)
    model = AcmModel(OracleEnc(), rate, kbps, nch, chmap)
    assert st.frame_bytes == model.fb == 2 * dict((r[3], r) for r in FRAMESIZES)[kbps][(32000, 44100, 48000).index(rate)]
    pos, start, produced, calls = 0, True, b"", 0
    launches0 = encoder.launch_count()
    while pos < len(raw):
        n = int(rng.choice([0, 1, 2, 100, 1536 * nch * 2 - 2, 5000, 20000, 100000]))
        room = int(rng.choice([0, 1, 7, 300, st.frame_bytes, st.frame_bytes + 1, 4000, 20000]))
        chunk = raw[pos:pos + n]
        u1, o1 = st.convert(chunk, room, start)
        u2, o2 = model.convert(chunk, room, start)
        assert u1 == u2 and o1 == o2, (calls, n, room, u1, u2, len(o1), len(o2))
        pos += u1
        produced += o1
        start = False
        calls += 1
    # drain the carried-over tail, then compare with the plain batch encode of the whole stream
    for _ in range(2):
        u1, o1 = st.convert(b"", 100000)
        u2, o2 = model.convert(b"", 100000)
        assert (u1, o1) == (u2, o2)
        produced += o1
        calls += 1
    whole = encoder.encode_host(pcm[None, :1536 * 23], rate, kbps * 1000, chmap=chmap)["frames"][0].reshape(-1)
    assert produced == whole.tobytes()
    # a START flag forgets buffered input and the carried frame (AC3ACM.cpp:1705-1710)
    st.convert(raw[:5000], 10, False)
    u1, o1 = st.convert(raw[:1536 * nch * 2], 100000, True)
    model.convert(raw[:5000], 10, False)
    u2, o2 = model.convert(raw[:1536 * nch * 2], 100000, True)
    assert (u1, o1) == (u2, o2) and len(o1) == st.frame_bytes
    assert encoder.launch_count() - launches0 <= calls + 4         # a call costs one launch however many frames it completes
    st.close()
    with pytest.raises(ValueError):
        engine.PcmToAc3Stream(encoder, 22050, 125 * 64, 2)
    with pytest.raises(ValueError):
        engine.PcmToAc3Stream(encoder, 48000, 125 * 100, 2)


@pytest.mark.gpu
def test_wav_files_to_ac3_and_back(engine, tmp_path):
    """ac3enc_b200 on WAV files of every channel count, as one batch; frames byte-identical to the per-frame
    encoder fed through the wrapper's channel map; a52dec_b200 -o wav6 brings the channels back in WAVE order."""
    ora = OracleEnc()
    cases = [(1, 32000, 64), (2, 48000, 192), (3, 44100, 192), (4, 48000, 192), (5, 48000, 192), (6, 48000, 192)]
    paths, pcms = [], []
    for nch, rate, kbps in cases:
        pcm = synth_pcm(9, nch, nch, 1536 * (4 + nch) + 333 * nch, rate, noise=0.01)
        p = tmp_path / ("in%d.wav" % nch)
        p.write_bytes(wav_bytes(pcm, rate, open_ended=(nch == 2)))
        paths.append(str(p))
        pcms.append(pcm)
    out = tmp_path / "out"
    out.mkdir()
    same_fmt = [i for i, c in enumerate(cases) if c[1:] == (48000, 192)]
    r = subprocess.run([ENC_CLI, "-b", "192", "-O", str(out), "-C", "3"] + [paths[i] for i in same_fmt],
                       capture_output=True, timeout=300)
    assert r.returncode == 0, r.stderr
    for i in (0, 2):
        r = subprocess.run([ENC_CLI, "-b", str(cases[i][2]), "-o", str(out / ("in%d.wav.ac3" % cases[i][0])), paths[i]],
                           capture_output=True, timeout=300)
        assert r.returncode == 0, r.stderr
    for (nch, rate, kbps), pcm in zip(cases, pcms):
        got = open(out / ("in%d.wav.ac3" % nch), "rb").read()
        nfr = pcm.shape[0] // 1536
        fb, want = ora.encode_stream(pcm[:nfr * 1536], rate, kbps * 1000, engine.wav_channel_map(nch))
        assert got == want.tobytes(), (nch, len(got), len(want))
    # round trip of the 5.1 file: same channel order out as in, codec noise only
    nch, rate, kbps = cases[5]
    r = subprocess.run([DEC_CLI, "-o", "wav6", "-a", str(out / "in6.wav.ac3")], capture_output=True, timeout=300)
    assert r.returncode == 0
    dec = np.frombuffer(r.stdout[68:], "<i2").reshape(-1, 6).astype(np.float64)
    src = pcms[5][:dec.shape[0] - 256].astype(np.float64)
    dec = dec[256:]                                                # the decoder's 256-sample latency
    for ch in (0, 1, 2, 4, 5):      # WAVE channel 3 is coded as LFE: 7 bins, not comparable for full-band input
        err = dec[:len(src), ch] - src[:, ch]
        assert np.sqrt((err ** 2).mean()) < 0.5 * np.sqrt((src[:, ch] ** 2).mean()) + 50, ch
