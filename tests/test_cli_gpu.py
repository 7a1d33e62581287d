"""SURVEY.md §8 row f1 (+ f4): `a52dec_b200`, the a52dec command line rebuilt on the batched engine
(ac-3-acm-codec_b200/cli/a52dec_b200.c), against the reference's own CLI (`oracle/_ref/a52dec_ref` = the
unmodified src/a52dec.c + libao + liba52 built by oracle/Makefile): same options, same output drivers, same
files.  Headers and sizes must be identical; int16 samples within +-1 LSB (north_star), float PCM within 1e-5
relative RMS."""
import os
import struct
import subprocess

import numpy as np
import pytest

from bitstream_writer import make_stream

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "a52dec_ref")
CLI = os.environ.get("A52DEC_CLI", os.path.join(ROOT, "ac-3-acm-codec_b200", "a52dec_b200"))   # override: self-check of this file
need = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/a52dec_ref missing (run make -C oracle ref)")


def run(exe, args, path=None, ok=(0,)):
    r = subprocess.run([exe] + args + ([path] if path else []), capture_output=True, timeout=300)
    assert r.returncode in ok, (exe, args, r.returncode, r.stderr[-400:])
    return r.stdout, r.stderr.decode(errors="replace"), r.returncode


def write(tmp_path, name, data):
    p = tmp_path / name
    np.ascontiguousarray(np.frombuffer(bytes(data), np.uint8) if isinstance(data, (bytes, bytearray)) else data).tofile(p)
    return str(p)


def wav_split(raw):
    """(header bytes, int16 samples) of a libao RIFF file: 44-byte PCM header or 68-byte extensible one."""
    raw = np.frombuffer(raw, np.uint8)
    n = 44 if raw[20] == 1 else 68
    return raw[:n], raw[n:].view(np.int16).astype(int)


def same_s16(a, b, what, natural=False):
    ha, sa = a
    hb, sb = b
    assert len(ha) == len(hb) and (ha == hb).all(), (what, "header", bytes(ha), bytes(hb))
    assert len(sa) == len(sb) and len(sa) > 0, (what, len(sa), len(sb))
    assert np.abs(sa - sb).max() <= 1, (what, np.abs(sa - sb).max())
    if natural:
        assert (sa == sb).mean() > 0.995, what


def same_f32(a, b, what):
    fa, fb = np.frombuffer(a, np.float32).astype(np.float64), np.frombuffer(b, np.float32).astype(np.float64)
    assert len(fa) == len(fb) and len(fa) > 0, what
    d = fa - fb
    assert np.sqrt((d * d).mean()) <= 1e-5 * np.sqrt((fa * fa).mean()), what


def ref_file(tmp_path, args, path):
    """The reference CLI with its stdout in a FILE: libao then rewinds and writes the final sizes."""
    out = tmp_path / ("ref_%d.out" % abs(hash((tuple(args), path))))
    with open(out, "wb") as f:
        r = subprocess.run([REF] + args + [path], stdout=f, stderr=subprocess.PIPE, timeout=300)
    assert r.returncode == 0, r.stderr[-300:]
    return open(out, "rb").read()


def cli_batch(tmp_path, drv, paths, extra=()):
    """ONE run of a52dec_b200 for all paths (-O: every input is a stream of the same batch)."""
    out = tmp_path / ("cli_%s_%d" % (drv, abs(hash(tuple(paths) + tuple(extra)))))
    out.mkdir()
    run(CLI, ["-o", drv, "-O", str(out)] + list(extra) + list(paths))
    ext = {"wav": "wav", "wavdolby": "wav", "wav6": "wav", "aif": "aif", "aifdolby": "aif", "peak": "txt",
           "peakdolby": "txt"}.get(drv, "raw")
    return [open(out / (os.path.basename(p) + "." + ext), "rb").read() for p in paths]


def same_aif(a, b, what):
    assert a[:54] == b[:54], (what, a[:54], b[:54])
    sa, sb = np.frombuffer(a[54:], ">i2").astype(int), np.frombuffer(b[54:], ">i2").astype(int)
    assert len(sa) == len(sb) > 0 and np.abs(sa - sb).max() <= 1, what


@need
def test_every_driver_on_natural_streams(tmp_path, c2, golden):
    cases = [("c2", c2["frames"][1, :24].reshape(-1)), ("stereo", np.tile(golden["enc20_stereo_bias.es"], 3)),
             ("441", golden["enc50_dolby_441.es"])]
    paths = [write(tmp_path, name + ".ac3", es) for name, es in cases]
    # every driver, the three files as one batch per driver; complete files (final headers) are compared
    for drv in ("wav", "wavdolby", "wav6"):
        for p, got in zip(paths, cli_batch(tmp_path, drv, paths)):
            same_s16(wav_split(ref_file(tmp_path, ["-o", drv], p)), wav_split(got), (p, drv), True)
    for drv in ("aif", "aifdolby"):
        for p, got in zip(paths, cli_batch(tmp_path, drv, paths)):
            same_aif(ref_file(tmp_path, ["-o", drv], p), got, (p, drv))
    for p, got in zip(paths, cli_batch(tmp_path, "float", paths)):
        same_f32(ref_file(tmp_path, ["-o", "float"], p), got, (p, "float"))
    for drv in ("peak", "peakdolby"):
        for p, got in zip(paths, cli_batch(tmp_path, drv, paths)):
            a, b = ref_file(tmp_path, ["-o", drv], p).decode(), got.decode()
            va, vb = float(a.split()[3]), float(b.split()[3])
            assert a.startswith("peak level = ") and abs(va - vb) <= 1e-4 + 1e-5 * va, (a, b)
    for drv in ("null", "null4", "null6"):
        assert cli_batch(tmp_path, drv, paths) == [b"", b"", b""]
        assert ref_file(tmp_path, ["-o", drv], paths[0]) == b""
    # one input without -O: stdout; a pipe cannot be rewound, the open-ended sizes stay (audio_out_wav.c:44-58)
    p = paths[1]
    ra, rb = run(REF, ["-o", "wav"], p)[0], run(CLI, ["-o", "wav"], p)[0]
    same_s16(wav_split(ra), wav_split(rb), "pipe", True)
    assert struct.unpack("<I", rb[4:8])[0] == 0xfffffffc
    # default driver is the first of the list (wav), stdin works as input
    es = open(p, "rb").read()
    r = subprocess.run([CLI], input=es, capture_output=True, timeout=300, check=True).stdout
    same_s16(wav_split(ra), wav_split(r), "stdin", True)


@need
def test_options_r_a_g(tmp_path, golden):
    p = write(tmp_path, "f.ac3", golden["syn51_stereo.es"])
    for opts in (["-r"], ["-a"], ["-g", "-6"], ["-g", "3.5", "-r", "-a"]):
        same_f32(run(REF, opts + ["-o", "float"], p)[0], run(CLI, opts + ["-o", "float"], p)[0], opts)
    for opts in (["-g", "-2.5", "-a"], ["-c", "-r"]):
        same_s16(wav_split(run(REF, opts + ["-o", "wav"], p)[0]), wav_split(run(CLI, opts + ["-o", "wav"], p)[0]), opts)
    for bad in (["-g", "97"], ["-o", "nosuch"], ["-t", "5"], ["-s", "9"]):
        assert run(CLI, bad, p, ok=(1,))[2] == run(REF, bad, p, ok=(1,))[2] == 1


MODES = [(0, 0), (1, 0), (1, 1), (2, 1), (3, 0), (3, 1), (4, 0), (4, 1), (5, 0), (5, 1), (6, 0), (6, 1), (7, 0), (7, 1)]


@need
def test_wav6_every_coded_mode(tmp_path, oracle):
    """wav6 leaves the request to the stream: every coded mode comes out in WAV channel order, extensible
    header with the right speaker mask - including libao's 2/1+LFE fall-through (convert2s16.c:270-285).
    All fourteen modes are streams of ONE batch per driver."""
    paths = []
    for acmod, lfe in MODES:
        es, fb = make_stream(4100 + acmod * 2 + lfe, acmod, lfe, 3, oracle.bit_allocate, frmsizecod=30)
        paths.append(write(tmp_path, "m%d%d.ac3" % (acmod, lfe), es))
    for (acmod, lfe), p, got in zip(MODES, paths, cli_batch(tmp_path, "wav6", paths)):
        b = wav_split(got)
        same_s16(wav_split(ref_file(tmp_path, ["-o", "wav6"], p)), b, (acmod, lfe))
        if (acmod, lfe) == (4, 1):
            assert (b[1].reshape(-1, 1024)[:, 4::5] == -32768).all()
    # the stereo drivers on the same streams (mono sources stay mono: one-channel header; aif and float take
    # two planes whatever the grant)
    for drv in ("wav", "wavdolby"):
        for m, p, got in zip(MODES, paths, cli_batch(tmp_path, drv, paths)):
            same_s16(wav_split(ref_file(tmp_path, ["-o", drv], p)), wav_split(got), (m, drv))
    for m, p, got in zip(MODES, paths, cli_batch(tmp_path, "aif", paths)):
        same_aif(ref_file(tmp_path, ["-o", "aif"], p), got, (m, "aif"))
    for m, p, got in zip(MODES, paths, cli_batch(tmp_path, "float", paths, extra=["-C", "2"])):
        same_f32(ref_file(tmp_path, ["-o", "float"], p), got, (m, "float"))


@need
def test_resync_damage_and_mode_changes(tmp_path, oracle, golden):
    rng = np.random.RandomState(5)
    es0, fb = make_stream(91, 7, 1, 6, oracle.bit_allocate)
    junk = lambda n: rng.randint(0, 256, n).astype(np.uint8)
    # garbage before, between and after frames; a truncated last frame
    es = np.concatenate([junk(301), es0[:2 * fb], junk(17), es0[2 * fb:5 * fb], junk(5), es0[5 * fb:6 * fb - 100]])
    p = write(tmp_path, "junk.ac3", es)
    ra, rb = run(REF, ["-o", "wav"], p), run(CLI, ["-o", "wav"], p)
    same_s16(wav_split(ra[0]), wav_split(rb[0]), "resync")
    assert ra[1].count("skip\n") == rb[1].count("skip\n") > 300
    # bit errors: frames abandoned in the same block, same number of blocks written
    for it in range(6):
        es = es0.copy()
        for _ in range(3):
            es[fb + int(rng.randint(6, fb - 1))] ^= 1 << int(rng.randint(8))
        p = write(tmp_path, "bad%d.ac3" % it, es)
        ra, rb = run(REF, ["-o", "wav"], p), run(CLI, ["-o", "wav"], p)
        ha, sa = wav_split(ra[0])
        hb, sb = wav_split(rb[0])
        assert (ha == hb).all() and len(sa) == len(sb), it
        assert ra[1].count("error\n") == rb[1].count("error\n"), it
        assert np.abs(sa - sb).max() <= 1, it
    # a sample-rate change after the header went out: libao refuses those frames (audio_out_wav.c:63-66)
    es = np.concatenate([golden["enc20_stereo_bias.es"], golden["enc50_dolby_441.es"], golden["enc20_stereo_bias.es"]])
    p = write(tmp_path, "rates.ac3", es)
    for drv in ("wav", "aif"):
        ra, rb = run(REF, ["-o", drv], p), run(CLI, ["-o", drv], p)
        assert len(ra[0]) == len(rb[0]) and ra[1].count("error\n") == rb[1].count("error\n") > 0
        assert ra[0][:44] == rb[0][:44]
    # a channel-layout change under wav6: frames of the other layout are refused by play()
    a, _ = make_stream(7, 2, 0, 2, oracle.bit_allocate, frmsizecod=30)
    b, _ = make_stream(8, 7, 1, 2, oracle.bit_allocate, frmsizecod=30)
    p = write(tmp_path, "layout.ac3", np.concatenate([a, b]))
    ra, rb = run(REF, ["-o", "wav6"], p), run(CLI, ["-o", "wav6"], p)
    assert len(ra[0]) == len(rb[0]) and ra[1].count("error\n") == rb[1].count("error\n") == 2
    same_s16(wav_split(ra[0]), wav_split(rb[0]), "layout")


def pes(stream_id, payload, mpeg2=True, stuffing=0, pts=True, sub=None):
    """One PES packet; `sub` = DVD substream header byte (private stream 1 in a program stream)."""
    body = bytearray()
    if sub is not None:
        body += bytes([sub, 1, 0, 1])
    body += bytes(payload)
    if mpeg2:
        opt = bytes([0x21, 0, 1, 0, 1]) if pts else b""
        opt += b"\xff" * stuffing
        head = bytes([0x81, 0x80 if pts else 0, len(opt)]) + opt
    else:
        head = b"\xff" * stuffing + bytes([0x40, 0x20]) + (bytes([0x21, 0, 1, 0, 1]) if pts else b"\x0f")
    return b"\x00\x00\x01" + bytes([stream_id]) + struct.pack(">H", len(head) + len(body)) + head + bytes(body)


def pack_header(mpeg2=True, stuffing=0):
    if mpeg2:
        return b"\x00\x00\x01\xba" + bytes([0x44, 0, 4, 0, 4, 1, 1, 0x89, 0xc3, 0xf8 | stuffing]) + b"\xff" * stuffing
    return b"\x00\x00\x01\xba" + bytes([0x21, 0, 1, 0, 1, 0x80, 0x27, 0x11])


@need
def test_program_stream_and_pes_demux(tmp_path, golden, c2):
    rng = np.random.RandomState(3)
    es = bytes(c2["frames"][2, :8].reshape(-1))
    other = bytes(np.tile(golden["enc20_stereo_bias.es"], 2))
    for mpeg2 in (True, False):
        ps = bytearray()
        i = j = 0
        while i < len(es):
            n = int(rng.randint(200, 2000))
            ps += pack_header(mpeg2, int(rng.randint(0, 4)))
            ps += pes(0xbd, es[i:i + n], mpeg2, int(rng.randint(0, 6)), bool(rng.randint(2)), sub=0x80)
            i += n
            if j < len(other):                                       # a second audio track, video, padding
                ps += pes(0xbd, other[j:j + 700], mpeg2, 0, True, sub=0x82)
                j += 700
            ps += pes(0xe0, rng.bytes(300), mpeg2) + pes(0xbe, b"\xff" * 50, mpeg2, pts=False) + b"\x00" * int(rng.randint(0, 5))
        ps += b"\x00\x00\x01\xb9" + b"trailing bytes after the program end code"
        p = write(tmp_path, "a.vob", ps)
        for opts in (["-s"], ["-s0x82", "-o", "float"]) if mpeg2 else (["-s", "-o", "wav6"], ["-s2"]):
            ra, rb = run(REF, opts, p), run(CLI, opts, p)
            if "float" in opts:
                same_f32(ra[0], rb[0], (mpeg2, opts))
            else:
                same_s16(wav_split(ra[0]), wav_split(rb[0]), (mpeg2, opts), True)
    # -T: bare MPEG-2 PES packets, no substream header
    pp = bytearray()
    i = 0
    while i < len(es):
        n = int(rng.randint(100, 3000))
        pp += pes(0xbd, es[i:i + n], True, int(rng.randint(0, 4)), bool(rng.randint(2))) + b"\x00" * int(rng.randint(0, 3))
        i += n
    p = write(tmp_path, "a.pes", pp)
    same_s16(wav_split(run(REF, ["-T"], p)[0]), wav_split(run(CLI, ["-T"], p)[0]), "pes", True)
    # the reference's fatal cases end the run with exit status 1 after the audio decoded so far
    p = write(tmp_path, "b.pes", bytes(pp) + pes(0xc0, b"x" * 40))
    ra, rb = run(REF, ["-T"], p, ok=(1,)), run(CLI, ["-T"], p, ok=(1,))
    assert "bad stream id" in ra[1] and "bad stream id" in rb[1]
    p = write(tmp_path, "video.m2v", b"\x00\x00\x01\xb3" + b"\x00" * 64)
    assert run(REF, ["-s"], p, ok=(1,))[2] == run(CLI, ["-s"], p, ok=(1,))[2] == 1


def ts_packets(pid, pes_bytes, rng, cc=0):
    """188-byte packets of one PES packet: payload-unit-start on the first, adaptation stuffing at random."""
    out = bytearray()
    i, first = 0, True
    while i < len(pes_bytes):
        room = 184
        stuff = int(rng.randint(0, 60)) if rng.rand() < 0.3 else 0
        left = len(pes_bytes) - i
        if left < room:
            stuff = max(stuff, room - left)
        take = min(left, room - stuff)
        stuff = room - take
        hdr = bytes([0x47, (0x40 if first else 0) | (pid >> 8), pid & 0xff, (0x30 if stuff else 0x10) | (cc & 15)])
        af = b""
        if stuff:
            af = bytes([stuff - 1]) + (bytes([0]) + b"\xff" * (stuff - 2) if stuff > 1 else b"")
        out += hdr + af + pes_bytes[i:i + take]
        i += take
        cc += 1
        first = False
    return bytes(out), cc


@need
def test_transport_stream_demux(tmp_path, c2):
    rng = np.random.RandomState(9)
    es = bytes(c2["frames"][3, :8].reshape(-1))
    ts = bytearray()
    i, cc = 0, 0
    while i < len(es):
        n = int(rng.randint(150, 4000))
        pk, cc = ts_packets(0x123, pes(0xbd, es[i:i + n], True, int(rng.randint(0, 5)), bool(rng.randint(2))), rng, cc)
        ts += pk
        i += n
        # other pids in between, and a packet of the audio pid that carries no payload
        ts += bytes([0x47, 0x01, 0x00, 0x10]) + rng.bytes(184)
        ts += bytes([0x47, 0x01, 0x23, 0x20, 183, 0]) + b"\xff" * 182
    p = write(tmp_path, "a.ts", b"\x11\x22" + bytes(ts))          # two stray bytes: "bad sync byte" resync
    ra, rb = run(REF, ["-t", "0x123"], p), run(CLI, ["-t", "0x123"], p)
    same_s16(wav_split(ra[0]), wav_split(rb[0]), "ts", True)
    assert ra[1].count("bad sync byte") == rb[1].count("bad sync byte") == 2
    same_f32(run(REF, ["-t", "291", "-o", "float"], p)[0], run(CLI, ["-t", "291", "-o", "float"], p)[0], "ts float")


@need
def test_batch_of_files_and_chunked_carry(tmp_path, c2, golden, oracle):
    """-O: several files = one batch; -C: frames per engine call (the carry record links the calls)."""
    srcs = {"a.ac3": c2["frames"][0, :20].reshape(-1), "b.ac3": np.tile(golden["enc20_stereo_bias.es"], 2),
            "c.ac3": make_stream(5, 7, 1, 7, oracle.bit_allocate, features=dict(blksw=0.5, cpl=0.9, dynrng=0.5))[0]}
    paths = [write(tmp_path, k, v) for k, v in srcs.items()]
    single = {os.path.basename(p): run(CLI, ["-o", "wav"], p)[0] for p in paths}
    for extra in ([], ["-C", "3"]):
        out = tmp_path / ("out%d" % len(extra))
        out.mkdir()
        run(CLI, ["-o", "wav", "-O", str(out)] + extra + paths)
        for k in srcs:
            # a file can be rewound: final sizes in the header (audio_out_wav.c:158-166); a pipe keeps the
            # open-ended ones
            got = open(out / (k + ".wav"), "rb").read()
            assert got[44:] == single[k][44:] and got[8:40] == single[k][8:40], (k, extra)
            assert struct.unpack("<I", got[4:8])[0] == len(got) - 8 and struct.unpack("<I", got[40:44])[0] == len(got) - 44
            assert struct.unpack("<I", single[k][4:8])[0] == 0xfffffffc
    for k, p in zip(srcs, paths):
        same_s16(wav_split(run(REF, ["-o", "wav"], p)[0]), wav_split(single[k]), k)
    with open(tmp_path / "ref.wav", "wb") as f:
        subprocess.run([REF, "-o", "wav6", paths[0]], stdout=f, stderr=subprocess.DEVNULL, check=True, timeout=300)
    with open(tmp_path / "cli.wav", "wb") as f:
        subprocess.run([CLI, "-o", "wav6", paths[0]], stdout=f, stderr=subprocess.DEVNULL, check=True, timeout=300)
    a, b = open(tmp_path / "ref.wav", "rb").read(), open(tmp_path / "cli.wav", "rb").read()
    assert a[:68] == b[:68] and struct.unpack("<I", b[64:68])[0] == len(b) - 68
    same_s16(wav_split(a), wav_split(b), "wav6 file", True)
