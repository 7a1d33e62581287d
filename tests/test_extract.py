"""SURVEY.md §8 row f4, second half: `a52dec_b200 -x` = the reference's extract_a52 tool (src/extract_a52.c): the
program-stream / PES / transport-stream demultiplexers writing the AC-3 elementary stream instead of decoding
it.  Pure host code (no GPU is touched), so this runs in the CPU suite against `oracle/_ref/extract_a52_ref`,
the reference tool built unmodified by oracle/Makefile, on multiplexes written by the muxers of
tests/test_cli_gpu.py."""
import os
import subprocess

import numpy as np
import pytest

from test_cli_gpu import pes, pack_header, ts_packets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "extract_a52_ref")
CLI = os.path.join(ROOT, "ac-3-acm-codec_b200", "a52dec_b200")
pytestmark = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(CLI)),
                                reason="oracle/_ref/extract_a52_ref or the tool missing (make -C oracle ref; build.sh)")


def both(args, path, ok=(0,)):
    a = subprocess.run([REF] + args + [path], capture_output=True, timeout=120)
    b = subprocess.run([CLI, "-x"] + args + [path], capture_output=True, timeout=120)
    assert a.returncode in ok and b.returncode == a.returncode, (args, a.returncode, b.returncode, b.stderr[-200:])
    return a.stdout, b.stdout


def test_extract_program_stream_pes_and_ts(tmp_path, golden, c2):
    rng = np.random.RandomState(21)
    es = bytes(c2["frames"][1, :6].reshape(-1))
    other = bytes(np.tile(golden["enc20_stereo_bias.es"], 2))
    for mpeg2 in (True, False):
        ps = bytearray()
        i = j = 0
        while i < len(es):
            n = int(rng.randint(100, 2500))
            ps += pack_header(mpeg2, int(rng.randint(0, 4)))
            ps += pes(0xbd, es[i:i + n], mpeg2, int(rng.randint(0, 6)), bool(rng.randint(2)), sub=0x80)
            i += n
            if j < len(other):
                ps += pes(0xbd, other[j:j + 900], mpeg2, 0, True, sub=0x83)
                j += 900
            ps += pes(0xe0, rng.bytes(200), mpeg2) + pes(0xbe, b"\xff" * 30, mpeg2, pts=False) + b"\x00" * int(rng.randint(0, 4))
        ps += b"\x00\x00\x01\xb9" + b"after the end code"
        p = tmp_path / "a.vob"
        p.write_bytes(bytes(ps))
        for args, want in (([], es), (["-s0"], es), (["-s0x83"], other[:j]), (["-s5"], b"")):
            a, b = both(args, str(p))
            assert a == b == want, (mpeg2, args, len(a), len(b), len(want))
    # -T: bare MPEG-2 PES packets
    pp = bytearray()
    i = 0
    while i < len(es):
        n = int(rng.randint(50, 3000))
        pp += pes(0xbd, es[i:i + n], True, int(rng.randint(0, 4)), bool(rng.randint(2))) + b"\x00" * int(rng.randint(0, 3))
        i += n
    p = tmp_path / "a.pes"
    p.write_bytes(bytes(pp))
    a, b = both(["-T"], str(p))
    assert a == b == es
    # -t: transport stream, headers straddling packets, adaptation fields, other pids, stray bytes
    ts = bytearray()
    i, cc = 0, 0
    while i < len(es):
        n = int(rng.randint(150, 4000))
        pk, cc = ts_packets(0x44, pes(0xbd, es[i:i + n], True, int(rng.randint(0, 5)), bool(rng.randint(2))), rng, cc)
        ts += pk
        i += n
        ts += bytes([0x47, 0x01, 0x00, 0x10]) + rng.bytes(184)
    p = tmp_path / "a.ts"
    p.write_bytes(b"\x55" + bytes(ts))
    a, b = both(["-t", "0x44"], str(p))
    assert a == b == es
    a, b = both(["-t", "0x45"], str(p))
    assert a == b == b""
    # the reference's fatal cases: same exit status, same bytes written before it
    p = tmp_path / "b.pes"
    p.write_bytes(bytes(pp) + pes(0xc0, b"x" * 40))
    a, b = both(["-T"], str(p), ok=(1,))
    assert a == b == es
    p = tmp_path / "v.m2v"
    p.write_bytes(b"\x00\x00\x01\xb3" + b"\x00" * 64)
    a, b = both([], str(p), ok=(1,))
    assert a == b == b""
