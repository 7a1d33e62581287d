"""CPU-only checks of the drop-in boundary: the shared library loads without a GPU, exports every
symbol include/*.h declares and nothing else (reference: test/globals:15-22 - every exported symbol
of liba52 must match ^a52_), the host-side pure functions match the oracle, and the compute entry
points fail loudly (no CPU fallback) when no CUDA device exists.
"""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for h in ("a52.h", "a52_batch.h", "ac3enc.h", "ac3enc_batch.h"):
        p = os.path.join(ROOT, "include", h)
        if not os.path.exists(p):
            continue
        src = open(p).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b((?:a52|AC3|ac3)_[A-Za-z0-9_]+)\s*\(", src))
    return names


def test_library_loads_and_exports_declared_symbols(engine):
    L = engine.load_library()
    decl = declared_functions()
    assert {"a52_init", "a52_samples", "a52_syncinfo", "a52_frame", "a52_dynrng", "a52_block", "a52_free",
            "a52_batch_decode"} <= decl
    for name in sorted(decl):
        assert hasattr(L, name), "declared in include/ but not exported: " + name
    assert set(engine.EXPORTS) <= decl


def test_only_api_symbols_exported(engine):
    out = subprocess.check_output(["nm", "-D", "--defined-only", engine.LIB_PATH]).decode()
    syms = [ln.split()[-1] for ln in out.splitlines() if ln.strip()]
    bad = [s for s in syms if not re.match(r"^(a52_|AC3_|ac3_batch_|ac3_stream_|ac3_wav_|ac3_acm_)", s)]
    assert not bad, bad


def test_syncinfo_matches_oracle(engine, oracle):
    L = engine.load_library()
    rng = np.random.RandomState(5)
    fl, sr, br = C.c_int(0), C.c_int(0), C.c_int(0)
    for it in range(4000):
        h = rng.randint(0, 256, 7).astype(np.uint8)
        if it % 8:
            h[0], h[1] = 0x0B, 0x77
        n = L.a52_syncinfo(h.ctypes.data, C.byref(fl), C.byref(sr), C.byref(br))
        o = oracle.syncinfo(h)
        assert n == o[0]
        if n:
            assert (fl.value, sr.value, br.value) == o[1:]


def test_frame_indexer_resync(engine, oracle, c2):
    from util import frame_offsets
    fr = c2["frames"][0, :6].reshape(-1)
    rng = np.random.RandomState(2)
    junk = rng.randint(0, 256, 333).astype(np.uint8)
    junk[junk == 0x0B] = 0
    es = np.concatenate([junk[:100], fr[:1792 * 2], junk[100:], fr[1792 * 2:], junk[:50]])
    got = engine.index_frames(es)
    want = frame_offsets(es, oracle)
    assert len(got) == 6 and (got == want).all()
    assert len(engine.index_frames(np.zeros(0, np.uint8))) == 0
    assert len(engine.index_frames(fr[:1791])) == 0          # truncated single frame


def test_encoder_frame_bytes(engine):
    L = engine.load_library()
    # AC3_encode_init's acceptance rules (ac3enc.cpp:1019-1077), host side only
    assert L.ac3_batch_frame_bytes(48000, 448000, 6) == 1792 and L.ac3_batch_frame_bytes(48000, 192000, 2) == 768
    assert L.ac3_batch_frame_bytes(44100, 320000, 5) == 2 * (320000 * 1536 // (44100 * 16))
    assert L.ac3_batch_frame_bytes(24000, 64000, 2) == 512
    for bad in ((48000, 448000, 7), (48000, 448000, 0), (47999, 448000, 2), (48000, 449000, 2)):
        assert L.ac3_batch_frame_bytes(*bad) == 0


def test_frame_stride(engine):
    L = engine.load_library()
    for flags, nout in [(2, 2), (7 | 16, 6), (1, 1), (10 | 32, 2), (6 | 16, 5)]:
        assert L.a52_batch_frame_stride(flags, engine.PCM_F32_PLANAR) == 1536 * nout * 4
        assert L.a52_batch_frame_stride(flags, engine.PCM_S16_INTERLEAVED) == 1536 * nout * 2


def test_no_cpu_fallback_without_gpu(engine):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = engine.load_library()
    assert not L.a52_init(0)                       # NULL: no device, no decode
    assert not L.a52_batch_create(0)
    assert not L.ac3_batch_create(0)
    assert L.AC3_encode_init(48000, 448000, 6) == 0
    with pytest.raises(RuntimeError):
        engine.BatchEncoder(0)
    with pytest.raises(RuntimeError):
        engine.BatchDecoder(0)


def test_a52dec_cli_usage_and_no_cpu_fallback(engine, tmp_path, golden):
    """The command line tool lists libao's driver names in libao's order
    (audio_out.c:55-92, without the sound-card ones), rejects what a52dec.c:155-238 rejects, and - like
    the library - has no CPU path to fall back to."""
    exe = os.path.join(ROOT, "ac-3-acm-codec_b200", "a52dec_b200")
    assert os.path.exists(exe), "run sh ac-3-acm-codec_b200/build.sh"
    r = subprocess.run([exe, "-h"], capture_output=True, timeout=60)
    assert r.returncode == 1
    names = [ln.strip() for ln in r.stderr.decode().splitlines() if ln.startswith("\t\t\t")]
    assert names == ["wav", "wavdolby", "wav6", "aif", "aifdolby", "peak", "peakdolby", "null", "null4", "null6", "float"]
    for bad in (["-g", "97"], ["-o", "nosuch"], ["-t", "5"], ["-s", "9"], ["-C", "0"]):
        assert subprocess.run([exe] + bad, capture_output=True, timeout=60, stdin=subprocess.DEVNULL).returncode == 1
    import torch
    if not torch.cuda.is_available():
        p = tmp_path / "s.ac3"
        golden["enc20_stereo_bias.es"].tofile(p)
        r = subprocess.run([exe, "-o", "float", str(p)], capture_output=True, timeout=120)
        assert r.returncode == 1 and r.stdout == b"" and b"init failed" in r.stderr
