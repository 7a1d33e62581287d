"""Drop-in at the link level: the reference's own command-line driver (src/a52dec.c + libao, compiled
UNMODIFIED by oracle/Makefile) linked against liba52_b200.so decodes through the GPU; its output is compared
with the same driver linked against the reference liba52 (BASELINE.json configs[0]: "decoded by a52dec CLI
(-o wav) on CPU, single stream")."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "a52dec_ref")
B200 = os.path.join(ROOT, "oracle", "_ref", "a52dec_b200")


def run(exe, args, path):
    return subprocess.run([exe] + args + [path], capture_output=True, timeout=300, check=True).stdout


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(B200)), reason="oracle/_ref CLI builds missing")
def test_a52dec_cli_wav_and_float(tmp_path, c2, golden):
    cases = [("c2", c2["frames"][1, :24].reshape(-1)), ("stereo", np.tile(golden["enc20_stereo_bias.es"], 3)),
             ("feature51", golden["syn51_stereo.es"]), ("441", golden["enc50_dolby_441.es"])]
    for name, es in cases:
        p = tmp_path / (name + ".ac3")
        np.ascontiguousarray(es).tofile(p)
        # -o wav: stereo int16 through libao's convert2s16 (bias 384): +-1 LSB
        a = np.frombuffer(run(REF, ["-o", "wav"], str(p)), np.uint8)
        b = np.frombuffer(run(B200, ["-o", "wav"], str(p)), np.uint8)
        assert len(a) == len(b) and len(a) > 44, name
        assert (a[:44] == b[:44]).all(), name                            # identical RIFF header
        sa, sb = a[44:].view(np.int16).astype(int), b[44:].view(np.int16).astype(int)
        assert np.abs(sa - sb).max() <= 1, (name, np.abs(sa - sb).max())
        if name in ("c2", "stereo"):                                      # natural signals: almost all samples identical
            assert (sa == sb).mean() > 0.995, name
        # -o float: raw float stereo, the reference's own regression format (test/compare.c thresholds)
        fa = np.frombuffer(run(REF, ["-o", "float"], str(p)), np.float32).astype(np.float64)
        fb = np.frombuffer(run(B200, ["-o", "float"], str(p)), np.float32).astype(np.float64)
        assert len(fa) == len(fb)
        d = fa - fb
        assert np.sqrt((d * d).mean()) / np.sqrt((fa * fa).mean()) < 1e-5, name
        if name in ("c2", "stereo"):                  # test/compare.c:66-72: max abs difference * 32768 below 0.01 .. 0.05
            assert np.abs(d).max() * 32768 < 0.05, name
    # -r (a52_dynrng (state, NULL, NULL), a52dec.c:290-291) and 5.1 output (wav6 is not in this libao: null6 only)
    p = tmp_path / "feature51.ac3"
    fa = np.frombuffer(run(REF, ["-r", "-o", "float"], str(p)), np.float32).astype(np.float64)
    fb = np.frombuffer(run(B200, ["-r", "-o", "float"], str(p)), np.float32).astype(np.float64)
    d = fa - fb
    assert len(fa) == len(fb) and np.sqrt((d * d).mean()) / np.sqrt((fa * fa).mean()) < 1e-5
