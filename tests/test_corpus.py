"""The bench corpus generator (tests/corpus.py): numpy and torch synthesise the same samples bit for bit, and the
reference encoder turns them into the committed bytes (tests/golden/c2_corpus.json)."""
import os

import numpy as np
import pytest

import corpus

HAVE_REF = os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ac3enc_ref.so"))


def test_numpy_and_torch_synthesis_agree():
    a = corpus.synth_numpy([0, 7, 255], 1536 * 6)
    b = corpus.synth_torch([0, 7, 255], "cpu", 1536 * 6).numpy()
    assert a.dtype == np.int16 and a.shape == (3, 1536 * 6, 6)
    assert (a == b).all()
    # the recipe of SURVEY.md 8(d): three sines of amplitude 0.1 .. 0.3 + noise, scaled by 0.9
    rms = np.sqrt((a.astype(np.float64) ** 2).mean(axis=1)) / 32767.0
    assert (rms > 0.1).all() and (rms < 0.45).all()
    assert np.abs(a).max() <= 32767


@pytest.mark.skipif(not HAVE_REF, reason="reference encoder not built (make -C oracle ref)")
def test_reference_encoder_reproduces_committed_streams():
    want = corpus.load_digest()["sha256_stream"]
    got = corpus.encode_cpu([0, 1, 255], procs=3)
    assert got.shape == (3, 313, 1792)
    for i, k in enumerate((0, 1, 255)):
        assert corpus.digest_of(got[i]) == want[str(k)]


def test_tiling_covers_every_unique_stream():
    base, idx = corpus.tile_index(4096, 313)
    assert set(base.tolist()) == set(range(256))
    assert (idx[0] == np.arange(313)).all() and idx[256][0] == 7 and idx.max() == 312
