"""pytest configuration: the `gpu` marker, shared fixtures.

`-m "not gpu"` : oracle vs golden vectors / reference build, host logic, C-ABI export checks.
`-m gpu`       : parity tests proper - CUDA path (through the C ABI) vs the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def engine():
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ge.PKG_DIR, "liba52_b200.so")) or \
            not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        ge.build()
    return ge.load_engine()


@pytest.fixture(scope="session")
def oracle(engine):
    from refbind import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference build (oracle/_ref); present in the build container and,
    as prebuilt .so files, on the GPU box."""
    from refbind import RefA52, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return RefA52()


@pytest.fixture(scope="session")
def refenc():
    from refbind import RefAc3Enc, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return RefAc3Enc()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "decode_vectors.npz"))


@pytest.fixture(scope="session")
def c2():
    return np.load(os.path.join(ROOT, "tests", "golden", "c2_fixture.npz"))


@pytest.fixture(scope="session")
def decoder(engine):
    dec = engine.BatchDecoder(0)
    yield dec
    dec.close()


def pytest_sessionfinish(session, exitstatus):
    """Under a bounds-checked build of the library (tools/gpu_checked.sh) report what the device checks saw."""
    if not os.environ.get("A52_B200_LIB"):
        return
    try:
        import __graft_entry__ as ge
        L = ge.load_engine().load_library()
        v = L.a52_batch_violations()
        print("\n[a52_batch_violations] first violated device check: %d (0 = none, -2 = build without checks)" % v)
        if v not in (0, -2):
            session.exitstatus = 1
    except Exception as e:  # noqa
        print("\n[a52_batch_violations] unavailable: %r" % (e,))
