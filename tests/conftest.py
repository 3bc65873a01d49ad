import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "stable-diffusion.mojo_b200"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def relerr(a, b):
    """max |a-b| / max |b| : the tolerance metric used by every parity test."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.fixture(scope="session")
def golden_small():
    return np.load(os.path.join(GOLDEN, "small.npz"))


@pytest.fixture(scope="session")
def oracle_lib():
    """Builds oracle/_build/libref_loops.so on demand (test infrastructure)."""
    import subprocess
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    import tsd_oracle
    return tsd_oracle.clib()


@pytest.fixture(scope="session")
def ctx():
    """One tsd_ctx on cuda:0 for the GPU tests.  Fails loudly if the CUDA library is missing."""
    from tsd_b200.api import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def unet_weights():
    import synth
    return synth.SynthWeights(synth.diffusion_specs(), 1234)


@pytest.fixture(scope="session")
def decoder_weights():
    import synth
    return synth.SynthWeights(synth.decoder_specs(), 1235)
