import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
FEATURES = os.path.join(GOLDEN, "refdata", "Features.txt")
RANGE = os.path.join(GOLDEN, "refdata", "range21062012_allfeatures")
REFERENCE_DATA = "/root/reference/data"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the in-place compiled reference (oracle/_ref)")


@pytest.fixture(scope="session")
def tmp_models(tmp_path_factory):
    """Substitute libsvm models generated on the fly (deterministic), keyed by n_sv."""
    from haf_grasping_b200 import synth

    d = tmp_path_factory.mktemp("models")
    cache = {}

    def get(n_sv=256, seed=7, labels=(1, -1), rho=0.0):
        key = (n_sv, seed, labels, rho)
        if key not in cache:
            p = str(d / ("synth_%d_%d_%d_%d_%g.model" % (n_sv, seed, labels[0], labels[1], rho)))
            synth.write_synth_model(p, n_sv=n_sv, seed=seed, labels=labels, rho=rho)
            cache[key] = p
        return cache[key]

    return get


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import orc

    orc.build(ref=os.path.isdir("/root/reference"))
    return orc
