import gzip
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "verify_vectors.json.gz"), "rb") as f:
        doc = json.loads(f.read())
    for v in doc["vectors"]:
        v["root_b"] = bytes.fromhex(v["root"])
        v["key_b"] = bytes.fromhex(v["key"])
        v["proof_b"] = [bytes.fromhex(n) for n in v["proof"]]
        v["value_b"] = None if v["value"] is None else bytes.fromhex(v["value"])
        v["expect_status"] = v["status"]   # what this repository answers is the reference's verdict, always
    return doc


@pytest.fixture(scope="session")
def rebuild_golden():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "rebuild_vectors.json.gz"), "rb") as f:
        return json.loads(f.read())


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle, build
    build()
    return Oracle()


@pytest.fixture(scope="session")
def verifier():
    """The CUDA product path through the C ABI.  Fails loudly (no fallback) when unusable."""
    import zk_state_proofs_b200 as z
    return z.Verifier([0])
