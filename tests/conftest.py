import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def repo_root():
    return ROOT


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "assembly_ref.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def ref_mats():
    """Reference plant QP matrices from the oracle assembly (validated against the golden)."""
    import oracle
    cfg = oracle.load_config(os.path.join(ROOT, "config", "MPC_API.json"))
    return oracle.mpc_build(**cfg), cfg
