import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
EMU = os.path.join(ROOT, "tests", "emu")
if EMU not in sys.path:
    sys.path.insert(0, EMU)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def port():
    """our CPU restatement (the checker)."""
    from oracle import pyoracle as po
    po.build(ref=False, port=True)
    return po.Port()
