import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_seq(name: str) -> bytes:
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def seqs():
    return {
        "H5_HA": golden_seq("NC_007362.1.txt"),
        "H1_HA": golden_seq("NC_026433.1.txt"),
        "CY137594": golden_seq("CY137594.txt"),
        "KJ907631": golden_seq("KJ907631.1.txt"),
        "KJ907623": golden_seq("KJ907623.1.txt"),
    }
