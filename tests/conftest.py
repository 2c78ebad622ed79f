import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running CPU check")


@pytest.fixture(scope="session")
def vectors():
    """The reference's own known-answer data (tests/golden/make_golden.py)."""
    return json.loads((ROOT / "tests/golden/reference_vectors.json").read_text())


def sq(name: str) -> int:
    return (ord(name[0]) - ord("a")) + 8 * (int(name[1]) - 1)
