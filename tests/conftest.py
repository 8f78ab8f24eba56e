import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than ~20 s on one CPU core")


@pytest.fixture(scope="session")
def oracle():
    from _refio import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import json
    from _refio import GOLDEN
    return json.loads((GOLDEN / "golden.json").read_text())


@pytest.fixture(scope="session")
def fixtures():
    """name -> list[Structure] for every packed fixture db, plus 'queries' as a dict by name."""
    from _refio import GOLDEN, read_packed
    out = {p.stem: read_packed(p) for p in GOLDEN.glob("*.satsdb")}
    out["queries_by_name"] = {q.name: q for q in out["queries"]}
    return out
