import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"
for path in (ROOT, ROOT / "oracle"):
    if str(path) not in sys.path:
        sys.path.insert(0, str(path))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def golden_cases(kind):
    out = []
    for path in sorted(GOLDEN.iterdir()):
        params_file = path / "params.json"
        if params_file.is_file():
            params = json.loads(params_file.read_text())
            if params["kind"] == kind:
                out.append(pytest.param(path, params, id=path.name))
    return out


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
