import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle():
    from oracle import oracle as orc
    orc.build()


@pytest.fixture(scope="session")
def golden_dir():
    import pathlib
    return pathlib.Path(__file__).parent / "golden"


class _ParityLog:
    """err / floor pairs of the GPU parity tests; written as JSON to $GMD_PARITY_OUT at session end (the run on the
    B200 box sets it to gpurun_out/parity.json, committed as profiles/parity_r2.json)"""

    def __init__(self):
        self.rows = []

    def add(self, test, **kw):
        import numpy as np

        def plain(x):
            if isinstance(x, (list, tuple, np.ndarray)):
                return [plain(y) for y in x]
            if isinstance(x, (np.floating, np.integer)):
                return x.item()
            return x
        row = {"test": test}
        row.update({k: plain(v) for k, v in kw.items()})
        self.rows.append(row)
        print("PARITY", row)


_PARITY = _ParityLog()


@pytest.fixture(scope="session")
def parity_log():
    return _PARITY


def pytest_sessionfinish(session, exitstatus):
    out = os.environ.get("GMD_PARITY_OUT")
    if out and _PARITY.rows:
        import json
        prev = []
        if os.path.exists(out):
            try:
                prev = json.load(open(out))
            except Exception:
                prev = []
        names = {r["test"] for r in _PARITY.rows}
        with open(out, "w") as f:
            json.dump([r for r in prev if r.get("test") not in names] + _PARITY.rows, f, indent=1)
