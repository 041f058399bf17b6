import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def ctx():
    import octane_b200 as ob
    c = ob.Context(0)          # OctaneError(ENODEV) without a GPU: the product has no fallback
    yield c
    c.close()


def load_golden(name):
    """tests/golden/<name>.npz; OCTANE_GOLDEN_EXTRA names a second directory (fixtures generated earlier in the
    same GPU call, before they are committed)"""
    import numpy as np
    dirs = [GOLDEN] + ([os.environ["OCTANE_GOLDEN_EXTRA"]] if os.environ.get("OCTANE_GOLDEN_EXTRA") else [])
    for d in dirs:
        path = os.path.join(d, name + ".npz")
        if os.path.exists(path):
            return dict(np.load(path))
    pytest.skip(f"fixture {name}.npz not generated yet (tests/golden/make_golden*.py)")


def parity_report(test, **numbers):
    """append one line of measured differences to gpurun_out/parity_report.jsonl (when that directory exists):
    the figures DESIGN.md quotes come from there"""
    import json
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
        f.write(json.dumps(dict(test=test, **{k: (float(v) if hasattr(v, "__float__") else v) for k, v in numbers.items()})) + "\n")
