import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture
def golden():
    return load_golden


# ---- parity ledger: every GPU parity comparison records its measured error next to its bar --------------------
_PARITY_ROWS = []


def parity_log(what, err, tol):
    test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0].split("::")[-1]
    _PARITY_ROWS.append({"test": test, "what": what, "err": float(err), "tol": float(tol)})


def pytest_sessionfinish(session, exitstatus):
    """Writes gpurun_out/parity_ledger.json (copied to profiles/ by hand after a GPU run)."""
    if not _PARITY_ROWS:
        return
    import json

    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_ledger.json"), "w") as f:
            json.dump(_PARITY_ROWS, f, indent=1)
    except OSError:
        pass
