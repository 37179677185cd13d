"""pytest configuration: registers the `gpu` marker and puts the repo root on sys.path.

`-m "not gpu"` tests run on CPU (oracle vs golden vectors, host logic, C-ABI symbol checks,
gloo world_size-2 sharding); `-m gpu` tests are the parity tests proper and call the CUDA
path through the C-ABI library.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # a fresh checkout has no built artefacts: compile the CUDA library (nvcc cross-compiles
    # without a GPU) and the C oracle once, before collection
    from tiler_slider_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH) and not os.environ.get("TS_LIB_PATH"):
        _lib.build()
    from oracle import oracle as _orc
    _orc.build()


@pytest.fixture(scope="session")
def golden_scenarios():
    with open(os.path.join(GOLDEN, "scenarios.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_misc():
    with open(os.path.join(GOLDEN, "misc.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_rollouts():
    z = np.load(os.path.join(GOLDEN, "rollouts.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    out = []
    for m in meta:
        rec = dict(m)
        for k in ("blocked", "tiles", "targets", "actions", "pos", "flags", "count", "obs_final"):
            rec[k] = z[f"{m['tag']}_{k}"]
        out.append(rec)
    return out
