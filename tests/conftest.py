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


def load_golden(name):
    """-> (meta dict, {group: {key: ndarray}})"""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    groups = {}
    for k in z.files:
        if k == "meta":
            continue
        g, key = k.split("/", 1)
        groups.setdefault(g, {})[key] = z[k]
    return meta, groups


def golden_cases(kind):
    out = []
    for f in sorted(os.listdir(GOLDEN)):
        if f.startswith(kind + "_") and f.endswith(".npz"):
            out.append(f[:-4])
    return out


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import simmodel_b200  # noqa: F401
    from simmodel_b200 import _lib
    _lib.lib()            # must load: no fallback
    return torch.device("cuda:0")
