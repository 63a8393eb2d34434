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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The C-ABI library must exist for every test session (nvcc cross-compiles without a GPU)."""
    # loaded by path: importing the package needs libdmp.so to exist already
    import importlib.util
    spec = importlib.util.spec_from_file_location("_snac_b200_build", os.path.join(ROOT, "snac_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build_lib()


TRACE_NAMES = sorted(f[len("trace_"):-len(".npz")] for f in os.listdir(GOLDEN) if f.startswith("trace_"))


def load_trace(name):
    z = np.load(os.path.join(GOLDEN, "trace_%s.npz" % name))
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


def load_plans(dim, density="dense", split="train"):
    """Plan dataset fixtures in the reference's array format (float64)."""
    z = np.load(os.path.join(GOLDEN, "plans_packed.npz"))
    if dim == 1:
        return z["1d_%s" % split].astype(np.float64)
    a = z["%dd_%s_%s" % (dim, density, split)]
    bits = np.unpackbits(a, axis=1)[:, :400].reshape(len(a), 20, 20).astype(np.float64)
    out = np.zeros((len(a), 26, 26))
    out[:, 3:23, 3:23] = bits * (6.0 if dim == 3 else 1.0)
    return out


def trace_env_spec(meta):
    """(dim, dynamic, plan_choose, plans) for a golden trace."""
    dim = int(meta["dim"][0])
    dynamic = meta["kind"] == "dynamic"
    if dynamic:
        return dim, True, 0, load_plans(dim, meta["kw"]["density"], meta["kw"]["split"])
    return dim, False, meta["kw"]["plan_choose"], None
