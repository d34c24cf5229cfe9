import glob
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # keep the in-tree CUDA library in step with its sources (nvcc cross-compiles sm_100a without a GPU);
    # if nvcc is unavailable the tests that need the library fail loudly on their own
    try:
        from lanpaint_b200 import build as _b
        _b.build(force=False)
    except Exception as e:  # pragma: no cover
        print(f"[conftest] could not (re)build liblanpaint_b200.so: {e}")


def golden_names():
    names = (os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return sorted(n for n in names if not n.startswith(("aux_", "node_")))   # engine-seam cases only


def load_golden(name):
    """-> dict of numpy arrays + 'meta' dict; mask is expanded back to the full fp32 latent shape."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files if k != "meta"}
    g["meta"] = json.loads(str(z["meta"]))
    g["mask_full"] = np.broadcast_to(g["mask"].astype(np.float32), g["x"].shape).copy()
    return g


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
