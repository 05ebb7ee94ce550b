import glob
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith("multiclass"))


def load_golden(name):
    f = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(f["meta"]))
    return meta, {k: f[k] for k in f.files if k != "meta"}


@pytest.fixture(scope="session")
def lib_built():
    """Build the shared library once (no-op when the shipped .so is current)."""
    from wav2vec_contr_loss_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def cuda_device(lib_built):
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test selected but no CUDA device is visible")
    return torch.device("cuda:0")
