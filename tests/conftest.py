import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(path):
    import torch
    z = np.load(path, allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["variant"] = "gmvae" if os.path.basename(path).startswith("gmvae") else "vae"
    g["weights"] = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w/")}
    return g


@pytest.fixture(params=GOLDEN_FILES, ids=[os.path.basename(p)[:-4] for p in GOLDEN_FILES])
def golden(request):
    return load_golden(request.param)
