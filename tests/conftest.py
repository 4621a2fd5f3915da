import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "music-fader-nets_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_FILES = sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not os.path.basename(p).startswith(("sib_", "glsr_")))   # glsr_*: tests/test_oracle_glsr.py
SIBLING_GOLDEN_FILES = sorted(glob.glob(os.path.join(GOLDEN_DIR, "sib_*.npz")))     # sibling models (oracle/gen_golden.py make_sibling_case)
GOLDEN_SEEDS = {"gmvae_H16_Z8_B3_T12": 10, "vae_H16_Z8_B4_T10": 20, "gmvae_H32_Z16_B2_T9": 30, "gmvae_H64_Z8_B3_T10": 70}   # oracle/gen_golden.py main()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(path):
    import torch
    z = np.load(path, allow_pickle=False)
    g = {k: z[k] for k in z.files}
    g["name"] = os.path.basename(path)[:-4]
    g["seed"] = GOLDEN_SEEDS.get(g["name"])
    g["variant"] = "gmvae" if os.path.basename(path).startswith("gmvae") else "vae"
    g["weights"] = {k[2:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("w/")}
    return g


@pytest.fixture(params=GOLDEN_FILES, ids=[os.path.basename(p)[:-4] for p in GOLDEN_FILES])
def golden(request):
    return load_golden(request.param)


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built on demand here on the CPU box; prebuilt on the GPU box)."""
    import __graft_entry__ as ge
    ge.build()
    import fadernets_b200
    return fadernets_b200.LIB.load()
