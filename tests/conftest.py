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
    # a fresh checkout has no built artefacts (they are git-ignored): build the CUDA library (nvcc
    # cross-compiles without a GPU) and the CPU oracle once, exactly as __graft_entry__.build() does
    lib = os.path.join(ROOT, "3dpointcloudattack_b200", "libpcdist.so")
    orc = os.path.join(ROOT, "oracle", "libpcd_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(orc)):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_graft_entry", os.path.join(ROOT, "__graft_entry__.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_inf(a, b):
    """max|a-b| / max|b| -- the 'relative' of the 1e-5 gradient/distance tolerance."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return np.abs(a - b).max() / (den if den > 0 else 1.0)


@pytest.fixture(autouse=True)
def _fp32_victims():
    """Parity is defined against the reference in fp32 (allow_tf32 off): the tiny victims of the loop tests
    are cuDNN convolutions / cuBLAS GEMMs, which default to TF32 on this GPU and moved the logits by 1e-3."""
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
