import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import torch
    return torch.load(os.path.join(GOLDEN, "reference_outputs.pt"), weights_only=False)


@pytest.fixture(scope="session")
def idf_vector():
    import numpy as np
    import torch
    return torch.from_numpy(np.load(os.path.join(GOLDEN, "idf_vector_f32.npy")))
