import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", message=".*window was not provided.*")
warnings.filterwarnings("ignore", message=".*torch.meshgrid.*")

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))
    return load


@pytest.fixture(scope="session")
def weights():
    """state_dict cache keyed by (kind, seed)."""
    from oracle import uformer as O
    from image_in_speech_watermarking_b200 import synthetic as SY
    cache = {}

    def get(kind, seed=0):
        if (kind, seed) not in cache:
            cache[(kind, seed)] = SY.init_state_dict(O.state_dict_schema(), kind, seed)
        return cache[(kind, seed)]
    return get
