import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device and the built library: on a host without them they are skipped, not failed."""
    import torch

    from conette_audio_captioning_b200 import _lib

    if torch.cuda.is_available() and _lib.LIB_PATH.exists():
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (sm_100a) and conette_audio_captioning_b200/lib/libconette_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def small_sd():
    """Synthetic state dict with a small vocabulary (V = 4 + 7 + 300 + 7 = 318) and an EOS-biased classifier."""
    from conette_audio_captioning_b200 import synth

    return synth.make_state_dict(seed=1234, n_words=300, eos_bias=3.0)
