import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


@pytest.fixture(scope="session")
def small_sd():
    """Synthetic state dict with a small vocabulary (V = 4 + 7 + 300 + 7 = 318) and an EOS-biased classifier."""
    from conette_audio_captioning_b200 import synth

    return synth.make_state_dict(seed=1234, n_words=300, eos_bias=3.0)
