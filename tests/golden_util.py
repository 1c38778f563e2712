"""Helpers shared by the golden-fixture tests (fixtures are produced by oracle/make_golden.py from the real reference)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SD_KW = dict(seed=1234, n_words=300, eos_bias=3.0)
CHECK_KEYS = (
    "preprocessor.encoder.stages.2.4.pwconv1.weight",
    "preprocessor.encoder.bn0.running_mean",
    "model.decoder.layers.3.linear2.weight",
    "model.decoder.classifier.bias",
)


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def checksum(sd):
    return np.array([float(sd[k].double().sum()) for k in CHECK_KEYS] + [float(sd[k].double().abs().sum()) for k in CHECK_KEYS])


def assert_weights_match(sd, fixture):
    np.testing.assert_allclose(checksum(sd), fixture["weights_checksum"], rtol=1e-9, atol=1e-9,
                               err_msg="synthetic weights differ from the ones the golden fixtures were made with")


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))
