"""Oracle shim for ``nltk==3.8.1``.  TEST INFRASTRUCTURE ONLY.

Only ``nltk.corpus.stopwords.words("english")`` (reference pl_modules/common.py:11,270) and ``nltk.download`` are needed.
With a synthetic vocabulary the stop-word list only decides which token ids are exempt from the no-repeat rule; the
resulting ``forbid_rep_mask`` tensor is what both the oracle and the CUDA path consume.
"""
from . import corpus  # noqa: F401


def download(*args, **kwargs) -> bool:
    return True
