"""Oracle shim for the un-vendored dependency ``torchoutil~=0.3.0`` (reference requirements.txt:13).

TEST INFRASTRUCTURE ONLY.  Only the index / mask helpers the reference's inference path calls are restated
(SURVEY.md Appendix H lists their one-line semantics and call sites: beam.py:10-15, conette.py:9-13,
preprocessor.py:11, pad.py:8, get.py:9, common.py:15, model.py:10-11).
"""
from . import nn, utils  # noqa: F401
