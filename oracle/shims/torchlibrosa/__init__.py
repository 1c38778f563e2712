"""Oracle shim for the un-vendored dependency ``torchlibrosa==0.1.0`` (reference requirements.txt:12).

TEST INFRASTRUCTURE ONLY.  The reference imports ``torchlibrosa.stft.{Spectrogram,LogmelFilterBank}`` and
``torchlibrosa.augmentation.SpecAugmentation`` at ``src/conette/nn/encoders/convnext.py:10-11`` but the package is
neither vendored under /root/reference nor installed in this image (no network).  This shim restates the package's
published algorithm (Conv1d-DFT STFT, ``librosa.filters.mel`` Slaney filter bank, ``power_to_db``) with the same
module / parameter names and shapes, so the reference's *unmodified* ``convnext.py`` can run as the CPU oracle.
"""
from . import stft, augmentation  # noqa: F401
