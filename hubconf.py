"""torch.hub entry point, as the reference's hubconf.py:1-14: ``torch.hub.load(<repo dir>, "conette", source="local", ...)``."""
dependencies = ["torch"]

from conette_audio_captioning_b200.predict import conette  # noqa: E402,F401

__all__ = ["conette"]
