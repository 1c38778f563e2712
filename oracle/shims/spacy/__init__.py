"""Oracle shim for ``spacy==3.7.2``.  TEST INFRASTRUCTURE ONLY.

``spacy.load(name).tokenizer(text)`` -> tokens with ``.text`` (reference tokenization/tokenizers/spacy.py:22,41-47);
whitespace split is enough to *fit* the synthetic vocabulary; ``decode_rec`` never tokenises.
"""


class _Token:
    __slots__ = ("text",)

    def __init__(self, text: str) -> None:
        self.text = text


class _Model:
    def tokenizer(self, sentence: str) -> list:
        return [_Token(w) for w in sentence.split()]


def load(name: str, *args, **kwargs) -> _Model:
    return _Model()
