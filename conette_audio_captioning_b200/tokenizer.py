"""Host-side ids -> text, mirroring the reference tokenizer's *decode* path only.

The reference keeps detokenisation in Python (``AACTokenizer.decode_rec`` tokenization/aac_tokenizer.py:364-388 ->
``decode_batch`` :327-362 -> ``detokenize_batch`` :197-209 with the post-decoding normalisers of
tokenization/normalizers.py: strip special tokens, collapse spaces, fix spaces before punctuation, lowercase, strip).
The north star keeps that ("behind the existing Python preprocessing and tokenizer"), so this module is a thin,
dependency-free equivalent driven by the id->token table; a real reference ``AACTokenizer`` object can be passed to
``CoNeTTEModel`` instead and is then used as is.  Fitting / encoding text is training-side and out of scope.
"""
from __future__ import annotations

import re
from typing import Iterable, List, Sequence, Union

import torch

SPECIAL_TOKENS = ("<pad>", "<bos>", "<eos>", "<unk>")  # reference tokenization/constants.py:15

# NLTK 3.8.1 English stop-word list (public corpus data; reference pl_modules/common.py:270 reads it through nltk).
ENGLISH_STOPWORDS = tuple(
    (
        "i me my myself we our ours ourselves you you're you've you'll you'd your yours yourself yourselves he him his "
        "himself she she's her hers herself it it's its itself they them their theirs themselves what which who whom "
        "this that that'll these those am is are was were be been being have has had having do does did doing a an the "
        "and but if or because as until while of at by for with about against between into through during before after "
        "above below to from up down in out on off over under again further then once here there when where why how "
        "all any both each few more most other some such no nor not only own same so than too very s t can will just "
        "don don't should should've now d ll m o re ve y ain aren aren't couldn couldn't didn didn't doesn doesn't "
        "hadn hadn't hasn hasn't haven haven't isn isn't ma mightn mightn't mustn mustn't needn needn't shan shan't "
        "shouldn shouldn't wasn wasn't weren weren't won won't wouldn wouldn't"
    ).split()
)

_RE_SPECIAL = re.compile("(" + "|".join(SPECIAL_TOKENS) + ")")
_RE_SPACES = re.compile(" +")
_RE_SPACE_BEFORE_PUNCT = re.compile(r'\s+([,.!?;:"\'])')
_RE_HYPHEN_SPACES = re.compile(r"(\s*)(\-)(\s*)")


class IdTokenizer:
    """id -> token table with the reference's post-decoding normalisation (lowercase tokenizer, word level)."""

    pad_token_id, bos_token_id, eos_token_id, unk_token_id = 0, 1, 2, 3

    def __init__(self, itos: Sequence[str]) -> None:
        self.itos: List[str] = list(itos)
        self.stoi = {t: i for i, t in enumerate(self.itos)}
        if tuple(self.itos[:4]) != SPECIAL_TOKENS:
            raise ValueError(f"vocabulary must start with {SPECIAL_TOKENS} (found {self.itos[:4]})")

    def get_vocab_size(self) -> int:
        return len(self.itos)

    def has(self, token: str) -> bool:
        return token in self.stoi

    def token_to_id(self, token: str) -> int:
        return self.stoi[token]

    def id_to_token(self, idx: int) -> str:
        return self.itos[int(idx)]

    @staticmethod
    def _normalize(sentence: str) -> str:
        # order follows _get_post_decoding_normalizers (aac_tokenizer.py:953-963): CleanSpecialTokens,
        # CleanSpacesBeforePunctuation, Strip, CleanDoubleSpaces, CleanHyphenSpaces, Lowercase
        sentence = _RE_SPECIAL.sub("", sentence)
        sentence = _RE_SPACE_BEFORE_PUNCT.sub(r"\1", sentence)
        sentence = sentence.strip()
        sentence = _RE_SPACES.sub(" ", sentence)
        sentence = _RE_HYPHEN_SPACES.sub(r"\2", sentence)
        return sentence.lower()

    def decode_batch(self, sentences: Iterable[Iterable[int]]) -> List[str]:
        return [self._normalize(" ".join(self.itos[int(t)] for t in sent)) for sent in sentences]

    # ---- batched fast path ---------------------------------------------------------------------
    # With ~12 ms of GPU time per 64-clip batch the regex pipeline above (3.6 ms for 64 + 192 sentences) is a visible part
    # of a captioning call.  For a sentence whose tokens are all "plain" (lower-case, none of the characters any normaliser
    # looks at) or one of the four special tokens, the normalisers reduce to "drop the specials, join with one space":
    # that is done with one vectorised table gather for the whole id tensor.  Any other sentence takes the regex path.
    def _tables(self):
        if getattr(self, "_tab", None) is None:
            import numpy as np

            plain = re.compile(r"^[a-z0-9_]+$")
            skip = np.array([t in SPECIAL_TOKENS for t in self.itos], dtype=bool)
            ok = np.array([bool(plain.match(t)) for t in self.itos], dtype=bool) | skip
            self._tab = (np.array(self.itos, dtype=object), skip, ok)
        return self._tab

    def decode_tensor(self, ids: torch.Tensor) -> Union[str, list]:
        """(..., L) integer tensor -> nested lists of sentences, identical to ``decode_rec`` on the same ids."""
        import numpy as np

        itos_np, skip, ok = self._tables()
        arr = ids.detach().cpu().numpy()
        flat = arr.reshape(-1, arr.shape[-1]) if arr.ndim > 1 else arr.reshape(1, -1)
        toks = itos_np[flat]
        keep = ~skip[flat]
        fast = ok[flat].all(axis=1)
        out: List[str] = []
        for r in range(flat.shape[0]):
            if fast[r]:
                out.append(" ".join(toks[r][keep[r]]))
            else:
                out.append(self._normalize(" ".join(toks[r])))
        res = np.empty(len(out), dtype=object)
        res[:] = out
        return res.reshape(arr.shape[:-1]).tolist() if arr.ndim > 1 else out[0]

    def decode_rec(self, nested: Union[torch.Tensor, list]) -> Union[str, list]:
        if isinstance(nested, torch.Tensor):
            if nested.ndim >= 1 and nested.shape[-1] > 0 and not nested.is_floating_point():
                return self.decode_tensor(nested)
            nested = nested.tolist()
        if len(nested) > 0 and not isinstance(nested[0], (list, tuple)):
            return self.decode_batch([nested])[0]
        if all(len(s) == 0 or not isinstance(s[0], (list, tuple)) for s in nested):
            return self.decode_batch(nested)
        return [self.decode_rec(s) for s in nested]


def make_forbid_rep_mask(itos: Sequence[str], mode: str = "content_words", stopwords: "Sequence[str] | None" = None) -> "Tensor | None":
    """Host mirror of reference ``get_forbid_rep_mask`` (pl_modules/common.py:222-299)."""
    if mode == "none":
        return None
    if mode == "all":
        return torch.ones(len(itos), dtype=torch.bool)
    if mode == "content_words":
        if stopwords is None:
            stopwords = ENGLISH_STOPWORDS
        sw = set(stopwords)
        return torch.tensor([tok not in sw for tok in itos], dtype=torch.bool)
    raise ValueError(
        f"Invalid argument forbid_rep_mode={mode!r}. (expected one of ('none', 'all', 'content_words'))"
    )
