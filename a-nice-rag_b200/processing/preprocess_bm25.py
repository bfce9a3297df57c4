"""Query tokeniser with the contract of the reference's ``src/processing/preprocess_bm25.py``
(``preprocess_text`` :33-52): lower-case, strip ``string.punctuation``, tokenise, drop English
stop-words / numeric tokens / tokens of length <= 1, optionally WordNet-lemmatise.

String work stays on the CPU; kernel work starts at tokens.  With NLTK
installed this calls the same NLTK functions as the reference.  Offline (no NLTK, no
corpora) it uses a whitespace tokeniser -- equal to ``word_tokenize`` once punctuation has
been stripped -- and NLTK's English stop-word list restated below; WordNet lemmatisation has
no offline equivalent, so ``use_lemmatization=True`` then leaves tokens unchanged and says so
once in the log.  Callers that need exact lemmas pass pre-tokenised queries
(``bm25_search_preprocessed``), which is what the reference's evaluator does.
"""
from __future__ import annotations

import logging
import string
from typing import List

logger = logging.getLogger(__name__)

try:  # pragma: no cover - NLTK is absent in the offline image
    from nltk.corpus import stopwords as _nltk_stopwords
    from nltk.stem import WordNetLemmatizer as _WordNetLemmatizer
    from nltk.tokenize import word_tokenize as _word_tokenize
    _nltk_stopwords.words("english")
    _word_tokenize("probe text")
    _HAVE_NLTK = True
except Exception:
    _HAVE_NLTK = False

# NLTK 3.8 `stopwords.words("english")` (179 entries)
ENGLISH_STOPWORDS = frozenset("""
i me my myself we our ours ourselves you you're you've you'll you'd your yours yourself
yourselves he him his himself she she's her hers herself it it's its itself they them their
theirs themselves what which who whom this that that'll these those am is are was were be been
being have has had having do does did doing a an the and but if or because as until while of at
by for with about against between into through during before after above below to from up down
in out on off over under again further then once here there when where why how all any both
each few more most other some such no nor not only own same so than too very s t can will just
don don't should should've now d ll m o re ve y ain aren aren't couldn couldn't didn didn't
doesn doesn't hadn hadn't hasn hasn't haven haven't isn isn't ma mightn mightn't mustn mustn't
needn needn't shan shan't shouldn shouldn't wasn wasn't weren weren't won won't wouldn
wouldn't
""".split())

_PUNCT_TABLE = str.maketrans("", "", string.punctuation)
_UNICODE_QUOTES = str.maketrans({c: " " for c in "\u2018\u2019\u201c\u201d\u00ab\u00bb"})
# the Treebank tokenizer behind word_tokenize splits these contractions (the ones that survive
# punctuation stripping); both halves are then ordinary tokens ("cannot" -> "can", "not")
_TREEBANK_SPLITS = {"cannot": ("can", "not"), "gimme": ("gim", "me"), "gonna": ("gon", "na"),
                    "gotta": ("got", "ta"), "lemme": ("lem", "me"), "wanna": ("wan", "na")}
_warned = False


def preprocess_text(text: str, use_lemmatization: bool = False) -> List[str]:
    if not text:
        return []
    text = text.lower().translate(_PUNCT_TABLE)
    if _HAVE_NLTK:  # pragma: no cover
        tokens = _word_tokenize(text)
        stop = set(_nltk_stopwords.words("english"))
    else:
        tokens = []
        # word_tokenize isolates typographic quotes (string.punctuation is ASCII only)
        for tok in text.translate(_UNICODE_QUOTES).split():
            tokens.extend(_TREEBANK_SPLITS.get(tok, (tok,)))
        stop = ENGLISH_STOPWORDS
    tokens = [t for t in tokens if t not in stop and not t.isnumeric() and len(t) > 1]
    if use_lemmatization:
        if _HAVE_NLTK:  # pragma: no cover
            lem = _WordNetLemmatizer()
            tokens = [lem.lemmatize(t) for t in tokens]
        else:
            global _warned
            if not _warned:
                logger.warning("WordNet is unavailable offline: tokens are not lemmatised")
                _warned = True
    return tokens
